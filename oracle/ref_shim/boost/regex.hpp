// Stand-in for a header of an absent third-party library (Boost / OpenBabel), written for oracle/_ref only:
// it lets the reference's own Vina headers compile where they lie under /root/reference. No arithmetic lives here.
#pragma once
#include <regex>
#include <string>
namespace boost {
class regex : public std::regex {
 public:
  static constexpr std::regex::flag_type perl = std::regex::ECMAScript;
  regex() {}
  regex(const std::string& s, std::regex::flag_type f = std::regex::ECMAScript) : std::regex(s, f) {}
  void assign(const std::string& s, std::regex::flag_type f = std::regex::ECMAScript) { std::regex::assign(s, f); }
};
using smatch = std::smatch;
inline bool regex_match(const std::string& s, smatch& m, const regex& r) { return std::regex_match(s, m, static_cast<const std::regex&>(r)); }
inline bool regex_match(const std::string& s, const regex& r) { return std::regex_match(s, static_cast<const std::regex&>(r)); }
}
