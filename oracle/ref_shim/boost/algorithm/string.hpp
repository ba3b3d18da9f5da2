// Stand-in for a header of an absent third-party library (Boost / OpenBabel), written for oracle/_ref only:
// it lets the reference's own headers compile where they lie under /root/reference. No arithmetic lives here.
#pragma once
#include <string>
#include "boost/algorithm/string/predicate.hpp"
namespace boost { namespace algorithm {
// templates, as in Boost: the reference's own non-template starts_with (common.h) wins overload resolution where both are visible
template <class A, class B> bool starts_with(const A& s, const B& p) { return std::string(s).compare(0, std::string(p).size(), p) == 0; }
template <class A, class B> bool ends_with(const A& s_, const B& p_) { const std::string s(s_), p(p_); return s.size() >= p.size() && s.compare(s.size() - p.size(), p.size(), p) == 0; }
inline std::string replace_last_copy(const std::string& s, const std::string& what, const std::string& with) {
  const size_t at = s.rfind(what);
  if (at == std::string::npos) return s;
  std::string r = s; r.replace(at, what.size(), with); return r;
}
} }
