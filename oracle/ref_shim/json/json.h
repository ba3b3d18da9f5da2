// Stand-in for jsoncpp (absent third-party dependency), written for oracle/_ref only: the flat string / number / bool object that
// lib/torch_model.cpp reads from a model file's "metadata" entry.  No arithmetic beyond strtod.
#pragma once
#include <cstdlib>
#include <map>
#include <string>
namespace Json {
class Value {
  std::map<std::string, std::string> kv_;   // raw token per key: a decoded string, or the literal text of a number / true / false
  std::string raw_;
  friend class Reader;
 public:
  Value() {}
  explicit Value(const std::string& r) : raw_(r) {}
  bool isMember(const std::string& k) const { return kv_.count(k) != 0; }
  Value operator[](const std::string& k) const { auto it = kv_.find(k); return it == kv_.end() ? Value() : Value(it->second); }
  double asDouble() const { return std::strtod(raw_.c_str(), nullptr); }
  float asFloat() const { return (float)asDouble(); }
  bool asBool() const { return raw_ == "true" || (raw_ != "false" && !raw_.empty() && asDouble() != 0); }
  std::string asString() const { return raw_; }
};
class Reader {
  static void ws(const std::string& s, size_t& i) { while (i < s.size() && (s[i] == ' ' || s[i] == '\n' || s[i] == '\t' || s[i] == '\r')) i++; }
  static bool str(const std::string& s, size_t& i, std::string& out) {
    if (i >= s.size() || s[i] != '"') return false;
    for (i++; i < s.size() && s[i] != '"'; i++) {
      if (s[i] != '\\') { out += s[i]; continue; }
      if (++i >= s.size()) return false;
      switch (s[i]) { case 'n': out += '\n'; break; case 't': out += '\t'; break; case 'r': out += '\r'; break;
                      case 'u': return false; default: out += s[i]; }
    }
    return i++ < s.size();
  }
 public:
  bool parse(const std::string& s, Value& root) {
    size_t i = 0; ws(s, i);
    if (i >= s.size() || s[i] != '{') return false;
    i++;
    for (;;) {
      ws(s, i);
      if (i < s.size() && s[i] == '}') return true;
      std::string k, v;
      if (!str(s, i, k)) return false;
      ws(s, i);
      if (i >= s.size() || s[i] != ':') return false;
      i++; ws(s, i);
      if (i < s.size() && s[i] == '"') { if (!str(s, i, v)) return false; }
      else { while (i < s.size() && s[i] != ',' && s[i] != '}') v += s[i++]; while (!v.empty() && (v.back() == ' ' || v.back() == '\n')) v.pop_back(); }
      root.kv_[k] = v;
      ws(s, i);
      if (i < s.size() && s[i] == ',') i++;
    }
  }
};
}  // namespace Json
