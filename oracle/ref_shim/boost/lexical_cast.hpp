// Stand-in for a header of an absent third-party library (Boost / OpenBabel), written for oracle/_ref only:
// it lets the reference's own Vina headers compile where they lie under /root/reference. No arithmetic lives here.
#pragma once
#include <sstream>
#include <string>
#include <stdexcept>
namespace boost {
struct bad_lexical_cast : std::runtime_error { bad_lexical_cast() : std::runtime_error("bad lexical cast") {} };
template <class T, class S> T lexical_cast(const S& s) {
  std::stringstream ss; ss << s; T t; ss >> t; if (ss.fail()) throw bad_lexical_cast(); return t; }
}
