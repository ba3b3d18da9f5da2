// Stand-in for a header of an absent third-party library (Boost / OpenBabel), written for oracle/_ref only:
// it lets the reference's own headers compile where they lie under /root/reference. No arithmetic lives here.
#pragma once
#include <sstream>
#include <string>
#include "boost/iostreams/filtering_stream.hpp"
namespace boost { namespace iostreams {
template <class Ch> struct basic_array_source { const Ch* b; size_t n; basic_array_source(const Ch* p, size_t len) : b(p), n(len) {} };
// a seekable read-only stream over the bytes (torch::jit::load seeks)
template <> class stream<basic_array_source<char>> : public std::istringstream {
 public:
  explicit stream(const basic_array_source<char>& s) : std::istringstream(std::string(s.b, s.n), std::ios::in | std::ios::binary) {}
};
} }
