// Stand-in for a header of an absent third-party library (Boost / OpenBabel), written for oracle/_ref only:
// it lets the reference's own Vina headers compile where they lie under /root/reference. No arithmetic lives here.
#pragma once
#include <iostream>
#include <sstream>
namespace boost { namespace iostreams {
struct input {}; struct output {}; struct gzip_decompressor {}; struct gzip_compressor {}; struct null_sink {};
// oracle/_ref performs no file I/O: this only has the members file.h names
template <class Mode> class filtering_stream : public std::stringstream { int n = 0; public:
  template <class T> void push(const T&) { ++n; } void pop() { --n; } bool empty() const { return n == 0; } };
template <class Dev> class stream : public std::stringstream { public: stream() {} template <class A> explicit stream(const A&) {} };
} }
