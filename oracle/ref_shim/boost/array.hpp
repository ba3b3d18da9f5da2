// Stand-in for a header of an absent third-party library (Boost / OpenBabel), written for oracle/_ref only:
// it lets the reference's own Vina headers compile where they lie under /root/reference. No arithmetic lives here.
#pragma once
#include <cstddef>
namespace boost {
template <class T, std::size_t N> struct array {
  T elems[N];
  typedef T value_type; typedef T* iterator; typedef const T* const_iterator;
  T& operator[](std::size_t i) { return elems[i]; }
  const T& operator[](std::size_t i) const { return elems[i]; }
  static std::size_t size() { return N; }
  T* begin() { return elems; } T* end() { return elems + N; }
  const T* begin() const { return elems; } const T* end() const { return elems + N; }
  void assign(const T& v) { for (std::size_t i = 0; i < N; ++i) elems[i] = v; }
};
}
