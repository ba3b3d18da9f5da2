// Stand-in for a header of an absent third-party library (Boost / OpenBabel), written for oracle/_ref only:
// it lets the reference's own Vina headers compile where they lie under /root/reference. No arithmetic lives here.
#pragma once
#include <vector>
#include <memory>
#include <cstddef>
#include <algorithm>
namespace boost {
// owning vector of heap objects with value-style element access (the subset the reference uses)
template <class T> class ptr_vector {
  std::vector<std::unique_ptr<T>> v;
 public:
  ptr_vector() {}
  ptr_vector(const ptr_vector&) = delete;
  ptr_vector& operator=(const ptr_vector&) = delete;
  void push_back(T* p) { v.emplace_back(p); }
  std::size_t size() const { return v.size(); }
  bool empty() const { return v.empty(); }
  T& operator[](std::size_t i) { return *v[i]; }
  const T& operator[](std::size_t i) const { return *v[i]; }
  T& back() { return *v.back(); }
  const T& back() const { return *v.back(); }
  T& front() { return *v.front(); }
  void pop_back() { v.pop_back(); }
  void clear() { v.clear(); }
  void resize(std::size_t n) { v.resize(n); }
  // boost: std::sort over the pointers, comparing the pointees with operator<
  void sort() { std::sort(v.begin(), v.end(), [](const std::unique_ptr<T>& a, const std::unique_ptr<T>& b) { return *a < *b; }); }
};
}
