// Stand-in for a header of an absent third-party library (Boost / OpenBabel), written for oracle/_ref only:
// it lets the reference's own Vina headers compile where they lie under /root/reference. No arithmetic lives here.
#pragma once
#include <optional>
namespace boost {
template <class T> class optional : public std::optional<T> { public:
  optional() {} optional(const T& t) : std::optional<T>(t) {}
  optional& operator=(const T& t) { std::optional<T>::operator=(t); return *this; }
  const T& get() const { return **this; } T& get() { return **this; }
  T get_value_or(const T& d) const { return this->has_value() ? **this : d; }
  bool is_initialized() const { return this->has_value(); } };
}
