// Stand-in for a header of an absent third-party library (Boost), written for oracle/_ref only: it lets the reference's own
// Vina sources compile where they lie under /root/reference.
//
// THIS one is not a neutral container mapping, so read it: Boost's mt19937 + distribution algorithms cannot be reproduced here
// (Boost is absent), and neither the oracle nor the device kernels claim its stream. What this header provides
// under the names the reference uses is the generator the restatement and the device kernels DO use -- xorshift32 with the
// mappings of oracle/vina_mc_ref.c (rng_next / rng_fl / rng_int) -- so that the reference's monte_carlo.cpp, mutate.cpp and
// conf.h run on the same random stream as the restatement and whole Monte-Carlo chains can be compared. The type is called
// mt19937 only because lib/random.h:29 spells it that way.
#pragma once
#include <cmath>
#include <cstdint>
namespace boost {
class mt19937 {
  uint32_t s;
 public:
  typedef uint32_t result_type;
  explicit mt19937(uint32_t seed = 1u) : s(seed ? seed : 1u) {}
  uint32_t operator()() { uint32_t x = s; x ^= x << 13; x ^= x >> 17; x ^= x << 5; s = x; return x; }
  uint32_t state() const { return s; }
  static constexpr uint32_t min() { return 1u; }
  static constexpr uint32_t max() { return 0xffffffffu; }
};
template <class T = double> class uniform_real {
  T a, b;
 public:
  typedef T result_type;
  uniform_real(T a_ = 0, T b_ = 1) : a(a_), b(b_) {}
  template <class E> T operator()(E& e) { return a + (b - a) * ((T)(e() >> 8) * (T)(1.0f / 16777216.0f)); }
};
template <class T = int> class uniform_int {
  T a, b;
 public:
  typedef T result_type;
  uniform_int(T a_ = 0, T b_ = 9) : a(a_), b(b_) {}
  template <class E> T operator()(E& e) { return a + (T)(e() % (uint32_t)(b - a + 1)); }
};
// only conf::randomize's random_orientation draws normals (lib/quaternion.cu:81-94); Box-Muller, one normal per two draws
template <class T = double> class normal_distribution {
  T mean, sigma;
 public:
  typedef T result_type;
  normal_distribution(T m = 0, T s = 1) : mean(m), sigma(s) {}
  template <class E> T operator()(E& e) {
    const double u1 = ((double)(e() >> 8) + 0.5) * (1.0 / 16777216.0), u2 = (double)(e() >> 8) * (1.0 / 16777216.0);
    return mean + sigma * (T)(std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2));
  }
};
template <class E, class D> class variate_generator {
  E e;
  D d;
 public:
  variate_generator(E e_, D d_) : e(e_), d(d_) {}
  typename D::result_type operator()() { return d(e); }
};
}
