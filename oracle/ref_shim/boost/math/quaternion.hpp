// Stand-in for a header of an absent third-party library (Boost / OpenBabel), written for oracle/_ref only:
// it lets the reference's own Vina headers compile where they lie under /root/reference. No arithmetic lives here.
#pragma once
namespace boost { namespace math {
// storage only: the reference converts to/from this type but does its quaternion arithmetic in its own qt class
template <class T> class quaternion { T a, b, c, d; public: quaternion(T a_ = 0, T b_ = 0, T c_ = 0, T d_ = 0) : a(a_), b(b_), c(c_), d(d_) {}
  T R_component_1() const { return a; } T R_component_2() const { return b; } T R_component_3() const { return c; } T R_component_4() const { return d; } };
} }
