// Stand-in for a header of an absent third-party library (Boost / OpenBabel), written for oracle/_ref only:
// it lets the reference's own Vina headers compile where they lie under /root/reference. No arithmetic lives here.
#pragma once
#include <type_traits>
#define BOOST_DEDUCED_TYPENAME typename
#define BOOST_STATIC_CONSTANT(t, a) static const t a
namespace boost {
using std::is_fundamental; using std::is_class; using std::is_array; using std::is_enum;
template <class B, class D> using is_base_and_derived = std::is_base_of<B, D>;
namespace mpl { struct integral_c_tag {}; template <int N> struct int_ { static const int value = N; typedef int_ type; };
template <class C, class A, class B> struct eval_if { typedef typename std::conditional<C::value, A, B>::type::type type; };
template <class A, class B> struct or_ : std::integral_constant<bool, A::value || B::value> {}; }
namespace serialization { struct basic_traits {}; enum level_type { not_serializable = 0, primitive_type = 1, object_serializable = 2 };
template <class T> struct implementation_level_impl; } }
namespace boost { namespace serialization { class access {}; template <class B, class D> B& base_object(D& d) { return d; }
template <class A, class T> void split_free(A&, T&, unsigned) {} template <class A, class T> void split_member(A&, T&, unsigned) {}
template <class T> T& make_nvp(const char*, T& t) { return t; }
template <class T> struct version { static const int value = 0; };
struct item_version_type { unsigned v; explicit item_version_type(unsigned v_ = 0) : v(v_) {} };
namespace detail { template <class A, class T> struct stack_construct { T t; stack_construct(A&, item_version_type) {} T& reference() { return t; } }; }
}
namespace archive { struct library_version_type { unsigned v; explicit library_version_type(unsigned v_ = 0) : v(v_) {}
  bool operator<(const library_version_type& o) const { return v < o.v; } };
// never instantiated by oracle/_ref (no archive is read or written there)
class binary_iarchive; class binary_oarchive; class text_iarchive; class text_oarchive; }
}
#define BOOST_SERIALIZATION_NVP(x) x
#define BOOST_SERIALIZATION_SPLIT_FREE(T)
#define BOOST_SERIALIZATION_SPLIT_MEMBER()
#define BOOST_CLASS_VERSION(T, N)
#define BOOST_STATIC_ASSERT(x) static_assert(x, #x)
