// Stand-in for a header of an absent third-party library (Boost / OpenBabel), written for oracle/_ref only:
// it lets the reference's own Vina headers compile where they lie under /root/reference. No arithmetic lives here.
#pragma once
#include <functional>
#include <string>
namespace boost {
template <class T> void hash_combine(std::size_t& seed, const T& v) { seed ^= std::hash<T>()(v) + 0x9e3779b9 + (seed << 6) + (seed >> 2); }
// like boost::hash: the key type supplies hash_value(), found by argument-dependent lookup
template <class K> struct hash { std::size_t operator()(const K& k) const { return hash_value(k); } };
template <> struct hash<std::string> { std::size_t operator()(const std::string& k) const { return std::hash<std::string>()(k); } };
}
