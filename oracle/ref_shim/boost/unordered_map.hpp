// Stand-in for a header of an absent third-party library (Boost / OpenBabel), written for oracle/_ref only:
// it lets the reference's own Vina headers compile where they lie under /root/reference. No arithmetic lives here.
#pragma once
#include <unordered_map>
#include "boost/functional/hash.hpp"
namespace boost { template <class K, class V, class H = boost::hash<K>> using unordered_map = std::unordered_map<K, V, H>; }
