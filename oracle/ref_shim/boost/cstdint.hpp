// Stand-in for a header of an absent third-party library (Boost / OpenBabel), written for oracle/_ref only:
// it lets the reference's own Vina headers compile where they lie under /root/reference. No arithmetic lives here.
#pragma once
#include <cstdint>
namespace boost { using std::uintmax_t; }
