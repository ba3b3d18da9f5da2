// Stand-in for a header of an absent third-party library (Boost / OpenBabel), written for oracle/_ref only:
// it lets the reference's own Vina headers compile where they lie under /root/reference. No arithmetic lives here.
#pragma once
#define OB_VERSION_CHECK(a, b, c) (((a) << 16) | ((b) << 8) | (c))
#define OB_VERSION OB_VERSION_CHECK(3, 1, 1)
