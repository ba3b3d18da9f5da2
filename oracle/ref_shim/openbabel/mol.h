// Stand-in for a header of an absent third-party library (Boost / OpenBabel), written for oracle/_ref only:
// it lets the reference's own Vina headers compile where they lie under /root/reference. No arithmetic lives here.
#pragma once
#include "atom.h"
namespace OpenBabel {
// oracle/_ref builds its molecules by hand (as test/gnina/test_cache.cu does) and writes no structure files
class OBMol { public: unsigned NumAtoms() const { return 0; } OBMol& operator+=(const OBMol&) { return *this; }
  void SetChainsPerceived(bool = true) {} void ConnectTheDots() {} };
}
