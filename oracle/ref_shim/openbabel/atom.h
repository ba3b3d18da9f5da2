// Stand-in for a header of an absent third-party library (Boost / OpenBabel), written for oracle/_ref only:
// it lets the reference's own Vina headers compile where they lie under /root/reference. No arithmetic lives here.
#pragma once
#include <vector>
namespace OpenBabel {
// only so that the inline typing helper of atom_constants.h parses; oracle/_ref never types molecules through OpenBabel
class OBAtom { public: std::vector<OBAtom*> nbrs; unsigned GetAtomicNum() const { return 0; } bool IsAromatic() const { return false; }
  bool IsHbondAcceptor() const { return false; } };
namespace OBElements { inline const char* GetSymbol(unsigned) { return ""; } }
}
#define FOR_NBORS_OF_ATOM(n, a) for (OpenBabel::OBAtom * n : (a).nbrs)
