// Stand-in for a header of an absent third-party library (Boost / OpenBabel), written for oracle/_ref only:
// it lets the reference's own Vina headers compile where they lie under /root/reference. No arithmetic lives here.
#pragma once
#include <string>
#include <ostream>
#include "mol.h"
namespace OpenBabel {
class OBConversion { public: enum Option_type { INOPTIONS, OUTOPTIONS, GENOPTIONS };
  bool SetOutFormat(const char*) { return false; } bool SetInFormat(const char*) { return false; }
  void AddOption(const char*, Option_type, const char* = nullptr) {}
  bool ReadString(OBMol*, std::string) { return false; } bool Write(OBMol*, std::ostream* = nullptr) { return false; } };
}
