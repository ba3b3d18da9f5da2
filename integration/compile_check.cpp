// Compiles the adapters of integration/ against the REFERENCE's own headers (tests/test_integration_adapters.py):
//   g++ -std=c++17 -fsyntax-only -I oracle/ref_shim -I <gnina>/gninasrc/lib -I <gnina> -I <cuda>/include -I include integration/compile_check.cpp
#include "cnn_b200_scorer.h"
#include "docking_b200.h"

// every member is instantiated by an out-of-line use
std::shared_ptr<DLScorer> make_b200_scorer(const cnn_options& opts, int device, const std::string& blob_dir) {
  return std::make_shared<CNNB200Scorer>(opts, device, blob_dir);
}
