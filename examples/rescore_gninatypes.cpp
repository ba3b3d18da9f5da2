// rescore_gninatypes — the smallest complete host program on top of include/gnina_b200.hpp: CNN rescoring of pre-typed
// ligands against one pre-typed receptor, the `gnina --score_only --cnn <models>` loop without OpenBabel.
//
//   rescore_gninatypes <weights_dir> <models: comma list | default | fast | X_ensemble> <receptor.gninatypes> <ligand.gninatypes>...
//
// Each ligand file is one pose (gninatyper's record format, gninatyper.cpp:30-36).  Poses are queued in a
// gb::PoseBatcher, scored in batches on the device and printed in input order:
//   <file> CNNscore CNNaffinity CNNvariance
// Build: g++ -std=c++17 -O2 -I include examples/rescore_gninatypes.cpp -L gnina_b200 -lgnina_b200 -Wl,-rpath,$PWD/gnina_b200
#include <cstdio>
#include <dirent.h>
#include <sstream>
#include "gnina_b200.hpp"

static std::vector<std::string> builtin_models(const std::string& dir) {  // names of the blobs shipped in weights_dir
  std::vector<std::string> out;
  if (DIR* d = opendir(dir.c_str())) {
    while (dirent* e = readdir(d)) {
      const std::string f = e->d_name;
      if (f.size() > 4 && f.compare(f.size() - 4, 4, ".gbw") == 0) out.push_back(f.substr(0, f.size() - 4));
    }
    closedir(d);
  }
  return out;
}

int main(int argc, char** argv) {
  if (argc < 5) {
    std::fprintf(stderr, "usage: %s <weights_dir> <models> <receptor.gninatypes> <ligand.gninatypes>...\n", argv[0]);
    return 2;
  }
  try {
    std::vector<std::string> names;
    if (std::string(argv[2]) != "default") {
      std::stringstream ss(argv[2]);
      for (std::string n; std::getline(ss, n, ',');) names.push_back(n);
    }
    names = gb::expand_model_names(names, builtin_models(argv[1]));   // "", fast, default1.0, X_ensemble as gnina spells them
    gb::CNNScorer scorer(argv[1], names);
    const gb::TypedAtoms rec = gb::read_gninatypes(argv[3]);
    scorer.set_receptor(rec.xyz.data(), rec.type.data(), (int)rec.size());
    std::vector<std::string> files(argv + 4, argv + argc);
    gb::PoseBatcher queue(scorer, 1024, [&](size_t ticket, float score, float affinity, float, float variance) {
      std::printf("%s %.5f %.5f %.5f\n", files[ticket].c_str(), score, affinity, variance);
    });
    for (const auto& f : files) {
      const gb::TypedAtoms lig = gb::read_gninatypes(f);
      queue.add(lig.xyz.data(), lig.type.data(), (int)lig.size());
    }
    queue.flush();
  } catch (const gb::usage_error& e) {
    std::fprintf(stderr, "usage error: %s\n", e.what());
    return 2;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
