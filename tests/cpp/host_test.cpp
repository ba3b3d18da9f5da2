// C++ host-side check: builds against include/gnina_b200.hpp + libgnina_b200.so.
//   host_test --names                      : model-name expansion only (no device needed)
//   host_test <weights_dir> <case.bin>     : score the poses in case.bin through gb::CNNScorer / gb::NonCacheCNN and
//                                            print the results as text (compared with the oracle by the pytest)
// case.bin (little endian): int32 n_rec, n_lig_atoms, n_poses; float rec_xyz[3 n_rec]; int32 rec_type[n_rec];
//                           float lig_xyz[3 n_lig]; int32 lig_type[n_lig]; int32 offsets[n_poses + 1]
#include <cstdio>
#include <cstring>
#include <fstream>
#include "gnina_b200.hpp"

template <typename T>
static std::vector<T> rd(std::ifstream& f, size_t n) {
  std::vector<T> v(n);
  f.read(reinterpret_cast<char*>(v.data()), sizeof(T) * n);
  return v;
}

int main(int argc, char** argv) {
  if (argc >= 2 && !strcmp(argv[1], "--names")) {
    auto d = gb::expand_model_names({});
    auto f = gb::expand_model_names({"fast"});
    auto e = gb::expand_model_names({"crossdock_default2018_ensemble"}, {"crossdock_default2018", "crossdock_default2018_1", "dense"});
    printf("%s %s %s | %s | %zu\n", d[0].c_str(), d[1].c_str(), d[2].c_str(), f[0].c_str(), e.size());
    return 0;
  }
  if (argc < 3) return 2;
  try {
    std::ifstream f(argv[2], std::ios::binary);
    int32_t hdr[3];
    f.read(reinterpret_cast<char*>(hdr), sizeof hdr);
    auto rec_xyz = rd<float>(f, 3 * (size_t)hdr[0]);
    auto rec_t = rd<int32_t>(f, hdr[0]);
    auto lig_xyz = rd<float>(f, 3 * (size_t)hdr[1]);
    auto lig_t = rd<int32_t>(f, hdr[1]);
    auto offs = rd<int32_t>(f, hdr[2] + 1);
    try {
      gb::CNNScorer bad(argv[1], {"no_such_model"});
      return 3;
    } catch (const gb::usage_error& e) {
      printf("usage_error: %s\n", e.what());
    }
    gb::CNNScorer s(argv[1], {"crossdock_default2018"});
    s.set_receptor(rec_xyz.data(), rec_t.data(), hdr[0]);
    auto r = s.score_batch(lig_xyz.data(), lig_t.data(), offs.data(), hdr[2]);
    for (int p = 0; p < hdr[2]; p++) printf("pose %d %.7f %.6f %.6f\n", p, r.score[p], r.affinity[p], r.loss[p]);
    auto c = s.fresh_copy();
    float aff, loss, var;
    std::vector<float> grad;
    const float sc = c->score(lig_xyz.data(), lig_t.data(), offs[1], true, aff, loss, var, &grad);
    printf("single %.7f %.6f %.6f gradsum %.6f\n", sc, aff, loss, [&] { double t = 0; for (float g : grad) t += std::fabs(g); return t; }());
    // non_cache_cnn::eval_deriv with a search box that cuts through the ligand: penalties + derivative signs
    gb::GridDims gd;
    for (int i = 0; i < 3; i++) { gd[i].begin = -1.0f; gd[i].end = 1.0f; gd[i].n = 8; }
    const float ctr[3] = {0, 0, 0};
    gb::NonCacheCNN nc(s, gd, ctr, 10.0f);
    std::vector<float> mf;
    const float e = nc.eval(lig_xyz.data(), lig_t.data(), offs[1], &mf);
    const float e0 = nc.eval(lig_xyz.data(), lig_t.data(), offs[1]);
    printf("noncache %.5f %.5f forces %zu\n", e, e0, mf.size());
  } catch (const std::exception& e) {
    printf("ERROR %s\n", e.what());
    return 1;
  }
  return 0;
}
