// C++ host-side check: builds against include/gnina_b200.hpp + libgnina_b200.so.
//   host_test --names                      : model-name expansion only (no device needed)
//   host_test --host <tmp.gninatypes>      : PoseBatcher with a stub runner + gninatypes round trip (no device needed)
//   host_test --minimize-host              : gb::minimize_poses (lock-step quasi-Newton, include/gnina_b200_minimize.hpp) on a small built-in
//                                            ligand with an analytic energy (no device needed)
//   host_test --minimize <weights_dir>     : the same driver over gb::CnnBatchEnergy = gb_cnn_score_grad (device; not part of the suite yet)
//   host_test --dock <weights_dir>         : gb::DockingPool / gb::dock_ligand (include/gnina_b200_dock.hpp) on the built-in ligand (device;
//                                            not part of the suite yet)
//   host_test <weights_dir> <case.bin>     : score the poses in case.bin through gb::CNNScorer / gb::NonCacheCNN and
//                                            print the results as text (compared with the oracle by the pytest)
// case.bin (little endian): int32 n_rec, n_lig_atoms, n_poses; float rec_xyz[3 n_rec]; int32 rec_type[n_rec];
//                           float lig_xyz[3 n_lig]; int32 lig_type[n_lig]; int32 offsets[n_poses + 1]
#include <algorithm>
#include <array>
#include <cstdio>
#include <cstring>
#include <fstream>
#include "gnina_b200.hpp"
#include "gnina_b200_minimize.hpp"
#include "gnina_b200_dock.hpp"

// a 9-atom ligand with two torsions: root (atoms 0-3), a segment hanging off atom 3 (atoms 4-6) and one off atom 6 (atoms 7-8)
struct TinyLigand {
  std::vector<float> local{0, 0, 0, 1.5f, 0, 0, 2.2f, 1.3f, 0, 3.7f, 1.4f, 0.2f, 0, 0, 0, 0.8f, 1.2f, 0.4f, 2.2f, 1.0f, 0.9f, 0, 0, 0, 1.1f, -0.9f, 0.5f};
  std::vector<int32_t> type{2, 2, 6, 2, 2, 10, 2, 2, 1}, parent{-1, 0, 1}, begin{0, 4, 7}, end{4, 7, 9};
  std::vector<float> rel_origin{0, 0, 0, 4.6f, 2.4f, 0.5f, 3.0f, 1.9f, 1.4f}, rel_axis{0, 0, 0, 0.6f, 0.8f, 0, 0, 0.6f, 0.8f};
  gb_ligand_topology topo{};
  TinyLigand() {
    topo.n_atoms = 9; topo.n_segments = 3; topo.n_pairs = 0;
    topo.local_xyz = local.data(); topo.smina_type = type.data(); topo.seg_parent = parent.data(); topo.seg_atom_begin = begin.data();
    topo.seg_atom_end = end.data(); topo.seg_rel_origin = rel_origin.data(); topo.seg_rel_axis = rel_axis.data();
    topo.pair_a = nullptr; topo.pair_b = nullptr; topo.gyration_radius = 2.f;
  }
};
static std::vector<float> tiny_starts(const gb::LigandTree& tree, int n) {
  std::vector<float> x((size_t)n * tree.conf_floats(), 0.f);
  for (int i = 0; i < n; i++) {
    float* c = &x[(size_t)i * tree.conf_floats()];
    c[0] = 3.f * std::sin(1.3f * i); c[1] = 2.f * std::cos(0.7f * i); c[2] = 0.5f * i - 2.f;
    const float a = 0.4f * i; c[3] = std::cos(a / 2); c[4] = std::sin(a / 2) * 0.6f; c[5] = std::sin(a / 2) * 0.8f; c[6] = 0;
    c[7] = 0.3f * i - 1.f; c[8] = 1.f - 0.25f * i;
  }
  return x;
}

template <typename T>
static std::vector<T> rd(std::ifstream& f, size_t n) {
  std::vector<T> v(n);
  f.read(reinterpret_cast<char*>(v.data()), sizeof(T) * n);
  return v;
}

int main(int argc, char** argv) {
  if (argc >= 2 && !strcmp(argv[1], "--minimize-host")) {
    TinyLigand L;
    gb::LigandTree tree(L.topo);
    const int n = 8, na = tree.n_atoms;
    std::vector<float> x = tiny_starts(tree, n), x1 = x;
    int calls = 0;
    auto energy = [&](const float* coords, const int*, int k, float* e, float* f) {      // springs to the origin on the heavy atoms
      calls++;
      for (int j = 0; j < k; j++) {
        float en = 0;
        for (int a = 0; a < na; a++)
          for (int q = 0; q < 3; q++) {
            const float d = coords[((size_t)j * na + a) * 3 + q];
            const bool heavy = tree.heavy(a);
            if (heavy) en += 0.05f * d * d;
            f[((size_t)j * na + a) * 3 + q] = heavy ? 0.1f * d : 0.f;
          }
        e[j] = en;
      }
    };
    std::vector<float> c0((size_t)3 * na), so(9), sa(9), e0(n), f0((size_t)3 * na);
    for (int i = 0; i < n; i++) { tree.set_conf(&x[(size_t)i * tree.conf_floats()], c0.data(), so.data(), sa.data()); int z = i; energy(c0.data(), &z, 1, &e0[i], f0.data()); }
    calls = 0;
    gb::MinimizeParams mp;
    std::vector<int> ev; int rounds = 0;
    std::vector<float> e = gb::minimize_poses(tree, energy, x.data(), n, mp, &ev, &rounds);
    int total = 0; bool down = true, alone = true;
    float sum0 = 0, sum1 = 0;
    for (int i = 0; i < n; i++) { total += ev[i]; down = down && e[i] <= e0[i]; sum0 += e0[i]; sum1 += e[i]; }
    down = down && sum1 < 0.6f * sum0;                                                      // the rigid shape keeps a floor
    for (int i : {0, 5}) {                                                                  // a pose minimised alone ends at the same point
      std::vector<float> xi(x1.begin() + (size_t)i * tree.conf_floats(), x1.begin() + (size_t)(i + 1) * tree.conf_floats());
      std::vector<float> ei = gb::minimize_poses(tree, energy, xi.data(), 1, mp);
      alone = alone && ei[0] == e[i] && !memcmp(xi.data(), &x[(size_t)i * tree.conf_floats()], 4 * (size_t)tree.conf_floats());
    }
    printf("minimize poses %d rounds %d evaluations %d descended %d same_alone %d batched %d\n", n, rounds, total, (int)down, (int)alone,
           (int)(rounds + 1 < total));
    return 0;
  }
  if (argc >= 3 && !strcmp(argv[1], "--dock")) {                                           // device: config 3 glue from C++ host threads
    gb::CNNScorer s(argv[2], {"crossdock_default2018"});
    std::vector<float> rec; std::vector<int32_t> rt;
    for (int i = 0; i < 200; i++) {                                                         // a shell of receptor atoms around the box
      const float a = 0.7f * i, b = 0.37f * i;
      rec.push_back(11.f * std::cos(a) * std::sin(b)); rec.push_back(11.f * std::sin(a) * std::sin(b)); rec.push_back(11.f * std::cos(b));
      rt.push_back(i % 3 == 0 ? 10 : 2);
    }
    s.set_receptor(rec.data(), rt.data(), (int)rt.size());
    TinyLigand L;
    gb::DockingPool pool(s, rec.data(), rt.data(), (int)rt.size(), 4);
    std::vector<const gb_ligand_topology*> ligs(8, &L.topo);
    const float c1[3] = {-6, -6, -6}, c2[3] = {6, 6, 6};
    gb::DockParams dp; dp.exhaustiveness = 8; dp.num_steps = 40; dp.num_saved_mins = 10;
    auto res = pool.dock(ligs, c1, c2, dp);
    for (size_t i = 0; i < res.size(); i++)
      printf("dock ligand %zu modes %zu best_cnnscore %g affinity %g\n", i, res[i].size(), res[i].empty() ? -1.f : res[i][0].cnnscore,
             res[i].empty() ? 0.f : res[i][0].e);
    return 0;
  }
  if (argc >= 3 && !strcmp(argv[1], "--minimize")) {                                       // device: CNN minimisation of 64 poses in lock step
    gb::CNNScorer s(argv[2], {"crossdock_default2018"});
    TinyLigand L;
    gb::LigandTree tree(L.topo);
    const float rec[6] = {8, 0, 0, -8, 0, 0}; const int32_t rt[2] = {2, 10};
    s.set_receptor(rec, rt, 2);
    const int n = 64;
    std::vector<float> x = tiny_starts(tree, n);
    const float b[3] = {-12, -12, -12}, en[3] = {12, 12, 12};
    gb::CnnBatchEnergy energy(s.handle(), tree, b, en, 10.f, s.info(0).dimension);
    energy.set_centers(x.data(), n);
    gb::MinimizeParams mp; mp.maxiters = 50;
    std::vector<int> ev; int rounds = 0;
    std::vector<float> e = gb::minimize_poses(tree, energy, x.data(), n, mp, &ev, &rounds);
    int total = 0; for (int v : ev) total += v;
    printf("cnn_minimize poses %d rounds %d evaluations %d loss0 %g\n", n, rounds, total, e[0]);
    return 0;
  }
  if (argc >= 2 && !strcmp(argv[1], "--names")) {
    auto d = gb::expand_model_names({});
    auto f = gb::expand_model_names({"fast"});
    auto e = gb::expand_model_names({"crossdock_default2018_ensemble"}, {"crossdock_default2018", "crossdock_default2018_1", "dense"});
    printf("%s %s %s | %s | %zu\n", d[0].c_str(), d[1].c_str(), d[2].c_str(), f[0].c_str(), e.size());
    return 0;
  }
  if (argc >= 3 && !strcmp(argv[1], "--host")) {
    // typed-atom records: write, read back, and reject a truncated file
    const float xyz[6] = {1.5f, -2.25f, 3.0f, 0.f, 1e-3f, -7.f};
    const int32_t ty[2] = {2, 13};
    gb::write_gninatypes(argv[2], xyz, ty, 2);
    auto a = gb::read_gninatypes(argv[2]);
    bool ok = a.size() == 2 && a.xyz[1] == -2.25f && a.xyz[5] == -7.f && a.type[1] == 13;
    { FILE* f = fopen(argv[2], "ab"); fputc(0, f); fclose(f); }
    try { gb::read_gninatypes(argv[2]); ok = false; } catch (const gb::usage_error&) {}
    try { gb::read_gninatypes(std::string(argv[2]) + ".missing"); ok = false; } catch (const gb::usage_error&) {}
    printf("gninatypes %s\n", ok ? "ok" : "FAILED");
    // pose queue: stub runner scores a pose as (n_atoms, first x, centre x or -1, batch size)
    std::vector<std::array<float, 5>> got;
    auto runner = [](const float* x, const int32_t*, const int32_t* o, int n, const float* c) {
      gb::Scores r;
      for (int i = 0; i < n; i++) {
        r.score.push_back((float)(o[i + 1] - o[i])); r.affinity.push_back(x[3 * o[i]]);
        r.loss.push_back(c ? c[3 * i] : -1.f); r.variance.push_back((float)n);
      }
      return r;
    };
    auto deliver = [&](size_t t, float s, float af, float l, float v) { got.push_back({(float)t, s, af, l, v}); };
    {
      gb::PoseBatcher q(runner, 3, deliver);
      for (int p = 0; p < 7; p++) {
        std::vector<float> px(3 * (size_t)(p + 1), (float)(10 + p));
        std::vector<int32_t> pt(p + 1, 2);
        const size_t t = q.add(px.data(), pt.data(), p + 1);
        if (t != (size_t)p) ok = false;
      }
      printf("queued %zu batches %zu delivered %zu\n", q.queued(), q.batches_run(), got.size());
    }  // destructor flushes the ragged tail
    for (auto& g : got) printf("ticket %.0f n %.0f x %.0f c %.0f batch %.0f\n", g[0], g[1], g[2], g[3], g[4]);
    const float ctr[3] = {4.f, 5.f, 6.f};
    got.clear();
    { gb::PoseBatcher q(runner, 8, deliver, ctr); const float x1[3] = {1, 2, 3}; const int32_t t1[1] = {2}; q.add(x1, t1, 1); q.flush(); q.flush(); }
    printf("fixed centre %.0f deliveries %zu\n", got.empty() ? -1.f : got[0][3], got.size());
    {  // merge_output_containers through the C++ wrapper: two chains found the same pose (rmsd 0.5), one found another
      const float me[4] = {-5.f, -1.f, -6.f, 0.f};                        // [chain 2][S 2]
      const float mc[4 * 3] = {0, 0, 0, 9, 9, 9, 0.5f, 0, 0, 0, 0, 0};    // one atom per pose
      const int32_t mn[2] = {2, 1};
      auto kept = gb::VinaScorer::merge_outputs(me, mc, mn, 2, 2, 1, 50);
      printf("merge");
      for (auto k : kept) printf(" %d", k);
      printf("\n");
    }
    return ok ? 0 : 1;
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
    {  // the pose queue in front of the real scorer reproduces the direct batch call, in order
      double worst = 0; size_t n_del = 0;
      gb::PoseBatcher q(s, 2, [&](size_t t, float sc2, float af2, float, float) {
        worst = std::max(worst, (double)std::fabs(sc2 - r.score[t]) + std::fabs(af2 - r.affinity[t])); n_del++;
      });
      for (int p = 0; p < hdr[2]; p++) q.add(lig_xyz.data() + 3 * (size_t)offs[p], lig_t.data() + offs[p], offs[p + 1] - offs[p]);
      q.flush();
      printf("batcher delivered %zu batches %zu maxdiff %.3g\n", n_del, q.batches_run(), worst);
    }
    // non_cache_cnn::eval_deriv with a search box that cuts through the ligand: penalties + derivative signs
    gb::GridDims gd;
    for (int i = 0; i < 3; i++) { gd[i].begin = -1.0f; gd[i].end = 1.0f; gd[i].n = 8; }
    const float ctr[3] = {0, 0, 0};
    gb::NonCacheCNN nc(s, gd, ctr, 10.0f);
    std::vector<float> mf;
    const float e = nc.eval(lig_xyz.data(), lig_t.data(), offs[1], &mf);
    const float e0 = nc.eval(lig_xyz.data(), lig_t.data(), offs[1]);
    printf("noncache %.5f %.5f forces %zu\n", e, e0, mf.size());
    {  // gb::MultiDeviceScorer: poses sharded over devices (here every visible device, twice -- also exercises two
       // handles per GPU); same results as the single scorer, in pose order
      std::vector<int> devs;
      for (int d = 0; d < gb_device_count(); d++) { devs.push_back(d); devs.push_back(d); }
      gb::MultiDeviceScorer md(argv[1], {"crossdock_default2018"}, devs);
      md.set_receptor(rec_xyz.data(), rec_t.data(), hdr[0]);
      auto r2 = md.score_batch(lig_xyz.data(), lig_t.data(), offs.data(), hdr[2]);
      double worst = 0;
      for (int p2 = 0; p2 < hdr[2]; p2++) worst = std::max(worst, (double)std::fabs(r2.score[p2] - r.score[p2]) + std::fabs(r2.affinity[p2] - r.affinity[p2]));
      const auto sb = gb::MultiDeviceScorer::shard_bounds(10, 4);
      printf("multidevice %zu maxdiff %.3g shards %d %d %d %d %d\n", md.n_devices(), worst, sb[0], sb[1], sb[2], sb[3], sb[4]);
    }
    {  // --cnn_mix_emp_force / --cnn_mix_emp_energy (lib/non_cache_cnn.cpp:113-166): with a box that contains the ligand
       // the blended energy is (CNN loss + w * empirical) / (1 + w), the identity test/gnina/test_min.py:45-61 checks
      gb::GridDims wide;
      for (int i = 0; i < 3; i++) { wide[i].begin = -40.0f; wide[i].end = 40.0f; wide[i].n = 8; }
      gb::VinaScorer vs;
      vs.set_receptor(rec_xyz.data(), rec_t.data(), hdr[0]);
      gb::NonCacheCNN plain(s, wide, ctr, 10.0f), mixed(s, wide, ctr, 10.0f);
      const float w = 0.5f;
      mixed.set_empirical(&vs, w, true, true);
      std::vector<float> f0, f1, ee, ed;
      const float e_plain = plain.eval(lig_xyz.data(), lig_t.data(), offs[1], &f0);
      const float e_mixed = mixed.eval(lig_xyz.data(), lig_t.data(), offs[1], &f1);
      const float b3[3] = {-40.f, -40.f, -40.f}, e3[3] = {40.f, 40.f, 40.f};
      vs.noncache_atoms(lig_xyz.data(), lig_t.data(), offs[1], b3, e3, 1000.f, ee, ed);
      double emp = 0, worst = 0;
      for (float x : ee) emp += x;
      for (size_t i = 0; i < f0.size(); i++) worst = std::max(worst, (double)std::fabs(f1[i] - (f0[i] + w * ed[i]) / (1 + w)));
      printf("mixing %.6f %.6f %.6f %.3g\n", e_plain, e_mixed, emp, worst);
    }
  } catch (const std::exception& e) {
    printf("ERROR %s\n", e.what());
    return 1;
  }
  return 0;
}
