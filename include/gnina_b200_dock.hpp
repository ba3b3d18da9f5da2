// gnina_b200_dock.hpp -- C++ host side of BASELINE config 3 (header-only, C++17): the docking branch of do_search
// (main/main.cpp:312-400) for one ligand as a sequence of C-ABI calls, and a pool that keeps many ligands in flight from C++ host
// threads.  This is gnina_b200/docking.py (dock_ligand, DockingPool -- the version the GPU tests and the measurements of
// profiles/README.md run) restated in the reference's host language; it compiles with tests/cpp/host_test.cpp (`--dock` mode) but has
// NOT been run on a GPU yet (written after round 2's GPU minutes were spent).
//
//   search_box            setup_grid_dims (main/main.cpp:625-634): ONE box for affinity grids, random starts and out-of-box penalties
//   dock_ligand           cache build -> all chains in one launch (parallel_mc.cpp:183-214) -> merge_output_containers (:165-181) ->
//                         refine_structure of every kept pose (main.cpp:131-171,324-331) -> ONE CNN batch call (:333) -> the docking
//                         branch's affinity (non_cache::eval + num_tors_div, :340-344) -> sort by CNNscore (:348-361) ->
//                         remove_redundant (:182-192) -> poses outside the box skipped, num_modes (:371-378)
//   DockingPool           one ligand's chains are only `exhaustiveness` warps: ligands are kept in flight concurrently, one host
//                         thread + one Vina handle + one CNN clone (fresh_copy) each, as parallel_mc gives every task its own model
#pragma once
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <random>
#include <thread>
#include <vector>
#include "gnina_b200.hpp"

namespace gb {

struct SearchBox { float begin[3], end[3]; int32_t n[3]; };
inline SearchBox search_box(const float corner1[3], const float corner2[3], float granularity = 0.375f) {
  SearchBox b;
  for (int i = 0; i < 3; i++) {
    const float center = (corner1[i] + corner2[i]) / 2.f, span = corner2[i] - corner1[i];
    b.n[i] = (int32_t)std::ceil(span / granularity);
    const float real_span = granularity * b.n[i];
    b.begin[i] = center - real_span / 2;
    b.end[i] = b.begin[i] + real_span;
  }
  return b;
}

struct DockParams {
  int exhaustiveness = 8, num_saved_mins = 50, num_modes = 9;
  int num_steps = 0, maxiters = 0;             // 0 = the reference's formulas (main/main.cpp:442-443,454)
  float out_min_rmsd = 1.f, num_tors = -1.f;   // num_tors < 0: the number of torsion segments
  uint32_t seed = 1;
  bool refine = true;
};
struct DockedPose {
  std::vector<float> conf, coords;             // conf [7+T], all movable atoms [n_atoms][3]
  float e = 0, search_e = 0, cnnscore = 0, cnnaffinity = 0, cnnvariance = 0;
  bool within = true;
};

inline float rmsd_heavy(const std::vector<float>& a, const std::vector<float>& b, const int32_t* type, int n_atoms) {
  float acc = 0; int nh = 0;                   // rmsd_upper_bound over get_heavy_atom_movable_coords (lib/coords.cpp:24-31)
  for (int i = 0; i < n_atoms; i++) {
    if (type[i] < 2) continue;
    for (int k = 0; k < 3; k++) { const float d = a[3 * i + k] - b[3 * i + k]; acc += d * d; }
    nh++;
  }
  return nh ? std::sqrt(acc / nh) : 0.f;
}

// vina: receptor set; cnn: the same receptor set
inline std::vector<DockedPose> dock_ligand(VinaScorer& vina, CNNScorer& cnn, const gb_ligand_topology& lig, const float corner1[3],
                                           const float corner2[3], const DockParams& p = DockParams()) {
  const int na = lig.n_atoms, T = lig.n_segments - 1, nx = 7 + T;
  const SearchBox box = search_box(corner1, corner2);
  std::vector<int32_t> needed;
  for (int i = 0; i < na; i++)
    if (lig.smina_type[i] > 1 && std::find(needed.begin(), needed.end(), lig.smina_type[i]) == needed.end()) needed.push_back(lig.smina_type[i]);
  std::sort(needed.begin(), needed.end());
  vina.set_ligand(lig);
  vina.cache_build(box.begin, box.end, box.n, needed);
  gb_mc_params P{};
  P.num_steps = p.num_steps > 0 ? p.num_steps : (int32_t)(70 * 3 * (50 + (na + 10 * (6 + T))) / 2);   // main.cpp:442-443
  P.maxiters = p.maxiters > 0 ? p.maxiters : (25 + na) / 3;                                            // ssd_par.evals, :454
  P.num_saved_mins = p.num_saved_mins; P.temperature = 1.2f; P.mutation_amplitude = 2.f; P.min_rmsd = 1.f;   // :458
  P.hunt_cap[0] = P.hunt_cap[1] = P.hunt_cap[2] = 10.f;                                                 // :460
  std::mt19937 gen(p.seed);
  std::vector<uint32_t> seeds(p.exhaustiveness);
  for (uint32_t& s : seeds) s = 1 + gen() % 1000000;                       // random_int(0, 1000000, generator), parallel_mc.cpp:197-199
  VinaScorer::ChainOutputs ch = vina.parallel_mc(P, box.begin, box.end, seeds);
  const int S = ch.S, nc = ch.n_chains;
  const float v3[3] = {1000.f, 1000.f, 1000.f};
  // heavy-atom coordinates of every minimum for the merge
  std::vector<float> all;
  vina.eval_deriv(ch.conf.data(), nc * S, v3, 1e3f, nullptr, &all);
  int nh = 0;
  for (int i = 0; i < na; i++) nh += lig.smina_type[i] > 1;
  std::vector<float> heavy((size_t)nc * S * nh * 3);
  for (size_t q = 0; q < (size_t)nc * S; q++) {
    size_t k = 0;
    for (int i = 0; i < na; i++)
      if (lig.smina_type[i] > 1) { for (int j = 0; j < 3; j++) heavy[(q * nh + k) * 3 + j] = all[(q * na + i) * 3 + j]; k++; }
  }
  const std::vector<int32_t> kept = VinaScorer::merge_outputs(ch.e.data(), heavy.data(), ch.n_out.data(), nc, S, nh, p.num_saved_mins);
  const int m = (int)kept.size();
  if (!m) return {};
  std::vector<float> confs((size_t)m * nx), search_e(m);
  for (int i = 0; i < m; i++) {
    std::copy(&ch.conf[(size_t)kept[i] * nx], &ch.conf[(size_t)(kept[i] + 1) * nx], &confs[(size_t)i * nx]);
    search_e[i] = ch.e[kept[i]];
  }
  std::vector<int32_t> ok(m, 1);
  std::vector<float> e_ref = search_e;
  if (p.refine) e_ref = vina.refine(confs.data(), m, P.maxiters, v3, box.begin, box.end, &ok);
  std::vector<float> coords;
  vina.eval_deriv(confs.data(), m, v3, 1e3f, nullptr, &coords);
  // ONE CNN batch call and ONE final-scoring call over all kept poses
  std::vector<int32_t> types((size_t)m * na), offs(m + 1);
  for (int i = 0; i < m; i++) { std::copy(lig.smina_type, lig.smina_type + na, &types[(size_t)i * na]); offs[i] = i * na; }
  offs[m] = m * na;
  const Scores sc = cnn.score_batch(coords.data(), types.data(), offs.data(), m);
  const std::vector<float> nt(m, p.num_tors >= 0 ? p.num_tors : (float)T);
  const std::vector<float> aff = vina.docking_affinity(coords.data(), types.data(), offs.data(), m, nt.data(), box.begin, box.end);
  std::vector<DockedPose> poses(m);
  for (int i = 0; i < m; i++) {
    DockedPose& d = poses[i];
    d.conf.assign(&confs[(size_t)i * nx], &confs[(size_t)(i + 1) * nx]);
    d.coords.assign(&coords[(size_t)i * na * 3], &coords[(size_t)(i + 1) * na * 3]);
    d.search_e = search_e[i]; d.within = ok[i] != 0;
    d.e = ok[i] ? aff[i] : std::numeric_limits<float>::max();             // main.cpp:163-164, :335
    d.cnnscore = sc.score[i]; d.cnnaffinity = sc.affinity[i]; d.cnnvariance = sc.variance[i];
  }
  std::stable_sort(poses.begin(), poses.end(), [](const DockedPose& a, const DockedPose& b) { return a.cnnscore > b.cnnscore; });
  std::vector<DockedPose> out;                                             // remove_redundant (main.cpp:182-192), then :371-378
  for (DockedPose& d : poses) {
    bool far = true;
    for (const DockedPose& q : out) if (!(rmsd_heavy(d.coords, q.coords, lig.smina_type, na) > p.out_min_rmsd)) { far = false; break; }
    if (far) out.push_back(std::move(d));
  }
  std::vector<DockedPose> ranked;
  for (DockedPose& d : out) {
    if (!(d.e < 0.1f * std::numeric_limits<float>::max())) continue;
    if ((int)ranked.size() >= p.num_modes) break;
    ranked.push_back(std::move(d));
  }
  return ranked;
}

// Many ligands against one receptor: n_workers host threads, each with its own Vina handle and CNN clone; results in input order
class DockingPool {
  std::vector<std::unique_ptr<VinaScorer>> vina_;
  std::vector<std::unique_ptr<CNNScorer>> cnn_;

 public:
  DockingPool(CNNScorer& cnn, const float* rec_xyz, const int32_t* rec_type, int n_rec, int n_workers, int device = 0) {
    for (int w = 0; w < n_workers; w++) {
      vina_.emplace_back(new VinaScorer(device));
      vina_.back()->set_receptor(rec_xyz, rec_type, n_rec);
      cnn_.push_back(cnn.fresh_copy());                                    // shares the device weights and the current receptor
    }
  }
  std::vector<std::vector<DockedPose>> dock(const std::vector<const gb_ligand_topology*>& ligs, const float corner1[3], const float corner2[3],
                                            DockParams p = DockParams()) {
    std::vector<std::vector<DockedPose>> out(ligs.size());
    std::atomic<size_t> next{0};
    std::vector<std::exception_ptr> err(vina_.size());
    std::vector<std::thread> th;
    for (size_t w = 0; w < vina_.size(); w++)
      th.emplace_back([&, w] {
        try {
          gb_initialize_cuda(0);                                           // every thread that uses a scorer (dl_scorer.h:20-21)
          for (size_t i; (i = next++) < ligs.size();) {
            DockParams q = p; q.seed = p.seed + (uint32_t)i;
            out[i] = dock_ligand(*vina_[w], *cnn_[w], *ligs[i], corner1, corner2, q);
          }
        } catch (...) { err[w] = std::current_exception(); }
      });
    for (std::thread& t : th) t.join();
    for (const std::exception_ptr& e : err) if (e) std::rethrow_exception(e);
    return out;
  }
};

}  // namespace gb
