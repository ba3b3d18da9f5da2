// gninasrc/lib/docking_b200.h -- the adapters a gnina maintainer adds on the docking side: they implement / replace gnina's own types
// (parallel_mc's call operator, igrid, the refine_structure loop, the final scoring of do_search) and forward to the C ABI of
// include/gnina_b200.h.  Compiled against the reference's real headers by tests/test_integration_adapters.py; B200Ligand (pure host
// code) is also RUN there against reference models (oracle/_ref).
#ifndef DOCKING_B200_H_
#define DOCKING_B200_H_

#include <cstdint>
#include <stdexcept>
#include <vector>

#include "conf.h"
#include "igrid.h"
#include "model.h"
#include "monte_carlo.h"
#include "non_cache.h"
#include "random.h"
#include "gnina_b200.h"

namespace b200 {

inline void check(int rc) {
  if (rc == GB_ERR_USAGE) throw usage_error(gb_last_error());
  if (rc != GB_OK) throw std::runtime_error(gb_last_error());
}

// read access to segment::relative_axis / relative_origin (private; friend struct segment_node of tree_gpu.h reads them the same way)
template <class Tag, typename Tag::type M> struct Peek { friend typename Tag::type peek(Tag) { return M; } };
struct SegRelAxis { typedef vec segment::*type; friend type peek(SegRelAxis); };
struct SegRelOrigin { typedef vec segment::*type; friend type peek(SegRelOrigin); };
template struct Peek<SegRelAxis, &segment::relative_axis>;
template struct Peek<SegRelOrigin, &segment::relative_origin>;

// model -> gb_ligand_topology, once per ligand.  heterotree<rigid_body> (lib/tree.h:341-400): node 0 = the rigid root, the branches in
// DFS pre-order = the order of conf.torsions (tree.h:361-366).  Atom coordinates are the frame-local ones atom_frame::set_coords reads
// (m.atoms[i].coords, tree.h:124-127).
struct B200Ligand {
  std::vector<float> local_xyz, rel_origin, rel_axis;
  std::vector<int32_t> type, parent, begin, end, pair_a, pair_b;
  int32_t n_heavy = 0;
  gb_ligand_topology topo{};

  explicit B200Ligand(const model& m) {
    VINA_CHECK(m.ligands.size() == 1);
    const ligand& lig = m.ligands[0];
    VINA_CHECK(lig.begin == 0);                              // rigid receptor: the ligand's atoms are the movable atoms 0 .. end
    add_node(m, lig.node, vec(0, 0, 0), vec(0, 0, 0), -1);
    for (const branch& b : lig.children) walk(m, b, 0);
    for (const interacting_pair& p : lig.pairs) { pair_a.push_back((int32_t)p.a); pair_b.push_back((int32_t)p.b); }   // interacting_pairs.h:7-19
    for (int32_t t : type) n_heavy += t >= 2;
    topo.n_atoms = (int32_t)type.size(); topo.n_segments = (int32_t)parent.size(); topo.n_pairs = (int32_t)pair_a.size();
    topo.local_xyz = local_xyz.data(); topo.smina_type = type.data(); topo.seg_parent = parent.data();
    topo.seg_atom_begin = begin.data(); topo.seg_atom_end = end.data(); topo.seg_rel_origin = rel_origin.data();
    topo.seg_rel_axis = rel_axis.data(); topo.pair_a = pair_a.data(); topo.pair_b = pair_b.data();
    topo.gyration_radius = (float)m.gyration_radius(0);      // of the pose the model holds now: what mutate_conf's first rotation sees
  }
  B200Ligand(const B200Ligand&) = delete;

 private:
  template <class Node> void add_node(const model& m, const Node& node, const vec& rel_o, const vec& rel_a, int par) {
    parent.push_back(par);
    begin.push_back((int32_t)node.begin);
    end.push_back((int32_t)node.end);
    for (sz i = node.begin; i < node.end; i++) {
      VINA_CHECK(i == type.size());                          // segments own consecutive atom ranges in DFS order
      for (int k = 0; k < 3; k++) local_xyz.push_back((float)m.atoms[i].coords[k]);
      type.push_back((int32_t)m.atoms[i].sm);
    }
    for (int k = 0; k < 3; k++) { rel_origin.push_back((float)rel_o[k]); rel_axis.push_back((float)rel_a[k]); }
  }
  void walk(const model& m, const branch& t, int par) {
    const int me = (int)parent.size();
    add_node(m, t.node, t.node.*peek(SegRelOrigin()), t.node.*peek(SegRelAxis()), par);   // tree.h:208-216
    for (const branch& c : t.children) walk(m, c, me);
  }
};

inline int conf_floats(const conf& c) { return 7 + (int)c.ligands[0].torsions.size(); }
inline void pack_conf(const conf& c, float* x) {
  const ligand_conf& l = c.ligands[0];
  for (int k = 0; k < 3; k++) x[k] = (float)l.rigid.position[k];
  x[3] = (float)l.rigid.orientation.R_component_1(); x[4] = (float)l.rigid.orientation.R_component_2();
  x[5] = (float)l.rigid.orientation.R_component_3(); x[6] = (float)l.rigid.orientation.R_component_4();
  for (sz i = 0; i < l.torsions.size(); i++) x[7 + i] = (float)l.torsions[i];
}
inline void unpack_conf(const float* x, conf& c) {
  ligand_conf& l = c.ligands[0];
  l.rigid.position = vec(x[0], x[1], x[2]);
  l.rigid.orientation = qt(x[3], x[4], x[5], x[6]);
  for (sz i = 0; i < l.torsions.size(); i++) l.torsions[i] = x[7 + i];
}
inline void box_of(const grid_dims& gd, float* b, float* e) {
  for (int i = 0; i < 3; i++) { b[i] = (float)gd[i].begin; e[i] = (float)gd[i].end; }
}

// parallel_mc::operator() (lib/parallel_mc.cpp:183-214): every chain of the ligand in ONE launch, then merge_output_containers
// (:165-181).  `h` has the receptor (gb_vina_set_receptor with m.grid_atoms) and the affinity grids of this ligand's atom types
// (cache_b200 below) set.
struct parallel_mc_b200 {
  monte_carlo mc;
  sz num_tasks = 8;
  gb_vina* h = nullptr;

  void operator()(const model& m0, output_container& out, const vec& corner1, const vec& corner2, rng& generator) const {
    model m(m0);
    B200Ligand L(m);
    check(gb_vina_set_ligand(h, &L.topo));
    std::vector<uint32_t> seeds(num_tasks);
    for (uint32_t& s : seeds) s = (uint32_t)random_int(0, 1000000, generator);                     // parallel_mc.cpp:197-199
    gb_mc_params P{};
    P.num_steps = (int32_t)mc.num_steps;
    P.maxiters = (int32_t)(mc.ssd_par.minparm.maxiters ? mc.ssd_par.minparm.maxiters : mc.ssd_par.evals);   // monte_carlo.cpp:108-110
    P.num_saved_mins = (int32_t)mc.num_saved_mins;
    P.temperature = (float)mc.temperature; P.mutation_amplitude = (float)mc.mutation_amplitude; P.min_rmsd = (float)mc.min_rmsd;
    for (int k = 0; k < 3; k++) P.hunt_cap[k] = (float)mc.hunt_cap[k];
    const int S = P.num_saved_mins, nx = 7 + L.topo.n_segments - 1, na = L.topo.n_atoms, nc = (int)num_tasks;
    std::vector<float> e((size_t)nc * S), confs((size_t)nc * S * nx);
    std::vector<int32_t> n_out(nc);
    const float c1[3] = {(float)corner1[0], (float)corner1[1], (float)corner1[2]}, c2[3] = {(float)corner2[0], (float)corner2[1], (float)corner2[2]};
    check(gb_vina_mc(h, &P, c1, c2, seeds.data(), nc, 1e3f, e.data(), confs.data(), n_out.data()));
    // get_heavy_atom_movable_coords of every minimum: the coordinates model::set gives for the conformations
    std::vector<float> ee((size_t)nc * S), all((size_t)nc * S * na * 3), heavy((size_t)nc * S * L.n_heavy * 3);
    const float v3[3] = {1000.f, 1000.f, 1000.f};
    check(gb_vina_eval_deriv(h, confs.data(), nc * S, v3, 1e3f, ee.data(), nullptr, all.data()));
    for (size_t q = 0; q < (size_t)nc * S; q++) {
      size_t k = 0;
      for (int i = 0; i < na; i++)
        if (L.type[i] >= 2) { for (int j = 0; j < 3; j++) heavy[(q * L.n_heavy + k) * 3 + j] = all[(q * na + i) * 3 + j]; k++; }
    }
    std::vector<int32_t> kept(mc.num_saved_mins);
    int32_t nk = 0;
    check(gb_vina_merge_outputs(e.data(), heavy.data(), n_out.data(), nc, S, L.n_heavy, 2.0f, (int)mc.num_saved_mins, kept.data(), &nk));
    conf c = m.get_initial_conf(false);
    for (int i = 0; i < nk; i++) {                                                                  // conf.h:520-537
      unpack_conf(&confs[(size_t)kept[i] * nx], c);
      output_type* o = new output_type(c, e[kept[i]]);
      for (int a = 0; a < L.n_heavy; a++) { const float* p = &heavy[((size_t)kept[i] * L.n_heavy + a) * 3]; o->coords.push_back(vec(p[0], p[1], p[2])); }
      out.push_back(o);
    }
  }
};

// igrid over the device cache (lib/igrid.h:32-46; cache::eval / eval_deriv, lib/cache.cpp:50-83), for callers that want single evaluations
struct cache_b200 : public igrid {
  gb_vina* h;
  fl slope;
  cache_b200(gb_vina* h_, const grid_dims& gd, fl slope_, const std::vector<smt>& needed) : h(h_), slope(slope_) {
    float b[3], e[3];
    box_of(gd, b, e);
    const int32_t n[3] = {(int32_t)gd[0].n, (int32_t)gd[1].n, (int32_t)gd[2].n};
    std::vector<int32_t> t(needed.begin(), needed.end());
    check(gb_vina_cache_build(h, b, e, n, t.data(), (int)t.size()));                                // cache::populate, lib/cache.cpp:104-184
  }
  fl eval(model& m, fl v) const override { return run(m, v, false); }
  fl eval_deriv(model& m, fl v, const grid&) const override { return run(m, v, true); }

 private:
  fl run(model& m, fl v, bool deriv) const {
    const int n = (int)m.num_movable_atoms();
    std::vector<float> xyz(3 * (size_t)n), d(3 * (size_t)n);
    std::vector<int32_t> t(n);
    for (int i = 0; i < n; i++) { t[i] = (int32_t)m.atoms[i].sm; for (int k = 0; k < 3; k++) xyz[3 * i + k] = (float)m.coords[i][k]; }
    const int32_t off[2] = {0, n};
    float e = 0;
    check(gb_vina_cache_eval(h, xyz.data(), t.data(), off, 1, (float)slope, (float)v, &e, deriv ? d.data() : nullptr));
    if (deriv) for (int i = 0; i < n; i++) m.minus_forces[i] = vec(d[3 * i], d[3 * i + 1], d[3 * i + 2]);
    return e;
  }
};

// main/main.cpp:131-171 refine_structure for EVERY kept pose of the ligand in one call (the loop of :324-331); gd = the search box of
// the non_cache the reference passes
inline void refine_structure_b200(gb_vina* h, output_container& out_cont, const vec& cap, const minimization_params& mp, const grid_dims& gd) {
  if (out_cont.empty()) return;
  const int nx = conf_floats(out_cont[0].c), n = (int)out_cont.size();
  std::vector<float> confs((size_t)n * nx), e(n);
  std::vector<int32_t> ok(n);
  for (int i = 0; i < n; i++) pack_conf(out_cont[i].c, &confs[(size_t)i * nx]);
  const float v3[3] = {(float)cap[0], (float)cap[1], (float)cap[2]};
  float b[3], en[3];
  box_of(gd, b, en);
  gb_minimization_params p{(int32_t)mp.maxiters, mp.type == minimization_params::BFGSAccurateLineSearch ? 1 : 0, mp.early_term ? 1 : 0};
  check(gb_vina_refine_minimize(h, confs.data(), n, &p, v3, b, en, e.data(), ok.data(), nullptr));
  for (int i = 0; i < n; i++) { unpack_conf(&confs[(size_t)i * nx], out_cont[i].c); out_cont[i].e = ok[i] ? e[i] : max_fl; }   // :163-164
}

// the "Affinity" of the docking branch for all poses at once (main/main.cpp:340-344: eval_adjusted with ig = nc_new): coordinates of
// every pose (all movable atoms, model::set order) -> out_cont[i].e; num_tors = conf_independent_inputs(m).num_tors (lib/terms.cpp:74-106)
inline void score_docked_b200(gb_vina* h, const model& m, output_container& out_cont, const std::vector<std::vector<float>>& pose_xyz,
                              const vec& cap, const grid_dims& gd, fl slope, float num_tors) {
  const int n = (int)out_cont.size(), na = (int)m.num_movable_atoms();
  std::vector<float> xyz, nt(n, num_tors), aff(n);
  std::vector<int32_t> t, off(1, 0);
  for (int i = 0; i < n; i++) {
    xyz.insert(xyz.end(), pose_xyz[i].begin(), pose_xyz[i].end());
    for (int a = 0; a < na; a++) t.push_back((int32_t)m.atoms[a].sm);
    off.push_back((int32_t)t.size());
  }
  float b[3], en[3];
  box_of(gd, b, en);
  check(gb_vina_score_noncache(h, xyz.data(), t.data(), off.data(), n, nt.data(), (float)cap[1], (float)slope, b, en, nullptr, aff.data()));
  for (int i = 0; i < n; i++) if (not_max(out_cont[i].e)) out_cont[i].e = aff[i];
}

}  // namespace b200

#endif
