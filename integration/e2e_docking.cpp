// integration/e2e_docking.cpp -- the drop-in shown end to end in ONE program: gnina's own classes (model, precalculate, cache,
// parallel_mc, output_container -- the reference's sources, linked from oracle/_ref) next to the adapters of docking_b200.h that forward the
// same calls to libgnina_b200.so.
//   e2e_docking cpu   : builds a reference `model` by hand (as test/gnina/test_tree.cu does), converts it with b200::B200Ligand, runs the
//                       REFERENCE's parallel_mc on the host and prints its best poses -- no device needed (tests/test_integration_adapters.py)
//   e2e_docking gpu   : additionally runs b200::parallel_mc_b200 + refine_structure_b200 + score_docked_b200 on the device and prints the two
//                       result lists side by side (the start orientations are drawn differently -- normals in the reference, the unit
//                       ball in the library -- so the lists are compared as search results, not pose by pose).  NOT YET RUN: written
//                       after round 2's GPU minutes were spent.
// Build: oracle/Makefile.ref target _ref/e2e_docking (links oracle/_ref/libgnina_vina_ref.so and gnina_b200/libgnina_b200.so).
#include <cstdio>
#include <cstring>
#include <memory>
#include <random>

#include "cache.h"
#include "custom_terms.h"
#include "parallel_mc.h"
#include "precalculate.h"
#include "weighted_terms.h"
#include "docking_b200.h"

static void build_model(model& m, std::mt19937& rs) {
  // receptor: a shell of atoms around the search box; ligand: root (4 atoms) + two torsion segments
  std::uniform_real_distribution<float> u(-1.f, 1.f);
  const smt rec_types[3] = {smina_atom_type::AliphaticCarbonXSHydrophobe, smina_atom_type::OxygenXSAcceptor, smina_atom_type::NitrogenXSDonor};
  for (int i = 0; i < 400; i++) {
    vec d(u(rs), u(rs), u(rs));
    const fl nrm = d.norm();
    if (nrm < 0.2) { i--; continue; }
    atom a; a.sm = rec_types[i % 3]; a.charge = 0; a.coords = (fl)(9.0 + 3.0 * std::fabs(u(rs))) / nrm * d;
    m.grid_atoms.push_back(a);
  }
  const float pos[9][3] = {{0, 0, 0}, {1.5f, 0, 0}, {2.2f, 1.3f, 0}, {3.7f, 1.4f, 0.2f}, {4.6f, 2.4f, 0.5f}, {5.4f, 3.6f, 0.9f}, {6.8f, 3.4f, 1.4f},
                           {7.6f, 4.3f, 1.9f}, {8.7f, 3.4f, 2.4f}};
  const smt lt[9] = {smina_atom_type::AliphaticCarbonXSHydrophobe, smina_atom_type::AliphaticCarbonXSNonHydrophobe, smina_atom_type::NitrogenXSDonor,
                     smina_atom_type::AliphaticCarbonXSHydrophobe, smina_atom_type::AliphaticCarbonXSHydrophobe, smina_atom_type::OxygenXSAcceptor,
                     smina_atom_type::AliphaticCarbonXSHydrophobe, smina_atom_type::AromaticCarbonXSHydrophobe, smina_atom_type::OxygenXSDonorAcceptor};
  const int seg_begin[3] = {0, 4, 7}, seg_end[3] = {4, 7, 9}, axis_root[3] = {0, 3, 6};
  m.m_num_movable_atoms = 9;
  m.minus_forces = std::vector<vec>(9, vec(0, 0, 0));
  for (int i = 0; i < 9; i++) {
    const int s = i < 4 ? 0 : (i < 7 ? 1 : 2);
    const vec c(pos[i][0], pos[i][1], pos[i][2]), o(pos[seg_begin[s]][0], pos[seg_begin[s]][1], pos[seg_begin[s]][2]);
    m.coords.push_back(c);
    atom a; a.sm = lt[i]; a.charge = 0; a.coords = c - o;
    m.atoms.push_back(a);
  }
  auto P = [&](int i) { return vec(pos[i][0], pos[i][1], pos[i][2]); };
  rigid_body root(P(0), (sz)seg_begin[0], (sz)seg_end[0]);
  m.ligands.push_back(ligand(flexible_body(root), 2));
  ligand& lig = m.ligands[0];
  segment s1(P(seg_begin[1]), (sz)seg_begin[1], (sz)seg_end[1], P(axis_root[1]), lig.node);
  lig.children.push_back(branch(s1));
  segment s2(P(seg_begin[2]), (sz)seg_begin[2], (sz)seg_end[2], P(axis_root[2]), lig.children[0].node);
  lig.children[0].children.push_back(branch(s2));
  lig.set_range();
  for (int a = 0; a < 4; a++) for (int b = 7; b < 9; b++) lig.pairs.push_back(interacting_pair(lt[a], lt[b], (sz)a, (sz)b));
}

static void print_container(const char* tag, const output_container& out) {
  for (sz i = 0; i < out.size() && i < 5; i++) printf("%s pose %zu e %.5f\n", tag, (size_t)i, (double)out[i].e);
}

int main(int argc, char** argv) {
  const bool gpu = argc >= 2 && !strcmp(argv[1], "gpu");
  std::mt19937 rs(7);
  model m;
  build_model(m, rs);
  custom_terms t;
  t.add("gauss(o=0,_w=0.5,_c=8)", -0.035579); t.add("gauss(o=3,_w=2,_c=8)", -0.005156); t.add("repulsion(o=0,_c=8)", 0.840245);
  t.add("hydrophobic(g=0.5,_b=1.5,_c=8)", -0.035069); t.add("non_dir_h_bond(g=-0.7,_b=0,_c=8)", -0.587439);
  t.add("num_tors_div", 5 * 0.05846 / 0.1 - 1);
  weighted_terms wt(&t, t.weights());
  precalculate_linear prec(wt, 32);
  grid_dims gd;
  for (int i = 0; i < 3; i++) { gd[i].n = 32; gd[i].begin = -6.05; gd[i].end = gd[i].begin + 0.375 * 32; }
  const vec corner1(gd[0].begin, gd[1].begin, gd[2].begin), corner2(gd[0].end, gd[1].end, gd[2].end);

  // the adapter's view of the ligand
  b200::B200Ligand L(m);
  printf("topology atoms %d segments %d pairs %d heavy %d gyration %.4f\n", L.topo.n_atoms, L.topo.n_segments, L.topo.n_pairs, L.n_heavy,
         (double)L.topo.gyration_radius);

  // the reference on the host
  parallel_mc par;
  par.mc.num_steps = 60; par.mc.ssd_par.evals = (25 + 9) / 3; par.mc.ssd_par.minparm.maxiters = par.mc.ssd_par.evals;
  par.mc.min_rmsd = 1.0; par.mc.num_saved_mins = 20; par.mc.hunt_cap = vec(10, 10, 10);
  par.num_tasks = 4; par.num_threads = 2; par.display_progress = false;
  cache c("scoring_function_version001", gd, 1e3);
  std::vector<smt> needed;
  m.get_movable_atom_types(needed);
  grid user_grid;
  c.populate(m, prec, needed, user_grid, false);
  szv_grid_cache gridcache(m, prec.cutoff_sqr());
  non_cache nc(gridcache, gd, &prec, 1e3);
  rng generator(static_cast<rng::result_type>(42));
  output_container ref_out;
  par(m, ref_out, prec, c, corner1, corner2, generator, user_grid, nc);
  print_container("reference", ref_out);
  if (!gpu) { printf("cpu ok\n"); return 0; }

  // the same search through the adapters on the device
  gb_vina* h = nullptr;
  b200::check(gb_vina_create(0, nullptr, 32.f, &h));
  std::vector<float> rx; std::vector<int32_t> rt;
  for (const atom& a : m.grid_atoms) { rt.push_back((int32_t)a.sm); for (int k = 0; k < 3; k++) rx.push_back((float)a.coords[k]); }
  b200::check(gb_vina_set_receptor(h, rx.data(), rt.data(), (int)rt.size()));
  b200::cache_b200 dc(h, gd, 1e3, needed);
  b200::parallel_mc_b200 bpar;
  bpar.mc = par.mc; bpar.num_tasks = par.num_tasks; bpar.h = h;
  rng generator2(static_cast<rng::result_type>(42));
  output_container dev_out;
  bpar(m, dev_out, corner1, corner2, generator2);
  print_container("device   ", dev_out);
  b200::refine_structure_b200(h, dev_out, vec(1000, 1000, 1000), par.mc.ssd_par.minparm, gd);
  print_container("refined  ", dev_out);
  gb_vina_destroy(h);
  return 0;
}
