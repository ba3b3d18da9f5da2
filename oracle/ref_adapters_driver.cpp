// oracle/ref_adapters_driver.cpp -- the docking-side integration adapters (integration/docking_b200.h) EXECUTED on the CPU inside the
// reference's own classes.  b200::cache_b200 is an `igrid` (lib/igrid.h) over gb_vina_cache_build / gb_vina_cache_eval, and
// b200::score_docked_b200 forwards the docking branch's final score to gb_vina_score_noncache.  Here those C-ABI names are redirected to
// a stand-in that honours include/gnina_b200.h's contract with the oracle's C restatement (oracle/vina_ref.c -- what the device kernels
// are checked against on the GPU), so that the adapter code itself (coordinate / type marshalling, minus_forces hand-back, box and
// num_tors plumbing) runs inside the reference's model::eval_deriv, quasi_newton and monte_carlo in place of its `cache`.
// TEST INFRASTRUCTURE (oracle/); linked into oracle/_ref/libgnina_cnn_ref.so (which already links liboracle.so and the Vina build).
#include <cstring>
#include <string>
#include <vector>

#include "cache.h"
#include "igrid.h"
#include "model.h"
#include "gnina_b200.h"

extern "C" {
typedef struct gvo_prec gvo_prec;
gvo_prec* gvo_prec_create(const float* weights6, float factor);
void gvo_prec_free(gvo_prec* p);
void gvo_cache_populate(const gvo_prec* p, const float* begin, const float* end, const int32_t* n, int n_rec, const float* rec_xyz,
                        const int32_t* rec_type, int t2, float* out);
float gvo_cache_eval(float* const* grids, const float* begin, const float* end, const int32_t* n, int n_lig, const float* lig_xyz,
                     const int32_t* lig_type, float slope, float v, float* deriv);
float gvo_noncache_eval(const gvo_prec* p, int n_rec, const float* rec_xyz, const int32_t* rec_type, int n_lig, const float* lig_xyz,
                        const int32_t* lig_type, float v, float slope, const float* begin, const float* end);
float gvo_num_tors_div(const gvo_prec* p, float e, float num_tors);
typedef struct {                     /* oracle/vina_mc_ref.c */
  int n_atoms, n_seg, n_pairs;
  const float* local_xyz; const int32_t* type; const int32_t* seg_parent; const int32_t *seg_begin, *seg_end;
  const float *seg_rel_origin, *seg_rel_axis; const int32_t *pair_a, *pair_b;
} gvo_lig;
typedef struct {
  float* const* grids; const float *begin, *end; const int32_t* n; float slope; const gvo_prec* prec; const void* splines;
  const float* rec_xyz; const int32_t* rec_type; int n_rec;
} gvo_field;
typedef struct { int num_steps, maxiters, num_saved_mins; float temperature, mutation_amplitude, min_rmsd; float hunt_cap[3]; float gyration_radius; } gvo_mc_params;
int gvo_mc_run_ex(const gvo_field* F, const gvo_lig* L, const gvo_mc_params* P, const float* corner1, const float* corner2, uint32_t seed,
                  float* out_e, float* out_conf, float* trace, const float* init_conf, const float* state_conf);
void gvo_lig_set_conf(const gvo_lig* L, const float* x, float* coords, float* seg_origin, float* seg_axis);
int gref_random_conf(void* mp, unsigned seed, const float* c1, const float* c2, float* x, unsigned* state_after);
int gb_vina_merge_outputs(const float* e, const float* coords, const int32_t* n_out, int n_chains, int S, int n_atoms, float min_rmsd,
                          int max_size, int32_t* kept, int32_t* n_kept);   /* the library's own host-side merge (libgnina_b200.so) */
float gvo_refine_structure_ex(const gvo_field* F0, const gvo_lig* L, float* x, float* g, int maxiters, const float* v, int* n_evals,
                              int* within_out, int accurate, int early_term);
void* gref_model_ptr(void* p);
void* gref_grid_wrap(void* ig);
}

struct gb_vina {
  gvo_prec* prec = nullptr;
  std::vector<float> rec; std::vector<int32_t> rec_t;
  float begin[3], end[3]; int32_t n[3];
  std::vector<std::vector<float>> grids = std::vector<std::vector<float>>(28);
  gvo_lig lig{};
  float gyration_radius = 0;
  void* model_handle = nullptr;      // for the reference-style start conformations of the chains (see mockgb_vina_mc)
  std::vector<float> state_conf;     // the conformation the model holds when the search starts (model-state semantics of the chains)
};
static std::string g_mock_err;
extern "C" {
const char* mockgbv_last_error(void) { return g_mock_err.c_str(); }
int mockgb_vina_cache_build(gb_vina* h, const float* begin, const float* end, const int32_t* n, const int32_t* types, int n_types) {
  for (int i = 0; i < 3; i++) { h->begin[i] = begin[i]; h->end[i] = end[i]; h->n[i] = n[i]; }
  const size_t pts = (size_t)(n[0] + 1) * (n[1] + 1) * (n[2] + 1);
  for (auto& g : h->grids) g.clear();
  for (int k = 0; k < n_types; k++) {
    if (types[k] < 0 || types[k] >= 28) { g_mock_err = "bad atom type"; return GB_ERR_USAGE; }
    h->grids[types[k]].resize(pts);
    gvo_cache_populate(h->prec, begin, end, n, (int)h->rec_t.size(), h->rec.data(), h->rec_t.data(), types[k], h->grids[types[k]].data());
  }
  return GB_OK;
}
int mockgb_vina_cache_eval(gb_vina* h, const float* xyz, const int32_t* t, const int32_t* offs, int n_poses, float slope, float v,
                           float* energy, float* deriv) {
  float* gp[28];
  for (int i = 0; i < 28; i++) gp[i] = h->grids[i].empty() ? nullptr : h->grids[i].data();
  for (int p = 0; p < n_poses; p++) {
    const int a = offs[p], n = offs[p + 1] - offs[p];
    for (int i = 0; i < n; i++)
      if (t[a + i] > 1 && t[a + i] < 28 && !gp[t[a + i]]) { g_mock_err = "no grid for a ligand atom type"; return GB_ERR_USAGE; }
    energy[p] = gvo_cache_eval(gp, h->begin, h->end, h->n, n, xyz + 3 * a, t + a, slope, v, deriv ? deriv + 3 * a : nullptr);
  }
  return GB_OK;
}
int mockgb_vina_score_noncache(gb_vina* h, const float* xyz, const int32_t* t, const int32_t* offs, int n_poses, const float* num_tors,
                               float vcap, float slope, const float* bb, const float* be, float* e_inter, float* affinity) {
  for (int p = 0; p < n_poses; p++) {
    const int a = offs[p], n = offs[p + 1] - offs[p];
    const float e = gvo_noncache_eval(h->prec, (int)h->rec_t.size(), h->rec.data(), h->rec_t.data(), n, xyz + 3 * a, t + a, vcap, slope, bb, be);
    if (e_inter) e_inter[p] = e;
    if (affinity) affinity[p] = gvo_num_tors_div(h->prec, e, num_tors[p]);
  }
  return GB_OK;
}
}
extern "C" {
int mockgb_vina_set_ligand(gb_vina* h, const gb_ligand_topology* t) {
  h->lig = gvo_lig{t->n_atoms, t->n_segments, t->n_pairs, t->local_xyz, t->smina_type, t->seg_parent, t->seg_atom_begin, t->seg_atom_end,
                   t->seg_rel_origin, t->seg_rel_axis, t->pair_a, t->pair_b};
  h->gyration_radius = t->gyration_radius;
  return GB_OK;
}
// refine_structure on the non_cache field (direct sums over the receptor, search box = [bb, be]); e = the last run's energy, within[i]
int mockgb_vina_refine_minimize(gb_vina* h, float* confs, int n, const gb_minimization_params* mp, const float* v3, const float* bb,
                                const float* be, float* e, int32_t* within, int32_t* n_evals) {
  float* no_grids[28] = {};
  const int32_t nn[3] = {1, 1, 1};
  const int nx = 7 + h->lig.n_seg - 1;
  std::vector<float> g(6 + h->lig.n_seg - 1);
  for (int i = 0; i < n; i++) {
    gvo_field F{no_grids, bb, be, nn, 10.f, h->prec, nullptr, h->rec.data(), h->rec_t.data(), (int)h->rec_t.size()};
    int ne = 0, ok = 0;
    e[i] = gvo_refine_structure_ex(&F, &h->lig, confs + (size_t)i * nx, g.data(), mp->maxiters, v3, &ne, &ok, mp->accurate_line_search, mp->early_term);
    if (within) within[i] = ok;
    if (n_evals) n_evals[i] = ne;
  }
  return GB_OK;
}
}
extern "C" {
// every chain = the restatement's monte_carlo chain on the cache grids.  The library draws a chain's start conformation itself (unit-ball
// orientation); HERE the start is drawn the reference's way (conf::randomize on the chain's generator, lib/conf.h) so that what the
// adapter builds can be compared with parallel_mc::operator() pose by pose -- the adapter's plumbing is what this stand-in is for.
int mockgb_vina_mc(gb_vina* h, const gb_mc_params* P, const float* c1, const float* c2, const uint32_t* seeds, int n_chains, float slope,
                   float* e, float* confs, int32_t* n_out) {
  float* gp[28];
  for (int i = 0; i < 28; i++) gp[i] = h->grids[i].empty() ? nullptr : h->grids[i].data();
  gvo_field F{gp, h->begin, h->end, h->n, slope, h->prec, nullptr, nullptr, nullptr, 0};
  gvo_mc_params Q{P->num_steps, P->maxiters, P->num_saved_mins, P->temperature, P->mutation_amplitude, P->min_rmsd,
                  {P->hunt_cap[0], P->hunt_cap[1], P->hunt_cap[2]}, h->gyration_radius};
  const int S = P->num_saved_mins, nx = 7 + h->lig.n_seg - 1;
  std::vector<float> x0(nx);
  for (int c = 0; c < n_chains; c++) {
    unsigned state = 0;
    if (gref_random_conf(h->model_handle, seeds[c], c1, c2, x0.data(), &state)) { g_mock_err = "random_conf failed"; return GB_ERR_INTERNAL; }
    n_out[c] = gvo_mc_run_ex(&F, &h->lig, &Q, c1, c2, state, e + (size_t)c * S, confs + (size_t)c * S * nx, nullptr, x0.data(),
                             h->state_conf.empty() ? nullptr : h->state_conf.data());
  }
  return GB_OK;
}
// only the coordinates are asked for by parallel_mc_b200 (model::set of every minimum)
int mockgb_vina_eval_deriv(gb_vina* h, const float* confs, int n, const float*, float, float* e, float* change, float* coords) {
  if (change) { g_mock_err = "the stand-in returns coordinates only"; return GB_ERR_USAGE; }
  const int nx = 7 + h->lig.n_seg - 1, na = h->lig.n_atoms;
  std::vector<float> so(3 * (size_t)h->lig.n_seg), sa(3 * (size_t)h->lig.n_seg);
  for (int i = 0; i < n; i++) { gvo_lig_set_conf(&h->lig, confs + (size_t)i * nx, coords + (size_t)i * na * 3, so.data(), sa.data()); if (e) e[i] = 0; }
  return GB_OK;
}
}
#define gb_vina_mc mockgb_vina_mc
#define gb_vina_eval_deriv mockgb_vina_eval_deriv
#define gb_vina_set_ligand mockgb_vina_set_ligand
#define gb_vina_refine_minimize mockgb_vina_refine_minimize
#define gb_last_error mockgbv_last_error
#define gb_vina_cache_build mockgb_vina_cache_build
#define gb_vina_cache_eval mockgb_vina_cache_eval
#define gb_vina_score_noncache mockgb_vina_score_noncache
#include "docking_b200.h"

static grid_dims dims_of(const float* begin, const float* end, const int* n) {
  grid_dims gd;
  for (int i = 0; i < 3; i++) { gd[i].begin = begin[i]; gd[i].end = end[i]; gd[i].n = (sz)n[i]; }
  return gd;
}

extern "C" {
// gb_vina_create + gb_vina_set_receptor of the stand-in: default weights, table factor 32, hydrogens dropped like the library does
void* gadp_vina_create(const float* rec_xyz, const int* rec_t, int n) {
  gb_vina* h = new gb_vina;
  h->prec = gvo_prec_create(nullptr, 32.f);
  for (int i = 0; i < n; i++) {
    if (rec_t[i] < 2) continue;
    h->rec_t.push_back(rec_t[i]);
    for (int k = 0; k < 3; k++) h->rec.push_back(rec_xyz[3 * i + k]);
  }
  return h;
}
void gadp_vina_destroy(void* p) { gb_vina* h = (gb_vina*)p; gvo_prec_free(h->prec); delete h; }
// b200::cache_b200(h, gd, slope, needed) as a grid handle of oracle/ref_driver.cpp (gref_model_eval_deriv, gref_bfgs, gref_mc take it)
void* gadp_cache_b200(void* h, const float* begin, const float* end, const int* n, float slope, const int* needed, int n_needed) {
  try {
    std::vector<smt> nd;
    for (int i = 0; i < n_needed; i++) nd.push_back((smt)needed[i]);
    return gref_grid_wrap(static_cast<igrid*>(new b200::cache_b200((gb_vina*)h, dims_of(begin, end, n), slope, nd)));
  } catch (const std::exception& e) { g_mock_err = e.what(); return nullptr; }
}
// b200::score_docked_b200 on n poses given by their coordinates; e [n] in (what refine_structure left: max_fl = never inside) / out
int gadp_score_docked(void* h, void* model_handle, const float* pose_xyz, int n_poses, const float* begin, const float* end, const int* n,
                      float slope, const float* cap3, float num_tors, float* e) {
  try {
    model& m = *(model*)gref_model_ptr(model_handle);
    const int na = (int)m.num_movable_atoms();
    output_container out;
    std::vector<std::vector<float>> xyz(n_poses);
    for (int i = 0; i < n_poses; i++) {
      out.push_back(new output_type(conf(m.get_size(), false), e[i]));
      xyz[i].assign(pose_xyz + (size_t)i * na * 3, pose_xyz + (size_t)(i + 1) * na * 3);
    }
    b200::score_docked_b200((gb_vina*)h, m, out, xyz, vec(cap3[0], cap3[1], cap3[2]), dims_of(begin, end, n), slope, num_tors);
    for (int i = 0; i < n_poses; i++) e[i] = out[i].e;
    return 0;
  } catch (const std::exception& ex) { g_mock_err = ex.what(); return 1; }
}
// b200::refine_structure_b200 on n conformations of the model's ligand (topology through b200::B200Ligand, as parallel_mc_b200 sets it):
// confs [n][7+T] in/out, e [n] out (max_fl = never inside the box)
int gadp_refine_structure(void* h, void* model_handle, float* confs, int n_confs, const float* begin, const float* end, const int* n,
                          const float* cap3, int maxiters, int accurate, int early_term, float* e) {
  try {
    model& m = *(model*)gref_model_ptr(model_handle);
    b200::B200Ligand L(m);
    b200::check(gb_vina_set_ligand((gb_vina*)h, &L.topo));
    const int nx = 7 + L.topo.n_segments - 1;
    output_container out;
    for (int i = 0; i < n_confs; i++) {
      conf c(m.get_size(), false);
      b200::unpack_conf(confs + (size_t)i * nx, c);
      out.push_back(new output_type(c, 0));
    }
    minimization_params mp;
    mp.maxiters = (unsigned)maxiters;
    mp.type = accurate ? minimization_params::BFGSAccurateLineSearch : minimization_params::BFGSFastLineSearch;
    mp.early_term = early_term != 0;
    b200::refine_structure_b200((gb_vina*)h, out, vec(cap3[0], cap3[1], cap3[2]), mp, dims_of(begin, end, n));
    for (int i = 0; i < n_confs; i++) { b200::pack_conf(out[i].c, confs + (size_t)i * nx); e[i] = out[i].e; }
    return 0;
  } catch (const std::exception& ex) { g_mock_err = ex.what(); return 1; }
}
// b200::parallel_mc_b200::operator() with the reference's generator type seeded like parallel_mc's caller; the model holds state_conf.
// -> the merged container (sorted by energy as parallel_mc::operator() leaves it): e, confs [n][7+T]
int gadp_parallel_mc(void* hv, void* model_handle, unsigned seed, const float* c1, const float* c2, int num_tasks, int num_steps, int maxiters,
                     int num_saved_mins, float min_rmsd, const float* hunt_cap, const float* state_conf, int nx_state, int max_out, float* out_e,
                     float* out_conf, int* n_out) {
  try {
    gb_vina* h = (gb_vina*)hv;
    h->model_handle = model_handle;
    h->state_conf.assign(state_conf, state_conf + nx_state);
    model& m = *(model*)gref_model_ptr(model_handle);
    b200::parallel_mc_b200 par;
    par.h = h; par.num_tasks = (sz)num_tasks;
    par.mc.num_steps = (unsigned)num_steps; par.mc.ssd_par.minparm.maxiters = (unsigned)maxiters; par.mc.num_saved_mins = (sz)num_saved_mins;
    par.mc.min_rmsd = min_rmsd; par.mc.hunt_cap = vec(hunt_cap[0], hunt_cap[1], hunt_cap[2]);
    rng gen(seed);
    output_container out;
    par(m, out, vec(c1[0], c1[1], c1[2]), vec(c2[0], c2[1], c2[2]), gen);
    out.sort();                                                   // parallel_mc.cpp:213
    const int nx = b200::conf_floats(out.empty() ? m.get_initial_conf(false) : out[0].c);
    *n_out = (int)std::min<sz>(out.size(), (sz)max_out);
    for (int i = 0; i < *n_out; i++) { out_e[i] = out[i].e; b200::pack_conf(out[i].c, out_conf + (size_t)i * nx); }
    return 0;
  } catch (const std::exception& ex) { g_mock_err = ex.what(); return 1; }
}
const char* gadp_last_error() { return g_mock_err.c_str(); }
}
