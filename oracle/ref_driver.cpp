// oracle/_ref driver (TEST INFRASTRUCTURE, never on the product path): a C ABI over the REFERENCE's own Vina code,
// compiled from the sources where they lie under /root/reference/gninasrc/lib (oracle/Makefile.ref). Nothing in this file
// computes a score: every number it returns comes out of the reference's classes -- custom_terms / weighted_terms /
// precalculate_{linear,splines,exact}, cache, non_cache, naive_non_cache, model (set, eval_deriv, eval_adjusted),
// quasi_newton (bfgs.h), monte_carlo (+ mutate.cpp, coords.cpp), conf / tree.h / quaternion.h.
//
// How the model is built: by hand, the way the reference's own unit tests do it (test/gnina/test_cache.cu:93-113,
// test/gnina/test_tree.cu:36-62): coords / atoms / grid_atoms / m_num_movable_atoms / minus_forces filled in directly, the
// ligand's torsion tree from rigid_body / segment constructors, interacting pairs pushed as data. No OpenBabel, no parser.
//
// Absent third-party headers (Boost, OpenBabel) are replaced by the stand-ins under oracle/ref_shim/ (containers, regex,
// optional, ... mapped to std::; no arithmetic). One of them matters for what is compared: boost/random.hpp's engine is the
// xorshift32 generator the oracle and the device kernels use, with their uniform-real / uniform-int mappings, so that the
// REFERENCE's monte_carlo / mutate_conf / conf::randomize code runs on the same random stream as the restatement. The
// reference's normal_distribution (random_orientation) is Box-Muller over the same stream -- see gref_random_conf.
//
// A handful of private data members are READ (cache::grids, grid::data, segment::relative_*, weighted_terms::weights) through
// the explicit-instantiation idiom below, which the language allows to name private members; no reference source is modified
// or copied.
#include <algorithm>
#include <array>
#include <cassert>
#include <cmath>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <functional>
#include <iomanip>
#include <iostream>
#include <limits>
#include <memory>
#include <mutex>
#include <numeric>
#include <optional>
#include <random>
#include <regex>
#include <sstream>
#include <string>
#include <thread>
#include <unordered_map>
#include <utility>
#include <vector>
#include <cuda_runtime.h>
#include "Eigen/Core"
#include "Eigen/Dense"
#include "boost/ptr_container/ptr_vector.hpp"
#include "boost/optional.hpp"
#include "boost/random.hpp"
#include "boost/iostreams/filtering_stream.hpp"
#include "boost/functional/hash.hpp"
#include "boost/unordered_map.hpp"
#include "boost/array.hpp"

#include "cache.h"
#include "coords.h"
#include "custom_terms.h"
#include "everything.h"
#include "model.h"
#include "monte_carlo.h"
#include "mutate.h"
#include "naive_non_cache.h"
#include "non_cache.h"
#include "precalculate.h"
#include "quasi_newton.h"
#include "szv_grid.h"
#include "weighted_terms.h"
#include "non_cache_cnn.h"
#include "parallel_mc.h"
#include "gnina_b200.hpp"   // this repo's host-side C++ (header only): gb::NonCacheCNNT is run against non_cache_cnn below
#include "gnina_b200_minimize.hpp"   // gb::minimize_poses (lock-step quasi-Newton) is run against quasi_newton + non_cache_cnn below
#include "docking_b200.h"   // integration/: the model -> gb_ligand_topology adapter a gnina maintainer adds; exercised below

// read access to private data members: an explicit template instantiation may name them ([temp.spec]/6)
template <class Tag, typename Tag::type M> struct Peek { friend typename Tag::type peek(Tag) { return M; } };
#define GREF_PEEK(Tag, Class, Type, member) \
  struct Tag { typedef Type Class::*type; friend type peek(Tag); }; \
  template struct Peek<Tag, &Class::member>
GREF_PEEK(CacheGrids, cache, std::vector<grid>, grids);
GREF_PEEK(GridData, grid, array3d<fl>, data);
GREF_PEEK(SegRelAxis, segment, vec, relative_axis);
GREF_PEEK(SegRelOrigin, segment, vec, relative_origin);
GREF_PEEK(WtWeights, weighted_terms, flv, weights);
GREF_PEEK(WtConfIndepStart, weighted_terms, sz, conf_indep_start);

namespace {

struct RefSF {
  custom_terms t;
  std::unique_ptr<weighted_terms> wt;
  std::unique_ptr<precalculate> prec[3];  // 0 linear, 1 splines, 2 exact
};

struct RefModel {
  model m;
  std::unique_ptr<szv_grid_cache> gcache;  // non_cache's receptor-atom lists
  int n_seg = 0;
  std::vector<vec> seg_origin;  // construction-time origins (identity orientation)
  std::vector<vec> rel_origin, rel_axis;
};

struct RefGrid {
  std::unique_ptr<igrid> ig;
  int kind = 0;  // 0 cache, 1 non_cache, 2 naive_non_cache
};

conf make_conf(const model& m, const float* x) {
  conf c(m.get_size(), false);
  ligand_conf& l = c.ligands[0];
  l.rigid.position = vec(x[0], x[1], x[2]);
  l.rigid.orientation = qt(x[3], x[4], x[5], x[6]);
  for (sz i = 0; i < l.torsions.size(); i++) l.torsions[i] = x[7 + i];
  return c;
}
void read_conf(const conf& c, float* x) {
  const ligand_conf& l = c.ligands[0];
  for (int k = 0; k < 3; k++) x[k] = l.rigid.position[k];
  x[3] = l.rigid.orientation.R_component_1();
  x[4] = l.rigid.orientation.R_component_2();
  x[5] = l.rigid.orientation.R_component_3();
  x[6] = l.rigid.orientation.R_component_4();
  for (sz i = 0; i < l.torsions.size(); i++) x[7 + i] = l.torsions[i];
}
void read_change(const change& g, float* o) {
  const ligand_change& l = g.ligands[0];
  for (int k = 0; k < 3; k++) { o[k] = l.rigid.position[k]; o[3 + k] = l.rigid.orientation[k]; }
  for (sz i = 0; i < l.torsions.size(); i++) o[6 + i] = l.torsions[i];
}
grid_dims make_dims(const float* begin, const float* end, const int* n) {
  grid_dims gd;
  for (int i = 0; i < 3; i++) { gd[i].begin = begin[i]; gd[i].end = end[i]; gd[i].n = (sz)n[i]; }
  return gd;
}
atom_base typed(int t) { atom_base a; a.sm = (smt)t; a.charge = 0; return a; }

thread_local std::string g_err;
template <class F> int guarded(F&& f) {
  try { f(); return 0; }
  catch (const std::exception& e) { g_err = e.what(); }
  catch (const internal_error& e) { g_err = std::string("internal_error ") + e.file + ":" + std::to_string(e.line); }
  catch (...) { g_err = "unknown exception from the reference code"; }
  return 1;
}

}  // namespace

// ---- S3: test doubles so that the REFERENCE's non_cache_cnn and this repo's gb::NonCacheCNNT see the same "network" --------------
// loss = k * sum over heavy movable atoms |x_i - target|^2, gradient 2 k (x_i - target), hydrogens untouched: any smooth function
// would do -- what is compared is the code AROUND the network (out-of-box penalties of the search box and of the CNN box, hydrogen
// handling, the empirical mixing of --cnn_mix_emp_force / --cnn_mix_emp_energy)
namespace {
struct AnalyticLoss {
  float k = 0.01f;
  float target[3] = {0, 0, 0};
  float eval(const float* xyz, const int32_t* types, int n, float* grad) const {
    float loss = 0;
    for (int i = 0; i < n; i++) {
      const bool heavy = types[i] >= 2;
      for (int j = 0; j < 3; j++) {
        const float d = xyz[3 * i + j] - target[j];
        if (heavy) loss += k * d * d;
        if (grad) grad[3 * i + j] = heavy ? 2 * k * d : 0.f;
      }
    }
    return loss;
  }
};
class FakeDLScorer : public DLScorer {
  AnalyticLoss L;
  fl dim, res;
 public:
  FakeDLScorer(const cnn_options& o, const AnalyticLoss& l, fl dim_, fl res_) : DLScorer(o), L(l), dim(dim_), res(res_) {}
  bool initialized() const override { return true; }
  bool has_affinity() const override { return true; }
  float score(model& m, float& variance) override { float a, l; return score(m, false, a, l, variance); }
  float score(model& m, bool compute_gradient, float& affinity, float& loss, float& variance) override {
    const int n = (int)m.num_movable_atoms();
    std::vector<float> xyz(3 * (size_t)n), g(3 * (size_t)n);
    std::vector<int32_t> t(n);
    for (int i = 0; i < n; i++) { t[i] = (int32_t)m.atoms[i].sm; for (int j = 0; j < 3; j++) xyz[3 * i + j] = m.coords[i][j]; }
    m.clear_minus_forces();                                                  // CNNTorchScorer::score does (cnn_torch_scorer.cpp:115)
    loss = L.eval(xyz.data(), t.data(), n, compute_gradient ? g.data() : nullptr);
    if (compute_gradient) {                                                  // by-atom list -> add_minus_forces, as :167-173, 209-227
      std::vector<gfloat3> list(n);
      for (int i = 0; i < n; i++) list[i] = gfloat3(g[3 * i], g[3 * i + 1], g[3 * i + 2]);
      m.add_minus_forces(list);
    }
    affinity = 0; variance = 0;
    return std::exp(-loss);
  }
  void set_bounding_box(grid_dims& box) const override {                     // as cnn_torch_scorer.cpp:229-241
    const vec center = get_center();
    const fl n = dim / res, half = dim / 2.0;
    for (unsigned i = 0; i < 3; i++) { box[i].begin = center[i] - half; box[i].end = center[i] + half; box[i].n = n; }
  }
  std::shared_ptr<DLScorer> fresh_copy() const override { return nullptr; }
};
struct FakeScorer {                                                          // the Scorer of gb::NonCacheCNNT
  AnalyticLoss L;
  gb_model_info inf{};
  gb_model_info info(int = 0) const { return inf; }
  float score(const float* xyz, const int32_t* t, int n, bool want_gradient, float& affinity, float& loss, float& variance,
              std::vector<float>* gradient = nullptr) {
    std::vector<float> g(3 * (size_t)n);
    loss = L.eval(xyz, t, n, want_gradient ? g.data() : nullptr);
    if (gradient) *gradient = g;
    affinity = 0; variance = 0;
    return std::exp(-loss);
  }
};
// the Emp of gb::NonCacheCNNT with gb_vina_noncache_atoms's contract (include/gnina_b200.h), computed from the REFERENCE's precalculate
// and the model's receptor atoms in index order: per heavy atom the eval_deriv sum at the position clamped to [begin, end], then curl
struct RefEmp {
  const precalculate* p;
  const model* m;
  void noncache_atoms(const float* xyz, const int32_t* type, int n, const float* b, const float* en, float cap, std::vector<float>& e,
                      std::vector<float>& d) {
    e.assign(n, 0.f); d.assign(3 * (size_t)n, 0.f);
    for (int i = 0; i < n; i++) {
      if (type[i] < 2 || type[i] >= 28) continue;
      vec a(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
      for (int j = 0; j < 3; j++) { if (a[j] < b[j]) a[j] = b[j]; else if (a[j] > en[j]) a[j] = en[j]; }
      fl emp_e = 0; vec emp_d(0, 0, 0);
      atom_base A = typed(type[i]);
      for (const atom& g : m->grid_atoms) {
        if (g.is_hydrogen()) continue;
        vec r; r = a - g.coords;
        const fl r2 = sqr(r);
        if (r2 < p->cutoff_sqr()) { pr ed = p->eval_deriv(A, g, r2); emp_e += ed.first; emp_d += ed.second * r; }
      }
      curl(emp_e, emp_d, (fl)cap);
      e[i] = emp_e; for (int j = 0; j < 3; j++) d[3 * i + j] = emp_d[j];
    }
  }
};
}  // namespace

extern "C" {

// non_cache_cnn::eval / eval_deriv (lib/non_cache_cnn.cpp:33-54,79-169) of the REFERENCE on the model's current coordinates, and
// gb::NonCacheCNNT (include/gnina_b200.hpp) on the same atoms, both around the same analytic "network": -> e and minus_forces of both
int gref_noncache_cnn_compare(void* mp, void* sf, int kind, const float* begin, const float* end, const int* n, float slope, float dim,
                              float res, float k, const float* target, int mix_force, int mix_energy, float weight, float v,
                              int with_deriv, float* e_ref, float* f_ref, float* e_mine, float* f_mine) {
  RefModel* R = (RefModel*)mp; RefSF* S = (RefSF*)sf;
  return guarded([&] {
    model& m = R->m;
    AnalyticLoss L; L.k = k; for (int j = 0; j < 3; j++) L.target[j] = target[j];
    cnn_options o; o.mix_emp_force = mix_force != 0; o.mix_emp_energy = mix_energy != 0; o.empirical_weight = weight;
    FakeDLScorer dl(o, L, dim, res);
    const grid_dims gd = make_dims(begin, end, n);
    non_cache_cnn nc(*R->gcache, gd, S->prec[kind].get(), slope, dl);
    const vecv saved = m.coords;
    nc.adjust_center(m);                                     // set_center_from_model + set_bounding_box, as refine_structure does (main.cpp:136)
    const int na = (int)m.num_movable_atoms();
    grid user_grid;
    if (with_deriv) {
      *e_ref = nc.eval_deriv(m, v, user_grid);
      for (int i = 0; i < na; i++) for (int j = 0; j < 3; j++) f_ref[3 * i + j] = m.minus_forces[i][j];
    } else *e_ref = nc.eval(m, v);
    // this repo's host code on the same atoms
    FakeScorer fs; fs.L = L; fs.inf.dimension = dim; fs.inf.resolution = res;
    RefEmp emp{S->prec[kind].get(), &m};
    gb::GridDims g3;
    for (int i = 0; i < 3; i++) { g3[i].begin = begin[i]; g3[i].end = end[i]; g3[i].n = n[i]; }
    const vec c = dl.get_center();
    const float ctr[3] = {(float)c[0], (float)c[1], (float)c[2]};
    gb::NonCacheCNNT<FakeScorer, RefEmp> mine(fs, g3, ctr, slope);
    mine.set_empirical(&emp, weight, mix_force != 0, mix_energy != 0);
    std::vector<float> xyz(3 * (size_t)na), forces;
    std::vector<int32_t> t(na);
    for (int i = 0; i < na; i++) { t[i] = (int32_t)m.atoms[i].sm; for (int j = 0; j < 3; j++) xyz[3 * i + j] = saved[i][j]; }
    *e_mine = mine.eval(xyz.data(), t.data(), na, with_deriv ? &forces : nullptr, v);
    if (with_deriv) std::copy(forces.begin(), forces.end(), f_mine);
  });
}

// non_cache_cnn::eval / eval_deriv (lib/non_cache_cnn.cpp:33-54, 79-169) around ANY DLScorer on the pose the model holds, after
// adjust_center: -> e, minus_forces [n_movable][3], the CNN box centre
int gref_noncache_dl_eval(void* mp, void* sf, int kind, const float* begin, const float* end, const int* n, float slope, void* dl, float v,
                          int with_deriv, float* e, float* forces, float* center) {
  RefModel* R = (RefModel*)mp; RefSF* S = (RefSF*)sf;
  return guarded([&] {
    model& m = R->m;
    DLScorer& d = *(DLScorer*)dl;
    non_cache_cnn nc(*R->gcache, make_dims(begin, end, n), S->prec[kind].get(), slope, d);
    nc.adjust_center(m);
    grid user_grid;
    if (with_deriv) {
      *e = nc.eval_deriv(m, v, user_grid);
      for (sz i = 0; i < m.minus_forces.size(); i++) for (int j = 0; j < 3; j++) forces[3 * i + j] = m.minus_forces[i][j];
    } else *e = nc.eval(m, v);
    const vec c = d.get_center();
    for (int j = 0; j < 3; j++) center[j] = c[j];
  });
}

// quasi_newton::operator() with ig = non_cache_cnn around the analytic test double: what --minimize --cnn_scoring all runs per pose
// (main/main.cpp:264-268 -> refine_structure -> quasi_newton; here ONE quasi-Newton run, no slope escalation); x is updated
static void minimize_over(RefModel* R, RefSF* S, int kind, const float* begin, const float* end, const int* n, float slope, DLScorer& dl,
                          float* x, int maxiters, int accurate, int early_term, float* e) {
  model& m = R->m;
  non_cache_cnn nc(*R->gcache, make_dims(begin, end, n), S->prec[kind].get(), slope, dl);
  m.set(make_conf(m, x));
  nc.adjust_center(m);
  minimization_params mp_;
  mp_.maxiters = (unsigned)maxiters;
  mp_.type = accurate ? minimization_params::BFGSAccurateLineSearch : minimization_params::BFGSFastLineSearch;
  mp_.early_term = early_term != 0;
  quasi_newton qn(mp_);
  output_type out(make_conf(m, x), 0);
  change g(m.get_size(), false);
  grid user_grid;
  qn(m, *S->prec[kind], nc, out, g, vec(1000, 1000, 1000), user_grid);
  *e = out.e;
  read_conf(out.c, x);
}
int gref_minimize_cnn(void* mp, void* sf, int kind, const float* begin, const float* end, const int* n, float slope, float dim, float res,
                      float k, const float* target, float* x, int maxiters, int accurate, int early_term, float* e) {
  return guarded([&] {
    AnalyticLoss L; L.k = k; for (int j = 0; j < 3; j++) L.target[j] = target[j];
    cnn_options o;
    FakeDLScorer dl(o, L, dim, res);
    minimize_over((RefModel*)mp, (RefSF*)sf, kind, begin, end, n, slope, dl, x, maxiters, accurate, early_term, e);
  });
}
// the same run over ANY DLScorer -- oracle/ref_cnn_driver.cpp passes the reference's own CNNTorchScorer (real networks on libtorch)
int gref_minimize_dl(void* mp, void* sf, int kind, const float* begin, const float* end, const int* n, float slope, void* dl, float* x,
                     int maxiters, int accurate, int early_term, float* e) {
  return guarded([&] { minimize_over((RefModel*)mp, (RefSF*)sf, kind, begin, end, n, slope, *(DLScorer*)dl, x, maxiters, accurate, early_term, e); });
}

// this repo's C++ lock-step minimiser (include/gnina_b200_minimize.hpp) on n conformations of the model's ligand at once, with
// gb::NonCacheCNNT around the same analytic stand-in as the energy of every pose (its CNN box centred on the pose's start conformation, as
// adjust_center does); the kinematics use this host's sinf / cosf / acosf like the reference build.  confs [n][7+T] in/out
int gref_lockstep_minimize(void* mp, const float* begin, const float* end, const int* nbox, float slope, float dim, float res, float k,
                           const float* target, float* confs, int n, int maxiters, int accurate, int early_term, float* e_out, int* evals_out,
                           int* rounds_out, int* energy_calls_out) {
  RefModel* R = (RefModel*)mp;
  return guarded([&] {
    b200::B200Ligand BL(R->m);
    gb::LigandTree tree(BL.topo);
    gb::Transcendentals saved = gb::transcendentals();
    gb::transcendentals().sin = [](float x) { return sinf(x); };
    gb::transcendentals().cos = [](float x) { return cosf(x); };
    gb::transcendentals().acos = [](float x) { return acosf(x); };
    const int na = tree.n_atoms, ns = tree.n_seg, nx = tree.conf_floats();
    FakeScorer fs; fs.L.k = k; for (int j = 0; j < 3; j++) fs.L.target[j] = target[j];
    fs.inf.dimension = dim; fs.inf.resolution = res;
    gb::GridDims g3;
    for (int i = 0; i < 3; i++) { g3[i].begin = begin[i]; g3[i].end = end[i]; g3[i].n = nbox[i]; }
    // DLScorer::set_center_from_model (lib/dl_scorer.cpp:196-217) for every pose's start conformation
    std::vector<float> centers(3 * (size_t)n), c(3 * (size_t)na), so(3 * (size_t)ns), sa(3 * (size_t)ns);
    for (int i = 0; i < n; i++) {
      tree.set_conf(confs + (size_t)i * nx, c.data(), so.data(), sa.data());
      float cen[3] = {0, 0, 0}; unsigned cnt = 0;
      for (int a = 0; a < na; a++) if (tree.heavy(a)) { for (int j = 0; j < 3; j++) cen[j] += c[3 * a + j]; cnt++; }
      for (int j = 0; j < 3; j++) centers[3 * i + j] = cen[j] / (float)cnt;
    }
    int calls = 0;
    auto energy = [&](const float* coords, const int* pose, int kk, float* e, float* forces) {
      calls++;
      for (int j = 0; j < kk; j++) {
        gb::NonCacheCNNT<FakeScorer, RefEmp> nc(fs, g3, &centers[3 * pose[j]], slope);
        std::vector<float> f;
        e[j] = nc.eval(coords + (size_t)j * na * 3, tree.type.data(), na, &f);
        std::copy(f.begin(), f.end(), forces + (size_t)j * na * 3);
      }
    };
    gb::MinimizeParams mpar; mpar.maxiters = maxiters; mpar.accurate_line_search = accurate != 0; mpar.early_term = early_term != 0;
    std::vector<int> ev; int rounds = 0;
    std::vector<float> e = gb::minimize_poses(tree, energy, confs, n, mpar, &ev, &rounds);
    gb::transcendentals() = saved;
    std::copy(e.begin(), e.end(), e_out);
    std::copy(ev.begin(), ev.end(), evals_out);
    *rounds_out = rounds; *energy_calls_out = calls;
  });
}

// refine_structure (main/main.cpp:131-171, which is not a library source) replayed with the REFERENCE's parts on ig = non_cache_cnn:
// adjust_center once, then up to five quasi_newton runs with the slope 10, 100, ... until non_cache_cnn::within -- the refinement of
// --cnn_scoring refinement / all.  x is updated; *inside tells whether the pose ended within
static void refine_over(RefModel* R, RefSF* S, int kind, const float* begin, const float* end, const int* n, DLScorer& dl, float* x, int maxiters,
                        int accurate, int early_term, float* e, int* inside) {
  model& m = R->m;
  non_cache_cnn nc(*R->gcache, make_dims(begin, end, n), S->prec[kind].get(), 1e3, dl);
  m.set(make_conf(m, x));
  nc.adjust_center(m);
  minimization_params mp_;
  mp_.maxiters = (unsigned)maxiters;
  mp_.type = accurate ? minimization_params::BFGSAccurateLineSearch : minimization_params::BFGSFastLineSearch;
  mp_.early_term = early_term != 0;
  quasi_newton qn(mp_);
  output_type out(make_conf(m, x), 0);
  change g(m.get_size(), false);
  grid user_grid;
  fl slope = 10;
  for (int p = 0; p < 5; p++) {
    nc.setSlope(slope);
    qn(m, *S->prec[kind], nc, out, g, vec(1000, 1000, 1000), user_grid);
    m.set(out.c);
    if (nc.within(m)) break;
    slope *= 10;
  }
  *inside = nc.within(m) ? 1 : 0;
  *e = *inside ? out.e : max_fl;
  read_conf(out.c, x);
}
int gref_refine_cnn(void* mp, void* sf, int kind, const float* begin, const float* end, const int* n, float dim, float res, float k,
                    const float* target, float* x, int maxiters, int accurate, int early_term, float* e, int* inside) {
  return guarded([&] {
    AnalyticLoss L; L.k = k; for (int j = 0; j < 3; j++) L.target[j] = target[j];
    cnn_options o;
    FakeDLScorer dl(o, L, dim, res);
    refine_over((RefModel*)mp, (RefSF*)sf, kind, begin, end, n, dl, x, maxiters, accurate, early_term, e, inside);
  });
}
int gref_refine_dl(void* mp, void* sf, int kind, const float* begin, const float* end, const int* n, void* dl, float* x, int maxiters,
                   int accurate, int early_term, float* e, int* inside) {
  return guarded([&] { refine_over((RefModel*)mp, (RefSF*)sf, kind, begin, end, n, *(DLScorer*)dl, x, maxiters, accurate, early_term, e, inside); });
}

const char* gref_last_error() { return g_err.c_str(); }

// ---- G0: the smina type table (lib/atom_constants.h:45-133, the `data` array the typers and xs_radius() read) ------------------
// flags: bit 0 xs_hydrophobe, 1 xs_donor, 2 xs_acceptor, 3 ad_heteroatom; returns the type string_to_smina_type(name) maps back to
int gref_type_info(int t, char* name64, float* xs_radius_out, float* covalent_radius_out, int* flags) {
  const smina_atom_type::info& d = smina_atom_type::data[t];
  std::strncpy(name64, d.smina_name, 63); name64[63] = 0;
  *xs_radius_out = xs_radius((smt)t);
  *covalent_radius_out = covalent_radius((smt)t);
  *flags = (d.xs_hydrophobe ? 1 : 0) | (d.xs_donor ? 2 : 0) | (d.xs_acceptor ? 4 : 0) | (d.ad_heteroatom ? 8 : 0);
  return (int)string_to_smina_type(d.smina_name);
}

// ---- scoring function: the default Vina terms and weights of main/main.cpp:1324-1329 (= test_cache.cu:30-36) ----------------
void* gref_sf_create_weights(float factor_linear, float factor_splines, const float* w6);
void* gref_sf_create(float factor_linear, float factor_splines) { return gref_sf_create_weights(factor_linear, factor_splines, nullptr); }
// w6 (nullable): the five term weights and the num_tors_div weight of a custom scoring function (--custom_scoring style), same term set
void* gref_sf_create_weights(float factor_linear, float factor_splines, const float* w6) {
  RefSF* s = new RefSF;
  int rc = guarded([&] {
    s->t.add("gauss(o=0,_w=0.5,_c=8)", w6 ? (fl)w6[0] : (fl)-0.035579);
    s->t.add("gauss(o=3,_w=2,_c=8)", w6 ? (fl)w6[1] : (fl)-0.005156);
    s->t.add("repulsion(o=0,_c=8)", w6 ? (fl)w6[2] : (fl)0.840245);
    s->t.add("hydrophobic(g=0.5,_b=1.5,_c=8)", w6 ? (fl)w6[3] : (fl)-0.035069);
    s->t.add("non_dir_h_bond(g=-0.7,_b=0,_c=8)", w6 ? (fl)w6[4] : (fl)-0.587439);
    s->t.add("num_tors_div", w6 ? (fl)w6[5] : (fl)(5 * 0.05846 / 0.1 - 1));
    s->wt.reset(new weighted_terms(&s->t, s->t.weights()));
    s->prec[0].reset(new precalculate_linear(*s->wt, factor_linear));
    s->prec[1].reset(new precalculate_splines(*s->wt, factor_splines));
    s->prec[2].reset(new precalculate_exact(*s->wt));
  });
  if (rc) { delete s; return nullptr; }
  return s;
}
void gref_sf_destroy(void* p) { delete (RefSF*)p; }
float gref_cutoff_sqr(void* p) { return ((RefSF*)p)->prec[0]->cutoff_sqr(); }
// weighted_terms::eval_fast: the weighted sum of the five distance terms at distance r (NOT r^2), lib/weighted_terms.cpp:54-68
float gref_terms_eval(void* p, int t1, int t2, float r) {
  result_components c = ((RefSF*)p)->wt->eval_fast((smt)t1, (smt)t2, r);
  return c.eval(typed(t1), typed(t2));
}
// precalculate::eval(a, b, r2) = eval_fast(t1, t2, r2).eval(a, b)
float gref_prec_eval(void* p, int kind, int t1, int t2, float r2) {
  return ((RefSF*)p)->prec[kind]->eval(typed(t1), typed(t2), r2);
}
// precalculate::eval_deriv -> (e, dE/dr divided by r)
void gref_prec_eval_deriv(void* p, int kind, int t1, int t2, float r2, float* out2) {
  pr v = ((RefSF*)p)->prec[kind]->eval_deriv(typed(t1), typed(t2), r2);
  out2[0] = v.first; out2[1] = v.second;
}

// ---- model ---------------------------------------------------------------------------------------------------------------
// atoms [n_atoms] at their construction-time positions xyz (the conformation "position = root origin, identity orientation,
// torsions 0" reproduces them); segments in DFS pre-order, segment 0 = rigid root; axis_root[s] = atom whose position is the
// begin point of segment s's rotation axis (its end point is the segment origin = first atom of the segment).
void* gref_model_create(int n_atoms, const float* xyz, const int* types, int n_seg, const int* seg_parent, const int* seg_begin,
                        const int* seg_end, const int* axis_root, int n_pairs, const int* pair_a, const int* pair_b, int n_rec,
                        const float* rec_xyz, const int* rec_types) {
  RefModel* R = new RefModel;
  int rc = guarded([&] {
    model& m = R->m;
    m.m_num_movable_atoms = n_atoms;
    m.minus_forces = std::vector<vec>(n_atoms, vec(0, 0, 0));
    R->n_seg = n_seg;
    R->seg_origin.resize(n_seg);
    for (int s = 0; s < n_seg; s++) { const float* o = xyz + 3 * seg_begin[s]; R->seg_origin[s] = vec(o[0], o[1], o[2]); }
    std::vector<int> seg_of(n_atoms, 0);
    for (int s = 0; s < n_seg; s++) for (int i = seg_begin[s]; i < seg_end[s]; i++) seg_of[i] = s;
    for (int i = 0; i < n_atoms; i++) {
      vec c(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
      m.coords.push_back(c);
      m.atoms.push_back(atom());
      m.atoms[i].sm = (smt)types[i];
      m.atoms[i].charge = 0;
      m.atoms[i].coords = c - R->seg_origin[seg_of[i]];  // atom coordinates live in their segment's frame (atom_frame::set_coords)
    }
    for (int i = 0; i < n_rec; i++) {
      m.grid_atoms.push_back(atom());
      m.grid_atoms[i].sm = (smt)rec_types[i];
      m.grid_atoms[i].charge = 0;
      m.grid_atoms[i].coords = vec(rec_xyz[3 * i], rec_xyz[3 * i + 1], rec_xyz[3 * i + 2]);
    }
    rigid_body root(R->seg_origin[0], (sz)seg_begin[0], (sz)seg_end[0]);
    flexible_body flex(root);
    m.ligands.push_back(ligand(flex, (unsigned)(n_seg - 1)));
    ligand& lig = m.ligands[0];
    // nodes in pre-order: a node's children vector only grows while the node is on the current path, so the pointers of the
    // path stay valid
    std::vector<branch*> node(n_seg, nullptr);
    R->rel_origin.assign(n_seg, vec(0, 0, 0));
    R->rel_axis.assign(n_seg, vec(0, 0, 0));
    for (int s = 1; s < n_seg; s++) {
      const int p = seg_parent[s];
      const float* ar = xyz + 3 * axis_root[s];
      if (p == 0) {
        segment seg(R->seg_origin[s], (sz)seg_begin[s], (sz)seg_end[s], vec(ar[0], ar[1], ar[2]), lig.node);
        lig.children.push_back(branch(seg));
        node[s] = &lig.children.back();
      } else {
        if (!node[p]) throw std::runtime_error("segments are not in DFS pre-order");
        segment seg(R->seg_origin[s], (sz)seg_begin[s], (sz)seg_end[s], vec(ar[0], ar[1], ar[2]), node[p]->node);
        node[p]->children.push_back(branch(seg));
        node[s] = &node[p]->children.back();
      }
      R->rel_origin[s] = node[s]->node.*peek(SegRelOrigin());
      R->rel_axis[s] = node[s]->node.*peek(SegRelAxis());
    }
    lig.set_range();
    for (int k = 0; k < n_pairs; k++)
      lig.pairs.push_back(interacting_pair((smt)types[pair_a[k]], (smt)types[pair_b[k]], (sz)pair_a[k], (sz)pair_b[k]));
    R->gcache.reset(new szv_grid_cache(m, 64.f));
  });
  if (rc) { delete R; return nullptr; }
  return R;
}
void gref_model_destroy(void* p) { delete (RefModel*)p; }
// the `model` inside the handle, for oracle/ref_cnn_driver.cpp (the reference's CNN scorer takes a model&)
void* gref_model_ptr(void* p) { return &((RefModel*)p)->m; }
// what the reference's constructors computed: atom coordinates in their segment frames, segment origin relative to the parent
// origin, unit rotation axis
void gref_model_export(void* p, float* local_xyz, float* rel_origin, float* rel_axis) {
  RefModel* R = (RefModel*)p;
  for (sz i = 0; i < R->m.atoms.size(); i++) for (int k = 0; k < 3; k++) local_xyz[3 * i + k] = R->m.atoms[i].coords[k];
  for (int s = 0; s < R->n_seg; s++) for (int k = 0; k < 3; k++) { rel_origin[3 * s + k] = R->rel_origin[s][k]; rel_axis[3 * s + k] = R->rel_axis[s][k]; }
}
// model::set(conf) -> coordinates (lib/model.cpp:968-975 -> tree.h set_conf)
int gref_model_set(void* p, const float* x, float* out_xyz) {
  RefModel* R = (RefModel*)p;
  return guarded([&] {
    R->m.set(make_conf(R->m, x));
    for (sz i = 0; i < R->m.coords.size(); i++) for (int k = 0; k < 3; k++) out_xyz[3 * i + k] = R->m.coords[i][k];
  });
}
// overwrite the current coordinates (for evaluations on given positions)
void gref_model_put_coords(void* p, const float* xyz) {
  RefModel* R = (RefModel*)p;
  for (sz i = 0; i < R->m.coords.size(); i++) R->m.coords[i] = vec(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
}
void gref_model_get_coords(void* p, float* xyz) {
  RefModel* R = (RefModel*)p;
  for (sz i = 0; i < R->m.coords.size(); i++) for (int k = 0; k < 3; k++) xyz[3 * i + k] = R->m.coords[i][k];
}
float gref_gyration_radius(void* p) { return ((RefModel*)p)->m.gyration_radius(0); }
// model::clear_minus_forces + model::add_minus_forces (lib/model.cu:232-259) with a per-movable-atom force list -> what each atom received
void gref_add_minus_forces(void* p, const float* forces, int n, float* out) {
  model& m = ((RefModel*)p)->m;
  std::vector<gfloat3> f;
  for (int i = 0; i < n; i++) f.push_back(gfloat3(forces[3 * i], forces[3 * i + 1], forces[3 * i + 2]));
  m.clear_minus_forces();
  m.add_minus_forces(f);
  for (sz i = 0; i < m.minus_forces.size(); i++) for (int k = 0; k < 3; k++) out[3 * i + k] = m.minus_forces[i][k];
}
// model::movable_atoms_box (lib/model.cpp:751-776) of the coordinates the model holds -> begin, end, n
void gref_movable_atoms_box(void* p, float add, float granularity, float* begin, float* end, int* n) {
  const grid_dims gd = ((RefModel*)p)->m.movable_atoms_box(add, granularity);
  for (int i = 0; i < 3; i++) { begin[i] = gd[i].begin; end[i] = gd[i].end; n[i] = (int)gd[i].n; }
}
// heterotree::derivative on given coordinates and forces (tree.h:374-382) -> change [6 + T]
int gref_tree_derivative(void* p, const float* forces, float* out_change) {
  RefModel* R = (RefModel*)p;
  return guarded([&] {
    model& m = R->m;
    for (sz i = 0; i < m.minus_forces.size(); i++) m.minus_forces[i] = vec(forces[3 * i], forces[3 * i + 1], forces[3 * i + 2]);
    change g(m.get_size(), false);
    m.ligands.derivative(m.coords, m.minus_forces, g.ligands);
    read_change(g, out_change);
  });
}

// integration/docking_b200.h's B200Ligand run on this reference model: -> counts {n_atoms, n_segments, n_pairs, n_heavy}; arrays may be null
int gref_adapter_topology(void* p, int* counts, float* local_xyz, int* type, int* parent, int* begin, int* end, float* rel_origin,
                          float* rel_axis, int* pair_a, int* pair_b, float* gyration_radius) {
  RefModel* R = (RefModel*)p;
  return guarded([&] {
    b200::B200Ligand L(R->m);
    counts[0] = L.topo.n_atoms; counts[1] = L.topo.n_segments; counts[2] = L.topo.n_pairs; counts[3] = L.n_heavy;
    if (!local_xyz) return;
    std::copy(L.local_xyz.begin(), L.local_xyz.end(), local_xyz);
    std::copy(L.type.begin(), L.type.end(), type);
    std::copy(L.parent.begin(), L.parent.end(), parent);
    std::copy(L.begin.begin(), L.begin.end(), begin);
    std::copy(L.end.begin(), L.end.end(), end);
    std::copy(L.rel_origin.begin(), L.rel_origin.end(), rel_origin);
    std::copy(L.rel_axis.begin(), L.rel_axis.end(), rel_axis);
    std::copy(L.pair_a.begin(), L.pair_a.end(), pair_a);
    std::copy(L.pair_b.begin(), L.pair_b.end(), pair_b);
    *gyration_radius = L.topo.gyration_radius;
  });
}

// ---- grids -----------------------------------------------------------------------------------------------------------------
// cache: ctor + populate for the movable atom types of the model (lib/cache.cpp:104-184)
// an igrid made elsewhere (oracle/ref_adapters_driver.cpp: the integration adapter cache_b200) as a grid handle of this driver; takes it over
void* gref_grid_wrap(void* ig) { RefGrid* G = new RefGrid; G->ig.reset((igrid*)ig); G->kind = 3; return G; }
void* gref_cache_create(void* sf, int kind, void* mp, const float* begin, const float* end, const int* n, float slope) {
  RefSF* S = (RefSF*)sf; RefModel* R = (RefModel*)mp;
  RefGrid* G = new RefGrid;
  int rc = guarded([&] {
    cache* c = new cache("scoring_function_version001", make_dims(begin, end, n), slope);
    G->ig.reset(c);
    std::vector<smt> needed;
    R->m.get_movable_atom_types(needed);
    grid user_grid;
    c->populate(R->m, *S->prec[kind], needed, user_grid, false);
  });
  if (rc) { delete G; return nullptr; }
  return G;
}
// one populated grid, x fastest, (n0+1)(n1+1)(n2+1) floats; returns 0 when type t has no grid
int gref_cache_grid(void* gp, int t, float* out) {
  cache* c = dynamic_cast<cache*>(((RefGrid*)gp)->ig.get());
  if (!c) return 0;
  const grid& g = (c->*peek(CacheGrids()))[t];
  if (!g.initialized()) return 0;
  const array3d<fl>& d = g.*peek(GridData());
  for (sz z = 0; z < d.dim2(); z++) for (sz y = 0; y < d.dim1(); y++) for (sz x = 0; x < d.dim0(); x++)
    out[x + d.dim0() * (y + d.dim1() * z)] = d(x, y, z);
  return 1;
}
void* gref_noncache_create(void* sf, int kind, void* mp, const float* begin, const float* end, const int* n, float slope) {
  RefSF* S = (RefSF*)sf; RefModel* R = (RefModel*)mp;
  RefGrid* G = new RefGrid;
  G->kind = 1;
  int rc = guarded([&] { G->ig.reset(new non_cache(*R->gcache, make_dims(begin, end, n), S->prec[kind].get(), slope)); });
  if (rc) { delete G; return nullptr; }
  return G;
}
void gref_noncache_set_slope(void* gp, float slope) { dynamic_cast<non_cache*>(((RefGrid*)gp)->ig.get())->setSlope(slope); }
int gref_noncache_within(void* gp, void* mp, float margin) {
  return dynamic_cast<non_cache*>(((RefGrid*)gp)->ig.get())->within(((RefModel*)mp)->m, margin) ? 1 : 0;
}
void* gref_naive_create(void* sf, int kind) {
  RefGrid* G = new RefGrid;
  G->kind = 2;
  G->ig.reset(new naive_non_cache(((RefSF*)sf)->prec[kind].get()));
  return G;
}
void gref_grid_destroy(void* gp) { delete (RefGrid*)gp; }
// igrid::eval / eval_deriv on the model's CURRENT coordinates
int gref_ig_eval(void* gp, void* mp, float v, float* e) {
  return guarded([&] { *e = ((RefGrid*)gp)->ig->eval(((RefModel*)mp)->m, v); });
}
int gref_ig_eval_deriv(void* gp, void* mp, float v, float* e, float* minus_forces) {
  RefModel* R = (RefModel*)mp;
  return guarded([&] {
    grid user_grid;
    *e = ((RefGrid*)gp)->ig->eval_deriv(R->m, v, user_grid);
    for (sz i = 0; i < R->m.minus_forces.size(); i++) for (int k = 0; k < 3; k++) minus_forces[3 * i + k] = R->m.minus_forces[i][k];
  });
}

// ---- model-level evaluations ---------------------------------------------------------------------------------------------
// model::eval_deriv (lib/model.cu:202-225): set(conf), grid term, intramolecular pairs, tree derivative
int gref_model_eval_deriv(void* mp, void* sf, int kind, void* gp, const float* v3, const float* x, float* e, float* out_change) {
  RefModel* R = (RefModel*)mp; RefSF* S = (RefSF*)sf;
  return guarded([&] {
    conf c = make_conf(R->m, x);
    change g(R->m.get_size(), false);
    grid user_grid;
    *e = R->m.eval_deriv(*S->prec[kind], *((RefGrid*)gp)->ig, vec(v3[0], v3[1], v3[2]), c, g, user_grid);
    read_change(g, out_change);
  });
}
// model::eval_intramolecular + model::eval_adjusted (lib/model.cu:352-406) as main.cpp:219-232 uses them for "Affinity":
// e = eval_adjusted(wt, exact_prec, naive_non_cache, v, conf, intramolecular_energy)
int gref_model_affinity(void* mp, void* sf, const float* v3, const float* x, float* intramolecular, float* affinity) {
  RefModel* R = (RefModel*)mp; RefSF* S = (RefSF*)sf;
  return guarded([&] {
    conf c = make_conf(R->m, x);
    vec v(v3[0], v3[1], v3[2]);
    naive_non_cache nnc(S->prec[2].get());
    grid user_grid;
    *intramolecular = R->m.eval_intramolecular(*S->prec[2], v, c);
    *affinity = R->m.eval_adjusted(*S->wt, *S->prec[2], nnc, v, c, *intramolecular, user_grid);
  });
}
// weighted_terms::conf_independent with a given num_tors: num_tors_div::eval (lib/everything.h:795-809)
int gref_num_tors_div(void* sf, float e, float num_tors, float* out) {
  RefSF* S = (RefSF*)sf;
  return guarded([&] {
    conf_independent_inputs in;
    in.num_tors = num_tors;
    flv::const_iterator it = ((*S->wt).*peek(WtWeights())).begin() + (*S->wt).*peek(WtConfIndepStart());
    *out = S->t.eval_conf_independent(in, e, it);
  });
}
// quasi_newton::operator() (lib/quasi_newton.cpp:49-83 -> bfgs.h:358-502, fast line search) from conf x; x is updated
int gref_bfgs(void* mp, void* sf, int kind, void* gp, float* x, int maxiters, const float* v3, float* e, float* out_change, int accurate,
              int early_term) {
  RefModel* R = (RefModel*)mp; RefSF* S = (RefSF*)sf;
  return guarded([&] {
    minimization_params mp_;
    mp_.maxiters = (unsigned)maxiters;
    mp_.type = accurate ? minimization_params::BFGSAccurateLineSearch : minimization_params::BFGSFastLineSearch;
    mp_.early_term = early_term != 0;
    quasi_newton qn(mp_);
    output_type out(make_conf(R->m, x), 0);
    change g(R->m.get_size(), false);
    grid user_grid;
    qn(R->m, *S->prec[kind], *((RefGrid*)gp)->ig, out, g, vec(v3[0], v3[1], v3[2]), user_grid);
    *e = out.e;
    read_conf(out.c, x);
    read_change(g, out_change);
  });
}
// conf::randomize (lib/conf.h:441-447) on the shim's generator seeded with `seed`: -> conf, generator state afterwards
int gref_random_conf(void* mp, unsigned seed, const float* c1, const float* c2, float* x, unsigned* state_after) {
  RefModel* R = (RefModel*)mp;
  return guarded([&] {
    rng gen(seed);
    conf c(R->m.get_size(), false);
    c.randomize(vec(c1[0], c1[1], c1[2]), vec(c2[0], c2[1], c2[2]), gen);
    read_conf(c, x);
    *state_after = gen.state();
  });
}
// monte_carlo::operator() (lib/monte_carlo.cpp:99-148) -> the sorted output container: energies, confs [n][7+T]
int gref_mc(void* mp, void* sf, int kind, void* gp, const float* c1, const float* c2, unsigned seed, int num_steps, int maxiters,
            int num_saved_mins, float temperature, float amplitude, float min_rmsd, const float* hunt_cap, const float* state_conf,
            int max_out, float* out_e, float* out_conf, int* n_out) {
  RefModel* R = (RefModel*)mp; RefSF* S = (RefSF*)sf;
  return guarded([&] {
    monte_carlo mc;
    mc.num_steps = (unsigned)num_steps;
    mc.temperature = temperature;
    mc.hunt_cap = vec(hunt_cap[0], hunt_cap[1], hunt_cap[2]);
    mc.min_rmsd = min_rmsd;
    mc.num_saved_mins = (sz)num_saved_mins;
    mc.mutation_amplitude = amplitude;
    mc.ssd_par.minparm.maxiters = (unsigned)maxiters;
    rng gen(seed);
    output_container out;
    grid user_grid;
    igrid& ig = *((RefGrid*)gp)->ig;
    R->m.set(make_conf(R->m, state_conf));  // the conformation the model object holds when the chain starts (mutate_conf reads it)
    mc(R->m, out, *S->prec[kind], ig, vec(c1[0], c1[1], c1[2]), vec(c2[0], c2[1], c2[2]), nullptr, gen, user_grid, ig);
    const int T = R->n_seg - 1;
    *n_out = (int)std::min<sz>(out.size(), (sz)max_out);
    for (int i = 0; i < *n_out; i++) { out_e[i] = out[i].e; read_conf(out[i].c, out_conf + (size_t)i * (7 + T)); }
  });
}
// parallel_mc::operator() (lib/parallel_mc.cpp:183-214): task seeds random_int(0, 1000000, generator), one monte_carlo chain per task on
// the reference's own thread pool (lib/parallel.h), merge_output_containers (min_rmsd forced to 2) and the final sort
int gref_parallel_mc(void* mp, void* sf, int kind, void* gp, const float* c1, const float* c2, unsigned seed, int num_tasks, int num_threads,
                     int num_steps, int maxiters, int num_saved_mins, float temperature, float amplitude, float min_rmsd, const float* hunt_cap,
                     const float* state_conf, const float* box_begin, const float* box_end, const int* box_n, int max_out, float* out_e,
                     float* out_conf, int* n_out) {
  RefModel* R = (RefModel*)mp; RefSF* S = (RefSF*)sf;
  return guarded([&] {
    parallel_mc par;
    par.num_tasks = (sz)num_tasks; par.num_threads = (sz)num_threads; par.display_progress = false;
    par.mc.num_steps = (unsigned)num_steps; par.mc.temperature = temperature;
    par.mc.hunt_cap = vec(hunt_cap[0], hunt_cap[1], hunt_cap[2]);
    par.mc.min_rmsd = min_rmsd; par.mc.num_saved_mins = (sz)num_saved_mins; par.mc.mutation_amplitude = amplitude;
    par.mc.ssd_par.minparm.maxiters = (unsigned)maxiters;
    rng gen(seed);
    output_container out;
    grid user_grid;
    igrid& ig = *((RefGrid*)gp)->ig;
    R->m.set(make_conf(R->m, state_conf));
    non_cache nc(*R->gcache, make_dims(box_begin, box_end, box_n), S->prec[kind].get(), 1e3);   // only passed through (no CNN)
    par(R->m, out, *S->prec[kind], ig, vec(c1[0], c1[1], c1[2]), vec(c2[0], c2[1], c2[2]), gen, user_grid, nc);
    const int T = R->n_seg - 1;
    *n_out = (int)std::min<sz>(out.size(), (sz)max_out);
    for (int i = 0; i < *n_out; i++) { out_e[i] = out[i].e; read_conf(out[i].c, out_conf + (size_t)i * (7 + T)); }
  });
}
// add_to_output_container (lib/coords.cpp:43-56) replayed on a sequence of (energy, coords) entries -> kept energies in order
int gref_container_replay(int n_items, int n_coords, const float* e, const float* coords, float min_rmsd, int max_size, float* out_e,
                          int* n_out) {
  return guarded([&] {
    output_container out;
    conf dummy;
    for (int i = 0; i < n_items; i++) {
      output_type t(dummy, e[i]);
      for (int a = 0; a < n_coords; a++) { const float* q = coords + ((size_t)i * n_coords + a) * 3; t.coords.push_back(vec(q[0], q[1], q[2])); }
      add_to_output_container(out, t, min_rmsd, (sz)max_size);
    }
    *n_out = (int)out.size();
    for (int i = 0; i < *n_out; i++) out_e[i] = out[i].e;
  });
}

}  // extern "C"
