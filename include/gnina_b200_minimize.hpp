// gnina_b200_minimize.hpp -- C++ host side of BASELINE config 5 / `--cnn_scoring refinement`: quasi-Newton minimisation of MANY poses
// in lock step (header-only, C++17).  gnina's host code is C++; gnina_b200/minimize.py is the same algorithm in numpy for the tests.
//
// The reference minimises one pose at a time: quasi_newton (lib/quasi_newton.cpp:49-83 -> lib/bfgs.h:358-502) calls
// non_cache_cnn::eval_deriv once per function evaluation = one CNN forward + backward for ONE pose (main/main.cpp:264-268).  Here every
// unfinished pose has exactly one pending function evaluation per round (its first evaluation or a line-search trial); the pending
// conformations of ALL poses go to the energy functor in one batch (-> one gb_cnn_score_grad call), then every pose advances by bfgs.h's
// rules: fast_line_search (:73-91) or accurate_line_search (:107-180, with the reference's float / double mix), bfgs_update (:52-66, upper
// triangle only), --minimize_early_term (:455-462), the restore of x_orig (:494-498).  Per pose the sequence of evaluations, and with it
// the result, is the reference's: checked bit for bit against the reference's own quasi_newton + non_cache_cnn compiled in oracle/_ref
// (oracle/ref_driver.cpp gref_lockstep_*, tests/test_oracle_vs_reference_build.py).
//
// conf = position[3], orientation quaternion[4] (a,b,c,d), torsions[T]; change = force[3], torque[3], torsion derivatives[T].
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <stdexcept>
#include <vector>
#include "gnina_b200.h"

namespace gb {

// float32 -> float32 transcendentals of the kinematics: correctly rounded (evaluated in double, rounded once), like the device kernels.
// Tests that compare with the reference compiled on the same host swap in sinf / cosf / acosf (what that build executes).
struct Transcendentals {
  float (*sin)(float) = [](float x) { return (float)std::sin((double)x); };
  float (*cos)(float) = [](float x) { return (float)std::cos((double)x); };
  float (*acos)(float) = [](float x) { return (float)std::acos((double)x); };
};
inline Transcendentals& transcendentals() { static Transcendentals t; return t; }

namespace mdetail {
constexpr float kPi = 3.1415926535897931f, kEps = 1.1920929e-07f;
inline void normalize_angle(float& x) {                       // lib/common.h normalize_angle
  if (x > 3 * kPi) { const float n = (x - kPi) / (2 * kPi); x -= 2 * kPi * std::ceil(n); normalize_angle(x); }
  else if (x < -3 * kPi) { const float n = (-x - kPi) / (2 * kPi); x += 2 * kPi * std::ceil(n); normalize_angle(x); }
  else if (x > kPi) x -= 2 * kPi;
  else if (x < -kPi) x += 2 * kPi;
}
inline void angle_to_q(const float* axis, float angle, float* q) {          // quaternion.h:284-291
  normalize_angle(angle);
  const float c = transcendentals().cos(angle / 2), s = transcendentals().sin(angle / 2);
  q[0] = c; q[1] = s * axis[0]; q[2] = s * axis[1]; q[3] = s * axis[2];
}
inline void qmul(const float* l, const float* r, float* o) {
  const float a = l[0], b = l[1], c = l[2], d = l[3];
  o[0] = +a * r[0] - b * r[1] - c * r[2] - d * r[3];
  o[1] = +a * r[1] + b * r[0] + c * r[3] - d * r[2];
  o[2] = +a * r[2] - b * r[3] + c * r[0] + d * r[1];
  o[3] = +a * r[3] + b * r[2] - c * r[1] + d * r[0];
}
inline void qnorm_approx(float* q) {                                         // quaternion.h:243-257
  const float s = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  if (std::fabs(s - 1) < 1e-6f) return;
  const float a = std::sqrt(s);
  for (int i = 0; i < 4; i++) q[i] *= 1 / a;
}
inline void q_to_r3(const float* q, float* m) {                              // quaternion.h:327-364
  const float a = q[0], b = q[1], c = q[2], d = q[3];
  const float aa = a * a, ab = a * b, ac = a * c, ad = a * d, bb = b * b, bc = b * c, bd = b * d, cc = c * c, cd = c * d, dd = d * d;
  m[0] = (aa + bb - cc - dd); m[1] = 2 * (-ad + bc); m[2] = 2 * (ac + bd);
  m[3] = 2 * (ad + bc); m[4] = (aa - bb + cc - dd); m[5] = 2 * (-ab + cd);
  m[6] = 2 * (-ac + bd); m[7] = 2 * (ab + cd); m[8] = (aa - bb - cc + dd);
}
inline void mv(const float* m, const float* v, float* o) {
  o[0] = m[0] * v[0] + m[1] * v[1] + m[2] * v[2];
  o[1] = m[3] * v[0] + m[4] * v[1] + m[5] * v[2];
  o[2] = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
}
inline void cross(const float* a, const float* b, float* o) {
  o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
inline void quaternion_increment(float* q, const float* rot) {               // quaternion.cu:32-43,96-100
  const float angle = std::sqrt(rot[0] * rot[0] + rot[1] * rot[1] + rot[2] * rot[2]);
  float r[4] = {1, 0, 0, 0};
  if (angle > kEps) {
    const float axis[3] = {(1 / angle) * rot[0], (1 / angle) * rot[1], (1 / angle) * rot[2]};
    angle_to_q(axis, angle, r);
  }
  float o[4];
  qmul(r, q, o);
  std::memcpy(q, o, sizeof o);
  qnorm_approx(q);
}
inline int tri(int i, int j) { return i <= j ? i + j * (j + 1) / 2 : j + i * (i + 1) / 2; }   // triangular_matrix_index.h
}  // namespace mdetail

// heterotree<rigid_body> kinematics (lib/tree.h:218-233,300-310,361-382) over a copy of a gb_ligand_topology
class LigandTree {
 public:
  int n_atoms = 0, n_seg = 0;
  std::vector<float> local, rel_origin, rel_axis;
  std::vector<int32_t> type, parent, begin, end;

  explicit LigandTree(const gb_ligand_topology& t)
      : n_atoms(t.n_atoms), n_seg(t.n_segments), local(t.local_xyz, t.local_xyz + 3 * t.n_atoms),
        rel_origin(t.seg_rel_origin, t.seg_rel_origin + 3 * t.n_segments), rel_axis(t.seg_rel_axis, t.seg_rel_axis + 3 * t.n_segments),
        type(t.smina_type, t.smina_type + t.n_atoms), parent(t.seg_parent, t.seg_parent + t.n_segments),
        begin(t.seg_atom_begin, t.seg_atom_begin + t.n_segments), end(t.seg_atom_end, t.seg_atom_end + t.n_segments) {}
  int n_tors() const { return n_seg - 1; }
  int conf_floats() const { return 7 + n_tors(); }
  int change_floats() const { return 6 + n_tors(); }
  bool heavy(int i) const { return type[i] >= 2; }

  // model::set: coords [n_atoms][3], segment origins / axes [n_seg][3]
  void set_conf(const float* x, float* coords, float* so, float* sa) const {
    using namespace mdetail;
    std::vector<float> q(4 * (size_t)n_seg), M(9 * (size_t)n_seg);
    for (int s = 0; s < n_seg; s++) {
      float* o = so + 3 * s;
      if (s == 0) {
        std::memcpy(o, x, 12); std::memcpy(q.data(), x + 3, 16);
        sa[0] = sa[1] = sa[2] = 0;
      } else {
        const int p = parent[s];
        float t[3];
        mv(&M[9 * p], &rel_origin[3 * s], t);
        for (int k = 0; k < 3; k++) o[k] = so[3 * p + k] + t[k];
        mv(&M[9 * p], &rel_axis[3 * s], sa + 3 * s);
        float aq[4];
        angle_to_q(sa + 3 * s, x[7 + s - 1], aq);
        qmul(aq, &q[4 * p], &q[4 * s]);
        qnorm_approx(&q[4 * s]);
      }
      q_to_r3(&q[4 * s], &M[9 * s]);
      for (int i = begin[s]; i < end[s]; i++) {
        float t[3];
        mv(&M[9 * s], &local[3 * i], t);
        for (int k = 0; k < 3; k++) coords[3 * i + k] = o[k] + t[k];
      }
    }
  }
  // heterotree::derivative: minus_forces -> change
  void derivative(const float* coords, const float* forces, const float* so, const float* sa, float* change) const {
    std::vector<float> ft(6 * (size_t)n_seg, 0.f);
    for (int s = n_seg - 1; s >= 0; s--) {
      float* f = &ft[6 * s];
      for (int i = begin[s]; i < end[s]; i++) {              // sum_force_and_torque
        const float r[3] = {coords[3 * i] - so[3 * s], coords[3 * i + 1] - so[3 * s + 1], coords[3 * i + 2] - so[3 * s + 2]};
        float c[3];
        mdetail::cross(r, forces + 3 * i, c);
        for (int k = 0; k < 3; k++) { f[k] += forces[3 * i + k]; f[3 + k] += c[k]; }
      }
      for (int ch = s + 1; ch < n_seg; ch++) {               // branches_derivative, children in ascending order
        if (parent[ch] != s) continue;
        const float* cf = &ft[6 * ch];
        float r[3], c[3];
        for (int k = 0; k < 3; k++) { f[k] += cf[k]; r[k] = so[3 * ch + k] - so[3 * s + k]; }
        mdetail::cross(r, cf, c);
        for (int k = 0; k < 3; k++) f[3 + k] += c[k] + cf[3 + k];
      }
      if (s == 0) std::memcpy(change, f, 24);
      else change[6 + s - 1] = f[3] * sa[3 * s] + f[4] * sa[3 * s + 1] + f[5] * sa[3 * s + 2];
    }
  }
  // conf::increment (lib/conf.h:54-59,113-118,385-393)
  void increment(float* x, const float* p, float f) const {
    for (int k = 0; k < 3; k++) x[k] += f * p[k];
    const float rot[3] = {f * p[3], f * p[4], f * p[5]};
    mdetail::quaternion_increment(x + 3, rot);
    for (int t = 0; t < n_tors(); t++) {
      float a = f * p[6 + t];
      mdetail::normalize_angle(a);
      x[7 + t] += a;
      mdetail::normalize_angle(x[7 + t]);
    }
  }
  // compute_lambdamin (lib/bfgs.h:93-102) over conf::operator()(i) (lib/conf.h:459-473, quaternion_to_angle quaternion.cu:46-62)
  float lambdamin(const float* x, const float* p) const {
    using namespace mdetail;
    float ang[3] = {0, 0, 0};
    const float c = x[3];
    if (c > -1 && c < 1) {
      float angle = 2 * transcendentals().acos(c);
      if (angle > kPi) angle -= 2 * kPi;
      const float s = transcendentals().sin(angle / 2);
      if (!(std::fabs(s) < kEps)) { const float f = angle / s; ang[0] = x[4] * f; ang[1] = x[5] * f; ang[2] = x[6] * f; }
    }
    float test = 0;
    for (int i = 0; i < change_floats(); i++) {
      const float xi = i < 3 ? x[i] : (i < 6 ? ang[i - 3] : x[7 + i - 6]);
      const float ax = std::fabs(xi);
      const float temp = std::fabs(p[i]) / ((ax < 1.0f) ? 1.0f : ax);
      if (temp > test) test = temp;
    }
    return test;
  }
};

struct MinimizeParams {        // minimization_params (lib/common.h:50-60); --minimize: accurate, maxiters 10000 (main.cpp:1157-1160)
  int maxiters = 10000;
  bool accurate_line_search = true, early_term = false;
};

// Energy: void(const float* coords /* [k][n_atoms][3] */, const int* pose /* [k] rows of confs */, int k, float* e /* [k] */,
//              float* minus_forces /* [k][n_atoms][3] */) -- called once per round with the pending conformations of all unfinished poses.
// confs [n][7+T] in/out.  Returns the final energies; evals / rounds (nullable) report the work.
template <class Energy>
std::vector<float> minimize_poses(const LigandTree& tree, Energy&& energy, float* confs, int n, const MinimizeParams& mp,
                                  std::vector<int>* evals = nullptr, int* rounds = nullptr) {
  using namespace mdetail;
  const int nx = tree.conf_floats(), m = tree.change_floats(), na = tree.n_atoms, ns = tree.n_seg;
  struct Pose {
    std::vector<float> x, g, x_new, g_new, x_orig, g_orig, p, h;
    float f0 = 0, f_orig = 0, alpha = 1, pg = 0, alpha2 = 0, f2 = 0, alamin = 0;
    int step = 0, trial = 0, evals = 0;
    bool active = true;
  };
  std::vector<Pose> P(n);
  std::vector<float> coords, forces, so, sa, e;
  std::vector<int> who;
  // one batched evaluation of the conformations `pick(i)` of the poses in `who` -> energy + change per pose
  auto evaluate = [&](auto pick, auto store) {
    const int k = (int)who.size();
    coords.resize((size_t)k * na * 3); forces.assign((size_t)k * na * 3, 0.f);
    so.resize((size_t)k * ns * 3); sa.resize((size_t)k * ns * 3); e.resize(k);
    for (int j = 0; j < k; j++) tree.set_conf(pick(P[who[j]]), &coords[(size_t)j * na * 3], &so[(size_t)j * ns * 3], &sa[(size_t)j * ns * 3]);
    energy(coords.data(), who.data(), k, e.data(), forces.data());
    for (int j = 0; j < k; j++) {
      Pose& q = P[who[j]];
      tree.derivative(&coords[(size_t)j * na * 3], &forces[(size_t)j * na * 3], &so[(size_t)j * ns * 3], &sa[(size_t)j * ns * 3], store(q));
      q.evals++;
    }
  };
  auto dot = [&](const std::vector<float>& a, const std::vector<float>& b) { float s = 0; for (int i = 0; i < m; i++) s += a[i] * b[i]; return s; };
  auto start_iteration = [&](Pose& q) {                      // p = -H g, slope, first trial step
    for (int i = 0; i < m; i++) { float s = 0; for (int j = 0; j < m; j++) s += q.h[tri(i, j)] * q.g[j]; q.p[i] = -s; }
    q.pg = dot(q.p, q.g);
    q.alpha = 1; q.trial = 0; q.alpha2 = 0; q.f2 = 0;
    if (mp.accurate_line_search) {
      if (q.pg >= 0) { q.active = false; return; }           // not a descent direction: the search returns 0 and bfgs gives up
      q.alamin = kEps / tree.lambdamin(q.x.data(), q.p.data());
    }
  };
  for (int i = 0; i < n; i++) {
    Pose& q = P[i];
    q.x.assign(confs + (size_t)i * nx, confs + (size_t)(i + 1) * nx);
    q.g.assign(m, 0.f); q.g_new.assign(m, 0.f); q.p.assign(m, 0.f); q.x_new = q.x;
    q.h.assign((size_t)m * (m + 1) / 2, 0.f);
    for (int d = 0; d < m; d++) q.h[tri(d, d)] = 1;
    who.push_back(i);
  }
  evaluate([](Pose& q) { return q.x.data(); }, [](Pose& q) { return q.g.data(); });
  for (int i = 0; i < n; i++) {
    Pose& q = P[i];
    q.f0 = q.f_orig = e[i]; q.x_orig = q.x; q.g_orig = q.g;
    q.active = mp.maxiters > 0;
    if (q.active) start_iteration(q);
  }
  int n_rounds = 0;
  for (;;) {
    who.clear();
    for (int i = 0; i < n; i++) if (P[i].active) who.push_back(i);
    if (who.empty()) break;
    n_rounds++;
    for (int i : who) { Pose& q = P[i]; q.x_new = q.x; tree.increment(q.x_new.data(), q.p.data(), q.alpha); }
    evaluate([](Pose& q) { return q.x_new.data(); }, [](Pose& q) { return q.g_new.data(); });
    for (size_t j = 0; j < who.size(); j++) {
      Pose& q = P[who[j]];
      const float f1 = e[j];
      bool over;
      if (mp.accurate_line_search) {                         // accurate_line_search: fl = float, the literals 2.0 / 3.0 / .5 are double
        if (q.alpha < q.alamin || !std::isfinite(q.alpha)) { q.alpha = 0; over = true; }
        else if (f1 <= q.f0 + 1.0e-4f * q.alpha * q.pg) over = true;
        else {
          float tmplam;
          if (q.alpha == 1.0f) tmplam = (float)(-q.pg / (2.0 * (f1 - q.f0 - q.pg)));
          else {
            const float rhs1 = f1 - q.f0 - q.alpha * q.pg, rhs2 = q.f2 - q.f0 - q.alpha2 * q.pg;
            const float a = (rhs1 / (q.alpha * q.alpha) - rhs2 / (q.alpha2 * q.alpha2)) / (q.alpha - q.alpha2);
            const float b = (-q.alpha2 * rhs1 / (q.alpha * q.alpha) + q.alpha * rhs2 / (q.alpha2 * q.alpha2)) / (q.alpha - q.alpha2);
            if (a == 0.0f) tmplam = (float)(-q.pg / (2.0 * b));
            else {
              const float disc = (float)(b * b - 3.0 * a * q.pg);
              if (disc < 0) tmplam = (float)(0.5 * q.alpha);
              else if (b <= 0) tmplam = (float)((-b + std::sqrt(disc)) / (3.0 * a));
              else tmplam = -q.pg / (b + std::sqrt(disc));
            }
            if (tmplam > .5 * q.alpha) tmplam = (float)(.5 * q.alpha);
          }
          q.alpha2 = q.alpha; q.f2 = f1;
          const float tenth = 0.1f * q.alpha;
          q.alpha = (tmplam < tenth) ? tenth : tmplam;       // std::max(tmplam, 0.1 alpha)
          over = false;
        }
      } else {                                               // fast_line_search
        const bool accepted = f1 - q.f0 < 0.0001f * q.alpha * q.pg;
        if (!accepted) { q.alpha *= 0.5f; q.trial++; }
        over = accepted || q.trial >= 10;
      }
      if (!over) continue;
      if (q.alpha == 0) { q.active = false; continue; }
      std::vector<float> y(m);
      for (int i = 0; i < m; i++) y[i] = q.g_new[i] - q.g[i];
      const float prevf0 = q.f0;
      q.f0 = f1;
      q.x = q.x_new;
      if (mp.early_term && std::fabs((double)(prevf0 - q.f0)) < 1e-5) { q.active = false; continue; }   // before g is replaced
      q.g = q.g_new;
      if (!(dot(q.g, q.g) >= 1e-4f)) { q.active = false; continue; }
      const float yp = dot(y, q.p);
      if (q.step == 0) {
        const float yy = dot(y, y);
        if (std::fabs(yy) > kEps) for (int d = 0; d < m; d++) q.h[tri(d, d)] = q.alpha * yp / yy;
      }
      if (!(q.alpha * yp < kEps)) {                          // bfgs_update
        std::vector<float> mhy(m);
        for (int i = 0; i < m; i++) { float s = 0; for (int k = 0; k < m; k++) s += q.h[tri(i, k)] * y[k]; mhy[i] = -s; }
        const float yhy = -dot(y, mhy);
        const float r = 1 / (q.alpha * yp);
        for (int i = 0; i < m; i++)
          for (int k = i; k < m; k++)
            q.h[tri(i, k)] += q.alpha * r * (mhy[i] * q.p[k] + mhy[k] * q.p[i]) + q.alpha * q.alpha * (r * r * yhy + r) * q.p[i] * q.p[k];
      }
      if (++q.step >= mp.maxiters) { q.active = false; continue; }
      start_iteration(q);
    }
  }
  std::vector<float> out(n);
  if (evals) evals->assign(n, 0);
  for (int i = 0; i < n; i++) {
    Pose& q = P[i];
    if (!(q.f0 <= q.f_orig)) { q.f0 = q.f_orig; q.x = q.x_orig; q.g = q.g_orig; }   // succeeds for NaNs too
    out[i] = q.f0;
    std::memcpy(confs + (size_t)i * nx, q.x.data(), 4 * (size_t)nx);
    if (evals) (*evals)[i] = q.evals;
  }
  if (rounds) *rounds = n_rounds;
  return out;
}

// The energy of minimize_poses for `--minimize --cnn_scoring all` / `--cnn_scoring refinement`: non_cache_cnn::eval_deriv
// (lib/non_cache_cnn.cpp:79-169) for the k pending poses of one ligand -- ONE gb_cnn_score_grad call for the CNN loss and its atom
// gradients, then per heavy atom the out-of-box penalties of the search box and of the CNN's cubic grid, whose centre is fixed per pose
// before the minimisation (adjust_center :57-68 -> DLScorer::set_center_from_model, lib/dl_scorer.cpp:196-217: mean of the heavy atoms).
class CnnBatchEnergy {
  gb_cnn* h_;
  const LigandTree& tree_;
  float begin_[3], end_[3], slope_, half_;
  bool reference_force_routing_ = true;
  std::vector<float> centers_, loss_, grad_;
  std::vector<int32_t> types_, offs_;

  float bounds(const float* lo, const float* hi, const float* a, float* deriv) const {   // check_bounds_deriv, lib/non_cache.cpp:102-123
    float pen = 0;
    for (int j = 0; j < 3; j++) {
      if (a[j] < lo[j]) { deriv[j] = -1 * slope_; pen += std::fabs(a[j] - lo[j]); }
      else if (a[j] > hi[j]) { deriv[j] = 1 * slope_; pen += std::fabs(a[j] - hi[j]); }
    }
    return pen * slope_;
  }

 public:
  CnnBatchEnergy(gb_cnn* h, const LigandTree& tree, const float box_begin[3], const float box_end[3], float slope, float cnn_dimension)
      : h_(h), tree_(tree), slope_(slope), half_(cnn_dimension / 2.0f) {
    for (int j = 0; j < 3; j++) { begin_[j] = box_begin[j]; end_[j] = box_end[j]; }
  }
  void set_slope(float s) { slope_ = s; }               // refine_structure escalates it (main/main.cpp:145-154)
  // true (default): the minimiser sees the forces CNNTorchScorer::score leaves in the model -- getGradient's by-atom list consumed
  // COMPACTLY over the non-hydrogen atoms by model::add_minus_forces (lib/cnn_torch_scorer.cpp:209-227, lib/model.cu:247-259): the
  // j-th heavy atom receives entry j.  Identity for ligands without hydrogens; with hydrogens it is what makes a pose move exactly
  // as under gnina (checked against the reference's own code with real networks, tests/test_oracle_cnn_vs_reference_build.py).
  // false: the true per-atom gradient.
  void set_reference_force_routing(bool on) { reference_force_routing_ = on; }
  // the CNN box of every pose from its start conformation
  void set_centers(const float* confs, int n) {
    const int na = tree_.n_atoms, ns = tree_.n_seg, nx = tree_.conf_floats();
    std::vector<float> c(3 * (size_t)na), so(3 * (size_t)ns), sa(3 * (size_t)ns);
    centers_.assign(3 * (size_t)n, 0.f);
    for (int i = 0; i < n; i++) {
      tree_.set_conf(confs + (size_t)i * nx, c.data(), so.data(), sa.data());
      float cen[3] = {0, 0, 0}; unsigned cnt = 0;
      for (int a = 0; a < na; a++) if (tree_.heavy(a)) { for (int j = 0; j < 3; j++) cen[j] += c[3 * a + j]; cnt++; }
      for (int j = 0; j < 3; j++) centers_[3 * (size_t)i + j] = cen[j] / (float)cnt;
    }
  }
  void operator()(const float* coords, const int* pose, int k, float* e, float* forces) {
    const int na = tree_.n_atoms;
    types_.resize((size_t)k * na); offs_.resize(k + 1); loss_.resize(k); grad_.resize(3 * (size_t)k * na);
    for (int j = 0; j < k; j++) { std::memcpy(&types_[(size_t)j * na], tree_.type.data(), 4 * (size_t)na); offs_[j] = j * na; }
    offs_[k] = k * na;
    if (gb_cnn_score_grad(h_, coords, types_.data(), offs_.data(), k, nullptr, nullptr, nullptr, loss_.data(), nullptr, grad_.data(), nullptr) != GB_OK)
      throw std::runtime_error(gb_last_error());
    for (int j = 0; j < k; j++) {
      const float* cen = &centers_[3 * (size_t)pose[j]];
      const float lo[3] = {cen[0] - half_, cen[1] - half_, cen[2] - half_}, hi[3] = {cen[0] + half_, cen[1] + half_, cen[2] + half_};
      float en = loss_[j];
      int n_heavy_before = 0;
      for (int a = 0; a < na; a++) {
        float* f = forces + ((size_t)j * na + a) * 3;
        if (!tree_.heavy(a)) { f[0] = f[1] = f[2] = 0; continue; }
        const int src = reference_force_routing_ ? n_heavy_before : a;   // which entry of the by-atom gradient this atom receives
        n_heavy_before++;
        const float* x = coords + ((size_t)j * na + a) * 3;
        float d1[3] = {0, 0, 0}, d2[3] = {0, 0, 0};
        float pen = bounds(begin_, end_, x, d1);
        pen += bounds(lo, hi, x, d2);
        en += pen;
        for (int q = 0; q < 3; q++) f[q] = grad_[((size_t)j * na + src) * 3 + q] + (d1[q] + d2[q]);
      }
      e[j] = en;
    }
  }
};

}  // namespace gb
