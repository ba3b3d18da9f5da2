/*
 * oracle/vina_mc_ref.c — TEST INFRASTRUCTURE ONLY (CPU oracle of the docking inner loop, rows V6-V11).
 *
 * Scalar float restatement of
 *   V7  conformation -> coordinates   heterotree::set_conf, segment::set_conf     lib/tree.h:218-233,361-366
 *                                     quaternion_to_r3 / angle_to_quaternion        lib/quaternion.h:284-364
 *   V5  cache::eval_deriv (via vina_ref.c)                                         lib/cache.cpp:65-83
 *   V6  eval_interacting_pairs_deriv (precalculate_linear::eval_deriv + curl)      lib/model.cu:38-60
 *   V8  forces -> change              heterotree::derivative, branches_derivative   lib/tree.h:300-310,374-382
 *   V9  bfgs + fast_line_search + bfgs_update, conf::increment                      lib/bfgs.h:52-91,358-502, lib/conf.h:54-59,113-118
 *   V10 monte_carlo::operator(), mutate_conf, metropolis_accept, add_to_output_container
 *                                     lib/monte_carlo.cpp:38-47,99-148, lib/mutate.cpp:35-73, lib/coords.cpp:25-56
 *       non_cache::eval_deriv / within, refine_structure                           lib/non_cache.cpp:84-174, main/main.cpp:131-171
 * The reference draws from boost::mt19937 through Boost distributions (lib/random.cpp), whose sources are not in the
 * tree: its random STREAM cannot be reproduced, so this oracle (and the device code, identically) uses a small
 * counter-free generator (xorshift32).
 * PINNED against the reference's own code compiled here (oracle/_ref; oracle/Makefile.ref, ref_driver.cpp), which runs on the
 * same generator through the stand-in oracle/ref_shim/boost/random.hpp: with this host's sinf / cosf / expf (gvo_use_libm(1))
 * model::set, model::eval_deriv (cache and non_cache), quasi_newton after any number of iterations, refine_structure and WHOLE
 * Monte-Carlo chains (model-state mode of mc_impl) are BIT-IDENTICAL to the reference: tests/test_oracle_vs_reference_build.py
 * (live) and tests/test_oracle_vina_golden.py (tests/golden/vina_ref_kat.npz, generated from that build).  Not pinned: Boost's
 * random stream itself.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct gvo_prec gvo_prec; /* vina_ref.c */
void gvo_prec_eval_deriv(const gvo_prec *p, int t1, int t2, float r2, float *e, float *dor);
typedef struct gvo_splines gvo_splines;
void gvo_splines_eval_deriv(const gvo_splines *s, int t1, int t2, float r2, float *e, float *dor);
float gvo_grid_evaluate(const float *data, const float *begin, const float *end, const int32_t *n, const float *loc,
                        float slope, float v, float *deriv);

typedef struct {
  int n_atoms, n_seg, n_pairs;
  const float *local_xyz;     /* [n_atoms][3] in the frame of the atom's segment */
  const int32_t *type;        /* smina types */
  const int32_t *seg_parent;  /* [n_seg], -1 for the root; segments in DFS pre-order (= torsion order) */
  const int32_t *seg_begin, *seg_end;
  const float *seg_rel_origin, *seg_rel_axis; /* [n_seg][3] */
  const int32_t *pair_a, *pair_b;
} gvo_lig;

typedef struct {
  float *const *grids; /* [28] */
  const float *begin, *end;
  const int32_t *n;
  float slope;
  const gvo_prec *prec;
  const gvo_splines *splines; /* NULL: precalculate_linear ; else precalculate_splines for the pair terms */
  /* non_cache (lib/non_cache.cpp): when rec_xyz != NULL the intermolecular term is summed over the receptor atoms
   * directly (begin/end are then the search box gd of check_bounds_deriv) instead of read from the cache grids --
   * what refine_structure (main/main.cpp:131-171) minimises after the search */
  const float *rec_xyz;
  const int32_t *rec_type;
  int n_rec;
} gvo_field;

#define PI_F 3.14159265358979323846f
static const float kEps = 1.1920929e-07f, kMax = 3.402823466e+38f;
static int is_h(int t) { return t == 0 || t == 1; }

static void normalize_angle(float *x) { /* quaternion.h:261-281 */
  if (*x > 3 * PI_F) { float n = (*x - PI_F) / (2 * PI_F); *x -= 2 * PI_F * ceilf(n); normalize_angle(x); }
  else if (*x < -3 * PI_F) { float n = (-*x - PI_F) / (2 * PI_F); *x += 2 * PI_F * ceilf(n); normalize_angle(x); }
  else if (*x > PI_F) *x -= 2 * PI_F;
  else if (*x < -PI_F) *x += 2 * PI_F;
}
/* sin / cos / exp of a float argument, CORRECTLY ROUNDED (evaluated in double, rounded once).  The reference calls
 * std::sin / std::cos / std::exp on `fl` = float, i.e. whatever sinf / cosf / expf its libm provides; those are not
 * specified bit for bit (glibc's differ from the correctly rounded value for 1.3 % of the arguments, by one ulp;
 * measured) and a GPU has no glibc.  The correctly rounded value is what every libm approximates and what both this
 * restatement and the device code can compute identically, so that BFGS and Monte-Carlo trajectories can be compared
 * step by step instead of statistically. */
/* gvo_use_libm(1): call this host's sinf / cosf / expf instead -- what the reference compiled HERE (oracle/_ref) executes -- so
 * that the comparison with oracle/_ref can demand bit-identical coordinates and energies; the default (0) is the correctly
 * rounded value that the device kernels reproduce. */
static int g_use_libm = 0;
void gvo_use_libm(int on) { g_use_libm = on; }
static float sin_cr(float x) { return g_use_libm ? sinf(x) : (float)sin((double)x); }
static float cos_cr(float x) { return g_use_libm ? cosf(x) : (float)cos((double)x); }
static float exp_cr(float x) { return g_use_libm ? expf(x) : (float)exp((double)x); }
static void angle_to_q(const float *axis, float angle, float *q) {
  normalize_angle(&angle);
  float c = cos_cr(angle / 2), s = sin_cr(angle / 2);
  q[0] = c; q[1] = s * axis[0]; q[2] = s * axis[1]; q[3] = s * axis[2];
}
static void qmul(const float *l, const float *r, float *o) {
  const float a = l[0], b = l[1], c = l[2], d = l[3];
  o[0] = +a * r[0] - b * r[1] - c * r[2] - d * r[3];
  o[1] = +a * r[1] + b * r[0] + c * r[3] - d * r[2];
  o[2] = +a * r[2] - b * r[3] + c * r[0] + d * r[1];
  o[3] = +a * r[3] + b * r[2] - c * r[1] + d * r[0];
}
static void qnorm_approx(float *q) { /* quaternion.h:243-257 */
  float s = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  if (fabsf(s - 1) < 1e-6f) return;
  float a = sqrtf(s);
  for (int i = 0; i < 4; i++) q[i] *= 1 / a;
}
static void q_to_r3(const float *q, float *m) {
  const float a = q[0], b = q[1], c = q[2], d = q[3];
  const float aa = a * a, ab = a * b, ac = a * c, ad = a * d, bb = b * b, bc = b * c, bd = b * d, cc = c * c, cd = c * d, dd = d * d;
  m[0] = (aa + bb - cc - dd); m[1] = 2 * (-ad + bc); m[2] = 2 * (ac + bd);
  m[3] = 2 * (ad + bc); m[4] = (aa - bb + cc - dd); m[5] = 2 * (-ab + cd);
  m[6] = 2 * (-ac + bd); m[7] = 2 * (ab + cd); m[8] = (aa - bb - cc + dd);
}
static void mv(const float *m, const float *v, float *o) {
  o[0] = m[0] * v[0] + m[1] * v[1] + m[2] * v[2];
  o[1] = m[3] * v[0] + m[4] * v[1] + m[5] * v[2];
  o[2] = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
}
static void quaternion_increment(float *q, const float *rot) { /* quaternion.cu:32-43,96-100 */
  float angle = sqrtf(rot[0] * rot[0] + rot[1] * rot[1] + rot[2] * rot[2]);
  float r[4] = {1, 0, 0, 0};
  if (angle > kEps) {
    float axis[3] = {(1 / angle) * rot[0], (1 / angle) * rot[1], (1 / angle) * rot[2]};
    angle_to_q(axis, angle, r);
  }
  float o[4];
  qmul(r, q, o);
  memcpy(q, o, sizeof o);
  qnorm_approx(q);
}

/* conf = position[3], orientation[4], torsions[n_seg-1].  Outputs: coords, per-segment origin / axis (lab frame). */
void gvo_lig_set_conf(const gvo_lig *L, const float *conf, float *coords, float *seg_origin, float *seg_axis) {
  float *q = (float *)malloc(sizeof(float) * 4 * L->n_seg), *M = (float *)malloc(sizeof(float) * 9 * L->n_seg);
  for (int s = 0; s < L->n_seg; s++) {
    float *o = seg_origin + 3 * s;
    if (s == 0) {
      memcpy(o, conf, 12);
      memcpy(q, conf + 3, 16);
      seg_axis[0] = seg_axis[1] = seg_axis[2] = 0;
    } else {
      const int p = L->seg_parent[s];
      float t[3];
      mv(M + 9 * p, L->seg_rel_origin + 3 * s, t);
      for (int k = 0; k < 3; k++) o[k] = seg_origin[3 * p + k] + t[k];
      mv(M + 9 * p, L->seg_rel_axis + 3 * s, seg_axis + 3 * s);
      float aq[4];
      angle_to_q(seg_axis + 3 * s, conf[7 + s - 1], aq);
      qmul(aq, q + 4 * p, q + 4 * s);
      qnorm_approx(q + 4 * s);
    }
    q_to_r3(q + 4 * s, M + 9 * s);
    for (int i = L->seg_begin[s]; i < L->seg_end[s]; i++) {
      float t[3];
      mv(M + 9 * s, L->local_xyz + 3 * i, t);
      for (int k = 0; k < 3; k++) coords[3 * i + k] = o[k] + t[k];
    }
  }
  free(q); free(M);
}

static void cross(const float *a, const float *b, float *o) {
  o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
/* tree<T>::derivative / branches_derivative, recursion in the reference's order; returns force/torque of node s */
static void node_derivative(const gvo_lig *L, int s, const float *coords, const float *forces, const float *seg_origin,
                            const float *seg_axis, float *change, float *ft /* 6 */) {
  memset(ft, 0, 24);
  for (int i = L->seg_begin[s]; i < L->seg_end[s]; i++) { /* sum_force_and_torque */
    float r[3] = {coords[3 * i] - seg_origin[3 * s], coords[3 * i + 1] - seg_origin[3 * s + 1], coords[3 * i + 2] - seg_origin[3 * s + 2]}, c[3];
    cross(r, forces + 3 * i, c);
    for (int k = 0; k < 3; k++) { ft[k] += forces[3 * i + k]; ft[3 + k] += c[k]; }
  }
  for (int ch = s + 1; ch < L->n_seg; ch++) {
    if (L->seg_parent[ch] != s) continue;
    float cft[6], r[3], c[3];
    node_derivative(L, ch, coords, forces, seg_origin, seg_axis, change, cft);
    for (int k = 0; k < 3; k++) { ft[k] += cft[k]; r[k] = seg_origin[3 * ch + k] - seg_origin[3 * s + k]; }
    cross(r, cft, c);
    for (int k = 0; k < 3; k++) ft[3 + k] += c[k] + cft[3 + k];
  }
  if (s == 0) memcpy(change, ft, 24);
  else change[6 + s - 1] = ft[3] * seg_axis[3 * s] + ft[4] * seg_axis[3 * s + 1] + ft[5] * seg_axis[3 * s + 2];
}

/* model::eval_deriv with ig = cache: returns e, fills change[6+T] (and coords if non-NULL) */
/* non_cache::eval_deriv for one movable atom (lib/non_cache.cpp:126-174): clamp to the box (check_bounds_deriv
 * :102-123, penalty slope * L1 distance), sum e and dor * r over receptor atoms with r^2 < cutoff^2 (the szv_grid only
 * pre-selects candidates), curl(e, deriv, v), add the out-of-box derivative.  deriv may be NULL (non_cache::eval). */
float gvo_noncache_atom(const gvo_field *F, int t1, const float *a, float v, float *deriv) {
  float adj[3], oob[3] = {0, 0, 0}, pen = 0;
  for (int j = 0; j < 3; j++) {
    adj[j] = a[j];
    if (a[j] < F->begin[j]) { adj[j] = F->begin[j]; oob[j] = -1; pen += fabsf(a[j] - F->begin[j]); }
    else if (a[j] > F->end[j]) { adj[j] = F->end[j]; oob[j] = 1; pen += fabsf(a[j] - F->end[j]); }
  }
  pen *= F->slope;
  float e = 0, d[3] = {0, 0, 0};
  for (int b = 0; b < F->n_rec; b++) {
    const int t2 = F->rec_type[b];
    if (t2 < 0 || t2 >= 28 || is_h(t2)) continue; /* grid_atoms: heavy receptor atoms */
    const float r[3] = {adj[0] - F->rec_xyz[3 * b], adj[1] - F->rec_xyz[3 * b + 1], adj[2] - F->rec_xyz[3 * b + 2]};
    const float r2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
    if (r2 < 64.f) {
      float pe, dor;
      if (F->splines) gvo_splines_eval_deriv(F->splines, t1, t2, r2, &pe, &dor); /* non_cache holds the run's precalculate: --minimize's splines */
      else gvo_prec_eval_deriv(F->prec, t1, t2, r2, &pe, &dor);
      e += pe;
      for (int q = 0; q < 3; q++) d[q] += dor * r[q];
    }
  }
  if (e > 0 && v < 0.1f * kMax) { /* curl, lib/curl.h:30-42 */
    const float tmp = (v < kEps) ? 0 : (v / (v + e));
    e *= tmp;
    for (int q = 0; q < 3; q++) d[q] *= tmp * tmp;
  }
  if (deriv) for (int q = 0; q < 3; q++) deriv[q] = d[q] + F->slope * oob[q];
  return e + pen;
}

/* non_cache::within (lib/non_cache.cpp:84-100): every heavy movable atom inside the box (margin 0.0001 as refine uses) */
int gvo_within(const gvo_field *F, const gvo_lig *L, const float *conf, float margin) {
  float *coords = (float *)malloc(12 * L->n_atoms), *so = (float *)malloc(12 * L->n_seg), *sa = (float *)malloc(12 * L->n_seg);
  gvo_lig_set_conf(L, conf, coords, so, sa);
  int ok = 1;
  for (int i = 0; i < L->n_atoms && ok; i++) {
    if (is_h(L->type[i])) continue;
    for (int j = 0; j < 3; j++)
      if (coords[3 * i + j] < F->begin[j] - margin || coords[3 * i + j] > F->end[j] + margin) ok = 0;
  }
  free(coords); free(so); free(sa);
  return ok;
}

float gvo_lig_eval_deriv(const gvo_field *F, const gvo_lig *L, const float *conf, const float *v, float *change, float *coords_out) {
  const int n = L->n_atoms;
  float *coords = (float *)malloc(12 * n), *forces = (float *)calloc(3 * n, 4), *so = (float *)malloc(12 * L->n_seg),
        *sa = (float *)malloc(12 * L->n_seg);
  gvo_lig_set_conf(L, conf, coords, so, sa);
  float e = 0;
  for (int i = 0; i < n; i++) { /* cache::eval_deriv */
    const int t = L->type[i];
    if (t < 0 || t >= 28 || is_h(t)) continue;
    if (F->rec_xyz) e += gvo_noncache_atom(F, t, coords + 3 * i, v[1], forces + 3 * i);
    else e += gvo_grid_evaluate(F->grids[t], F->begin, F->end, F->n, coords + 3 * i, F->slope, v[1], forces + 3 * i);
  }
  float ie = 0;
  for (int k = 0; k < L->n_pairs; k++) { /* eval_interacting_pairs_deriv with v[0] */
    const int a = L->pair_a[k], b = L->pair_b[k];
    float r[3] = {coords[3 * b] - coords[3 * a], coords[3 * b + 1] - coords[3 * a + 1], coords[3 * b + 2] - coords[3 * a + 2]};
    float r2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
    if (r2 < 64.f) {
      float pe, dor, f[3];
      if (F->splines) gvo_splines_eval_deriv(F->splines, L->type[a], L->type[b], r2, &pe, &dor);
      else gvo_prec_eval_deriv(F->prec, L->type[a], L->type[b], r2, &pe, &dor);
      for (int q = 0; q < 3; q++) f[q] = dor * r[q];
      if (pe > 0 && v[0] < 0.1f * kMax) { float tmp = (v[0] < kEps) ? 0 : (v[0] / (v[0] + pe)); pe *= tmp; for (int q = 0; q < 3; q++) f[q] *= tmp * tmp; }
      ie += pe;
      for (int q = 0; q < 3; q++) { forces[3 * a + q] -= f[q]; forces[3 * b + q] += f[q]; }
    }
  }
  e += ie;
  float ft[6];
  node_derivative(L, 0, coords, forces, so, sa, change, ft);
  if (coords_out) memcpy(coords_out, coords, 12 * n);
  free(coords); free(forces); free(so); free(sa);
  return e;
}

/* update_energy: ig->eval(m, v[1]) = cache::eval (intermolecular grid energy only) */
float gvo_lig_eval_grid(const gvo_field *F, const gvo_lig *L, const float *conf, float v1, float *coords_out) {
  const int n = L->n_atoms;
  float *coords = (float *)malloc(12 * n), *so = (float *)malloc(12 * L->n_seg), *sa = (float *)malloc(12 * L->n_seg);
  gvo_lig_set_conf(L, conf, coords, so, sa);
  float e = 0;
  for (int i = 0; i < n; i++) {
    const int t = L->type[i];
    if (t < 0 || t >= 28 || is_h(t)) continue;
    if (F->rec_xyz) e += gvo_noncache_atom(F, t, coords + 3 * i, v1, 0);
    else e += gvo_grid_evaluate(F->grids[t], F->begin, F->end, F->n, coords + 3 * i, F->slope, v1, 0);
  }
  if (coords_out) memcpy(coords_out, coords, 12 * n);
  free(coords); free(so); free(sa);
  return e;
}

/* conf::increment (ligand): position += f*p ; quaternion_increment(orientation, f*p.orientation) ; torsions */
static void conf_increment(float *x, const float *p, float f, int T) {
  for (int k = 0; k < 3; k++) x[k] += f * p[k];
  float rot[3] = {f * p[3], f * p[4], f * p[5]};
  quaternion_increment(x + 3, rot);
  for (int t = 0; t < T; t++) {
    float a = f * p[6 + t];
    normalize_angle(&a);
    x[7 + t] += a;
    normalize_angle(&x[7 + t]);
  }
}
static int tri(int i, int j) { return i <= j ? i + j * (j + 1) / 2 : j + i * (i + 1) / 2; }

/* bfgs (lib/bfgs.h:358-502) with fast_line_search; x in/out (7+T floats), g out (6+T); returns f0 */
static float bfgs_impl(const gvo_field *F, const gvo_lig *L, float *x, float *g, int maxiters, const float *v, int *n_evals, float *x_last,
                       int accurate, int early_term);
float gvo_bfgs(const gvo_field *F, const gvo_lig *L, float *x, float *g, int maxiters, const float *v, int *n_evals) {
  return bfgs_impl(F, L, x, g, maxiters, v, n_evals, 0, 0, 0);
}
/* minimization_params (lib/common.h:50-60): accurate = BFGSAccurateLineSearch (what --minimize selects, main/main.cpp:1160,1186),
 * early_term = --minimize_early_term */
float gvo_bfgs_ex(const gvo_field *F, const gvo_lig *L, float *x, float *g, int maxiters, const float *v, int *n_evals, int accurate,
                  int early_term) {
  return bfgs_impl(F, L, x, g, maxiters, v, n_evals, 0, accurate, early_term);
}
static float acos_cr(float x) { return g_use_libm ? acosf(x) : (float)acos((double)x); }
/* quaternion_to_angle (lib/quaternion.cu:46-62) */
static void q_to_angle(const float *q, float *out) {
  const float c = q[0];
  out[0] = out[1] = out[2] = 0;
  if (c > -1 && c < 1) {
    float angle = 2 * acos_cr(c);
    if (angle > PI_F) angle -= 2 * PI_F;
    const float s = sin_cr(angle / 2);
    if (fabsf(s) < kEps) return;
    const float f = angle / s;
    out[0] = q[1] * f; out[1] = q[2] * f; out[2] = q[3] * f;
  }
}
/* conf::operator()(index) (lib/conf.h:459-473): position, orientation as a rotation vector, torsions */
static float conf_at(const float *x, int i, const float *ang) { return i < 3 ? x[i] : (i < 6 ? ang[i - 3] : x[7 + i - 6]); }
/* x_last (nullable, 7+T): the conformation of the LAST function evaluation = what model::set left in the model's coordinates
 * when quasi_newton returns. bfgs does not re-evaluate at the x it returns: after a line search that used up its 10 trials,
 * or when x_orig is restored (:494-498), the model holds another conformation than the returned one. */
static float bfgs_impl(const gvo_field *F, const gvo_lig *L, float *x, float *g, int maxiters, const float *v, int *n_evals, float *x_last,
                       int accurate, int early_term) {
  const int T = L->n_seg - 1, n = 6 + T, nx = 7 + T;
  float *h = (float *)calloc((size_t)n * (n + 1) / 2, 4), *g_new = (float *)malloc(4 * n), *x_new = (float *)malloc(4 * nx),
        *p = (float *)malloc(4 * n), *y = (float *)malloc(4 * n), *mhy = (float *)malloc(4 * n), *x_orig = (float *)malloc(4 * nx),
        *g_orig = (float *)malloc(4 * n);
  for (int i = 0; i < n; i++) h[tri(i, i)] = 1;
  int evals = 0;
  float f0 = gvo_lig_eval_deriv(F, L, x, v, g, 0); evals++;
  if (x_last) memcpy(x_last, x, 4 * nx);
  const float f_orig = f0;
  memcpy(g_orig, g, 4 * n); memcpy(x_orig, x, 4 * nx);
  int didreset = 0;
  for (int step = 0; step < maxiters; step++) {
    for (int i = 0; i < n; i++) { float s = 0; for (int j = 0; j < n; j++) s += h[tri(i, j)] * g[j]; p[i] = -s; }
    float f1 = 0, alpha = 1, pg = 0;
    if (!accurate) {
      for (int i = 0; i < n; i++) pg += p[i] * g[i];
      for (int trial = 0; trial < 10; trial++) { /* fast_line_search :73-91 */
        memcpy(x_new, x, 4 * nx);
        conf_increment(x_new, p, alpha, T);
        f1 = gvo_lig_eval_deriv(F, L, x_new, v, g_new, 0); evals++;
        if (x_last) memcpy(x_last, x_new, 4 * nx);
        if (f1 - f0 < 0.0001f * alpha * pg) break;
        alpha *= 0.5f;
      }
    } else { /* accurate_line_search :107-180 (after lnsrch of Numerical Recipes); fl = float, the literals 2.0 / 3.0 / .5 are double */
      float a, alpha2 = 0, alamin, b, disc, f2 = 0, rhs1, rhs2, slope = 0, test = 0, tmplam = 0, ang[3];
      const float ALF = (float)1.0e-4, FIRST = 1.0f;
      for (int i = 0; i < n; i++) slope += g[i] * p[i];
      if (slope >= 0) { /* not a descent direction */
        memcpy(x_new, x, 4 * nx); memset(g_new, 0, 4 * n); alpha = 0;
      } else {
        q_to_angle(x + 3, ang);
        for (int i = 0; i < n; i++) { /* compute_lambdamin :93-102 */
          const float ax = fabsf(conf_at(x, i, ang));
          const float temp = fabsf(p[i]) / ((ax < 1.0f) ? 1.0f : ax); /* std::max(std::fabs(x(i)), 1.0f) */
          if (temp > test) test = temp;
        }
        alamin = kEps / test;
        alpha = FIRST;
        for (;;) {
          memcpy(x_new, x, 4 * nx);
          conf_increment(x_new, p, alpha, T);
          f1 = gvo_lig_eval_deriv(F, L, x_new, v, g_new, 0); evals++;
          if (x_last) memcpy(x_last, x_new, 4 * nx);
          if (alpha < alamin || !isfinite(alpha)) { memcpy(x_new, x, 4 * nx); memset(g_new, 0, 4 * n); alpha = 0; break; }
          if (f1 <= f0 + ALF * alpha * slope) break;
          if (alpha == FIRST) tmplam = (float)(-slope / (2.0 * (f1 - f0 - slope)));
          else {
            rhs1 = f1 - f0 - alpha * slope;
            rhs2 = f2 - f0 - alpha2 * slope;
            a = (rhs1 / (alpha * alpha) - rhs2 / (alpha2 * alpha2)) / (alpha - alpha2);
            b = (-alpha2 * rhs1 / (alpha * alpha) + alpha * rhs2 / (alpha2 * alpha2)) / (alpha - alpha2);
            if (a == 0.0) tmplam = (float)(-slope / (2.0 * b));
            else {
              disc = (float)(b * b - 3.0 * a * slope);
              if (disc < 0) tmplam = (float)(0.5 * alpha);
              else if (b <= 0) tmplam = (float)((-b + sqrtf(disc)) / (3.0 * a));
              else tmplam = -slope / (b + sqrtf(disc));
            }
            if (tmplam > .5 * alpha) tmplam = (float)(.5 * alpha);
          }
          alpha2 = alpha; f2 = f1;
          { const float tenth = 0.1f * alpha; alpha = (tmplam < tenth) ? tenth : tmplam; } /* std::max: a NaN tmplam stays NaN */
        }
      }
    }
    if (alpha == 0) break;
    for (int i = 0; i < n; i++) y[i] = g_new[i] - g[i];
    const float prevf0 = f0;
    f0 = f1;
    memcpy(x, x_new, 4 * nx);
    if (early_term && fabs((double)(prevf0 - f0)) < 1e-5) break; /* :455-462, before g is replaced */
    memcpy(g, g_new, 4 * n);
    float gn = 0;
    for (int i = 0; i < n; i++) gn += g[i] * g[i];
    if (!(gn >= 1e-4f)) break;
    if (step == 0 || didreset) {
      float yy = 0, yp = 0;
      for (int i = 0; i < n; i++) { yy += y[i] * y[i]; yp += y[i] * p[i]; }
      didreset = 0;
      if (fabsf(yy) > kEps) for (int i = 0; i < n; i++) h[tri(i, i)] = alpha * yp / yy;
    }
    { /* bfgs_update :52-66 */
      float yp = 0;
      for (int i = 0; i < n; i++) yp += y[i] * p[i];
      if (!(alpha * yp < kEps)) {
        for (int i = 0; i < n; i++) { float s = 0; for (int j = 0; j < n; j++) s += h[tri(i, j)] * y[j]; mhy[i] = -s; }
        float yhy = 0;
        for (int i = 0; i < n; i++) yhy += y[i] * mhy[i];
        yhy = -yhy;
        const float r = 1 / (alpha * yp);
        for (int i = 0; i < n; i++)
          for (int j = i; j < n; j++)
            h[tri(i, j)] += alpha * r * (mhy[i] * p[j] + mhy[j] * p[i]) + +alpha * alpha * (r * r * yhy + r) * p[i] * p[j];
      }
    }
  }
  if (!(f0 <= f_orig)) { f0 = f_orig; memcpy(x, x_orig, 4 * nx); memcpy(g, g_orig, 4 * n); }
  if (n_evals) *n_evals = evals;
  free(h); free(g_new); free(x_new); free(p); free(y); free(mhy); free(x_orig); free(g_orig);
  return f0;
}

/* ---- random numbers: xorshift32, shared bit for bit with the device code ------------------------------ */
/* refine_structure (main/main.cpp:131-171) on a non_cache field: up to 5 BFGS runs (quasi_newton::operator(),
 * lib/quasi_newton.cpp:49-83) with the out-of-box slope 10, 100, ... until every heavy atom is within the box
 * (non_cache::within, margin 1e-4); returns the last run's energy, x is refined in place, *within_out tells whether the
 * final pose is inside (the reference sets out.e = max_fl otherwise). */
float gvo_refine_structure_ex(const gvo_field *F0, const gvo_lig *L, float *x, float *g, int maxiters, const float *v, int *n_evals,
                              int *within_out, int accurate, int early_term);
float gvo_refine_structure(const gvo_field *F0, const gvo_lig *L, float *x, float *g, int maxiters, const float *v, int *n_evals,
                           int *within_out) {
  return gvo_refine_structure_ex(F0, L, x, g, maxiters, v, n_evals, within_out, 0, 0);
}
/* with the minimization_params of --minimize / --local_only (main/main.cpp:264-268 passes par.mc.ssd_par.minparm = the user's) */
float gvo_refine_structure_ex(const gvo_field *F0, const gvo_lig *L, float *x, float *g, int maxiters, const float *v, int *n_evals,
                              int *within_out, int accurate, int early_term) {
  gvo_field F = *F0;
  float slope = 10.f, e = 0;
  int evals = 0, ok = 0;
  for (int p = 0; p < 5; p++) {
    int ne = 0;
    F.slope = slope;
    e = gvo_bfgs_ex(&F, L, x, g, maxiters, v, &ne, accurate, early_term);
    evals += ne;
    ok = gvo_within(&F, L, x, 0.0001f);
    if (ok) break;
    slope *= 10.f;
  }
  if (n_evals) *n_evals = evals;
  if (within_out) *within_out = ok;
  return e;
}

static uint32_t rng_next(uint32_t *s) { uint32_t x = *s; x ^= x << 13; x ^= x >> 17; x ^= x << 5; *s = x; return x; }
static float rng_fl(uint32_t *s, float a, float b) { return a + (b - a) * ((float)(rng_next(s) >> 8) * (1.0f / 16777216.0f)); }
static int rng_int(uint32_t *s, int a, int b) { return a + (int)(rng_next(s) % (uint32_t)(b - a + 1)); }
static void rng_sphere(uint32_t *s, float *o) { /* random_inside_sphere, random.cpp:65-74 */
  for (;;) {
    o[0] = rng_fl(s, -1, 1); o[1] = rng_fl(s, -1, 1); o[2] = rng_fl(s, -1, 1);
    if (o[0] * o[0] + o[1] * o[1] + o[2] * o[2] < 1) return;
  }
}
void gvo_random_conf(uint32_t *seed, const float *c1, const float *c2, int T, float *x) { /* conf::randomize */
  for (int k = 0; k < 3; k++) x[k] = rng_fl(seed, c1[k], c2[k]);
  /* random_orientation: four uniform draws on the unit ball shell would need normals; use rejection on the 4-ball */
  for (;;) {
    float q[4], s = 0;
    for (int k = 0; k < 4; k++) { q[k] = rng_fl(seed, -1, 1); s += q[k] * q[k]; }
    if (s < 1 && s > 1e-3f) { float inv = 1 / sqrtf(s); for (int k = 0; k < 4; k++) x[3 + k] = q[k] * inv; break; }
  }
  for (int t = 0; t < T; t++) x[7 + t] = rng_fl(seed, -PI_F, PI_F);
}

typedef struct {
  int num_steps, maxiters, num_saved_mins;
  float temperature, mutation_amplitude, min_rmsd;
  float hunt_cap[3];
  float gyration_radius;
} gvo_mc_params;

/* mutate_conf, lib/mutate.cpp:35-73 (single ligand, no flexible residues) */
static void mutate_conf(float *x, int T, float amplitude, float gr, uint32_t *s) {
  const int which = rng_int(s, 0, 2 + T - 1);
  float r[3];
  if (which == 0) { rng_sphere(s, r); for (int k = 0; k < 3; k++) x[k] += amplitude * r[k]; return; }
  if (which == 1) {
    if (gr > kEps) { rng_sphere(s, r); float rot[3] = {amplitude / gr * r[0], amplitude / gr * r[1], amplitude / gr * r[2]}; quaternion_increment(x + 3, rot); }
    return;
  }
  x[7 + which - 2] = rng_fl(s, -PI_F, PI_F);
}

/* model::gyration_radius (lib/model.cpp:1002-1014) of the conformation the model currently holds: heavy atoms, about the root origin */
static float gyration_radius_of(const gvo_lig *L, const float *x_state, float *coords, float *so, float *sa) {
  gvo_lig_set_conf(L, x_state, coords, so, sa);
  float acc = 0; int cnt = 0;
  for (int i = 0; i < L->n_atoms; i++) if (!is_h(L->type[i])) {
    const float a = coords[3 * i] - x_state[0], b = coords[3 * i + 1] - x_state[1], c = coords[3 * i + 2] - x_state[2];
    acc += a * a + b * b + c * c; cnt++;
  }
  return cnt > 0 ? sqrtf(acc / cnt) : 0;
}
float gvo_gyration_radius(const gvo_lig *L, const float *conf) {
  float *coords = (float *)malloc(12 * L->n_atoms), *so = (float *)malloc(12 * L->n_seg), *sa = (float *)malloc(12 * L->n_seg);
  const float r = gyration_radius_of(L, conf, coords, so, sa);
  free(coords); free(so); free(sa);
  return r;
}
/* cache::eval / non_cache::eval on given coordinates (update_energy, monte_carlo.cpp:44-47, evaluates whatever the model holds) */
static float grid_energy_on(const gvo_field *F, const gvo_lig *L, const float *coords, float v1) {
  float e = 0;
  for (int i = 0; i < L->n_atoms; i++) {
    const int t = L->type[i];
    if (t < 0 || t >= 28 || is_h(t)) continue;
    if (F->rec_xyz) e += gvo_noncache_atom(F, t, coords + 3 * i, v1, 0);
    else e += gvo_grid_evaluate(F->grids[t], F->begin, F->end, F->n, coords + 3 * i, F->slope, v1, 0);
  }
  return e;
}

/* monte_carlo::operator() (lib/monte_carlo.cpp:99-148).  output container entry: e, conf[7+T], heavy coords; returns the number of
 * entries kept (sorted by e).
 *   init_conf  (nullable): start from this conformation with generator state `seed` instead of drawing conf::randomize here
 *              (the reference's random_orientation draws normals; oracle/_ref hands over its own draw);
 *   state_conf (nullable): MODEL-STATE mode, what the reference's code does to the letter: the model object keeps the coordinates
 *              of the last conformation that was SET (by an evaluation inside quasi_newton or by an explicit m.set), and
 *              (1) mutate_conf takes the gyration radius of THOSE coordinates (mutate.cpp:55, model.cpp:1002-1014),
 *              (2) update_energy evaluates the grid on THOSE coordinates, which are the last line-search trial's, not always the
 *                  returned conformation's (bfgs.h does not re-evaluate at the x it returns).
 *              state_conf is the conformation the model holds on entry.  NULL = the stateless variant the device kernels
 *              implement: constant gyration radius P->gyration_radius, energies re-evaluated at the returned conformation. */
static int mc_impl(const gvo_field *F, const gvo_lig *L, const gvo_mc_params *P, const float *corner1, const float *corner2,
                   uint32_t seed, float *out_e, float *out_conf /* [num_saved_mins][7+T] */, float *trace /* [num_steps] or NULL */,
                   const float *init_conf, const float *state_conf) {
  const int T = L->n_seg - 1, nx = 7 + T, n = 6 + T, na = L->n_atoms;
  int nh = 0;
  for (int i = 0; i < na; i++) nh += !is_h(L->type[i]);
  uint32_t s = seed ? seed : 1u;
  const float av[3] = {1000, 1000, 1000};
  float *tmp = (float *)malloc(4 * nx), *cand = (float *)malloc(4 * nx), *g = (float *)malloc(4 * n), *coords = (float *)malloc(12 * na);
  float *oc = (float *)malloc((size_t)4 * P->num_saved_mins * (3 * nh)), *hv = (float *)malloc(12 * nh);
  float *xs = (float *)malloc(4 * nx), *so = (float *)malloc(12 * L->n_seg), *sa = (float *)malloc(12 * L->n_seg);
  const int stateful = state_conf != 0;
  if (stateful) memcpy(xs, state_conf, 4 * nx);
  int n_out = 0;
  if (init_conf) memcpy(tmp, init_conf, 4 * nx);
  else gvo_random_conf(&s, corner1, corner2, T, tmp);
  float tmp_e = 0, best_e = kMax;
  for (int step = 0; step < P->num_steps; step++) {
    memcpy(cand, tmp, 4 * nx);
    mutate_conf(cand, T, P->mutation_amplitude, stateful ? gyration_radius_of(L, xs, coords, so, sa) : P->gyration_radius, &s);
    float cand_e;
    if (stateful) {
      bfgs_impl(F, L, cand, g, P->maxiters, P->hunt_cap, 0, xs, 0, 0);
      gvo_lig_set_conf(L, xs, coords, so, sa);
      cand_e = grid_energy_on(F, L, coords, av[1]);
    } else {
      gvo_bfgs(F, L, cand, g, P->maxiters, P->hunt_cap, 0);
      cand_e = gvo_lig_eval_grid(F, L, cand, av[1], 0);
    }
    int accept = step == 0 || cand_e < tmp_e;
    if (!accept) { /* metropolis_accept :38-42 */
      const float pr = exp_cr((tmp_e - cand_e) / P->temperature);
      accept = rng_fl(&s, 0, 1) < pr;
    }
    if (accept) {
      memcpy(tmp, cand, 4 * nx); tmp_e = cand_e;
      if (stateful) memcpy(xs, tmp, 4 * nx); /* m.set(tmp.c) :126 */
      if (tmp_e < best_e || n_out < P->num_saved_mins) {
        if (stateful) {
          bfgs_impl(F, L, tmp, g, P->maxiters, av, 0, xs, 0, 0);
          gvo_lig_set_conf(L, xs, coords, so, sa);
          tmp_e = grid_energy_on(F, L, coords, av[1]);
          memcpy(xs, tmp, 4 * nx);                 /* m.set(tmp.c) :134 */
          gvo_lig_set_conf(L, tmp, coords, so, sa); /* get_heavy_atom_movable_coords :136 */
        } else {
          gvo_bfgs(F, L, tmp, g, P->maxiters, av, 0);
          tmp_e = gvo_lig_eval_grid(F, L, tmp, av[1], coords);
        }
        int k = 0;
        for (int i = 0; i < na; i++) if (!is_h(L->type[i])) { memcpy(hv + 3 * k, coords + 3 * i, 12); k++; }
        /* add_to_output_container, lib/coords.cpp:43-56 */
        int ci = n_out; float cr = kMax;
        for (int o = 0; o < n_out; o++) {
          float acc = 0;
          for (int q = 0; q < 3 * nh; q++) { float d = hv[q] - oc[(size_t)o * 3 * nh + q]; acc += d * d; }
          float r = nh > 0 ? sqrtf(acc / nh) : 0;
          if (o == 0 || r < cr) { ci = o; cr = r; }
        }
        int slot = -1;
        if (ci < n_out && cr < P->min_rmsd) { if (tmp_e < out_e[ci]) slot = ci; }
        else if (n_out < P->num_saved_mins) slot = n_out++;
        else if (n_out > 0 && tmp_e < out_e[n_out - 1]) slot = n_out - 1;
        if (slot >= 0) {
          out_e[slot] = tmp_e; memcpy(out_conf + (size_t)slot * nx, tmp, 4 * nx); memcpy(oc + (size_t)slot * 3 * nh, hv, 12 * nh);
          /* out.sort(): insertion sort by e (stable) */
          for (int a = 1; a < n_out; a++)
            for (int b = a; b > 0 && out_e[b] < out_e[b - 1]; b--) {
              float te = out_e[b]; out_e[b] = out_e[b - 1]; out_e[b - 1] = te;
              for (int q = 0; q < nx; q++) { float t2 = out_conf[(size_t)b * nx + q]; out_conf[(size_t)b * nx + q] = out_conf[(size_t)(b - 1) * nx + q]; out_conf[(size_t)(b - 1) * nx + q] = t2; }
              for (int q = 0; q < 3 * nh; q++) { float t2 = oc[(size_t)b * 3 * nh + q]; oc[(size_t)b * 3 * nh + q] = oc[(size_t)(b - 1) * 3 * nh + q]; oc[(size_t)(b - 1) * 3 * nh + q] = t2; }
            }
        }
        if (tmp_e < best_e) best_e = tmp_e;
      }
    }
    if (trace) trace[step] = tmp_e; /* the chain's current energy (monte_carlo.cpp's tmp.e) after this step */
  }
  free(tmp); free(cand); free(g); free(coords); free(oc); free(hv); free(xs); free(so); free(sa);
  return n_out;
}
int gvo_mc_run_traced(const gvo_field *F, const gvo_lig *L, const gvo_mc_params *P, const float *corner1, const float *corner2,
                      uint32_t seed, float *out_e, float *out_conf, float *trace) {
  return mc_impl(F, L, P, corner1, corner2, seed, out_e, out_conf, trace, 0, 0);
}
int gvo_mc_run_ex(const gvo_field *F, const gvo_lig *L, const gvo_mc_params *P, const float *corner1, const float *corner2,
                  uint32_t seed, float *out_e, float *out_conf, float *trace, const float *init_conf, const float *state_conf) {
  return mc_impl(F, L, P, corner1, corner2, seed, out_e, out_conf, trace, init_conf, state_conf);
}
int gvo_mc_run(const gvo_field *F, const gvo_lig *L, const gvo_mc_params *P, const float *corner1, const float *corner2,
               uint32_t seed, float *out_e, float *out_conf) {
  return gvo_mc_run_traced(F, L, P, corner1, corner2, seed, out_e, out_conf, 0);
}
