/*
 * oracle/vina_ref.c — TEST INFRASTRUCTURE ONLY (CPU oracle of the smina/Vina scoring rows V1, V2, V4, V5, V12).
 *
 * Restates, in scalar C with the reference's float (fl = float, lib/common.h:47) operation order:
 *   V1  terms            gauss / repulsion / hydrophobic / non_dir_h_bond      lib/everything.h:149-247,480-506
 *       default weights  main/main.cpp:1324-1329 ; weighted sum                  lib/weighted_terms.cpp:54-68
 *   V2  precalculate_linear (factor 32: n = 2051 samples in r^2)                 lib/precalculate.h:82-272
 *       precalculate_exact::eval_fast                                            lib/precalculate.h:452-463
 *   V4  cache::populate                                                          lib/cache.cpp:104-184
 *   V5  grid::evaluate_aux + cache::eval / eval_deriv, curl                      lib/grid.cpp:96-186, lib/cache.cpp:50-83,
 *                                                                                lib/curl.h:30-42
 *   V12 naive_non_cache::eval (exact terms, per-atom curl) + num_tors_div        lib/naive_non_cache.cpp:29-57,
 *                                                                                lib/everything.h:795-809
 *       non_cache::eval (the docking branch's final intermolecular energy)     lib/non_cache.cpp:52-83
 * PINNED: the reference's tests hold no absolute numbers for these functions (gninacheck compares two live implementations,
 * test_gnina.py only inequalities), but the reference's own sources compile here with stand-in Boost / OpenBabel headers
 * (oracle/Makefile.ref, oracle/ref_driver.cpp -> oracle/_ref).  Against that build every function of this file is BIT-IDENTICAL
 * (terms, precalculate_linear, precalculate_exact, cache::populate, grid evaluation, naive_non_cache::eval, non_cache::eval,
 * num_tors_div) except the spline coefficients (Thomas solve in double vs the reference's dense float inverse: 1e-6):
 * tests/test_oracle_vs_reference_build.py (live, where /root/reference exists) and tests/test_oracle_vina_golden.py (against
 * tests/golden/vina_ref_kat.npz, known answers generated from that build, everywhere).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define NT 28
/* lib/atom_constants.h:101-133: xs_radius, xs_hydrophobe, xs_donor, xs_acceptor */
static const float xs_radius[NT] = {0.37f, 0.37f, 1.9f, 1.9f, 1.9f, 1.9f, 1.8f, 1.8f, 1.8f, 1.8f, 1.7f, 1.7f, 1.7f, 1.7f,
                                    2.0f,  2.0f,  2.1f, 1.5f, 1.8f, 2.0f, 2.2f, 1.2f, 1.2f, 1.2f, 1.2f, 1.2f, 1.2f, 1.92f};
static const int xs_hydrophobe[NT] = {0, 0, 1, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 1};
static const int xs_donor[NT] = {0, 0, 0, 0, 0, 0, 0, 1, 1, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 0};
static const int xs_acceptor[NT] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};

void gvo_type_props(int t, float *radius, int *hydrophobe, int *donor, int *acceptor) {
  *radius = xs_radius[t]; *hydrophobe = xs_hydrophobe[t]; *donor = xs_donor[t]; *acceptor = xs_acceptor[t];
}

static float slope_step(float x_bad, float x_good, float x) { /* everything.h:207-216 */
  if (x_bad < x_good) {
    if (x <= x_bad) return 0;
    if (x >= x_good) return 1;
  } else {
    if (x >= x_bad) return 0;
    if (x <= x_good) return 1;
  }
  return (x - x_bad) / (x_good - x_bad);
}
static float gaussian(float x, float width) { float q = x / width; return expf(-(q * q)); } /* :48-50 */

/* weighted_terms::eval_fast with the default term set; w[0..4] in the order they are added in main.cpp:1324-1328 */
float gvo_eval_terms(const float *w, int t1, int t2, float r) {
  const float R = xs_radius[t1] + xs_radius[t2]; /* optimal_distance */
  float acc = 0.f;
  acc += w[0] * gaussian(r - (R + 0.f), 0.5f);
  acc += w[1] * gaussian(r - (R + 3.f), 2.f);
  {
    float d = r - (R + 0.f);
    acc += w[2] * (d > 0 ? 0.f : d * d);
  }
  acc += w[3] * ((xs_hydrophobe[t1] && xs_hydrophobe[t2]) ? slope_step(1.5f, 0.5f, r - R) : 0.f);
  {
    int hb = (xs_donor[t1] && xs_acceptor[t2]) || (xs_donor[t2] && xs_acceptor[t1]);
    acc += w[4] * (hb ? slope_step(0.f, -0.7f, r - R) : 0.f);
  }
  return acc;
}

/* ---- precalculate_linear --------------------------------------------------------------------------- */
typedef struct {
  float w[6];      /* 5 term weights + num_tors_div weight */
  float factor, cutoff_sqr;
  int n;           /* samples per pair */
  float *rs;       /* n+2 */
  float *fast;     /* [pair][n] */
  float *smooth_e; /* [pair][n] */
  float *smooth_d; /* [pair][n] */
} gvo_prec;

static int tri_index(int t1, int t2) { /* t1 <= t2, triangular_matrix_index.h: i + j*(j+1)/2 */
  return t1 + t2 * (t2 + 1) / 2;
}

gvo_prec *gvo_prec_create(const float *weights6, float factor) {
  gvo_prec *p = (gvo_prec *)calloc(1, sizeof(gvo_prec));
  static const float dflt[6] = {-0.035579f, -0.005156f, 0.840245f, -0.035069f, -0.587439f, (float)(5 * 0.05846 / 0.1 - 1)};
  memcpy(p->w, weights6 ? weights6 : dflt, sizeof(p->w));
  p->factor = factor;
  p->cutoff_sqr = 8.f * 8.f;
  p->n = (int)(size_t)(factor * p->cutoff_sqr) + 3;
  const int n = p->n, npairs = NT * (NT + 1) / 2;
  p->rs = (float *)malloc(sizeof(float) * (n + 2));
  for (int i = 0; i < n + 2; i++) p->rs[i] = sqrtf((float)i / factor);
  p->fast = (float *)malloc(sizeof(float) * npairs * n);
  p->smooth_e = (float *)malloc(sizeof(float) * npairs * n);
  p->smooth_d = (float *)malloc(sizeof(float) * npairs * n);
  for (int t2 = 0; t2 < NT; t2++)
    for (int t1 = 0; t1 <= t2; t1++) {
      float *e = p->smooth_e + (size_t)tri_index(t1, t2) * n, *d = p->smooth_d + (size_t)tri_index(t1, t2) * n,
            *f = p->fast + (size_t)tri_index(t1, t2) * n;
      for (int i = 0; i < n; i++) e[i] = gvo_eval_terms(p->w, t1, t2, p->rs[i]);
      for (int i = 0; i < n; i++) { /* init_from_smooth_fst, precalculate.h:135-158 */
        if (i == 0 || i == n - 1) d[i] = 0;
        else {
          float delta = p->rs[i + 1] - p->rs[i - 1];
          d[i] = (e[i + 1] - e[i - 1]) / (delta * p->rs[i]);
        }
        float f1 = e[i], f2 = (i + 1 >= n) ? 0 : e[i + 1];
        f[i] = (f2 + f1) / 2;
      }
    }
  return p;
}
void gvo_prec_free(gvo_prec *p) {
  if (!p) return;
  free(p->rs); free(p->fast); free(p->smooth_e); free(p->smooth_d); free(p);
}
int gvo_prec_n(const gvo_prec *p) { return p->n; }
void gvo_prec_table(const gvo_prec *p, int t1, int t2, float *fast, float *se, float *sd) {
  if (t1 > t2) { int t = t1; t1 = t2; t2 = t; }
  const size_t o = (size_t)tri_index(t1, t2) * p->n;
  memcpy(fast, p->fast + o, sizeof(float) * p->n);
  memcpy(se, p->smooth_e + o, sizeof(float) * p->n);
  memcpy(sd, p->smooth_d + o, sizeof(float) * p->n);
}
float gvo_prec_eval_fast(const gvo_prec *p, int t1, int t2, float r2) { /* precalculate.h:90-95,166-175 */
  if (t1 > t2) { int t = t1; t1 = t2; t2 = t; }
  return p->fast[(size_t)tri_index(t1, t2) * p->n + (size_t)(p->factor * r2)];
}
void gvo_prec_eval_deriv(const gvo_prec *p, int t1, int t2, float r2, float *e, float *dor) { /* :97-133 */
  if (t1 > t2) { int t = t1; t1 = t2; t2 = t; }
  const size_t o = (size_t)tri_index(t1, t2) * p->n;
  float r2f = p->factor * r2;
  size_t i1 = (size_t)r2f, i2 = i1 + 1;
  float rem = r2f - i1;
  float e1 = p->smooth_e[o + i1], e2 = p->smooth_e[o + i2], d1 = p->smooth_d[o + i1], d2 = p->smooth_d[o + i2];
  *e = e1 + rem * (e2 - e1);
  *dor = d1 + rem * (d2 - d1);
}
float gvo_exact_eval(const gvo_prec *p, int t1, int t2, float r2) { /* precalculate_exact::eval_fast */
  return gvo_eval_terms(p->w, t1, t2, sqrtf(r2));
}

/* ---- cache::populate ------------------------------------------------------------------------------------ */
/* grid for ligand type t2: (n[0]+1) x (n[1]+1) x (n[2]+1) points, x fastest (array3d, gpu_math.h:222);
 * point = init + factor_inv * index, factor = (dim-1)/range (grid.cpp:51-66, grid.h:54-57). */
static int is_h(int t) { return t == 0 || t == 1; }
void gvo_cache_populate(const gvo_prec *p, const float *begin, const float *end, const int32_t *n, int n_rec,
                        const float *rec_xyz, const int32_t *rec_type, int t2, float *out) {
  const int d0 = n[0] + 1, d1 = n[1] + 1, d2 = n[2] + 1;
  float finv[3];
  const int dims[3] = {d0, d1, d2};
  for (int i = 0; i < 3; i++) {
    float factor = (float)(dims[i] - 1.0) / (end[i] - begin[i]);
    finv[i] = 1 / factor;
  }
  for (int z = 0; z < d2; z++)
    for (int y = 0; y < d1; y++)
      for (int x = 0; x < d0; x++) {
        const float px = begin[0] + finv[0] * x, py = begin[1] + finv[1] * y, pz = begin[2] + finv[2] * z;
        float aff = 0.f;
        for (int a = 0; a < n_rec; a++) { /* grid_atoms in index order (szv_grid lists are index-ordered) */
          const int t1 = rec_type[a];
          if (t1 < 0 || t1 >= NT || is_h(t1)) continue; /* grid_atoms hold heavy receptor atoms */
          const float dx = rec_xyz[3 * a] - px, dy = rec_xyz[3 * a + 1] - py, dz = rec_xyz[3 * a + 2] - pz;
          const float r2 = dx * dx + dy * dy + dz * dz;
          if (r2 <= p->cutoff_sqr) aff += gvo_prec_eval_fast(p, t1, t2, r2);
        }
        out[x + (size_t)d0 * (y + (size_t)d1 * z)] = aff;
      }
}

/* ---- grid::evaluate_aux (+ curl) ---------------------------------------------------------------------- */
static const float kMaxFl = 3.402823466e+38f, kEps = 1.1920929e-07f;
static int not_max(float x) { return x < 0.1f * kMaxFl; }
float gvo_grid_evaluate(const float *data, const float *begin, const float *end, const int32_t *n, const float *loc,
                        float slope, float v, float *deriv /* nullable, 3 */) {
  const int dims[3] = {n[0] + 1, n[1] + 1, n[2] + 1};
  float s[3], miss[3] = {0, 0, 0}, factor[3], finv[3], dm1[3];
  int region[3];
  size_t a[3];
  for (int i = 0; i < 3; i++) {
    dm1[i] = (float)(dims[i] - 1.0);
    factor[i] = dm1[i] / (end[i] - begin[i]);
    finv[i] = 1 / factor[i];
    s[i] = (loc[i] - begin[i]) * factor[i];
    if (s[i] < 0) { miss[i] = -s[i]; region[i] = -1; a[i] = 0; s[i] = 0; }
    else if (s[i] >= dm1[i]) { miss[i] = s[i] - dm1[i]; region[i] = 1; a[i] = dims[i] - 2; s[i] = 1; }
    else { region[i] = 0; a[i] = (size_t)s[i]; s[i] -= a[i]; }
  }
  const float penalty = slope * (miss[0] * finv[0] + miss[1] * finv[1] + miss[2] * finv[2]);
#define D(X, Y, Z) data[(X) + (size_t)dims[0] * ((Y) + (size_t)dims[1] * (Z))]
  const size_t x0 = a[0], y0 = a[1], z0 = a[2], x1 = x0 + 1, y1 = y0 + 1, z1 = z0 + 1;
  const float f000 = D(x0, y0, z0), f100 = D(x1, y0, z0), f010 = D(x0, y1, z0), f110 = D(x1, y1, z0), f001 = D(x0, y0, z1),
              f101 = D(x1, y0, z1), f011 = D(x0, y1, z1), f111 = D(x1, y1, z1);
#undef D
  const float x = s[0], y = s[1], z = s[2], mx = 1 - x, my = 1 - y, mz = 1 - z;
  float f = f000 * mx * my * mz + f100 * x * my * mz + f010 * mx * y * mz + f110 * x * y * mz + f001 * mx * my * z +
            f101 * x * my * z + f011 * mx * y * z + f111 * x * y * z;
  if (deriv) {
    float g[3];
    g[0] = f000 * (-1) * my * mz + f100 * 1 * my * mz + f010 * (-1) * y * mz + f110 * 1 * y * mz + f001 * (-1) * my * z +
           f101 * 1 * my * z + f011 * (-1) * y * z + f111 * 1 * y * z;
    g[1] = f000 * mx * (-1) * mz + f100 * x * (-1) * mz + f010 * mx * 1 * mz + f110 * x * 1 * mz + f001 * mx * (-1) * z +
           f101 * x * (-1) * z + f011 * mx * 1 * z + f111 * x * 1 * z;
    g[2] = f000 * mx * my * (-1) + f100 * x * my * (-1) + f010 * mx * y * (-1) + f110 * x * y * (-1) + f001 * mx * my * 1 +
           f101 * x * my * 1 + f011 * mx * y * 1 + f111 * x * y * 1;
    if (f > 0 && not_max(v)) { /* curl, curl.h:30-35 */
      float tmp = (v < kEps) ? 0 : (v / (v + f));
      f *= tmp;
      for (int i = 0; i < 3; i++) g[i] *= tmp * tmp;
    }
    for (int i = 0; i < 3; i++) deriv[i] = factor[i] * ((region[i] == 0) ? g[i] : 0) + slope * region[i];
    return f + penalty;
  }
  if (f > 0 && not_max(v)) { float tmp = (v < kEps) ? 0 : (v / (v + f)); f *= tmp; }
  return f + penalty;
}

/* cache::eval / eval_deriv for one pose: grids[type] may be NULL for types that are not needed */
float gvo_cache_eval(float *const *grids, const float *begin, const float *end, const int32_t *n, int n_lig,
                     const float *lig_xyz, const int32_t *lig_type, float slope, float v, float *deriv /* nullable */) {
  float e = 0;
  for (int i = 0; i < n_lig; i++) {
    const int t = lig_type[i];
    if (t < 0 || t >= NT || is_h(t)) { if (deriv) deriv[3 * i] = deriv[3 * i + 1] = deriv[3 * i + 2] = 0; continue; }
    e += gvo_grid_evaluate(grids[t], begin, end, n, lig_xyz + 3 * i, slope, v, deriv ? deriv + 3 * i : 0);
  }
  return e;
}

/* ---- naive_non_cache::eval with precalculate_exact, then num_tors_div -------------------------------- */
float gvo_naive_exact(const gvo_prec *p, int n_rec, const float *rec_xyz, const int32_t *rec_type, int n_lig,
                      const float *lig_xyz, const int32_t *lig_type, float v) {
  float e = 0;
  for (int i = 0; i < n_lig; i++) {
    const int t1 = lig_type[i];
    if (t1 < 0 || t1 >= NT || is_h(t1)) continue;
    float this_e = 0;
    for (int j = 0; j < n_rec; j++) {
      const int t2 = rec_type[j];
      if (t2 < 0 || t2 >= NT || is_h(t2)) continue;
      const float dx = lig_xyz[3 * i] - rec_xyz[3 * j], dy = lig_xyz[3 * i + 1] - rec_xyz[3 * j + 1],
                  dz = lig_xyz[3 * i + 2] - rec_xyz[3 * j + 2];
      const float r2 = dx * dx + dy * dy + dz * dz;
      if (r2 < p->cutoff_sqr) this_e += gvo_exact_eval(p, t1, t2, r2);
    }
    if (this_e > 0 && not_max(v)) { float tmp = (v < kEps) ? 0 : (v / (v + this_e)); this_e *= tmp; }
    e += this_e;
  }
  return e;
}
/* non_cache::eval (lib/non_cache.cpp:52-83) -- what the DOCKING branch's final "Affinity" goes through (main/main.cpp:340-344:
 * eval_adjusted with ig = nc_new = non_cache(grid_cache, gd, &prec, slope)): per heavy ligand atom the coordinates are clamped to the
 * box (check_bounds :32-50), the pair terms come from the search's precalculate through precalculate::eval = eval_FAST (the
 * piecewise-constant table, precalculate.h:70-74,90-95 -- not the exact terms of --score_only), curl, plus slope x distance outside. */
float gvo_noncache_eval(const gvo_prec *p, int n_rec, const float *rec_xyz, const int32_t *rec_type, int n_lig, const float *lig_xyz,
                        const int32_t *lig_type, float v, float slope, const float *begin, const float *end) {
  float e = 0;
  for (int i = 0; i < n_lig; i++) {
    const int t1 = lig_type[i];
    if (t1 < 0 || t1 >= NT || is_h(t1)) continue;
    float adj[3], pen = 0;
    for (int j = 0; j < 3; j++) {
      const float a = lig_xyz[3 * i + j];
      adj[j] = a;
      if (a < begin[j]) { adj[j] = begin[j]; pen += fabsf(a - begin[j]); }
      else if (a > end[j]) { adj[j] = end[j]; pen += fabsf(a - end[j]); }
    }
    pen *= slope;
    float this_e = 0;
    for (int j = 0; j < n_rec; j++) {
      const int t2 = rec_type[j];
      if (t2 < 0 || t2 >= NT || is_h(t2)) continue; /* the szv_grid lists hold heavy receptor atoms (szv_grid.h:84-93) */
      const float dx = adj[0] - rec_xyz[3 * j], dy = adj[1] - rec_xyz[3 * j + 1], dz = adj[2] - rec_xyz[3 * j + 2];
      const float r2 = dx * dx + dy * dy + dz * dz;
      if (r2 < p->cutoff_sqr) this_e += gvo_prec_eval_fast(p, t1, t2, r2);
    }
    if (this_e > 0 && not_max(v)) { float tmp = (v < kEps) ? 0 : (v / (v + this_e)); this_e *= tmp; }
    e += this_e + pen;
  }
  return e;
}
float gvo_num_tors_div(const gvo_prec *p, float e, float num_tors) { /* everything.h:804-809, smooth_div :52-56 */
  const float w = (float)(0.1 * ((double)p->w[5] + 1)); /* "fl w = 0.1 * (read_iterator(i) + 1)": double product, float store */
  const float wnt = w * num_tors;                 /* "1 + w * in.num_tors / 5.0": fl * fl is a float product, the rest double */
  const float y = (float)(1 + (double)wnt / 5.0);
  if (fabsf(e) < kEps) return 0;
  if (fabsf(y) < kEps) return (e * y > 0) ? kMaxFl : -kMaxFl;
  return e / y;
}

/* ---- V3: precalculate_splines (lib/precalculate.h:380-449) + Spline (lib/splines.h:22-138) ---------------------
 * n = factor*cutoff intervals (factor 10 for --minimize, main/main.cpp:1162-1165): points (i*fraction, E(i*fraction))
 * for i < n plus (cutoff, 0); natural-to-clamped cubic spline with zero first derivative at both ends.  The reference
 * inverts the (tridiagonal, symmetric for even spacing) system densely with Eigen in float; here it is solved with
 * the Thomas algorithm in double and the coefficients are rounded to float (difference: float round-off of the
 * reference's inverse).  eval_deriv returns (value, d/dr divided by r). */
typedef struct gvo_splines { int n; float cutoff, fraction; float *abcd; /* [pair][n][4] */ unsigned char *valid; } gvo_splines;

gvo_splines *gvo_splines_create(const gvo_prec *p, float factor) {
  gvo_splines *s = (gvo_splines *)calloc(1, sizeof(gvo_splines));
  s->cutoff = 8.f;
  s->n = (int)(unsigned)(factor * s->cutoff);
  s->fraction = s->cutoff / (float)s->n;
  const int n = s->n, npairs = NT * (NT + 1) / 2, np = n + 1;
  s->abcd = (float *)calloc((size_t)npairs * n * 4, sizeof(float));
  s->valid = (unsigned char *)calloc(npairs, 1);
  double *y = (double *)malloc(sizeof(double) * np), *C = (double *)malloc(sizeof(double) * np), *dg = (double *)malloc(sizeof(double) * np),
         *up = (double *)malloc(sizeof(double) * np), *lo = (double *)malloc(sizeof(double) * np), *dd = (double *)malloc(sizeof(double) * np);
  float *xs = (float *)malloc(sizeof(float) * np);
  for (int t2 = 0; t2 < NT; t2++)
    for (int t1 = 0; t1 <= t2; t1++) {
      int nonzero = 0;
      for (int i = 0; i < n; i++) {
        xs[i] = i * s->fraction;
        float v = gvo_eval_terms(p->w, t1, t2, xs[i]);
        y[i] = v;
        if (v != 0) nonzero = 1;
      }
      xs[n] = s->cutoff; y[n] = 0;
      const int pi = tri_index(t1, t2);
      s->valid[pi] = (unsigned char)nonzero;
      if (!nonzero) continue;
      const int e = n;
      const double fr = (double)(float)(xs[1] - xs[0]), hlast = (double)(float)(xs[e] - xs[e - 1]);
      /* row vector ddy * A = C with A(i-1,i) = hi, A(i,i) = 2(fr+hi), A(i+1,i) = hi  (columns i = 1..e-1),
       * A(0,0) = 2 fr, A(1,0) = fr, A(e,e) = 2 hlast, A(e-1,e) = hlast.  Equation of column j:
       *   ddy(j-1) A(j-1,j) + ddy(j) A(j,j) + ddy(j+1) A(j+1,j) = C(j) */
      for (int j = 0; j <= e; j++) {
        double hj = (j == e - 1) ? hlast : fr;
        if (j == 0) { lo[j] = 0; dg[j] = 2 * fr; up[j] = fr; C[j] = 6 * ((y[1] - y[0]) / fr); }
        else if (j == e) { lo[j] = hlast; dg[j] = 2 * hlast; up[j] = 0; C[j] = 6 * (-(y[e] - y[e - 1]) / hlast); }
        else { lo[j] = hj; dg[j] = 2 * (fr + hj); up[j] = hj; C[j] = 6 * ((y[j + 1] - y[j]) / hj - (y[j] - y[j - 1]) / fr); }
      }
      /* Thomas */
      for (int j = 1; j <= e; j++) { double m = lo[j] / dg[j - 1]; dg[j] -= m * up[j - 1]; C[j] -= m * C[j - 1]; }
      dd[e] = C[e] / dg[e];
      for (int j = e - 1; j >= 0; j--) dd[j] = (C[j] - up[j] * dd[j + 1]) / dg[j];
      for (int i = 0; i < e; i++) {
        double hi = (i == e - 1) ? hlast : fr;
        float *o = s->abcd + ((size_t)pi * n + i) * 4;
        o[0] = (float)((dd[i + 1] - dd[i]) / (6 * hi));
        o[1] = (float)(dd[i] / 2);
        o[2] = (float)((y[i + 1] - y[i]) / hi - dd[i + 1] * hi / 6 - dd[i] * hi / 3);
        o[3] = (float)y[i];
      }
    }
  free(y); free(C); free(dg); free(up); free(lo); free(dd); free(xs);
  return s;
}
void gvo_splines_free(gvo_splines *s) { if (s) { free(s->abcd); free(s->valid); free(s); } }
int gvo_splines_n(const gvo_splines *s) { return s->n; }
void gvo_splines_table(const gvo_splines *s, int t1, int t2, float *abcd) {
  if (t1 > t2) { int t = t1; t1 = t2; t2 = t; }
  memcpy(abcd, s->abcd + (size_t)tri_index(t1, t2) * s->n * 4, sizeof(float) * 4 * s->n);
}
/* precalculate_splines::eval_deriv: (value, derivative / r) at r = sqrt(r2) */
void gvo_splines_eval_deriv(const gvo_splines *s, int t1, int t2, float r2, float *e, float *dor) {
  if (t1 > t2) { int t = t1; t1 = t2; t2 = t; }
  const int pi = tri_index(t1, t2);
  const float r = sqrtf(r2);
  *e = 0; *dor = 0;
  if (!s->valid[pi] || r >= s->cutoff) return;
  unsigned index = (unsigned)(r / s->fraction);
  if ((int)index >= s->n) index = s->n - 1;
  const float *c = s->abcd + ((size_t)pi * s->n + index) * 4;
  const float lx = r - index * s->fraction; /* SplineData.x = points[i].first = i*fraction */
  *e = ((c[0] * lx + c[1]) * lx + c[2]) * lx + c[3];
  *dor = ((3 * c[0] * lx + 2 * c[1]) * lx + c[2]) / r;
}
