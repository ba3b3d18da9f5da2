// Batched smina/Vina scoring rows on the device:
//   V1/V2 host: default term set (lib/everything.h:149-247,480-506; weights main/main.cpp:1324-1329) tabulated like
//         precalculate_linear (lib/precalculate.h:82-272, factor 32 -> 2051 samples in r^2 per type pair)
//   V4    cache::populate (lib/cache.cpp:104-184): one thread per grid point, receptor atoms in index order
//   V5    cache::eval / eval_deriv -> grid::evaluate_aux (lib/grid.cpp:96-186) + curl (lib/curl.h:30-35)
//   V12   naive_non_cache::eval with precalculate_exact (lib/naive_non_cache.cpp:29-57, lib/precalculate.h:452-463)
//         and num_tors_div (lib/everything.h:795-809): the printed "Affinity (kcal/mol)" of a rigid pose
// Summation orders follow the reference (per-atom partial, curl, then atoms in index order), so results differ from
// the CPU restatement only by the device's expf/sqrtf rounding.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>
#include <vector>
#include "gb_internal.h"

namespace gb {

static const int kHydrophobe[kNumSminaTypes] = {0, 0, 1, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 1};
static const int kDonor[kNumSminaTypes] = {0, 0, 0, 0, 0, 0, 0, 1, 1, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 0};
static const int kAcceptor[kNumSminaTypes] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};

struct VinaTypeProps {
  float radius[kNumSminaTypes];
  int hydrophobe[kNumSminaTypes], donor[kNumSminaTypes], acceptor[kNumSminaTypes];
};
__constant__ VinaTypeProps c_props;
__constant__ float c_w[6];

__host__ __device__ inline float slope_step_hd(float x_bad, float x_good, float x) {
  if (x_bad < x_good) {
    if (x <= x_bad) return 0.f;
    if (x >= x_good) return 1.f;
  } else {
    if (x >= x_bad) return 0.f;
    if (x <= x_good) return 1.f;
  }
  return (x - x_bad) / (x_good - x_bad);
}

// weighted_terms::eval_fast for the default term set, float arithmetic in the reference's order
template <typename P>
__host__ __device__ inline float eval_terms(const P& pr, const float* w, int t1, int t2, float r) {
  const float R = pr.radius[t1] + pr.radius[t2];
  float acc = 0.f;
  { const float q = (r - (R + 0.f)) / 0.5f; acc += w[0] * expf(-(q * q)); }
  { const float q = (r - (R + 3.f)) / 2.f; acc += w[1] * expf(-(q * q)); }
  { const float d = r - (R + 0.f); acc += w[2] * (d > 0 ? 0.f : d * d); }
  acc += w[3] * ((pr.hydrophobe[t1] && pr.hydrophobe[t2]) ? slope_step_hd(1.5f, 0.5f, r - R) : 0.f);
  const bool hb = (pr.donor[t1] && pr.acceptor[t2]) || (pr.donor[t2] && pr.acceptor[t1]);
  acc += w[4] * (hb ? slope_step_hd(0.f, -0.7f, r - R) : 0.f);
  return acc;
}

struct Vina {
  int device = 0;
  cudaStream_t stream = nullptr;
  // grow-only device workspace of the docking entry points (no cudaMalloc/cudaFree -- and their implicit device
  // synchronisation -- per call, so handles on different host threads overlap their kernels)
  void* dws[8] = {};
  size_t dws_cap[8] = {};
  template <typename T>
  T* ws(int i, size_t n) {
    size_t bytes = std::max<size_t>(n, 1) * sizeof(T);
    if (dws_cap[i] < bytes) {
      bytes = std::max(bytes, dws_cap[i] + dws_cap[i] / 2);
      if (dws[i]) cudaFree(dws[i]);
      dws[i] = nullptr; dws_cap[i] = 0;
      GB_CUDA(cudaMalloc(&dws[i], bytes));
      dws_cap[i] = bytes;
    }
    return reinterpret_cast<T*>(dws[i]);
  }
  // pinned staging for results: a device->host copy into PAGEABLE memory waits for the stream's kernel inside the
  // driver call, holding driver locks that stall the launches of other host threads (measured: 8 handles on 8 threads
  // ran their Monte-Carlo kernels back to back); copies into pinned memory are truly asynchronous
  void* pws[8] = {};
  size_t pws_cap[8] = {};
  template <typename T>
  T* pin(int i, size_t n) {
    size_t bytes = std::max<size_t>(n, 1) * sizeof(T);
    if (pws_cap[i] < bytes) {
      bytes = std::max(bytes, pws_cap[i] + pws_cap[i] / 2);
      if (pws[i]) cudaFreeHost(pws[i]);
      pws[i] = nullptr; pws_cap[i] = 0;
      GB_CUDA(cudaMallocHost(&pws[i], bytes));
      pws_cap[i] = bytes;
    }
    return reinterpret_cast<T*>(pws[i]);
  }
  float w[6];
  float factor = 32.f, cutoff_sqr = 64.f;
  int n = 0;
  std::vector<float> h_fast, h_se, h_sd;  // [pair][n]
  float* d_fast = nullptr;
  float2* d_smooth = nullptr;  // [pair][n] (e, dor) for eval_deriv
  // precalculate_splines (V3): n_sp intervals per pair, coefficients (a,b,c,d); valid = some knot is non-zero
  int n_sp = 0;
  float sp_fraction = 0.f;
  std::vector<float4> h_sp;
  std::vector<unsigned char> h_sp_valid;
  float4* d_sp = nullptr;
  int use_splines = 0;
  // ligand topology for the docking inner loop (gb_vina_set_ligand)
  struct LigDev {
    int n_atoms = 0, n_seg = 0, n_pairs = 0, max_depth = 0, n_heavy = 0;
    float gyration_radius = 0;
    float4* local = nullptr;   // x, y, z, smina type
    int* atom_seg = nullptr;
    int4* seg = nullptr;       // parent, atom begin, atom end, depth
    float4* seg_rel_origin = nullptr;
    float4* seg_rel_axis = nullptr;
    int2* pairs = nullptr;
    int* adj_off = nullptr;    // [n_atoms + 1] CSR over the pair list, both directions
    int* adj = nullptr;        // [2 n_pairs] partner atom | pair index << 8 | (this atom is the pair's first) << 30
    int* child_off = nullptr;  // [n_seg + 1] CSR of the torsion tree: children of every segment, ascending
    int* child = nullptr;      // [n_seg - 1]
    unsigned heavy_types = 0;  // bit t set: a movable heavy atom of smina type t (checked against the built grids)
  } lig;
  // receptor (heavy atoms, index order)
  float4* d_rec = nullptr;  // x, y, z, type
  int n_rec = 0;
  // cache
  float begin[3], end[3];
  int gn[3] = {0, 0, 0};
  float* d_grids[kNumSminaTypes] = {};       // grids of the CURRENT cache (null = type not built)
  float* grid_pool[kNumSminaTypes] = {};     // grow-only backing store, reused from ligand to ligand
  size_t grid_pool_cap[kNumSminaTypes] = {};
  // pose staging
  float4* d_lig = nullptr; int* d_off = nullptr; float* d_atom_e = nullptr; float* d_deriv = nullptr; float* d_pose_e = nullptr;
  float* d_tors = nullptr;
  size_t cap_atoms = 0, cap_poses = 0;
  ~Vina() {
    cudaSetDevice(device);
    if (stream) cudaStreamDestroy(stream);
    for (auto q : dws) cudaFree(q);
    for (auto q : pws) cudaFreeHost(q);
    cudaFree(d_fast); cudaFree(d_smooth); cudaFree(d_sp); cudaFree(d_rec);
    cudaFree(lig.local); cudaFree(lig.atom_seg); cudaFree(lig.seg); cudaFree(lig.seg_rel_origin); cudaFree(lig.seg_rel_axis);
    cudaFree(lig.pairs); cudaFree(lig.adj_off); cudaFree(lig.adj); cudaFree(lig.child_off); cudaFree(lig.child); cudaFree(d_lig); cudaFree(d_off); cudaFree(d_atom_e); cudaFree(d_deriv);
    cudaFree(d_pose_e); cudaFree(d_tors);
    for (auto g : grid_pool) cudaFree(g);
  }
};

static inline int tri_index(int t1, int t2) { return t1 + t2 * (t2 + 1) / 2; }  // t1 <= t2 (triangular_matrix_index.h)

static void build_tables(Vina& v) {
  VinaTypeProps pr;
  for (int t = 0; t < kNumSminaTypes; t++) {
    pr.radius[t] = kSminaXsRadius[t]; pr.hydrophobe[t] = kHydrophobe[t]; pr.donor[t] = kDonor[t]; pr.acceptor[t] = kAcceptor[t];
  }
  v.n = (int)(size_t)(v.factor * v.cutoff_sqr) + 3;  // precalculate.h:182
  const int n = v.n, npairs = kNumSminaTypes * (kNumSminaTypes + 1) / 2;
  std::vector<float> rs(n + 2);
  for (int i = 0; i < n + 2; i++) rs[i] = std::sqrt((float)i / v.factor);  // calculate_rs, :262-267
  v.h_fast.assign((size_t)npairs * n, 0.f); v.h_se.assign((size_t)npairs * n, 0.f); v.h_sd.assign((size_t)npairs * n, 0.f);
  for (int t2 = 0; t2 < kNumSminaTypes; t2++)
    for (int t1 = 0; t1 <= t2; t1++) {
      float* e = &v.h_se[(size_t)tri_index(t1, t2) * n];
      float* d = &v.h_sd[(size_t)tri_index(t1, t2) * n];
      float* f = &v.h_fast[(size_t)tri_index(t1, t2) * n];
      for (int i = 0; i < n; i++) e[i] = eval_terms(pr, v.w, t1, t2, rs[i]);
      for (int i = 0; i < n; i++) {  // init_from_smooth_fst, :135-158
        if (i == 0 || i == n - 1) d[i] = 0;
        else d[i] = (e[i + 1] - e[i - 1]) / ((rs[i + 1] - rs[i - 1]) * rs[i]);
        const float f1 = e[i], f2 = (i + 1 >= n) ? 0.f : e[i + 1];
        f[i] = (f2 + f1) / 2;
      }
    }
  GB_CUDA(cudaMalloc(&v.d_fast, v.h_fast.size() * sizeof(float)));
  GB_CUDA(cudaMemcpy(v.d_fast, v.h_fast.data(), v.h_fast.size() * sizeof(float), cudaMemcpyHostToDevice));
  {
    std::vector<float2> sm(v.h_se.size());
    for (size_t i = 0; i < sm.size(); i++) sm[i] = make_float2(v.h_se[i], v.h_sd[i]);
    GB_CUDA(cudaMalloc(&v.d_smooth, sm.size() * sizeof(float2)));
    GB_CUDA(cudaMemcpy(v.d_smooth, sm.data(), sm.size() * sizeof(float2), cudaMemcpyHostToDevice));
  }
  {  // V3: Spline::initialize (lib/splines.h:38-97) for every pair: knots i*fraction (i < n) + (cutoff, 0), zero end
     // slopes; the tridiagonal system is solved with the Thomas algorithm in double (the reference inverts it densely
     // in float with Eigen), coefficients stored as float like SplineData
    const float cutoff = 8.f, spf = 10.f;  // --minimize uses factor 10 (main/main.cpp:1162-1165)
    const int nsp = (int)(unsigned)(spf * cutoff);
    v.n_sp = nsp; v.sp_fraction = cutoff / (float)nsp;
    v.h_sp.assign((size_t)npairs * nsp, make_float4(0, 0, 0, 0));
    v.h_sp_valid.assign(npairs, 0);
    std::vector<double> y(nsp + 1), C(nsp + 1), dg(nsp + 1), up(nsp + 1), lo(nsp + 1), dd(nsp + 1);
    std::vector<float> xs(nsp + 1);
    for (int t2 = 0; t2 < kNumSminaTypes; t2++)
      for (int t1 = 0; t1 <= t2; t1++) {
        bool nonzero = false;
        for (int i = 0; i < nsp; i++) {
          xs[i] = i * v.sp_fraction;
          const float val = eval_terms(pr, v.w, t1, t2, xs[i]);
          y[i] = val;
          nonzero |= val != 0;
        }
        xs[nsp] = cutoff; y[nsp] = 0;
        const int pi = tri_index(t1, t2), e = nsp;
        v.h_sp_valid[pi] = nonzero;
        if (!nonzero) continue;
        const double fr = (double)(float)(xs[1] - xs[0]), hlast = (double)(float)(xs[e] - xs[e - 1]);
        for (int j2 = 0; j2 <= e; j2++) {
          const double hj = (j2 == e - 1) ? hlast : fr;
          if (j2 == 0) { lo[j2] = 0; dg[j2] = 2 * fr; up[j2] = fr; C[j2] = 6 * ((y[1] - y[0]) / fr); }
          else if (j2 == e) { lo[j2] = hlast; dg[j2] = 2 * hlast; up[j2] = 0; C[j2] = 6 * (-(y[e] - y[e - 1]) / hlast); }
          else { lo[j2] = hj; dg[j2] = 2 * (fr + hj); up[j2] = hj; C[j2] = 6 * ((y[j2 + 1] - y[j2]) / hj - (y[j2] - y[j2 - 1]) / fr); }
        }
        for (int j2 = 1; j2 <= e; j2++) { const double mm = lo[j2] / dg[j2 - 1]; dg[j2] -= mm * up[j2 - 1]; C[j2] -= mm * C[j2 - 1]; }
        dd[e] = C[e] / dg[e];
        for (int j2 = e - 1; j2 >= 0; j2--) dd[j2] = (C[j2] - up[j2] * dd[j2 + 1]) / dg[j2];
        for (int i = 0; i < e; i++) {
          const double hi = (i == e - 1) ? hlast : fr;
          v.h_sp[(size_t)pi * nsp + i] = make_float4((float)((dd[i + 1] - dd[i]) / (6 * hi)), (float)(dd[i] / 2),
                                                     (float)((y[i + 1] - y[i]) / hi - dd[i + 1] * hi / 6 - dd[i] * hi / 3), (float)y[i]);
        }
      }
    GB_CUDA(cudaMalloc(&v.d_sp, v.h_sp.size() * sizeof(float4)));
    GB_CUDA(cudaMemcpy(v.d_sp, v.h_sp.data(), v.h_sp.size() * sizeof(float4), cudaMemcpyHostToDevice));
  }
  GB_CUDA(cudaMemcpyToSymbol(c_props, &pr, sizeof(pr)));
  GB_CUDA(cudaMemcpyToSymbol(c_w, v.w, sizeof(v.w)));
}

// ---- V4 ------------------------------------------------------------------------------------------------------
constexpr int kMaxNeeded = 16;
struct NeededTypes { int n; int t[kMaxNeeded]; float* grid[kMaxNeeded]; };

__global__ void __launch_bounds__(128) cache_populate_kernel(const float4* __restrict__ rec, int n_rec,
                                                             const float* __restrict__ fast, int n_samples, float factor,
                                                             float cutoff_sqr, float bx, float by, float bz, float fx,
                                                             float fy, float fz, int d0, int d1, int d2, NeededTypes nt) {
  const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (idx >= (size_t)d0 * d1 * d2) return;
  const int x = idx % d0, y = (idx / d0) % d1, z = idx / ((size_t)d0 * d1);
  const float px = __fadd_rn(bx, __fmul_rn(fx, (float)x)), py = __fadd_rn(by, __fmul_rn(fy, (float)y)),
              pz = __fadd_rn(bz, __fmul_rn(fz, (float)z));
  float aff[kMaxNeeded];
#pragma unroll
  for (int j = 0; j < kMaxNeeded; j++) aff[j] = 0.f;
  for (int a = 0; a < n_rec; a++) {
    const float4 r = rec[a];
    const float dx = r.x - px, dy = r.y - py, dz = r.z - pz;
    const float r2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    if (r2 <= cutoff_sqr) {
      const int t1 = (int)r.w;
      const size_t i = (size_t)(__fmul_rn(factor, r2));
#pragma unroll
      for (int j = 0; j < kMaxNeeded; j++)
        if (j < nt.n) {
          const int t2 = nt.t[j];
          const int pair = t1 <= t2 ? t1 + t2 * (t2 + 1) / 2 : t2 + t1 * (t1 + 1) / 2;
          aff[j] = __fadd_rn(aff[j], fast[(size_t)pair * n_samples + i]);
        }
    }
  }
#pragma unroll
  for (int j = 0; j < kMaxNeeded; j++)
    if (j < nt.n) nt.grid[j][idx] = aff[j];
}

// ---- V5 ------------------------------------------------------------------------------------------------------
struct GridGeom { float begin[3], factor[3], finv[3], dm1[3]; int dims[3]; };
struct GridPtrs { const float* g[kNumSminaTypes]; };

__device__ inline float grid_evaluate_dev(const float* __restrict__ data, const GridGeom& G, float lx, float ly, float lz,
                                          float slope, float v, float* deriv) {
  const float loc[3] = {lx, ly, lz};
  float s[3], miss[3] = {0, 0, 0};
  int region[3], a[3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    s[i] = __fmul_rn(loc[i] - G.begin[i], G.factor[i]);
    if (s[i] < 0) { miss[i] = -s[i]; region[i] = -1; a[i] = 0; s[i] = 0; }
    else if (s[i] >= G.dm1[i]) { miss[i] = s[i] - G.dm1[i]; region[i] = 1; a[i] = G.dims[i] - 2; s[i] = 1; }
    else { region[i] = 0; a[i] = (int)s[i]; s[i] -= a[i]; }
  }
  const float penalty = __fmul_rn(slope, __fadd_rn(__fadd_rn(__fmul_rn(miss[0], G.finv[0]), __fmul_rn(miss[1], G.finv[1])),
                                                   __fmul_rn(miss[2], G.finv[2])));
  const size_t d0 = G.dims[0], d01 = (size_t)G.dims[0] * G.dims[1];
  const float* b = data + a[0] + d0 * a[1] + d01 * a[2];
  const float f000 = b[0], f100 = b[1], f010 = b[d0], f110 = b[d0 + 1], f001 = b[d01], f101 = b[d01 + 1],
              f011 = b[d01 + d0], f111 = b[d01 + d0 + 1];
  const float x = s[0], y = s[1], z = s[2], mx = 1 - x, my = 1 - y, mz = 1 - z;
#define M3(A, B, C, Dd) __fmul_rn(__fmul_rn(__fmul_rn(A, B), C), Dd)
#define S8(a0, a1, a2, a3, a4, a5, a6, a7) \
  __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(a0, a1), a2), a3), a4), a5), a6), a7)
  float f = S8(M3(f000, mx, my, mz), M3(f100, x, my, mz), M3(f010, mx, y, mz), M3(f110, x, y, mz), M3(f001, mx, my, z),
               M3(f101, x, my, z), M3(f011, mx, y, z), M3(f111, x, y, z));
  if (deriv) {
    float g[3];
    g[0] = S8(M3(f000, -1.f, my, mz), M3(f100, 1.f, my, mz), M3(f010, -1.f, y, mz), M3(f110, 1.f, y, mz),
              M3(f001, -1.f, my, z), M3(f101, 1.f, my, z), M3(f011, -1.f, y, z), M3(f111, 1.f, y, z));
    g[1] = S8(M3(f000, mx, -1.f, mz), M3(f100, x, -1.f, mz), M3(f010, mx, 1.f, mz), M3(f110, x, 1.f, mz),
              M3(f001, mx, -1.f, z), M3(f101, x, -1.f, z), M3(f011, mx, 1.f, z), M3(f111, x, 1.f, z));
    g[2] = S8(M3(f000, mx, my, -1.f), M3(f100, x, my, -1.f), M3(f010, mx, y, -1.f), M3(f110, x, y, -1.f),
              M3(f001, mx, my, 1.f), M3(f101, x, my, 1.f), M3(f011, mx, y, 1.f), M3(f111, x, y, 1.f));
    if (f > 0 && v < 0.1f * 3.402823466e+38f) {
      const float tmp = (v < 1.1920929e-07f) ? 0.f : __fdiv_rn(v, __fadd_rn(v, f));
      f = __fmul_rn(f, tmp);
      const float t2 = __fmul_rn(tmp, tmp);
#pragma unroll
      for (int i = 0; i < 3; i++) g[i] = __fmul_rn(g[i], t2);
    }
#pragma unroll
    for (int i = 0; i < 3; i++)
      deriv[i] = __fadd_rn(__fmul_rn(G.factor[i], (region[i] == 0) ? g[i] : 0.f), __fmul_rn(slope, (float)region[i]));
    return __fadd_rn(f, penalty);
  }
#undef M3
#undef S8
  if (f > 0 && v < 0.1f * 3.402823466e+38f) {
    const float tmp = (v < 1.1920929e-07f) ? 0.f : __fdiv_rn(v, __fadd_rn(v, f));
    f = __fmul_rn(f, tmp);
  }
  return __fadd_rn(f, penalty);
}

__global__ void cache_eval_kernel(const float4* __restrict__ lig, int n_atoms, GridPtrs gp, GridGeom G, float slope, float v,
                                  float* __restrict__ atom_e, float* __restrict__ deriv) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_atoms) return;
  const float4 a = lig[i];
  const int t = (int)a.w;
  float e = 0.f, d[3] = {0.f, 0.f, 0.f};
  if (t >= 2 && t < kNumSminaTypes && gp.g[t]) e = grid_evaluate_dev(gp.g[t], G, a.x, a.y, a.z, slope, v, deriv ? d : nullptr);
  atom_e[i] = e;
  if (deriv) { deriv[3 * i] = d[0]; deriv[3 * i + 1] = d[1]; deriv[3 * i + 2] = d[2]; }
}

// e += per-atom energy in atom order (the reference's accumulation), one thread per pose
__global__ void pose_sum_kernel(const float* __restrict__ atom_e, const int* __restrict__ off, int n_poses,
                                float* __restrict__ pose_e) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_poses) return;
  float e = 0.f;
  for (int i = off[p]; i < off[p + 1]; i++) e = __fadd_rn(e, atom_e[i]);
  pose_e[p] = e;
}

// ---- V12 -----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) naive_exact_kernel(const float4* __restrict__ lig, int n_atoms,
                                                          const float4* __restrict__ rec, int n_rec, float cutoff_sqr,
                                                          float v, float* __restrict__ atom_e) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_atoms) return;
  const float4 a = lig[i];
  const int t1 = (int)a.w;
  float this_e = 0.f;
  if (t1 >= 2 && t1 < kNumSminaTypes) {
    for (int j = 0; j < n_rec; j++) {
      const float4 b = rec[j];
      const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
      const float r2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
      if (r2 < cutoff_sqr) this_e = __fadd_rn(this_e, eval_terms(c_props, c_w, t1, (int)b.w, sqrtf(r2)));
    }
    if (this_e > 0 && v < 0.1f * 3.402823466e+38f) {
      const float tmp = (v < 1.1920929e-07f) ? 0.f : __fdiv_rn(v, __fadd_rn(v, this_e));
      this_e = __fmul_rn(this_e, tmp);
    }
  }
  atom_e[i] = this_e;
}

// non_cache::eval (lib/non_cache.cpp:52-83): the intermolecular energy of the DOCKING branch's final score (main/main.cpp:340-344:
// eval_adjusted with ig = nc_new).  Per heavy ligand atom: coordinates clamped to the search box (check_bounds :32-50), pair terms
// from the search's precalculate through precalculate::eval = eval_fast (the piecewise-constant table the affinity grids are built
// from), receptor atoms in index order, curl, + slope x distance outside the box.
__global__ void __launch_bounds__(128) noncache_eval_kernel(const float4* __restrict__ lig, int n_atoms, const float4* __restrict__ rec,
                                                            int n_rec, const float* __restrict__ fast, int n_samples, float factor,
                                                            float cutoff_sqr, float v, float slope, float b0, float b1, float b2,
                                                            float e0, float e1, float e2, float* __restrict__ atom_e) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_atoms) return;
  const float4 a = lig[i];
  const int t1 = (int)a.w;
  float out = 0.f;
  if (t1 >= 2 && t1 < kNumSminaTypes) {
    const float c[3] = {a.x, a.y, a.z}, bb[3] = {b0, b1, b2}, be[3] = {e0, e1, e2};
    float adj[3], pen = 0.f;
#pragma unroll
    for (int j = 0; j < 3; j++) {
      adj[j] = c[j];
      if (c[j] < bb[j]) { adj[j] = bb[j]; pen = __fadd_rn(pen, fabsf(c[j] - bb[j])); }
      else if (c[j] > be[j]) { adj[j] = be[j]; pen = __fadd_rn(pen, fabsf(c[j] - be[j])); }
    }
    pen = __fmul_rn(pen, slope);
    float this_e = 0.f;
    for (int j = 0; j < n_rec; j++) {
      const float4 b = rec[j];
      const int t2 = (int)b.w;
      if (t2 < 2 || t2 >= kNumSminaTypes) continue;
      const float dx = adj[0] - b.x, dy = adj[1] - b.y, dz = adj[2] - b.z;
      const float r2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
      if (r2 < cutoff_sqr) {
        const int pair = t1 <= t2 ? t1 + t2 * (t2 + 1) / 2 : t2 + t1 * (t1 + 1) / 2;
        this_e = __fadd_rn(this_e, fast[(size_t)pair * n_samples + (size_t)(__fmul_rn(factor, r2))]);
      }
    }
    if (this_e > 0 && v < 0.1f * 3.402823466e+38f) {
      const float tmp = (v < 1.1920929e-07f) ? 0.f : __fdiv_rn(v, __fadd_rn(v, this_e));
      this_e = __fmul_rn(this_e, tmp);
    }
    out = __fadd_rn(this_e, pen);
  }
  atom_e[i] = out;
}

}  // namespace gb

using namespace gb;
struct gb_vina { Vina v; };

#define GBV_BEGIN try {
#define GBV_END                                                                         \
  }                                                                                     \
  catch (const gb::Error& e) { gb::set_last_error(e.what()); return e.code; }            \
  catch (const std::exception& e) { gb::set_last_error(e.what()); return GB_ERR_INTERNAL; } \
  return GB_OK;

static void stage_poses(Vina& v, const float* lig_xyz, const int32_t* lig_type, const int32_t* off, int n_poses) {
  const int n_atoms = off[n_poses];
  if ((size_t)n_atoms > v.cap_atoms) {
    cudaFree(v.d_lig); cudaFree(v.d_atom_e); cudaFree(v.d_deriv);
    v.cap_atoms = std::max<size_t>({(size_t)n_atoms + 1024, 16384, v.cap_atoms * 2});  // floor + doubling: rare device-wide syncs
    GB_CUDA(cudaMalloc(&v.d_lig, v.cap_atoms * sizeof(float4)));
    GB_CUDA(cudaMalloc(&v.d_atom_e, v.cap_atoms * sizeof(float)));
    GB_CUDA(cudaMalloc(&v.d_deriv, v.cap_atoms * 3 * sizeof(float)));
  }
  if ((size_t)n_poses + 1 > v.cap_poses) {
    cudaFree(v.d_off); cudaFree(v.d_pose_e); cudaFree(v.d_tors);
    v.cap_poses = std::max<size_t>({(size_t)n_poses + 1024, 4096, v.cap_poses * 2});
    GB_CUDA(cudaMalloc(&v.d_off, v.cap_poses * sizeof(int)));
    GB_CUDA(cudaMalloc(&v.d_pose_e, v.cap_poses * sizeof(float)));
    GB_CUDA(cudaMalloc(&v.d_tors, v.cap_poses * sizeof(float)));
  }
  std::vector<float4> h(n_atoms);
  for (int i = 0; i < n_atoms; i++) h[i] = make_float4(lig_xyz[3 * i], lig_xyz[3 * i + 1], lig_xyz[3 * i + 2], (float)lig_type[i]);
  GB_CUDA(cudaMemcpyAsync(v.d_lig, h.data(), (size_t)n_atoms * sizeof(float4), cudaMemcpyHostToDevice, v.stream));
  GB_CUDA(cudaMemcpyAsync(v.d_off, off, ((size_t)n_poses + 1) * sizeof(int), cudaMemcpyHostToDevice, v.stream));
  GB_CUDA(cudaStreamSynchronize(v.stream));
}

extern "C" {

int gb_vina_create(int device, const float* weights6, float factor, gb_vina** out) {
  GBV_BEGIN
  GB_CHECK(out, "null argument");
  prefer_many_hw_queues();
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    throw Error(GB_ERR_NO_DEVICE, "gnina_b200: no CUDA device visible — this library has no CPU fallback");
  }
  GB_CUDA(cudaSetDevice(device));
  std::unique_ptr<gb_vina> h(new gb_vina);
  static const float dflt[6] = {-0.035579f, -0.005156f, 0.840245f, -0.035069f, -0.587439f, (float)(5 * 0.05846 / 0.1 - 1)};
  memcpy(h->v.w, weights6 ? weights6 : dflt, sizeof(h->v.w));
  h->v.device = device;
  h->v.factor = factor > 0 ? factor : 32.f;
  GB_CUDA(cudaStreamCreateWithFlags(&h->v.stream, cudaStreamNonBlocking));
  build_tables(h->v);
  *out = h.release();
  GBV_END
}

void gb_vina_destroy(gb_vina* h) { delete h; }

int gb_vina_table_size(const gb_vina* h) { return h ? h->v.n : 0; }

int gb_vina_prec_table(const gb_vina* h, int t1, int t2, float* fast, float* smooth_e, float* smooth_dor) {
  GBV_BEGIN
  GB_CHECK(h && t1 >= 0 && t2 >= 0 && t1 < kNumSminaTypes && t2 < kNumSminaTypes, "bad type");
  if (t1 > t2) std::swap(t1, t2);
  const size_t o = (size_t)tri_index(t1, t2) * h->v.n;
  if (fast) memcpy(fast, &h->v.h_fast[o], sizeof(float) * h->v.n);
  if (smooth_e) memcpy(smooth_e, &h->v.h_se[o], sizeof(float) * h->v.n);
  if (smooth_dor) memcpy(smooth_dor, &h->v.h_sd[o], sizeof(float) * h->v.n);
  GBV_END
}

int gb_vina_set_receptor(gb_vina* h, const float* xyz, const int32_t* smina_type, int n) {
  GBV_BEGIN
  GB_CHECK(h && n >= 0 && (n == 0 || (xyz && smina_type)), "bad receptor");
  Vina& v = h->v;
  GB_CUDA(cudaSetDevice(v.device));
  std::vector<float4> heavy;  // grid_atoms: heavy receptor atoms in index order
  for (int i = 0; i < n; i++)
    if (smina_type[i] >= 2 && smina_type[i] < kNumSminaTypes)
      heavy.push_back(make_float4(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], (float)smina_type[i]));
  cudaFree(v.d_rec);
  v.d_rec = nullptr;
  v.n_rec = (int)heavy.size();
  if (v.n_rec) {
    GB_CUDA(cudaMalloc(&v.d_rec, heavy.size() * sizeof(float4)));
    GB_CUDA(cudaMemcpy(v.d_rec, heavy.data(), heavy.size() * sizeof(float4), cudaMemcpyHostToDevice));
  }
  GBV_END
}

int gb_vina_cache_build(gb_vina* h, const float* begin, const float* end, const int32_t* n, const int32_t* types_needed,
                        int n_types) {
  GBV_BEGIN
  GB_CHECK(h && begin && end && n && types_needed && n_types > 0 && n_types <= kNumSminaTypes, "bad cache arguments");
  Vina& v = h->v;
  GB_CUDA(cudaSetDevice(v.device));
  for (auto& g : v.d_grids) g = nullptr;
  const size_t vol = (size_t)(n[0] + 1) * (n[1] + 1) * (n[2] + 1);
  float finv[3];
  for (int i = 0; i < 3; i++) {
    GB_CHECK(n[i] >= 1 && end[i] > begin[i], "bad grid dims");
    v.begin[i] = begin[i]; v.end[i] = end[i]; v.gn[i] = n[i];
    const float factor = (float)((n[i] + 1) - 1.0) / (end[i] - begin[i]);
    finv[i] = 1 / factor;
  }
  // the kernel keeps up to kMaxNeeded running sums per grid point: a ligand with more heavy types (the reference's own
  // random-molecule tests draw all 26) takes another pass over the receptor
  for (int j0 = 0; j0 < n_types; j0 += kMaxNeeded) {
    NeededTypes nt;
    nt.n = std::min(kMaxNeeded, n_types - j0);
    for (int j = 0; j < nt.n; j++) {
      const int t = types_needed[j0 + j];
      GB_CHECK(t >= 2 && t < kNumSminaTypes, "needed type must be a heavy smina type");
      if (v.grid_pool_cap[t] < vol) {
        cudaFree(v.grid_pool[t]);
        v.grid_pool[t] = nullptr; v.grid_pool_cap[t] = 0;
        GB_CUDA(cudaMalloc(&v.grid_pool[t], vol * sizeof(float)));
        v.grid_pool_cap[t] = vol;
      }
      v.d_grids[t] = v.grid_pool[t];
      nt.t[j] = t;
      nt.grid[j] = v.d_grids[t];
    }
    cache_populate_kernel<<<(unsigned)((vol + 127) / 128), 128, 0, v.stream>>>(v.d_rec, v.n_rec, v.d_fast, v.n, v.factor,
                                                                                v.cutoff_sqr, begin[0], begin[1], begin[2], finv[0],
                                                                                finv[1], finv[2], n[0] + 1, n[1] + 1, n[2] + 1, nt);
    GB_CUDA(cudaGetLastError());
  }
  GB_CUDA(cudaStreamSynchronize(v.stream));
  GBV_END
}

int gb_vina_cache_read(gb_vina* h, int type, float* out) {
  GBV_BEGIN
  GB_CHECK(h && out && type >= 0 && type < kNumSminaTypes && h->v.d_grids[type], "grid not built for this type");
  const Vina& v = h->v;
  const size_t vol = (size_t)(v.gn[0] + 1) * (v.gn[1] + 1) * (v.gn[2] + 1);
  GB_CUDA(cudaMemcpy(out, v.d_grids[type], vol * sizeof(float), cudaMemcpyDeviceToHost));
  GBV_END
}

int gb_vina_cache_eval(gb_vina* h, const float* lig_xyz, const int32_t* lig_type, const int32_t* pose_offsets, int n_poses,
                       float slope, float vcap, float* energy, float* deriv) {
  GBV_BEGIN
  GB_CHECK(h && n_poses >= 0 && energy, "bad arguments");
  Vina& v = h->v;
  GB_CUDA(cudaSetDevice(v.device));
  if (n_poses == 0) return GB_OK;
  stage_poses(v, lig_xyz, lig_type, pose_offsets, n_poses);
  const int n_atoms = pose_offsets[n_poses];
  GridGeom G;
  GridPtrs gp;
  for (int i = 0; i < 3; i++) {
    G.dims[i] = v.gn[i] + 1;
    G.dm1[i] = (float)(G.dims[i] - 1.0);
    G.begin[i] = v.begin[i];
    G.factor[i] = G.dm1[i] / (v.end[i] - v.begin[i]);
    G.finv[i] = 1 / G.factor[i];
  }
  for (int t = 0; t < kNumSminaTypes; t++) gp.g[t] = v.d_grids[t];
  for (int i = 0; i < n_atoms; i++) {
    const int t = lig_type[i];
    if (t >= 2 && t < kNumSminaTypes && !v.d_grids[t]) throw Error(GB_ERR_USAGE, "cache has no grid for a ligand atom type");
  }
  if (n_atoms) cache_eval_kernel<<<(n_atoms + 127) / 128, 128, 0, v.stream>>>(v.d_lig, n_atoms, gp, G, slope, vcap, v.d_atom_e,
                                                                              deriv ? v.d_deriv : nullptr);
  pose_sum_kernel<<<(n_poses + 127) / 128, 128, 0, v.stream>>>(v.d_atom_e, v.d_off, n_poses, v.d_pose_e);
  GB_CUDA(cudaMemcpyAsync(energy, v.d_pose_e, (size_t)n_poses * sizeof(float), cudaMemcpyDeviceToHost, v.stream));
  if (deriv && n_atoms) GB_CUDA(cudaMemcpyAsync(deriv, v.d_deriv, (size_t)n_atoms * 3 * sizeof(float), cudaMemcpyDeviceToHost, v.stream));
  GB_CUDA(cudaStreamSynchronize(v.stream));
  GB_CUDA(cudaGetLastError());
  GBV_END
}

// poses -> per-pose intermolecular energy (ordered atom sum) -> num_tors_div; mode 0 = naive_non_cache with the exact terms
// (--score_only / --minimize, main/main.cpp:233-236,282-285), mode 1 = non_cache with the search's tables (docking, :340-344)
static void score_poses_common(Vina& v, const float* lig_xyz, const int32_t* lig_type, const int32_t* pose_offsets, int n_poses,
                               const float* num_tors, float vcap, int mode, float slope, const float* bb, const float* be, float* e_inter,
                               float* affinity);

int gb_vina_score_noncache(gb_vina* h, const float* lig_xyz, const int32_t* lig_type, const int32_t* pose_offsets, int n_poses,
                           const float* num_tors, float vcap, float slope, const float* box_begin, const float* box_end, float* e_inter,
                           float* affinity) {
  GBV_BEGIN
  GB_CHECK(h && n_poses >= 0 && box_begin && box_end, "bad arguments");
  Vina& v = h->v;
  GB_CUDA(cudaSetDevice(v.device));
  GB_CHECK(v.d_rec && v.n_rec > 0, "gb_vina_set_receptor has not been called");
  if (n_poses == 0) return GB_OK;
  score_poses_common(v, lig_xyz, lig_type, pose_offsets, n_poses, num_tors, vcap, 1, slope, box_begin, box_end, e_inter, affinity);
  GBV_END
}

int gb_vina_score_exact(gb_vina* h, const float* lig_xyz, const int32_t* lig_type, const int32_t* pose_offsets, int n_poses,
                        const float* num_tors, float vcap, float* e_inter, float* affinity) {
  GBV_BEGIN
  GB_CHECK(h && n_poses >= 0, "bad arguments");
  Vina& v = h->v;
  GB_CUDA(cudaSetDevice(v.device));
  if (n_poses == 0) return GB_OK;
  score_poses_common(v, lig_xyz, lig_type, pose_offsets, n_poses, num_tors, vcap, 0, 0.f, nullptr, nullptr, e_inter, affinity);
  GBV_END
}

}  // extern "C"

static void score_poses_common(Vina& v, const float* lig_xyz, const int32_t* lig_type, const int32_t* pose_offsets, int n_poses,
                               const float* num_tors, float vcap, int mode, float slope, const float* bb, const float* be, float* e_inter,
                               float* affinity) {
  stage_poses(v, lig_xyz, lig_type, pose_offsets, n_poses);
  const int n_atoms = pose_offsets[n_poses];
  if (n_atoms && mode == 0)
    naive_exact_kernel<<<(n_atoms + 127) / 128, 128, 0, v.stream>>>(v.d_lig, n_atoms, v.d_rec, v.n_rec, v.cutoff_sqr, vcap, v.d_atom_e);
  if (n_atoms && mode == 1)
    noncache_eval_kernel<<<(n_atoms + 127) / 128, 128, 0, v.stream>>>(v.d_lig, n_atoms, v.d_rec, v.n_rec, v.d_fast, v.n, v.factor,
                                                                     v.cutoff_sqr, vcap, slope, bb[0], bb[1], bb[2], be[0], be[1], be[2],
                                                                     v.d_atom_e);
  pose_sum_kernel<<<(n_poses + 127) / 128, 128, 0, v.stream>>>(v.d_atom_e, v.d_off, n_poses, v.d_pose_e);
  std::vector<float> e(n_poses);
  GB_CUDA(cudaMemcpyAsync(e.data(), v.d_pose_e, (size_t)n_poses * sizeof(float), cudaMemcpyDeviceToHost, v.stream));
  GB_CUDA(cudaStreamSynchronize(v.stream));
  GB_CUDA(cudaGetLastError());
  for (int p = 0; p < n_poses; p++) {
    if (e_inter) e_inter[p] = e[p];
    if (affinity) {
      // num_tors_div (everything.h:804-809) with smooth_div (:52-56)
      const float w = (float)(0.1 * ((double)v.w[5] + 1));
      // "1 + w * in.num_tors / 5.0" with fl w, fl num_tors: the product is a FLOAT product, then double arithmetic (pinned by oracle/_ref)
      const float wnt = w * (num_tors ? num_tors[p] : 0.f);
      const float y = (float)(1 + (double)wnt / 5.0);
      const float x = e[p];
      const float eps = 1.1920929e-07f, maxfl = 3.402823466e+38f;
      affinity[p] = std::fabs(x) < eps ? 0.f : (std::fabs(y) < eps ? ((x * y > 0) ? maxfl : -maxfl) : x / y);
    }
  }
}

// =================================================================================================================
// Docking inner loop as a batched kernel: ONE WARP PER CHAIN / CONFORMATION.
//   V7  heterotree::set_conf (lib/tree.h:218-233,361-366): segment frames level by level (lane = segment), then atoms
//   V5  cache::eval_deriv per atom (lane-strided), V6 eval_interacting_pairs_deriv (lib/model.cu:38-60) with the
//       precalculate_linear (e, dor) tables, forces accumulated in shared memory
//   V8  heterotree::derivative (lib/tree.h:300-310,374-382): per-segment force/torque (lane = segment), children
//       folded into parents from the deepest level up
//   V9  bfgs + fast_line_search + bfgs_update (lib/bfgs.h:52-91,358-502) with the packed upper-triangular Hessian in
//       shared memory, conf::increment (lib/conf.h:54-59,113-118)
//   V10 monte_carlo::operator() (lib/monte_carlo.cpp:99-148), mutate_conf (lib/mutate.cpp:35-73), metropolis_accept,
//       add_to_output_container (lib/coords.cpp:25-56); V11: chains are independent -> one launch for all of them
//       (the reference farms them to a boost thread pool, lib/parallel_mc.cpp:183-214)
// Random numbers: xorshift32 exactly as oracle/vina_mc_ref.c (Boost's distributions are not reproducible here).
// Every sum follows the reference's sequential float association (atom energies in atom order, pair energies in pair order, forces
// per atom in pair-list order, children folded into parents in ascending order; no FMA contraction in this file), so single
// evaluations agree with the restatement to the last bits of the transcendentals and BFGS / Monte-Carlo trajectories are compared step
// by step; the restatement itself is bit-identical to the reference's own code compiled in oracle/_ref.
// =================================================================================================================
namespace gb {

constexpr int kDkMaxAtoms = 96, kDkMaxSeg = 32, kDkMaxN = 6 + kDkMaxSeg - 1, kDkWarps = 4;

struct LigPtrs {
  int n_atoms, n_seg, n_pairs, max_depth, n_heavy;
  float gyration_radius;
  const float4* local; const int* atom_seg; const int4* seg; const float4* rel_origin; const float4* rel_axis; const int2* pairs;
  const int* adj_off; const int* adj;      // adj: partner | pair index << 8 | first-of-pair << 30
  const int* child_off; const int* child;  // children of every segment (CSR, ascending)
};
// rec != nullptr selects non_cache (lib/non_cache.cpp): direct sums over the receptor's heavy atoms instead of the
// affinity grids, with the box [nc_begin, nc_end] for the out-of-box clamp and penalty
struct DockField {
  GridGeom G; GridPtrs gp; const float2* smooth; int n_samples; float factor, slope; const float4* sp; int n_sp; float sp_fraction;
  const float4* rec; int n_rec; float nc_begin[3], nc_end[3];
};

// Per-warp shared-memory workspace, carved out of dynamic shared memory and sized by the ACTUAL ligand (atoms,
// segments) instead of the maxima: a typical ligand (27 atoms, 7 segments) needs 2.3 KB per warp instead of 10 KB, so
// shared memory no longer caps the resident chains per SM (ncu r1l: the chain is latency-bound on dependent
// shared-memory operations; resident warps are what hides it).
struct WarpWs {
  float *coords, *forces;              // [3 na]
  float *so, *sa, *sq, *sm, *ft;       // [3 ns] [3 ns] [4 ns] [9 ns] [6 ns]
  float *x, *x_new, *x_orig;           // [n + 2]  (7 + T)
  float *g, *g_new, *g_orig, *p, *y, *mhy;  // [n]  (6 + T)
  float *h;                            // [n (n + 1) / 2]
  float *cand, *tmp;                   // [n + 2]
  float *ea, *pe;                      // [na] per-atom grid energies, [np] per-pair energies (summed in index order)
};
__host__ __device__ inline int dk_ws_floats(int na, int ns, int np) {
  const int n = 6 + ns - 1;
  const int n4 = (n + 3) & ~3;
  int f = 2 * ((3 * na + 3) & ~3) + 25 * ns + 3 * (n + 2) + 6 * n4 + n * (n + 1) / 2 + 2 * (n + 2) + ((na + 3) & ~3) + ((np + 3) & ~3);
  return (f + 3) & ~3;  // 16-byte multiple
}
__device__ inline void dk_ws_carve(WarpWs& W, float* base, int na, int ns, int np) {
  const int n = 6 + ns - 1;
  float* p = base;
  auto take = [&](int k) { float* r = p; p += k; return r; };
  W.ea = take((na + 3) & ~3); W.pe = take((np + 3) & ~3);   // first: 16-byte aligned for the vector loads of dk_sum_seq
  W.coords = take((3 * na + 3) & ~3); W.forces = take((3 * na + 3) & ~3);
  const int n4 = (n + 3) & ~3;
  W.g = take(n4); W.g_new = take(n4); W.g_orig = take(n4); W.p = take(n4); W.y = take(n4); W.mhy = take(n4);
  W.so = take(3 * ns); W.sa = take(3 * ns); W.sq = take(4 * ns); W.sm = take(9 * ns); W.ft = take(6 * ns);
  W.x = take(n + 2); W.x_new = take(n + 2); W.x_orig = take(n + 2);
  W.h = take(n * (n + 1) / 2);
  W.cand = take(n + 2); W.tmp = take(n + 2);
}
// sin / cos of a float argument, correctly rounded (evaluated in double): what glibc's sinf / cosf return in all but
// ~1e-4 of the cases, so that the torsion-tree kinematics reproduce the CPU restatement bit for bit.  B200 has a full-rate
// FP64 pipe; two calls per torsion and evaluation are noise next to the pair-table lookups.
__device__ inline void dk_sincos(float a, float* s, float* c) {
  double sd, cd;
  sincos((double)a, &sd, &cd);
  *s = (float)sd; *c = (float)cd;
}
// sum_{k<n} a[k] (or a[k] b[k]) in index order, every lane computing the same chain: the reference's sequential
// float association (lib/model.cu:38-60, lib/bfgs.h), not a butterfly -- a different order changes line-search decisions
// a is 16-byte aligned (W.ea, W.pe): one 128-bit shared load per four additions
__device__ inline float dk_sum_seq(const float* a, int n) {
  float s = 0.f;
  const float4* a4 = reinterpret_cast<const float4*>(a);
  const int n16 = n >> 4;
  // 16 elements per trip: four 128-bit loads issued together, then the sixteen dependent additions (r2s: with one quad per
  // trip the loop bookkeeping was 40 % of this function, and this function 20 % of the chain kernel's instructions)
#pragma unroll 1
  for (int k = 0; k < n16; k++) {
    const float4 v0 = a4[4 * k], v1 = a4[4 * k + 1], v2 = a4[4 * k + 2], v3 = a4[4 * k + 3];
    s += v0.x; s += v0.y; s += v0.z; s += v0.w;
    s += v1.x; s += v1.y; s += v1.z; s += v1.w;
    s += v2.x; s += v2.y; s += v2.z; s += v2.w;
    s += v3.x; s += v3.y; s += v3.z; s += v3.w;
  }
#pragma unroll 1
  for (int k = n16 << 4; k < n; k++) s += a[k];
  return s;
}
// a, b 16-byte aligned (the BFGS vectors of the workspace)
__device__ inline float dk_dot_seq(const float* a, const float* b, int n) {
  float s = 0.f;
  const float4* a4 = reinterpret_cast<const float4*>(a);
  const float4* b4 = reinterpret_cast<const float4*>(b);
  const int n4 = n >> 2;
#pragma unroll 1
  for (int k = 0; k < n4; k++) {
    const float4 u = a4[k], v = b4[k];
    s += u.x * v.x; s += u.y * v.y; s += u.z * v.z; s += u.w * v.w;
  }
#pragma unroll 1
  for (int k = n4 << 2; k < n; k++) s += a[k] * b[k];
  return s;
}

__device__ __forceinline__ float warp_sum(float v) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ inline void dk_normalize_angle(float& x) {
  const float pi = 3.14159265358979323846f;
  if (x > 3 * pi) { x -= 2 * pi * ceilf((x - pi) / (2 * pi)); }
  else if (x < -3 * pi) { x += 2 * pi * ceilf((-x - pi) / (2 * pi)); }
  if (x > pi) x -= 2 * pi;
  else if (x < -pi) x += 2 * pi;
}
__device__ inline void dk_angle_to_q(const float* axis, float angle, float* q) {
  dk_normalize_angle(angle);
  float c, s;
  dk_sincos(angle / 2, &s, &c);
  q[0] = c; q[1] = s * axis[0]; q[2] = s * axis[1]; q[3] = s * axis[2];
}
__device__ inline void dk_qmul(const float* l, const float* r, float* o) {
  const float a = l[0], b = l[1], c = l[2], d = l[3];
  o[0] = +a * r[0] - b * r[1] - c * r[2] - d * r[3];
  o[1] = +a * r[1] + b * r[0] + c * r[3] - d * r[2];
  o[2] = +a * r[2] - b * r[3] + c * r[0] + d * r[1];
  o[3] = +a * r[3] + b * r[2] - c * r[1] + d * r[0];
}
__device__ inline void dk_qnorm_approx(float* q) {
  const float s = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  if (fabsf(s - 1) < 1e-6f) return;
  const float inv = 1 / sqrtf(s);
  for (int i = 0; i < 4; i++) q[i] *= inv;
}
__device__ inline void dk_q_to_r3(const float* q, float* m) {
  const float a = q[0], b = q[1], c = q[2], d = q[3];
  const float aa = a * a, ab = a * b, ac = a * c, ad = a * d, bb = b * b, bc = b * c, bd = b * d, cc = c * c, cd = c * d, dd = d * d;
  m[0] = (aa + bb - cc - dd); m[1] = 2 * (-ad + bc); m[2] = 2 * (ac + bd);
  m[3] = 2 * (ad + bc); m[4] = (aa - bb + cc - dd); m[5] = 2 * (-ab + cd);
  m[6] = 2 * (-ac + bd); m[7] = 2 * (ab + cd); m[8] = (aa - bb - cc + dd);
}
__device__ inline void dk_mv(const float* m, float vx, float vy, float vz, float* o) {
  o[0] = m[0] * vx + m[1] * vy + m[2] * vz;
  o[1] = m[3] * vx + m[4] * vy + m[5] * vz;
  o[2] = m[6] * vx + m[7] * vy + m[8] * vz;
}
__device__ inline void dk_quaternion_increment(float* q, const float* rot) {
  const float angle = sqrtf(rot[0] * rot[0] + rot[1] * rot[1] + rot[2] * rot[2]);
  float r[4] = {1, 0, 0, 0};
  if (angle > 1.1920929e-07f) {
    const float axis[3] = {(1 / angle) * rot[0], (1 / angle) * rot[1], (1 / angle) * rot[2]};
    dk_angle_to_q(axis, angle, r);
  }
  float o[4];
  dk_qmul(r, q, o);
  for (int i = 0; i < 4; i++) q[i] = o[i];
  dk_qnorm_approx(q);
}

// V7: conf (W.x-like array xc) -> segment frames + atom coordinates in W.coords
__device__ void dk_set_conf(const LigPtrs& L, WarpWs& W, const float* xc, int lane) {
  // lane s owns segment s.  The half-angle sine / cosine of every torsion depends only on the conformation, not on the
  // parent frames: all lanes compute theirs at once here, instead of one lane per tree level inside the loop below
  // (ncu r1l: cosf / sinf ran with 1.2 active lanes and were 8 % of the chain's instructions).
  float tors_c = 1.f, tors_s = 0.f;
  if (lane >= 1 && lane < L.n_seg) {
    float ang = xc[7 + lane - 1];
    dk_normalize_angle(ang);
    dk_sincos(ang / 2, &tors_s, &tors_c);
  }
  #pragma unroll 1
  for (int d = 0; d <= L.max_depth; d++) {
    if (lane < L.n_seg) {
      const int4 sg = L.seg[lane];
      if (sg.w == d) {
        float* o = W.so + 3 * lane;
        float* q = W.sq + 4 * lane;
        if (lane == 0) {
          o[0] = xc[0]; o[1] = xc[1]; o[2] = xc[2];
          q[0] = xc[3]; q[1] = xc[4]; q[2] = xc[5]; q[3] = xc[6];
          W.sa[0] = W.sa[1] = W.sa[2] = 0.f;
        } else {
          const int pp = sg.x;
          const float4 ro = L.rel_origin[lane], ra = L.rel_axis[lane];
          float t[3];
          dk_mv(W.sm + 9 * pp, ro.x, ro.y, ro.z, t);
          for (int k = 0; k < 3; k++) o[k] = W.so[3 * pp + k] + t[k];
          dk_mv(W.sm + 9 * pp, ra.x, ra.y, ra.z, W.sa + 3 * lane);
          const float* ax = W.sa + 3 * lane;
          const float aq[4] = {tors_c, tors_s * ax[0], tors_s * ax[1], tors_s * ax[2]};  // angle_to_quaternion
          dk_qmul(aq, W.sq + 4 * pp, q);
          dk_qnorm_approx(q);
        }
        dk_q_to_r3(q, W.sm + 9 * lane);
      }
    }
    __syncwarp();
  }
  #pragma unroll 1
  for (int i = lane; i < L.n_atoms; i += 32) {
    const int sgi = L.atom_seg[i];
    const float4 a = L.local[i];
    float t[3];
    dk_mv(W.sm + 9 * sgi, a.x, a.y, a.z, t);
    for (int k = 0; k < 3; k++) W.coords[3 * i + k] = W.so[3 * sgi + k] + t[k];
  }
  __syncwarp();
}

// precalculate::eval_deriv of one type pair at squared distance r2 -> (e, dE/dr / r): the linear tables
// (lib/precalculate.h:97-133) or, when selected, the splines (:380-449)
__device__ __forceinline__ void dk_pair_terms(const DockField& F, int t1, int t2, float r2, float& pe, float& dor) {
  if (t1 > t2) { const int tt = t1; t1 = t2; t2 = tt; }
  if (F.sp) {  // precalculate_splines::eval_deriv (a pair whose knots are all zero has all-zero coefficients)
    const float r = sqrtf(r2);
    int idx = (int)(r / F.sp_fraction);
    if (idx >= F.n_sp) idx = F.n_sp - 1;
    const float4 c = F.sp[(size_t)(t1 + t2 * (t2 + 1) / 2) * F.n_sp + idx];
    const float lx = r - idx * F.sp_fraction;
    pe = ((c.x * lx + c.y) * lx + c.z) * lx + c.w;
    dor = ((3 * c.x * lx + 2 * c.y) * lx + c.z) / r;
  } else {
    const float r2f = F.factor * r2;
    const int i1 = (int)r2f;
    const float rem = r2f - i1;
    const float2* tb = F.smooth + (size_t)(t1 + t2 * (t2 + 1) / 2) * F.n_samples;
    const float2 s1 = tb[i1], s2 = tb[i1 + 1];
    pe = s1.x + rem * (s2.x - s1.x);
    dor = s1.y + rem * (s2.y - s1.y);
  }
}

// non_cache::eval / eval_deriv for one movable heavy atom (lib/non_cache.cpp:52-81,126-174): clamp to the box
// (check_bounds_deriv :102-123, penalty = slope x L1 distance), sum e and dor r over the receptor atoms with r^2 < 64 in
// index order (the reference's szv_grid only pre-selects candidates, in ascending order), curl, out-of-box derivative
__device__ float dk_noncache_atom(const DockField& F, int t1, float ax, float ay, float az, float v, float* deriv) {
  const float a[3] = {ax, ay, az};
  float adj[3], oob[3] = {0.f, 0.f, 0.f}, pen = 0.f;
#pragma unroll
  for (int j = 0; j < 3; j++) {
    adj[j] = a[j];
    if (a[j] < F.nc_begin[j]) { adj[j] = F.nc_begin[j]; oob[j] = -1.f; pen += fabsf(a[j] - F.nc_begin[j]); }
    else if (a[j] > F.nc_end[j]) { adj[j] = F.nc_end[j]; oob[j] = 1.f; pen += fabsf(a[j] - F.nc_end[j]); }
  }
  pen *= F.slope;
  float e = 0.f, d0 = 0.f, d1 = 0.f, d2 = 0.f;
  #pragma unroll 1
  for (int b = 0; b < F.n_rec; b++) {
    const float4 rb = F.rec[b];
    const float r0 = adj[0] - rb.x, r1 = adj[1] - rb.y, r2c = adj[2] - rb.z;
    const float r2 = r0 * r0 + r1 * r1 + r2c * r2c;
    if (r2 < 64.f) {
      float pe, dor;
      dk_pair_terms(F, t1, (int)rb.w, r2, pe, dor);
      e += pe;
      d0 += dor * r0; d1 += dor * r1; d2 += dor * r2c;
    }
  }
  if (e > 0 && v < 0.1f * 3.402823466e+38f) {  // curl (lib/curl.h:30-42)
    const float tmp = (v < 1.1920929e-07f) ? 0.f : (v / (v + e));
    e *= tmp;
    d0 *= tmp * tmp; d1 *= tmp * tmp; d2 *= tmp * tmp;
  }
  if (deriv) { deriv[0] = d0 + F.slope * oob[0]; deriv[1] = d1 + F.slope * oob[1]; deriv[2] = d2 + F.slope * oob[2]; }
  return e + pen;
}

// ig->eval / ig->eval_deriv for atom i (cache or non_cache); 0 and zero forces for hydrogens
__device__ __forceinline__ float dk_atom_field(const LigPtrs& L, const DockField& F, const WarpWs& W, int i, float v1, float* d) {
  const int t = (int)L.local[i].w;
  if (!(t >= 2 && t < kNumSminaTypes)) return 0.f;
  if (F.rec) return dk_noncache_atom(F, t, W.coords[3 * i], W.coords[3 * i + 1], W.coords[3 * i + 2], v1, d);
  return grid_evaluate_dev(F.gp.g[t], F.G, W.coords[3 * i], W.coords[3 * i + 1], W.coords[3 * i + 2], F.slope, v1, d);
}

// model::eval_deriv (lib/model.cu:202-225): returns e (all lanes), writes change[6+T] to gout.  Every sum follows the
// reference's association -- atom energies in atom order, pair energies in pair order, the forces on an atom in the
// order its pairs appear in the pair list, children folded into their parent in ascending order -- so that the result
// agrees with the sequential CPU code to round-off of the transcendental functions only (north_star: 1e-6).
// mode kEvGridHere: update_energy's cache::eval (lib/monte_carlo.cpp:44-47,113,135) -- the intermolecular energy alone, atoms
// summed in index order, ON THE COORDINATES THE WORKSPACE HOLDS (no set_conf: the reference's model object keeps the coordinates of
// the last conformation that was set, and igrid::eval reads those); nothing is written to gout.  mode kEvSetOnly: model::set(conf)
// alone.  Both share the kinematics and the per-atom field code with the full evaluation (ONE copy of this function per kernel:
// the chain kernel's instruction footprint is what bounds it, profiles/README.md r2o).
enum { kEvFull = 0, kEvGridHere = 1, kEvSetOnly = 2 };
__device__ float dk_eval_deriv(const LigPtrs& L, const DockField& F, WarpWs& W, const float* xc, const float* v, float* gout, int lane,
                               int mode = kEvFull) {
  if (mode != kEvGridHere) dk_set_conf(L, W, xc, lane);
  if (mode == kEvSetOnly) return 0.f;
  const bool grid_only = mode == kEvGridHere;
  #pragma unroll 1
  for (int i = lane; i < L.n_atoms; i += 32) {
    float d[3] = {0.f, 0.f, 0.f};
    W.ea[i] = dk_atom_field(L, F, W, i, v[1], d);
    W.forces[3 * i] = d[0]; W.forces[3 * i + 1] = d[1]; W.forces[3 * i + 2] = d[2];
  }
  __syncwarp();
  if (grid_only) {
    const float eg = dk_sum_seq(W.ea, L.n_atoms);
    __syncwarp();
    return eg;
  }
  // V6 intramolecular pairs, atom by atom: lane i walks atom i's partners (CSR over the pair list, both directions, in
  // pair order) and continues the force sum of ITS atom in registers, starting from the grid force -- exactly the
  // sequence of additions the reference performs on minus_forces[i].  Every pair is evaluated from both ends (2x the
  // table lookups, no shared-memory atomics: ncu r1l); the end that is the pair's first atom records the pair energy.
  #pragma unroll 1
  for (int i = lane; i < L.n_atoms; i += 32) {
    const float xi = W.coords[3 * i], yi = W.coords[3 * i + 1], zi = W.coords[3 * i + 2];
    const int ti = (int)L.local[i].w;
    float fxs = W.forces[3 * i], fys = W.forces[3 * i + 1], fzs = W.forces[3 * i + 2];
    const int q1 = L.adj_off[i + 1];
    #pragma unroll 1
    for (int q = L.adj_off[i]; q < q1; q++) {
      const int code = L.adj[q];
      const int j = code & 0xff, k = (code >> 8) & 0x3fffff;
      const bool first = (code >> 30) & 1;
      const float rx = W.coords[3 * j] - xi, ry = W.coords[3 * j + 1] - yi, rz = W.coords[3 * j + 2] - zi;
      const float r2 = rx * rx + ry * ry + rz * rz;
      float pe = 0.f;
      if (r2 < 64.f) {
        float dor;
        dk_pair_terms(F, ti, (int)L.local[j].w, r2, pe, dor);
        float fx = dor * rx, fy = dor * ry, fz = dor * rz;
        if (pe > 0 && v[0] < 0.1f * 3.402823466e+38f) {
          const float tmp = (v[0] < 1.1920929e-07f) ? 0.f : (v[0] / (v[0] + pe));
          pe *= tmp;
          fx *= tmp * tmp; fy *= tmp * tmp; fz *= tmp * tmp;
        }
        fxs -= fx; fys -= fy; fzs -= fz;  // pair (a, b), r = x_b - x_a: forces[a] -= dor r ; seen from b, r changes sign
      }
      if (first) W.pe[k] = pe;            // a pair beyond the cut-off adds nothing to the reference's running sum
    }
    W.forces[3 * i] = fxs; W.forces[3 * i + 1] = fys; W.forces[3 * i + 2] = fzs;
  }
  __syncwarp();
  float e = dk_sum_seq(W.ea, L.n_atoms);
  e += dk_sum_seq(W.pe, L.n_pairs);
  // V8: per-segment force / torque about the segment origin (sum_force_and_torque, lib/tree.h:133-140) ...
  if (lane < L.n_seg) {
    const int4 sg = L.seg[lane];
    float f[6] = {0, 0, 0, 0, 0, 0};
    #pragma unroll 1
    for (int i = sg.y; i < sg.z; i++) {
      const float rx = W.coords[3 * i] - W.so[3 * lane], ry = W.coords[3 * i + 1] - W.so[3 * lane + 1], rz = W.coords[3 * i + 2] - W.so[3 * lane + 2];
      const float qx = W.forces[3 * i], qy = W.forces[3 * i + 1], qz = W.forces[3 * i + 2];
      f[0] += qx; f[1] += qy; f[2] += qz;
      f[3] += ry * qz - rz * qy; f[4] += rz * qx - rx * qz; f[5] += rx * qy - ry * qx;
    }
    #pragma unroll 1
    for (int k = 0; k < 6; k++) W.ft[6 * lane + k] = f[k];
  }
  __syncwarp();
  // ... then every parent gathers its children in ascending order, deepest parents first (branches_derivative,
  // lib/tree.h:300-310): no atomics, the reference's order
  #pragma unroll 1
  for (int d = L.max_depth - 1; d >= 0; d--) {
    if (lane < L.n_seg && L.seg[lane].w == d) {
      float* ft = W.ft + 6 * lane;
      const int c1 = L.child_off[lane + 1];
      #pragma unroll 1
      for (int q = L.child_off[lane]; q < c1; q++) {
        const int ch = L.child[q];
        const float* c = W.ft + 6 * ch;
        const float rx = W.so[3 * ch] - W.so[3 * lane], ry = W.so[3 * ch + 1] - W.so[3 * lane + 1], rz = W.so[3 * ch + 2] - W.so[3 * lane + 2];
        ft[0] += c[0]; ft[1] += c[1]; ft[2] += c[2];
        ft[3] += (ry * c[2] - rz * c[1]) + c[3];
        ft[4] += (rz * c[0] - rx * c[2]) + c[4];
        ft[5] += (rx * c[1] - ry * c[0]) + c[5];
      }
    }
    __syncwarp();
  }
  if (lane == 0) for (int k = 0; k < 6; k++) gout[k] = W.ft[k];
  else if (lane < L.n_seg) gout[6 + lane - 1] = W.ft[6 * lane + 3] * W.sa[3 * lane] + W.ft[6 * lane + 4] * W.sa[3 * lane + 1] + W.ft[6 * lane + 5] * W.sa[3 * lane + 2];
  __syncwarp();
  return e;
}

__device__ inline int dk_tri(int i, int j) { return i <= j ? i + j * (j + 1) / 2 : j + i * (i + 1) / 2; }

__device__ void dk_conf_increment(float* x, const float* p, float f, int T, int lane) {
  if (lane == 0) {
    for (int k = 0; k < 3; k++) x[k] += f * p[k];
    const float rot[3] = {f * p[3], f * p[4], f * p[5]};
    dk_quaternion_increment(x + 3, rot);
  }
  #pragma unroll 1
  for (int t = lane; t < T; t += 32) {
    float a = f * p[6 + t];
    dk_normalize_angle(a);
    float nx = x[7 + t] + a;
    dk_normalize_angle(nx);
    x[7 + t] = nx;
  }
  __syncwarp();
}

// model::gyration_radius (lib/model.cpp:1002-1014) of the coordinates the workspace holds: heavy atoms about the root origin, summed
// in atom order (W.ea is free between evaluations; a hydrogen contributes an exact zero)
__device__ inline float dk_gyration_radius(const LigPtrs& L, WarpWs& W, int lane) {
  #pragma unroll 1
  for (int i = lane; i < L.n_atoms; i += 32) {
    float d2 = 0.f;
    if ((int)L.local[i].w >= 2) {
      const float a = W.coords[3 * i] - W.so[0], b = W.coords[3 * i + 1] - W.so[1], c = W.coords[3 * i + 2] - W.so[2];
      d2 = a * a + b * b + c * c;
    }
    W.ea[i] = d2;
  }
  __syncwarp();
  const float acc = dk_sum_seq(W.ea, L.n_atoms);
  __syncwarp();
  return L.n_heavy > 0 ? sqrtf(acc / (float)L.n_heavy) : 0.f;
}

// compute_lambdamin (lib/bfgs.h:93-102) with conf::operator()(i) (lib/conf.h:459-473: position, quaternion_to_angle(orientation),
// torsions; lib/quaternion.cu:46-62); acos / sin correctly rounded like every transcendental of this file.  All lanes, same value.
__device__ inline float dk_lambdamin(const float* x, const float* p, int n) {
  float ang[3] = {0.f, 0.f, 0.f};
  const float c = x[3];
  if (c > -1 && c < 1) {
    const float pi = 3.14159265358979323846f;
    float angle = 2 * (float)acos((double)c);
    if (angle > pi) angle -= 2 * pi;
    const float sn = (float)sin((double)(angle / 2));
    if (!(fabsf(sn) < 1.1920929e-07f)) {
      const float f = angle / sn;
      ang[0] = x[4] * f; ang[1] = x[5] * f; ang[2] = x[6] * f;
    }
  }
  float test = 0.f;
  #pragma unroll 1
  for (int i = 0; i < n; i++) {
    const float xi = i < 3 ? x[i] : (i < 6 ? ang[i - 3] : x[7 + i - 6]);
    const float ax = fabsf(xi);
    const float temp = fabsf(p[i]) / ((ax < 1.0f) ? 1.0f : ax);   // std::max(std::fabs(x(i)), 1.0f)
    if (temp > test) test = temp;
  }
  return test;
}

// bfgs (lib/bfgs.h:358-502); W.x in/out, W.g out; returns f0 (all lanes).
// Written around ONE evaluation site: the initial evaluation, every line-search trial and (grid_e != nullptr) the chain's
// update_energy after the minimisation are iterations of the same loop, so a kernel holds one copy of dk_eval_deriv (the
// arithmetic and its order are those of the reference).
// kMinimize = false: fast_line_search (:73-91), what the Monte-Carlo search uses -- the chain kernel's instantiation.
// kMinimize = true:  minimization_params at run time: `accurate` = accurate_line_search (:107-180, after Numerical Recipes' lnsrch; what
//                    --minimize selects, main/main.cpp:1160,1186), `early_term` = --minimize_early_term (:455-462).
template <bool kMinimize>
__device__ float dk_bfgs(const LigPtrs& L, const DockField& F, WarpWs& W, int maxiters, const float* v, int lane, int* n_evals,
                         float* grid_e = nullptr, float grid_v1 = 0.f, float* gr_state = nullptr, float* gr_final = nullptr,
                         bool accurate = false, bool early_term = false) {
  const int T = L.n_seg - 1, n = 6 + T, nx = 7 + T;
  #pragma unroll 1
  for (int k = lane; k < n * (n + 1) / 2; k += 32) W.h[k] = 0.f;
  __syncwarp();
  #pragma unroll 1
  for (int i = lane; i < n; i += 32) W.h[dk_tri(i, i)] = 1.f;
  int evals = 0, step = -1, trial = 0;   // step -1: the initial evaluation
  // finishing (chain kernel only, grid_e != nullptr) = what monte_carlo.cpp does around quasi_newton: 1 = update_energy, the grid
  // energy of the coordinates the LAST evaluation left in the model (bfgs.h does not re-evaluate at the x it returns: after ten
  // failed line-search trials, or when x_orig is restored, those are another conformation's); 2 = m.set(x) of the returned x
  float f0 = 0.f, f1 = 0.f, f_orig = 0.f, alpha = 1.f, pg = 0.f;
  float alpha2 = 0.f, f2 = 0.f, alamin = 0.f;   // accurate line search: previous trial, smallest step
  bool didreset = false;
  int finishing = 0;
  const float vg[3] = {grid_v1, grid_v1, grid_v1};
  #pragma unroll 1
  for (;;) {
    const bool init = step < 0;
    const float fe = dk_eval_deriv(L, F, W, (init || finishing) ? W.x : W.x_new, finishing ? vg : v, init ? W.g : W.g_new, lane,
                                   finishing == 1 ? kEvGridHere : finishing == 2 ? kEvSetOnly : kEvFull);
    if (finishing == 1) {
      *grid_e = fe;
      if (gr_state) *gr_state = dk_gyration_radius(L, W, lane);
      finishing = 2;
      continue;
    }
    if (finishing == 2) {
      if (gr_final) *gr_final = dk_gyration_radius(L, W, lane);
      break;
    }
    evals++;
    bool new_iter = false, done = false;
    if (init) {
      f0 = fe; f_orig = fe;
      #pragma unroll 1
      for (int i = lane; i < nx; i += 32) W.x_orig[i] = W.x[i];
      #pragma unroll 1
      for (int i = lane; i < n; i += 32) W.g_orig[i] = W.g[i];
      __syncwarp();
      new_iter = true;
    } else {
      f1 = fe;
      bool ls_over;
      if (kMinimize && accurate) {
        // accurate_line_search: fl = float, the literals 2.0 / 3.0 / .5 are double (pg is the slope g.p)
        if (alpha < alamin || !isfinite(alpha)) { alpha = 0.f; ls_over = true; }          // too small a step
        else if (f1 <= f0 + 1.0e-4f * alpha * pg) ls_over = true;                          // sufficient decrease
        else {
          float tmplam;
          if (alpha == 1.0f) tmplam = (float)(-pg / (2.0 * (f1 - f0 - pg)));
          else {
            const float rhs1 = f1 - f0 - alpha * pg, rhs2 = f2 - f0 - alpha2 * pg;
            const float a = (rhs1 / (alpha * alpha) - rhs2 / (alpha2 * alpha2)) / (alpha - alpha2);
            const float b = (-alpha2 * rhs1 / (alpha * alpha) + alpha * rhs2 / (alpha2 * alpha2)) / (alpha - alpha2);
            if (a == 0.0f) tmplam = (float)(-pg / (2.0 * b));
            else {
              const float disc = (float)(b * b - 3.0 * a * pg);
              if (disc < 0) tmplam = (float)(0.5 * alpha);
              else if (b <= 0) tmplam = (float)((-b + sqrtf(disc)) / (3.0 * a));
              else tmplam = -pg / (b + sqrtf(disc));
            }
            if (tmplam > .5 * alpha) tmplam = (float)(.5 * alpha);
          }
          alpha2 = alpha; f2 = f1;
          { const float tenth = 0.1f * alpha; alpha = (tmplam < tenth) ? tenth : tmplam; }   // std::max: a NaN tmplam stays NaN
          ls_over = false;
        }
      } else {
        const bool accepted = f1 - f0 < 0.0001f * alpha * pg;
        if (!accepted) { alpha *= 0.5f; trial++; }
        // over after ten trials too (alpha has been halved after the tenth failure as well, as the reference's loop does)
        ls_over = accepted || trial >= 10;
      }
      if (ls_over) {
        if (alpha == 0.f) done = true;
        else {
          #pragma unroll 1
          for (int i = lane; i < n; i += 32) W.y[i] = W.g_new[i] - W.g[i];
          const float prevf0 = f0;
          f0 = f1;
          #pragma unroll 1
          for (int i = lane; i < nx; i += 32) W.x[i] = W.x_new[i];
          __syncwarp();
          // --minimize_early_term: stop on a small decrease, BEFORE g is replaced (bfgs.h:455-464)
          const bool stop_early = kMinimize && early_term && fabs((double)(prevf0 - f0)) < 1e-5;
          if (!stop_early) {
            #pragma unroll 1
            for (int i = lane; i < n; i += 32) W.g[i] = W.g_new[i];
          }
          __syncwarp();
          const float gn = dk_dot_seq(W.g, W.g, n), yy = dk_dot_seq(W.y, W.y, n), yp = dk_dot_seq(W.y, W.p, n);
          if (stop_early || !(gn >= 1e-4f)) done = true;
          else {
            if (step == 0 || didreset) {
              didreset = false;
              if (fabsf(yy) > 1.1920929e-07f) {
                #pragma unroll 1
                for (int i = lane; i < n; i += 32) W.h[dk_tri(i, i)] = alpha * yp / yy;
              }
              __syncwarp();
            }
            if (!(alpha * yp < 1.1920929e-07f)) {  // bfgs_update
              #pragma unroll 1
              for (int i = lane; i < n; i += 32) {
                float s = 0.f;
                #pragma unroll 1
                for (int j2 = 0; j2 < n; j2++) s += W.h[dk_tri(i, j2)] * W.y[j2];
                W.mhy[i] = -s;
              }
              __syncwarp();
              const float yhy = -dk_dot_seq(W.y, W.mhy, n);
              const float r = 1 / (alpha * yp);
              #pragma unroll 1
              for (int k = lane; k < n * (n + 1) / 2; k += 32) {
                // invert k = i + j(j+1)/2, i <= j
                int j2 = (int)((sqrtf(8.f * k + 1.f) - 1.f) * 0.5f);
                while (j2 * (j2 + 1) / 2 > k) j2--;
                while ((j2 + 1) * (j2 + 2) / 2 <= k) j2++;
                const int i = k - j2 * (j2 + 1) / 2;
                W.h[k] += alpha * r * (W.mhy[i] * W.p[j2] + W.mhy[j2] * W.p[i]) + alpha * alpha * (r * r * yhy + r) * W.p[i] * W.p[j2];
              }
              __syncwarp();
            }
            new_iter = true;
          }
        }
      }
    }
    if (new_iter) {
      step++;
      if (step >= maxiters) done = true;
      else {
        #pragma unroll 1
        for (int i = lane; i < n; i += 32) {
          float s = 0.f;
          #pragma unroll 1
          for (int j2 = 0; j2 < n; j2++) s += W.h[dk_tri(i, j2)] * W.g[j2];
          W.p[i] = -s;
        }
        __syncwarp();
        pg = dk_dot_seq(W.p, W.g, n);
        alpha = 1.f; trial = 0;
        if (kMinimize && accurate) {
          alpha2 = 0.f; f2 = 0.f;
          if (pg >= 0) done = true;   // not a descent direction: accurate_line_search returns 0 without evaluating, bfgs gives up
          else alamin = 1.1920929e-07f / dk_lambdamin(W.x, W.p, n);
        }
      }
    }
    if (done) {
      if (!(f0 <= f_orig)) {
        f0 = f_orig;
        #pragma unroll 1
        for (int i = lane; i < nx; i += 32) W.x[i] = W.x_orig[i];
        #pragma unroll 1
        for (int i = lane; i < n; i += 32) W.g[i] = W.g_orig[i];
        __syncwarp();
      }
      if (!grid_e) break;
      finishing = 1;
      continue;
    }
    // next trial point: x_new = x (+) alpha p
    #pragma unroll 1
    for (int i = lane; i < nx; i += 32) W.x_new[i] = W.x[i];
    __syncwarp();
    dk_conf_increment(W.x_new, W.p, alpha, T, lane);
  }
  if (n_evals) *n_evals = evals;
  return f0;
}

// xorshift32 as in oracle/vina_mc_ref.c; the state lives in lane 0's register and results are broadcast
__device__ inline uint32_t dk_rng_next(uint32_t& s) { uint32_t x = s; x ^= x << 13; x ^= x >> 17; x ^= x << 5; s = x; return x; }
__device__ inline float dk_rng_fl(uint32_t& s, float a, float b) { return a + (b - a) * ((float)(dk_rng_next(s) >> 8) * (1.0f / 16777216.0f)); }
__device__ inline void dk_rng_sphere(uint32_t& s, float* o) {
  for (;;) {
    o[0] = dk_rng_fl(s, -1, 1); o[1] = dk_rng_fl(s, -1, 1); o[2] = dk_rng_fl(s, -1, 1);
    if (o[0] * o[0] + o[1] * o[1] + o[2] * o[2] < 1) return;
  }
}

__global__ void __launch_bounds__(32 * kDkWarps) dock_eval_kernel(LigPtrs L, DockField F, const float* __restrict__ confs, int n,
                                                                  float v0, float v1, float v2, float* __restrict__ e_out,
                                                                  float* __restrict__ change_out, float* __restrict__ coords_out,
                                                                  int mode, int maxiters, float* __restrict__ confs_out,
                                                                  int* __restrict__ evals_out, int accurate, int early_term) {
  extern __shared__ __align__(16) float dk_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * kDkWarps + warp;
  if (c >= n) return;
  WarpWs W;
  dk_ws_carve(W, dk_smem + (size_t)warp * dk_ws_floats(L.n_atoms, L.n_seg, L.n_pairs), L.n_atoms, L.n_seg, L.n_pairs);
  const int T = L.n_seg - 1, nx = 7 + T, ng = 6 + T;
  const float v[3] = {v0, v1, v2};
  for (int i = lane; i < nx; i += 32) W.x[i] = confs[(size_t)c * nx + i];
  __syncwarp();
  if (mode == 0) {
    const float e = dk_eval_deriv(L, F, W, W.x, v, W.g, lane);
    if (lane == 0) e_out[c] = e;
    for (int i = lane; i < ng; i += 32) change_out[(size_t)c * ng + i] = W.g[i];
    if (coords_out)
      for (int i = lane; i < 3 * L.n_atoms; i += 32) coords_out[(size_t)c * 3 * L.n_atoms + i] = W.coords[i];
  } else {
    int ne = 0;
    const float e = dk_bfgs<true>(L, F, W, maxiters, v, lane, &ne, nullptr, 0.f, nullptr, nullptr, accurate != 0, early_term != 0);
    if (lane == 0) { e_out[c] = e; if (evals_out) evals_out[c] = ne; }
    for (int i = lane; i < nx; i += 32) confs_out[(size_t)c * nx + i] = W.x[i];
    if (change_out)
      for (int i = lane; i < ng; i += 32) change_out[(size_t)c * ng + i] = W.g[i];
  }
}

// refine_structure (main/main.cpp:131-171): up to 5 quasi-Newton runs on the non_cache field with the out-of-box slope
// 10, 100, ... until every heavy atom is inside the box (non_cache::within, margin 1e-4); e = max_fl if it never is
__global__ void __launch_bounds__(32 * kDkWarps) dock_refine_kernel(LigPtrs L, DockField F0, float* __restrict__ confs, int n, float v0,
                                                                    float v1, float v2, int maxiters, float* __restrict__ e_out,
                                                                    int* __restrict__ within_out, int* __restrict__ evals_out,
                                                                    int accurate, int early_term) {
  extern __shared__ __align__(16) float dk_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * kDkWarps + warp;
  if (c >= n) return;
  WarpWs W;
  dk_ws_carve(W, dk_smem + (size_t)warp * dk_ws_floats(L.n_atoms, L.n_seg, L.n_pairs), L.n_atoms, L.n_seg, L.n_pairs);
  const int T = L.n_seg - 1, nx = 7 + T;
  const float v[3] = {v0, v1, v2};
  for (int i = lane; i < nx; i += 32) W.x[i] = confs[(size_t)c * nx + i];
  __syncwarp();
  DockField F = F0;
  float slope = 10.f, e = 0.f;
  int evals = 0, ok = 0;
  for (int p = 0; p < 5; p++) {
    F.slope = slope;
    int ne = 0;
    e = dk_bfgs<true>(L, F, W, maxiters, v, lane, &ne, nullptr, 0.f, nullptr, nullptr, accurate != 0, early_term != 0);
    evals += ne;
    dk_set_conf(L, W, W.x, lane);  // m.set(out.c)
    int inside = 1;
    for (int i = lane; i < L.n_atoms; i += 32) {
      if ((int)L.local[i].w < 2) continue;
      for (int j = 0; j < 3; j++)
        if (W.coords[3 * i + j] < F.nc_begin[j] - 0.0001f || W.coords[3 * i + j] > F.nc_end[j] + 0.0001f) inside = 0;
    }
    ok = __all_sync(0xffffffffu, inside);
    if (ok) break;
    slope *= 10.f;
  }
  if (lane == 0) { e_out[c] = ok ? e : 3.402823466e+38f; within_out[c] = ok; evals_out[c] = evals; }
  for (int i = lane; i < nx; i += 32) confs[(size_t)c * nx + i] = W.x[i];
}

// the empirical term of non_cache_cnn::eval_deriv (lib/non_cache_cnn.cpp:113-140) for free atoms: one thread per atom
__global__ void noncache_atoms_kernel(const float4* __restrict__ atoms, int n, DockField F, float v, float* __restrict__ e_out,
                                      float* __restrict__ d_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 a = atoms[i];
  const int t = (int)a.w;
  float d[3] = {0.f, 0.f, 0.f}, e = 0.f;
  if (t >= 2 && t < kNumSminaTypes) e = dk_noncache_atom(F, t, a.x, a.y, a.z, v, d);
  e_out[i] = e;
  d_out[3 * i] = d[0]; d_out[3 * i + 1] = d[1]; d_out[3 * i + 2] = d[2];
}

struct McDev { int num_steps, maxiters, num_saved_mins; float temperature, mutation_amplitude, min_rmsd; float hunt_cap[3]; };

// no min-blocks hint: capping the chain kernel at 64 registers measured 17 % slower (621 k vs 750 k MC steps/s)
__global__ void __launch_bounds__(32 * kDkWarps, 7) dock_mc_kernel(LigPtrs L, DockField F, McDev P, float c1x, float c1y, float c1z, float c2x,
                                                                float c2y, float c2z, const uint32_t* __restrict__ seeds, int n_chains,
                                                                float* __restrict__ out_e, float* __restrict__ out_conf,
                                                                float* __restrict__ out_heavy, int* __restrict__ n_out_arr,
                                                                float* __restrict__ trace) {
  extern __shared__ __align__(16) float dk_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * kDkWarps + warp;
  if (c >= n_chains) return;
  WarpWs W;
  dk_ws_carve(W, dk_smem + (size_t)warp * dk_ws_floats(L.n_atoms, L.n_seg, L.n_pairs), L.n_atoms, L.n_seg, L.n_pairs);
  const int T = L.n_seg - 1, nx = 7 + T, nh = L.n_heavy, S = P.num_saved_mins;
  float* oe = out_e + (size_t)c * S;
  float* oc = out_conf + (size_t)c * S * nx;
  float* oh = out_heavy + (size_t)c * S * 3 * nh;
  uint32_t rs = seeds[c] ? seeds[c] : 1u;  // only lane 0's copy advances; draws are broadcast
  const float av[3] = {1000.f, 1000.f, 1000.f};
  const float pi = 3.14159265358979323846f;
  // conf::randomize
  if (lane == 0) {
    const float c1[3] = {c1x, c1y, c1z}, c2[3] = {c2x, c2y, c2z};
    for (int k = 0; k < 3; k++) W.tmp[k] = dk_rng_fl(rs, c1[k], c2[k]);
    for (;;) {
      float q[4], s = 0;
      for (int k = 0; k < 4; k++) { q[k] = dk_rng_fl(rs, -1, 1); s += q[k] * q[k]; }
      if (s < 1 && s > 1e-3f) { const float inv = 1 / sqrtf(s); for (int k = 0; k < 4; k++) W.tmp[3 + k] = q[k] * inv; break; }
    }
    for (int t = 0; t < T; t++) W.tmp[7 + t] = dk_rng_fl(rs, -pi, pi);
  }
  __syncwarp();
  float tmp_e = 0.f, best_e = 3.402823466e+38f;
  int n_out = 0;
  // mutate_conf rotates by amplitude / model::gyration_radius, and the model object holds the coordinates of the LAST conformation that
  // was set (mutate.cpp:55, model.cpp:1002-1014): the input pose when the chain starts (L.gyration_radius, from the host), afterwards
  // whatever quasi_newton's last evaluation or monte_carlo's explicit m.set left there -- followed step by step below
  float gr = L.gyration_radius;
  for (int step = 0; step < P.num_steps; step++) {
    for (int i = lane; i < nx; i += 32) W.cand[i] = W.tmp[i];
    __syncwarp();
    if (lane == 0) {  // mutate_conf
      const int which = (int)(dk_rng_next(rs) % (uint32_t)(2 + T));
      float r[3];
      if (which == 0) { dk_rng_sphere(rs, r); for (int k = 0; k < 3; k++) W.cand[k] += P.mutation_amplitude * r[k]; }
      else if (which == 1) {
        if (gr > 1.1920929e-07f) {
          dk_rng_sphere(rs, r);
          const float rot[3] = {P.mutation_amplitude / gr * r[0], P.mutation_amplitude / gr * r[1], P.mutation_amplitude / gr * r[2]};
          dk_quaternion_increment(W.cand + 3, rot);
        }
      } else W.cand[7 + which - 2] = dk_rng_fl(rs, -pi, pi);
    }
    __syncwarp();
    // Two quasi-Newton runs per step at most -- the hunt with hunt_cap, then, for an accepted candidate that is promising,
    // the full-cap run (monte_carlo.cpp:111-137) -- through ONE dk_bfgs call site (pass 0 / pass 1), each followed by
    // update_energy (the grid energy of the coordinates the last evaluation left) and m.set of the returned conformation, both
    // folded into dk_bfgs.  A rejected candidate leaves the last evaluation's coordinates in the model (gr_state), an accepted one
    // is set explicitly (monte_carlo.cpp:126,134: gr_final).
    bool promising = false;
    #pragma unroll 1
    for (int pass = 0; pass < 2; pass++) {
      const float* src = pass == 0 ? W.cand : W.tmp;
      #pragma unroll 1
      for (int i = lane; i < nx; i += 32) W.x[i] = src[i];
      __syncwarp();
      float ge = 0.f, gr_state = 0.f, gr_final = 0.f;
      dk_bfgs<false>(L, F, W, P.maxiters, pass == 0 ? P.hunt_cap : av, lane, nullptr, &ge, av[1], &gr_state, &gr_final);
      gr = gr_state;
      if (pass == 0) {
        #pragma unroll 1
        for (int i = lane; i < nx; i += 32) W.cand[i] = W.x[i];
        __syncwarp();
        const float cand_e = ge;
        int accept = (step == 0 || cand_e < tmp_e) ? 1 : 0;
        if (!accept) {
          float u = 0.f;
          if (lane == 0) u = dk_rng_fl(rs, 0, 1);
          u = __shfl_sync(0xffffffffu, u, 0);
          accept = u < (float)exp((double)((tmp_e - cand_e) / P.temperature));  // correctly rounded, like glibc's expf
        }
        if (!accept) break;
        #pragma unroll 1
        for (int i = lane; i < nx; i += 32) W.tmp[i] = W.cand[i];
        __syncwarp();
        tmp_e = cand_e;
        gr = gr_final;   // m.set(tmp.c)
        if (!(tmp_e < best_e || n_out < S)) break;
      } else {
        #pragma unroll 1
        for (int i = lane; i < nx; i += 32) W.tmp[i] = W.x[i];
        __syncwarp();
        tmp_e = ge;
        gr = gr_final;   // m.set(tmp.c): W.coords hold the pose's coordinates for the container
        promising = true;
      }
    }
    if (promising) {
      {
        // add_to_output_container: rmsd of the heavy atoms to every kept pose
        int ci = n_out;
        float cr = 3.402823466e+38f;
        for (int o = 0; o < n_out; o++) {
          // squared coordinate differences lane-parallel into W.forces (free here), then summed in index order
          int hk = 0;
          for (int i = 0; i < L.n_atoms; i++) {  // heavy index = running count (uniform across lanes)
            if ((int)L.local[i].w >= 2) {
              if ((hk & 31) == lane) {
                const float* q = oh + ((size_t)o * nh + hk) * 3;
                const float dx = W.coords[3 * i] - q[0], dy = W.coords[3 * i + 1] - q[1], dz = W.coords[3 * i + 2] - q[2];
                W.forces[3 * hk] = dx * dx; W.forces[3 * hk + 1] = dy * dy; W.forces[3 * hk + 2] = dz * dz;
              }
              hk++;
            }
          }
          __syncwarp();
          const float acc = dk_sum_seq(W.forces, 3 * nh);
          __syncwarp();
          const float r = nh > 0 ? sqrtf(acc / nh) : 0.f;
          if (o == 0 || r < cr) { ci = o; cr = r; }
        }
        int slot = -1;
        if (ci < n_out && cr < P.min_rmsd) { if (tmp_e < oe[ci]) slot = ci; }
        else if (n_out < S) slot = n_out++;
        else if (n_out > 0 && tmp_e < oe[n_out - 1]) slot = n_out - 1;
        if (slot >= 0) {
          __syncwarp();
          if (lane == 0) oe[slot] = tmp_e;
          for (int i = lane; i < nx; i += 32) oc[(size_t)slot * nx + i] = W.tmp[i];
          int hk = 0;
          for (int i = 0; i < L.n_atoms; i++)
            if ((int)L.local[i].w >= 2) {
              if ((hk & 31) == lane) for (int k = 0; k < 3; k++) oh[((size_t)slot * nh + hk) * 3 + k] = W.coords[3 * i + k];
              hk++;
            }
          __syncwarp();
          // out.sort(): bubble the changed entry to its place (entries are otherwise sorted)
          for (int a = 1; a < n_out; a++)
            for (int b = a; b > 0; b--) {
              const float eb = oe[b], ea = oe[b - 1];
              if (!(eb < ea)) break;
              __syncwarp();
              if (lane == 0) { oe[b] = ea; oe[b - 1] = eb; }
              for (int i = lane; i < nx; i += 32) { const float t2 = oc[(size_t)b * nx + i]; oc[(size_t)b * nx + i] = oc[(size_t)(b - 1) * nx + i]; oc[(size_t)(b - 1) * nx + i] = t2; }
              for (int i = lane; i < 3 * nh; i += 32) { const float t2 = oh[(size_t)b * 3 * nh + i]; oh[(size_t)b * 3 * nh + i] = oh[(size_t)(b - 1) * 3 * nh + i]; oh[(size_t)(b - 1) * 3 * nh + i] = t2; }
              __syncwarp();
            }
        }
        if (tmp_e < best_e) best_e = tmp_e;
      }
    }
    if (trace && lane == 0) trace[(size_t)c * P.num_steps + step] = tmp_e;  // the chain's current energy after this step
  }
  if (lane == 0) n_out_arr[c] = n_out;
}

}  // namespace gb

static void make_field(const Vina& v, float slope, DockField& F) {
  for (int i = 0; i < 3; i++) {
    F.G.dims[i] = v.gn[i] + 1;
    F.G.dm1[i] = (float)(F.G.dims[i] - 1.0);
    F.G.begin[i] = v.begin[i];
    F.G.factor[i] = F.G.dm1[i] / (v.end[i] - v.begin[i]);
    F.G.finv[i] = 1 / F.G.factor[i];
  }
  for (int t = 0; t < kNumSminaTypes; t++) F.gp.g[t] = v.d_grids[t];
  F.smooth = v.d_smooth; F.n_samples = v.n; F.factor = v.factor; F.slope = slope;
  F.sp = v.use_splines ? v.d_sp : nullptr; F.n_sp = v.n_sp; F.sp_fraction = v.sp_fraction;
  F.rec = nullptr; F.n_rec = 0;
  for (int i = 0; i < 3; i++) { F.nc_begin[i] = v.begin[i]; F.nc_end[i] = v.end[i]; }
}
static LigPtrs lig_ptrs(const Vina& v) {
  const auto& l = v.lig;
  return LigPtrs{l.n_atoms, l.n_seg, l.n_pairs, l.max_depth, l.n_heavy, l.gyration_radius, l.local, l.atom_seg, l.seg, l.seg_rel_origin,
                 l.seg_rel_axis, l.pairs, l.adj_off, l.adj, l.child_off, l.child};
}
// every entry point that evaluates the affinity grids: a ligand atom type without a built grid would dereference a
// null device pointer, and the resulting illegal-address error is sticky for the whole process
static void check_dock_ready(const Vina& v, bool needs_grids = true) {
  GB_CHECK(v.lig.n_atoms > 0, "gb_vina_set_ligand has not been called");
  if (!needs_grids) return;
  GB_CHECK(v.gn[0] > 0, "gb_vina_cache_build has not been called");
  for (int t = 2; t < kNumSminaTypes; t++)
    if (((v.lig.heavy_types >> t) & 1u) && !v.d_grids[t])
      throw Error(GB_ERR_USAGE, std::string("ligand atom type ") + kSminaNames[t] + " has no affinity grid: pass it to gb_vina_cache_build");
}

extern "C" {

int gb_vina_set_precalc(gb_vina* h, int use_splines) {
  GBV_BEGIN
  GB_CHECK(h, "null argument");
  h->v.use_splines = use_splines ? 1 : 0;
  GBV_END
}
int gb_vina_spline_size(const gb_vina* h) { return h ? h->v.n_sp : 0; }
int gb_vina_spline_table(const gb_vina* h, int t1, int t2, float* abcd) {
  GBV_BEGIN
  GB_CHECK(h && abcd && t1 >= 0 && t2 >= 0 && t1 < kNumSminaTypes && t2 < kNumSminaTypes, "bad type");
  if (t1 > t2) std::swap(t1, t2);
  memcpy(abcd, &h->v.h_sp[(size_t)tri_index(t1, t2) * h->v.n_sp], sizeof(float4) * h->v.n_sp);
  GBV_END
}

int gb_vina_set_ligand(gb_vina* h, const gb_ligand_topology* t) {
  GBV_BEGIN
  GB_CHECK(h && t, "null argument");
  GB_CHECK(t->n_atoms > 0 && t->n_atoms <= kDkMaxAtoms, "ligand atom count out of range (max 96)");
  GB_CHECK(t->n_segments >= 1 && t->n_segments <= kDkMaxSeg, "segment count out of range (max 32)");
  Vina& v = h->v;
  GB_CUDA(cudaSetDevice(v.device));
  auto& l = v.lig;
  l.n_atoms = 0;  // not ready until the upload below succeeds
  const int na = t->n_atoms, ns = t->n_segments;
  GB_CHECK(t->n_pairs >= 0 && t->n_pairs <= kDkMaxAtoms * (kDkMaxAtoms - 1) / 2, "pair count out of range");
  std::vector<float4> local(na), ro(ns), ra(ns);
  std::vector<int> aseg(na, -1);
  std::vector<int4> seg(ns);
  int max_depth = 0, nh = 0;
  for (int s = 0; s < ns; s++) {
    const int par = t->seg_parent[s];
    GB_CHECK((s == 0 && par < 0) || (s > 0 && par >= 0 && par < s), "segments must be in DFS pre-order with parent < child");
    const int depth = s == 0 ? 0 : seg[par].w + 1;
    max_depth = std::max(max_depth, depth);
    seg[s] = make_int4(par, t->seg_atom_begin[s], t->seg_atom_end[s], depth);
    for (int i = t->seg_atom_begin[s]; i < t->seg_atom_end[s]; i++) { GB_CHECK(i >= 0 && i < na, "segment atom range"); aseg[i] = s; }
    ro[s] = make_float4(t->seg_rel_origin[3 * s], t->seg_rel_origin[3 * s + 1], t->seg_rel_origin[3 * s + 2], 0.f);
    ra[s] = make_float4(t->seg_rel_axis[3 * s], t->seg_rel_axis[3 * s + 1], t->seg_rel_axis[3 * s + 2], 0.f);
  }
  for (int i = 0; i < na; i++) {
    GB_CHECK(aseg[i] >= 0, "every atom must belong to a segment");
    local[i] = make_float4(t->local_xyz[3 * i], t->local_xyz[3 * i + 1], t->local_xyz[3 * i + 2], (float)t->smina_type[i]);
    nh += t->smina_type[i] >= 2;
  }
  std::vector<int2> pairs(std::max(t->n_pairs, 1));
  for (int k = 0; k < t->n_pairs; k++) pairs[k] = make_int2(t->pair_a[k], t->pair_b[k]);
  // device arrays are allocated once at their maximum size (96 atoms, 32 segments, 4560 pairs: < 60 KB) and reused
  // from ligand to ligand: no cudaFree (device-wide synchronisation) between the ligands of a screen
  auto up = [&](auto** d, const auto& hv, size_t max_elems) {
    if (!*d) GB_CUDA(cudaMalloc(d, max_elems * sizeof(hv[0])));
    GB_CUDA(cudaMemcpyAsync(*d, hv.data(), hv.size() * sizeof(hv[0]), cudaMemcpyHostToDevice, v.stream));
  };
  const size_t max_pairs = (size_t)kDkMaxAtoms * (kDkMaxAtoms - 1) / 2;
  up(&l.local, local, kDkMaxAtoms); up(&l.atom_seg, aseg, kDkMaxAtoms); up(&l.seg, seg, kDkMaxSeg);
  up(&l.seg_rel_origin, ro, kDkMaxSeg); up(&l.seg_rel_axis, ra, kDkMaxSeg); up(&l.pairs, pairs, max_pairs);
  // CSR adjacency of the pair list in both directions: the device sums the pair forces atom by atom (no atomics)
  std::vector<int> adj_off(na + 1, 0), adj(std::max(2 * t->n_pairs, 1), 0);
  for (int k = 0; k < t->n_pairs; k++) {
    GB_CHECK(t->pair_a[k] >= 0 && t->pair_a[k] < na && t->pair_b[k] >= 0 && t->pair_b[k] < na && t->pair_a[k] != t->pair_b[k], "pair atom index");
    adj_off[t->pair_a[k] + 1]++; adj_off[t->pair_b[k] + 1]++;
  }
  for (int i = 0; i < na; i++) adj_off[i + 1] += adj_off[i];
  {
    std::vector<int> fill(adj_off.begin(), adj_off.end() - 1);
    for (int k = 0; k < t->n_pairs; k++) {  // pair order is kept inside every atom's list
      adj[fill[t->pair_a[k]]++] = t->pair_b[k] | (k << 8) | (1 << 30);
      adj[fill[t->pair_b[k]]++] = t->pair_a[k] | (k << 8);
    }
  }
  up(&l.adj_off, adj_off, kDkMaxAtoms + 1); up(&l.adj, adj, 2 * max_pairs);
  std::vector<int> child_off(ns + 1, 0), child(std::max(ns - 1, 1), 0);
  for (int s2 = 1; s2 < ns; s2++) child_off[t->seg_parent[s2] + 1]++;
  for (int s2 = 0; s2 < ns; s2++) child_off[s2 + 1] += child_off[s2];
  {
    std::vector<int> fill(child_off.begin(), child_off.end() - 1);
    for (int s2 = 1; s2 < ns; s2++) child[fill[t->seg_parent[s2]]++] = s2;  // ascending within every parent
  }
  up(&l.child_off, child_off, kDkMaxSeg + 1); up(&l.child, child, kDkMaxSeg);
  unsigned heavy = 0;
  for (int i = 0; i < na; i++)
    if (t->smina_type[i] >= 2 && t->smina_type[i] < kNumSminaTypes) heavy |= 1u << t->smina_type[i];
  l.heavy_types = heavy;
  GB_CUDA(cudaStreamSynchronize(v.stream));  // the host vectors above go out of scope
  l.n_atoms = na; l.n_seg = ns; l.n_pairs = t->n_pairs; l.max_depth = max_depth; l.n_heavy = nh; l.gyration_radius = t->gyration_radius;
  GBV_END
}

// dynamic shared memory of the docking kernels: kDkWarps per-warp workspaces sized by the actual ligand.  A large
// ligand (96 atoms, 4560 pairs) needs more than the 48 KB a kernel gets by default: opt in once per device.
static size_t dock_smem_bytes(const Vina& v) {
  const size_t bytes = (size_t)kDkWarps * dk_ws_floats(v.lig.n_atoms, v.lig.n_seg, v.lig.n_pairs) * sizeof(float);
  constexpr int kMaxDyn = 160 * 1024;
  GB_CHECK(bytes <= (size_t)kMaxDyn, "ligand too large for the per-warp docking workspace");
  static std::mutex mu;
  static bool done[64] = {};
  std::lock_guard<std::mutex> lk(mu);
  if (v.device < 64 && !done[v.device]) {
    GB_CUDA(cudaFuncSetAttribute(dock_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDyn));
    GB_CUDA(cudaFuncSetAttribute(dock_refine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDyn));
    GB_CUDA(cudaFuncSetAttribute(dock_mc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDyn));
    done[v.device] = true;
  }
  return bytes;
}
// non_cache field: direct receptor sums, box = [bb, be] (grid_dims of the search box, main/main.cpp:474-497)
static void make_noncache_field(const Vina& v, float slope, const float* bb, const float* be, DockField& F) {
  GB_CHECK(v.d_rec && v.n_rec > 0, "gb_vina_set_receptor has not been called");
  GB_CHECK(bb && be, "non_cache needs the box");
  make_field(v, slope, F);
  F.rec = v.d_rec; F.n_rec = v.n_rec;
  for (int i = 0; i < 3; i++) { F.nc_begin[i] = bb[i]; F.nc_end[i] = be[i]; }
}

static int dock_eval_common(gb_vina* h, const float* confs, int n, const float* vcap, float slope, int mode, int maxiters, float* e,
                            float* change, float* coords, float* confs_out, int32_t* evals, const float* nc_begin = nullptr,
                            const float* nc_end = nullptr, int accurate = 0, int early_term = 0) {
  GBV_BEGIN
  GB_CHECK(h && confs && vcap && e && n >= 0, "bad arguments");
  Vina& v = h->v;
  GB_CUDA(cudaSetDevice(v.device));
  check_dock_ready(v, nc_begin == nullptr);
  if (n == 0) return GB_OK;
  const int T = v.lig.n_seg - 1, nx = 7 + T, ng = 6 + T, na = v.lig.n_atoms;
  // workspaces are sized by the MAXIMUM ligand dimensions: in a screen every ligand has another size, and a growing
  // buffer means cudaFree / cudaFreeHost = a device-wide synchronisation that stalls every other handle's kernels
  constexpr size_t NXA = 7 + kDkMaxSeg - 1, NGA = 6 + kDkMaxSeg - 1, NAA = kDkMaxAtoms;
  float* d_conf = v.ws<float>(0, (size_t)n * NXA);
  float* d_e = v.ws<float>(1, n);
  float* d_g = v.ws<float>(2, (size_t)n * NGA);
  float* d_c = coords ? v.ws<float>(3, (size_t)n * 3 * NAA) : nullptr;
  float* d_xo = mode == 1 ? v.ws<float>(4, (size_t)n * NXA) : nullptr;
  int* d_ev = mode == 1 ? v.ws<int>(5, n) : nullptr;
  float* p_conf = v.pin<float>(0, (size_t)n * NXA);
  memcpy(p_conf, confs, (size_t)n * nx * 4);
  GB_CUDA(cudaMemcpyAsync(d_conf, p_conf, (size_t)n * nx * 4, cudaMemcpyHostToDevice, v.stream));
  DockField F;
  if (nc_begin) make_noncache_field(v, slope, nc_begin, nc_end, F);
  else make_field(v, slope, F);
  const size_t dk_smem_bytes = dock_smem_bytes(v);
  dock_eval_kernel<<<(n + kDkWarps - 1) / kDkWarps, 32 * kDkWarps, dk_smem_bytes, v.stream>>>(lig_ptrs(v), F, d_conf, n, vcap[0], vcap[1], vcap[2], d_e, d_g,
                                                                                  d_c, mode, maxiters, d_xo, d_ev, accurate, early_term);
  GB_CUDA(cudaGetLastError());
  float* p_e = v.pin<float>(1, n);
  float* p_g = change ? v.pin<float>(2, (size_t)n * NGA) : nullptr;
  float* p_c = coords ? v.pin<float>(3, (size_t)n * 3 * NAA) : nullptr;
  float* p_xo = (mode == 1 && confs_out) ? v.pin<float>(4, (size_t)n * NXA) : nullptr;
  int* p_ev = (mode == 1 && evals) ? v.pin<int>(5, n) : nullptr;
  GB_CUDA(cudaMemcpyAsync(p_e, d_e, (size_t)n * 4, cudaMemcpyDeviceToHost, v.stream));
  if (p_g) GB_CUDA(cudaMemcpyAsync(p_g, d_g, (size_t)n * ng * 4, cudaMemcpyDeviceToHost, v.stream));
  if (p_c) GB_CUDA(cudaMemcpyAsync(p_c, d_c, (size_t)n * 3 * na * 4, cudaMemcpyDeviceToHost, v.stream));
  if (p_xo) GB_CUDA(cudaMemcpyAsync(p_xo, d_xo, (size_t)n * nx * 4, cudaMemcpyDeviceToHost, v.stream));
  if (p_ev) GB_CUDA(cudaMemcpyAsync(p_ev, d_ev, (size_t)n * 4, cudaMemcpyDeviceToHost, v.stream));
  GB_CUDA(cudaStreamSynchronize(v.stream));
  memcpy(e, p_e, (size_t)n * 4);
  if (p_g) memcpy(change, p_g, (size_t)n * ng * 4);
  if (p_c) memcpy(coords, p_c, (size_t)n * 3 * na * 4);
  if (p_xo) memcpy(confs_out, p_xo, (size_t)n * nx * 4);
  if (p_ev) memcpy(evals, p_ev, (size_t)n * 4);
  GBV_END
}

int gb_vina_eval_deriv(gb_vina* h, const float* confs, int n, const float* v3, float slope, float* e, float* change, float* coords) {
  return dock_eval_common(h, confs, n, v3, slope, 0, 0, e, change, coords, nullptr, nullptr);
}
int gb_vina_bfgs(gb_vina* h, float* confs, int n, int maxiters, const float* v3, float slope, float* e, float* change, int32_t* n_evals) {
  return dock_eval_common(h, confs, n, v3, slope, 1, maxiters, e, change, nullptr, confs, n_evals);
}
int gb_vina_minimize(gb_vina* h, float* confs, int n, const gb_minimization_params* mp, const float* v3, float slope, float* e,
                     float* change, int32_t* n_evals) {
  if (!mp) { gb::set_last_error("gb_vina_minimize: null params"); return GB_ERR_USAGE; }
  return dock_eval_common(h, confs, n, v3, slope, 1, mp->maxiters, e, change, nullptr, confs, n_evals, nullptr, nullptr,
                          mp->accurate_line_search, mp->early_term);
}

int gb_vina_eval_deriv_noncache(gb_vina* h, const float* confs, int n, const float* v3, float slope, const float* box_begin,
                                const float* box_end, float* e, float* change) {
  if (!box_begin || !box_end) { gb::set_last_error("gb_vina_eval_deriv_noncache: null box"); return GB_ERR_USAGE; }
  return dock_eval_common(h, confs, n, v3, slope, 0, 0, e, change, nullptr, nullptr, nullptr, box_begin, box_end);
}

int gb_vina_noncache_atoms(gb_vina* h, const float* xyz, const int32_t* smina_type, int n_atoms, const float* box_begin,
                           const float* box_end, float v, float* e, float* deriv) {
  GBV_BEGIN
  GB_CHECK(h && xyz && smina_type && box_begin && box_end && e && deriv && n_atoms >= 0, "bad arguments");
  Vina& vv = h->v;
  GB_CUDA(cudaSetDevice(vv.device));
  if (n_atoms == 0) return GB_OK;
  const int32_t off[2] = {0, n_atoms};
  stage_poses(vv, xyz, smina_type, off, 1);
  DockField F;
  make_noncache_field(vv, 0.f, box_begin, box_end, F);   // slope 0: the caller adds its own out-of-box terms
  noncache_atoms_kernel<<<(n_atoms + 127) / 128, 128, 0, vv.stream>>>(vv.d_lig, n_atoms, F, v, vv.d_atom_e, vv.d_deriv);
  GB_CUDA(cudaGetLastError());
  GB_CUDA(cudaMemcpyAsync(e, vv.d_atom_e, (size_t)n_atoms * sizeof(float), cudaMemcpyDeviceToHost, vv.stream));
  GB_CUDA(cudaMemcpyAsync(deriv, vv.d_deriv, (size_t)n_atoms * 3 * sizeof(float), cudaMemcpyDeviceToHost, vv.stream));
  GB_CUDA(cudaStreamSynchronize(vv.stream));
  GBV_END
}

static int refine_common(gb_vina* h, float* confs, int n, int maxiters, int accurate, int early_term, const float* v3,
                         const float* box_begin, const float* box_end, float* e, int32_t* within, int32_t* n_evals);
int gb_vina_refine(gb_vina* h, float* confs, int n, int maxiters, const float* v3, const float* box_begin, const float* box_end,
                   float* e, int32_t* within, int32_t* n_evals) {
  return refine_common(h, confs, n, maxiters, 0, 0, v3, box_begin, box_end, e, within, n_evals);
}
int gb_vina_refine_minimize(gb_vina* h, float* confs, int n, const gb_minimization_params* mp, const float* v3, const float* box_begin,
                            const float* box_end, float* e, int32_t* within, int32_t* n_evals) {
  if (!mp) { gb::set_last_error("gb_vina_refine_minimize: null params"); return GB_ERR_USAGE; }
  return refine_common(h, confs, n, mp->maxiters, mp->accurate_line_search, mp->early_term, v3, box_begin, box_end, e, within, n_evals);
}
static int refine_common(gb_vina* h, float* confs, int n, int maxiters, int accurate, int early_term, const float* v3,
                         const float* box_begin, const float* box_end, float* e, int32_t* within, int32_t* n_evals) {
  GBV_BEGIN
  GB_CHECK(h && confs && v3 && box_begin && box_end && e && n >= 0, "bad arguments");
  Vina& v = h->v;
  GB_CUDA(cudaSetDevice(v.device));
  check_dock_ready(v, false);
  if (n == 0) return GB_OK;
  const int T = v.lig.n_seg - 1, nx = 7 + T;
  constexpr size_t NXA = 7 + kDkMaxSeg - 1;
  float* d_conf = v.ws<float>(0, (size_t)n * NXA);
  float* d_e = v.ws<float>(1, n);
  int* d_in = v.ws<int>(5, n);
  int* d_ev = v.ws<int>(6, n);
  float* p_conf = v.pin<float>(0, (size_t)n * NXA);
  memcpy(p_conf, confs, (size_t)n * nx * 4);
  GB_CUDA(cudaMemcpyAsync(d_conf, p_conf, (size_t)n * nx * 4, cudaMemcpyHostToDevice, v.stream));
  DockField F;
  make_noncache_field(v, 10.f, box_begin, box_end, F);
  const size_t smem = dock_smem_bytes(v);
  dock_refine_kernel<<<(n + kDkWarps - 1) / kDkWarps, 32 * kDkWarps, smem, v.stream>>>(lig_ptrs(v), F, d_conf, n, v3[0], v3[1], v3[2], maxiters,
                                                                                    d_e, d_in, d_ev, accurate, early_term);
  GB_CUDA(cudaGetLastError());
  float* p_e = v.pin<float>(1, n);
  int* p_in = v.pin<int>(5, n);
  int* p_ev = v.pin<int>(6, n);
  GB_CUDA(cudaMemcpyAsync(p_conf, d_conf, (size_t)n * nx * 4, cudaMemcpyDeviceToHost, v.stream));
  GB_CUDA(cudaMemcpyAsync(p_e, d_e, (size_t)n * 4, cudaMemcpyDeviceToHost, v.stream));
  GB_CUDA(cudaMemcpyAsync(p_in, d_in, (size_t)n * 4, cudaMemcpyDeviceToHost, v.stream));
  GB_CUDA(cudaMemcpyAsync(p_ev, d_ev, (size_t)n * 4, cudaMemcpyDeviceToHost, v.stream));
  GB_CUDA(cudaStreamSynchronize(v.stream));
  memcpy(confs, p_conf, (size_t)n * nx * 4);
  memcpy(e, p_e, (size_t)n * 4);
  if (within) memcpy(within, p_in, (size_t)n * 4);
  if (n_evals) memcpy(n_evals, p_ev, (size_t)n * 4);
  GBV_END
}

int gb_vina_mc(gb_vina* h, const gb_mc_params* P, const float* corner1, const float* corner2, const uint32_t* seeds, int n_chains,
               float slope, float* out_e, float* out_conf, int32_t* n_out) {
  return gb_vina_mc_traced(h, P, corner1, corner2, seeds, n_chains, slope, out_e, out_conf, n_out, nullptr);
}

int gb_vina_mc_traced(gb_vina* h, const gb_mc_params* P, const float* corner1, const float* corner2, const uint32_t* seeds, int n_chains,
                      float slope, float* out_e, float* out_conf, int32_t* n_out, float* trace) {
  GBV_BEGIN
  GB_CHECK(h && P && corner1 && corner2 && seeds && out_e && out_conf && n_out && n_chains >= 0, "bad arguments");
  GB_CHECK(P->num_saved_mins >= 1 && P->num_saved_mins <= 64, "num_saved_mins out of range (1..64)");
  Vina& v = h->v;
  GB_CUDA(cudaSetDevice(v.device));
  check_dock_ready(v);
  if (n_chains == 0) return GB_OK;
  const int T = v.lig.n_seg - 1, nx = 7 + T, S = P->num_saved_mins, nh = std::max(v.lig.n_heavy, 1);
  uint32_t* d_seeds = v.ws<uint32_t>(0, n_chains);
  float* d_e = v.ws<float>(1, (size_t)n_chains * S);
  constexpr size_t NXA = 7 + kDkMaxSeg - 1;  // allocation by maximum ligand dimensions, see dock_eval_common
  float* d_c = v.ws<float>(2, (size_t)n_chains * S * NXA);
  float* d_h = v.ws<float>(3, (size_t)n_chains * S * 3 * kDkMaxAtoms);
  int* d_n = v.ws<int>(4, n_chains);
  uint32_t* p_seeds = v.pin<uint32_t>(0, n_chains);
  memcpy(p_seeds, seeds, (size_t)n_chains * 4);
  GB_CUDA(cudaMemcpyAsync(d_seeds, p_seeds, (size_t)n_chains * 4, cudaMemcpyHostToDevice, v.stream));
  GB_CUDA(cudaMemsetAsync(d_e, 0, (size_t)n_chains * S * 4, v.stream));
  GB_CUDA(cudaMemsetAsync(d_c, 0, (size_t)n_chains * S * nx * 4, v.stream));
  DockField F;
  make_field(v, slope, F);
  McDev M{P->num_steps, P->maxiters, S, P->temperature, P->mutation_amplitude, P->min_rmsd, {P->hunt_cap[0], P->hunt_cap[1], P->hunt_cap[2]}};
  const size_t dk_smem_bytes = dock_smem_bytes(v);
  float* d_trace = trace ? v.ws<float>(7, (size_t)n_chains * std::max(P->num_steps, 1)) : nullptr;
  dock_mc_kernel<<<(n_chains + kDkWarps - 1) / kDkWarps, 32 * kDkWarps, dk_smem_bytes, v.stream>>>(lig_ptrs(v), F, M, corner1[0], corner1[1], corner1[2],
                                                                                     corner2[0], corner2[1], corner2[2], d_seeds, n_chains, d_e,
                                                                                     d_c, d_h, d_n, d_trace);
  GB_CUDA(cudaGetLastError());
  float* p_e = v.pin<float>(1, (size_t)n_chains * S);
  float* p_c = v.pin<float>(2, (size_t)n_chains * S * NXA);
  int* p_n = v.pin<int>(3, n_chains);
  GB_CUDA(cudaMemcpyAsync(p_e, d_e, (size_t)n_chains * S * 4, cudaMemcpyDeviceToHost, v.stream));
  GB_CUDA(cudaMemcpyAsync(p_c, d_c, (size_t)n_chains * S * nx * 4, cudaMemcpyDeviceToHost, v.stream));
  GB_CUDA(cudaMemcpyAsync(p_n, d_n, (size_t)n_chains * 4, cudaMemcpyDeviceToHost, v.stream));
  float* p_tr = trace ? v.pin<float>(7, (size_t)n_chains * std::max(P->num_steps, 1)) : nullptr;
  if (p_tr) GB_CUDA(cudaMemcpyAsync(p_tr, d_trace, (size_t)n_chains * P->num_steps * 4, cudaMemcpyDeviceToHost, v.stream));
  GB_CUDA(cudaStreamSynchronize(v.stream));
  if (p_tr) memcpy(trace, p_tr, (size_t)n_chains * P->num_steps * 4);
  memcpy(out_e, p_e, (size_t)n_chains * S * 4);
  memcpy(out_conf, p_c, (size_t)n_chains * S * nx * 4);
  memcpy(n_out, p_n, (size_t)n_chains * 4);
  GBV_END
}

int gb_vina_merge_outputs(const float* e, const float* coords, const int32_t* n_out, int n_chains, int S, int n_atoms,
                          float min_rmsd, int max_size, int32_t* kept, int32_t* n_kept) {
  GBV_BEGIN
  GB_CHECK(e && coords && n_out && kept && n_kept && n_chains >= 0 && S >= 1 && n_atoms >= 0 && max_size >= 0, "bad arguments");
  std::vector<int> out;  // flat indices, kept sorted by energy
  const size_t stride = (size_t)n_atoms * 3;
  auto rmsd = [&](int a, int b) -> double {  // rmsd_upper_bound
    if (n_atoms == 0) return 0.0;
    const float *pa = coords + (size_t)a * stride, *pb = coords + (size_t)b * stride;
    double acc = 0;
    for (size_t i = 0; i < stride; i++) { const double d = (double)pa[i] - (double)pb[i]; acc += d * d; }
    return std::sqrt(acc / n_atoms);
  };
  for (int c = 0; c < n_chains; c++) {
    GB_CHECK(n_out[c] >= 0 && n_out[c] <= S, "n_out out of range");
    for (int k = 0; k < n_out[c]; k++) {
      const int t = c * S + k;
      size_t best = out.size();
      double best_r = 0;
      for (size_t i = 0; i < out.size(); i++) {  // find_closest
        const double r = rmsd(t, out[i]);
        if (i == 0 || r < best_r) { best = i; best_r = r; }
      }
      if (best < out.size() && best_r < min_rmsd) {
        if (e[t] < e[out[best]]) out[best] = t;
      } else if ((int)out.size() < max_size) {
        out.push_back(t);
      } else if (!out.empty() && e[t] < e[out.back()]) {
        out.back() = t;
      }
      std::stable_sort(out.begin(), out.end(), [&](int a, int b) { return e[a] < e[b]; });
    }
  }
  for (size_t i = 0; i < out.size(); i++) kept[i] = out[i];
  *n_kept = (int32_t)out.size();
  GBV_END
}

}  // extern "C"
