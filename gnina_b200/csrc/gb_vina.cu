// Batched smina/Vina scoring rows on the device:
//   V1/V2 host: default term set (lib/everything.h:149-247,480-506; weights main/main.cpp:1324-1329) tabulated like
//         precalculate_linear (lib/precalculate.h:82-272, factor 32 -> 2051 samples in r^2 per type pair)
//   V4    cache::populate (lib/cache.cpp:104-184): one thread per grid point, receptor atoms in index order
//   V5    cache::eval / eval_deriv -> grid::evaluate_aux (lib/grid.cpp:96-186) + curl (lib/curl.h:30-35)
//   V12   naive_non_cache::eval with precalculate_exact (lib/naive_non_cache.cpp:29-57, lib/precalculate.h:452-463)
//         and num_tors_div (lib/everything.h:795-809): the printed "Affinity (kcal/mol)" of a rigid pose
// Summation orders follow the reference (per-atom partial, curl, then atoms in index order), so results differ from
// the CPU restatement only by the device's expf/sqrtf rounding.
#include <cmath>
#include <cstring>
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
  float w[6];
  float factor = 32.f, cutoff_sqr = 64.f;
  int n = 0;
  std::vector<float> h_fast, h_se, h_sd;  // [pair][n]
  float* d_fast = nullptr;
  // receptor (heavy atoms, index order)
  float4* d_rec = nullptr;  // x, y, z, type
  int n_rec = 0;
  // cache
  float begin[3], end[3];
  int gn[3] = {0, 0, 0};
  float* d_grids[kNumSminaTypes] = {};
  // pose staging
  float4* d_lig = nullptr; int* d_off = nullptr; float* d_atom_e = nullptr; float* d_deriv = nullptr; float* d_pose_e = nullptr;
  float* d_tors = nullptr;
  size_t cap_atoms = 0, cap_poses = 0;
  ~Vina() {
    cudaSetDevice(device);
    if (stream) cudaStreamDestroy(stream);
    cudaFree(d_fast); cudaFree(d_rec); cudaFree(d_lig); cudaFree(d_off); cudaFree(d_atom_e); cudaFree(d_deriv);
    cudaFree(d_pose_e); cudaFree(d_tors);
    for (auto g : d_grids) cudaFree(g);
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
    v.cap_atoms = (size_t)n_atoms + 1024;
    GB_CUDA(cudaMalloc(&v.d_lig, v.cap_atoms * sizeof(float4)));
    GB_CUDA(cudaMalloc(&v.d_atom_e, v.cap_atoms * sizeof(float)));
    GB_CUDA(cudaMalloc(&v.d_deriv, v.cap_atoms * 3 * sizeof(float)));
  }
  if ((size_t)n_poses + 1 > v.cap_poses) {
    cudaFree(v.d_off); cudaFree(v.d_pose_e); cudaFree(v.d_tors);
    v.cap_poses = (size_t)n_poses + 1024;
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
  GB_CHECK(h && begin && end && n && types_needed && n_types > 0 && n_types <= kMaxNeeded, "bad cache arguments");
  Vina& v = h->v;
  GB_CUDA(cudaSetDevice(v.device));
  for (auto& g : v.d_grids) { cudaFree(g); g = nullptr; }
  NeededTypes nt;
  nt.n = n_types;
  const size_t vol = (size_t)(n[0] + 1) * (n[1] + 1) * (n[2] + 1);
  float finv[3];
  for (int i = 0; i < 3; i++) {
    GB_CHECK(n[i] >= 1 && end[i] > begin[i], "bad grid dims");
    v.begin[i] = begin[i]; v.end[i] = end[i]; v.gn[i] = n[i];
    const float factor = (float)((n[i] + 1) - 1.0) / (end[i] - begin[i]);
    finv[i] = 1 / factor;
  }
  for (int j = 0; j < n_types; j++) {
    const int t = types_needed[j];
    GB_CHECK(t >= 2 && t < kNumSminaTypes, "needed type must be a heavy smina type");
    GB_CUDA(cudaMalloc(&v.d_grids[t], vol * sizeof(float)));
    nt.t[j] = t;
    nt.grid[j] = v.d_grids[t];
  }
  cache_populate_kernel<<<(unsigned)((vol + 127) / 128), 128, 0, v.stream>>>(v.d_rec, v.n_rec, v.d_fast, v.n, v.factor,
                                                                              v.cutoff_sqr, begin[0], begin[1], begin[2], finv[0],
                                                                              finv[1], finv[2], n[0] + 1, n[1] + 1, n[2] + 1, nt);
  GB_CUDA(cudaGetLastError());
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

int gb_vina_score_exact(gb_vina* h, const float* lig_xyz, const int32_t* lig_type, const int32_t* pose_offsets, int n_poses,
                        const float* num_tors, float vcap, float* e_inter, float* affinity) {
  GBV_BEGIN
  GB_CHECK(h && n_poses >= 0, "bad arguments");
  Vina& v = h->v;
  GB_CUDA(cudaSetDevice(v.device));
  if (n_poses == 0) return GB_OK;
  stage_poses(v, lig_xyz, lig_type, pose_offsets, n_poses);
  const int n_atoms = pose_offsets[n_poses];
  if (n_atoms) naive_exact_kernel<<<(n_atoms + 127) / 128, 128, 0, v.stream>>>(v.d_lig, n_atoms, v.d_rec, v.n_rec, v.cutoff_sqr,
                                                                               vcap, v.d_atom_e);
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
      const float y = (float)(1 + (double)w * (double)(num_tors ? num_tors[p] : 0.f) / 5.0);
      const float x = e[p];
      const float eps = 1.1920929e-07f, maxfl = 3.402823466e+38f;
      affinity[p] = std::fabs(x) < eps ? 0.f : (std::fabs(y) < eps ? ((x * y > 0) ? maxfl : -maxfl) : x / y);
    }
  }
  GBV_END
}

}  // extern "C"
