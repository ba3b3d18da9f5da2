// Fused voxelise + 2x2x2 pool of the fast path (G1 of SURVEY.md §8a fused with the first pooling layer of N1 / N2).
//
// libmolgrid's density (GridMaker::forward, called at lib/torch_model.cpp:181) for an atom of radius r at distance d:
//     exp(-2 d^2 / r^2)                     d <= r
//     e^-2 (4 (d/r)^2 - 12 d/r + 9)         r < d < 1.5 r
//     0                                     otherwise
// summed per channel over the atoms of the pose; the network's first layer is a 2x2x2 pooling (average for the
// default2018 family, max for the dense family), so this kernel never materialises the 48^3 grid (12.4 MB per
// pose): it writes the pooled 24^3 x 32 fp16 tensor in the chunk-planar padded layout conv1 reads (gb_cnn_tc.cu).
//
// Work decomposition (round 2, "v3"; profiles/README.md r2a has the instruction counts):
//   CTA   = one pose x one tile of 8x8x8 pooled voxels (16^3 fine voxels, an 8 A cube), 512 threads
//   thread= one pooled voxel = 8 sub-voxel densities per atom within reach
//   atoms : pose list (channel sorted, built once per pose by build_pose_lists_kernel) -> tile list by an ordered
//           block-wide compaction with an exact sphere/box test -> per-warp candidates (sphere vs the warp's 4x4x2
//           voxel box, one ballot per 32 tile atoms) -> per-lane bit mask (pooled-voxel centre within reach + sub-voxel
//           half-diagonal) walked in list order, so every lane evaluates exactly the atoms that can reach its voxel.
//   maths : everything is done in the scaled variable v = kappa (d/r)^2, kappa = 2 log2(e), so exp(-2 d^2/r^2) =
//           ex2(-v) is one MUFU without a multiply; the eight v of a pooled voxel are S0 +- ex +- ey +- ez with
//           S0 = |delta|^2 + 3 h^2, e = 2 h delta (delta = centre offset, h = quarter voxel, both in scaled units of r);
//           the shell polynomial, its square and the accumulation run on packed FFMA2/FMUL2/FADD2 (two sub-voxels per
//           instruction); the core/shell select is a predicated MUFU over the shell value.
//   output: fp16 tile in shared memory as [channel pair][voxel] half2 words; each thread then emits the four 16-byte
//           channel chunks of its voxel with four 32-bit shared loads each.
#include <cstdlib>
#include <cuda_fp16.h>
#include "gb_ptx.cuh"
#include "gb_tc.h"

namespace gb {

namespace {

constexpr float kKappa = 2.885390081777927f;        // 2 log2(e)
constexpr float kSqrtKappa = 1.6986436005760381f;   // sqrt(kappa)
constexpr int kVoxChunk = 512;                      // pose-list atoms staged per block-wide compaction

// shell polynomial u(s) ~ (2 sqrt(s) - 3) / e on s = (d/r)^2 in [1, 2.25] (|err| < 4e-5, u(2.25) = 0 exactly so the
// clamp of s at 2.25 implements the cut-off), re-expressed in v = kappa s:  c_k' = c_k / kappa^k
constexpr double kU0 = -0.8559605479240417, kU1 = 0.6502527594566345, kU2 = -0.20902030169963837,
                 kU3 = 0.052687861025333405, kU4 = -0.005817742552608252;
constexpr double kK = 2.885390081777927;
constexpr float kC0 = (float)kU0, kC1 = (float)(kU1 / kK), kC2 = (float)(kU2 / (kK * kK)),
                kC3 = (float)(kU3 / (kK * kK * kK)), kC4 = (float)(kU4 / (kK * kK * kK * kK));
constexpr float kVClamp = (float)(2.25 * kK);

// ---- packed fp32x2 (sm_100: FFMA2 / FMUL2 / FADD2) ----
typedef unsigned long long f2;
__device__ __forceinline__ f2 pk(float lo, float hi) {
  f2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk(f2 a, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a)); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
  f2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
  f2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
  f2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// r = (v <= kappa) ? 2^-v : r     (one FSETP + one predicated MUFU.EX2)
__device__ __forceinline__ void core_over_shell(float& r, float v) {
  asm("{\n\t.reg .pred p;\n\t"
      ".reg .f32 nv;\n\t"
      "setp.le.f32 p, %1, 0f4038AA3B;\n\t"   // kappa = 2.885390081777927
      "neg.f32 nv, %1;\n\t"
      "@p ex2.approx.ftz.f32 %0, nv;\n\t}"
      : "+f"(r)
      : "f"(v));
}

// densities of the two sub-voxels (va, vb): packed result
__device__ __forceinline__ f2 density_pair(float va, float vb, f2 C0, f2 C1, f2 C2, f2 C3, f2 C4) {
  const float ta = fminf(va, kVClamp), tb = fminf(vb, kVClamp);
  const f2 T = pk(ta, tb);
  f2 U = fma2(T, C4, C3);
  U = fma2(U, T, C2);
  U = fma2(U, T, C1);
  U = fma2(U, T, C0);
  U = mul2(U, U);
  float ra, rb;
  upk(U, ra, rb);
  core_over_shell(ra, ta);
  core_over_shell(rb, tb);
  return pk(ra, rb);
}

__device__ __forceinline__ int block_ordered_slot512(bool pred, int* s_warp_counts, int& total) {
  const unsigned mask = __ballot_sync(0xffffffffu, pred);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) s_warp_counts[warp] = __popc(mask);
  __syncthreads();
  // inclusive scan of the 16 warp counts across lanes 0..15 (every warp does it redundantly)
  const int c = lane < 16 ? s_warp_counts[lane] : 0;
  int incl = c;
#pragma unroll
  for (int o = 1; o < 16; o <<= 1) {
    const int up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += up;
  }
  total = __shfl_sync(0xffffffffu, incl, 15);
  const int base = __shfl_sync(0xffffffffu, incl - c, warp);
  __syncthreads();   // s_warp_counts may be rewritten by the next chunk
  return base + __popc(mask & ((1u << lane) - 1u));
}

// squared distance from the point x to the interval [lo, hi]
__device__ __forceinline__ float gap2(float x, float lo, float hi) {
  const float g = fmaxf(fmaxf(lo - x, x - hi), 0.f);
  return g * g;
}

template <int kC8>
struct VoxSmemT {
  float4 A[kVoxChunk];       // (-bx w', -by w', -bz w', vs w'), b = atom - tile origin, w' = sqrt(kappa)/r, vs = 2 res
  float4 B[kVoxChunk];       // (bx, by, bz, reach) in Angstrom, reach = 1.5 r
  float thr2[kVoxChunk];     // kappa (1.5 + half-diagonal/r)^2 : pooled-centre test in scaled units
  int ch[kVoxChunk];
  uint32_t out[kC8 * 4 * 512];   // [channel pair][pooled voxel] half2
  int counts[16];
};
using VoxSmem = VoxSmemT<4>;

template <bool kMax, int kMinBlocks, int kC8 = 4>
__global__ void __launch_bounds__(512, kMinBlocks) voxelize_pool_f16_kernel(const float4* __restrict__ list_xyzr,
                                                                const int* __restrict__ list_ch,
                                                                const int* __restrict__ list_n, int cap,
                                                                const float* __restrict__ centers, float resolution,
                                                                float dimension, uint4* __restrict__ x0, int Lp, int G) {
  // the pooled grid of every supported model is 24^3 x 32 channels (48^3 fine voxels, 28 channels padded): compile-time
  // dimensions keep the index arithmetic free of integer divisions (ncu r2a: the prologue was 11 % of the instructions)
  // (kC8 = 6: the 35 channels of default2017, padded to 48)
  constexpr int D = 24, P = D + 2, C8 = kC8, tiles = D / 8;
  extern __shared__ __align__(16) uint8_t vox_smem_raw[];
  VoxSmemT<kC8>& S = *reinterpret_cast<VoxSmemT<kC8>*>(vox_smem_raw);
  const int p = blockIdx.y;
  const int t = blockIdx.x;
  const int tx = t / (tiles * tiles), ty = (t / tiles) % tiles, tz = t % tiles;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wx = warp & 1, wy = (warp >> 1) & 1, wz = warp >> 2;
  const int px = wx * 4 + (lane & 3), py = wy * 4 + ((lane >> 2) & 3), pz = wz * 2 + (lane >> 4);
  const int pv = (px * 8 + py) * 8 + pz;
  // zero this thread's own voxel (channels without atoms in reach are never flushed): 16 channel-pair words
#pragma unroll
  for (int e = 0; e < 4 * kC8; e++) S.out[e * 512 + pv] = 0u;
  const float half = dimension * 0.5f;
  const float vs = 2.f * resolution, hres = 0.5f * resolution;
  // tile origin = centre of the tile's pooled voxel (0,0,0), i.e. between fine voxels 16 t and 16 t + 1
  const float t0x = centers[3 * p] - half + (16 * tx + 0.5f) * resolution;
  const float t0y = centers[3 * p + 1] - half + (16 * ty + 0.5f) * resolution;
  const float t0z = centers[3 * p + 2] - half + (16 * tz + 0.5f) * resolution;
  const float fx = (float)px, fy = (float)py, fz = (float)pz;
  // boxes (tile-relative Angstrom) spanned by the fine-voxel centres of the tile / of this warp's 4x4x2 voxels
  const float tlo = -hres, thi = 7.f * vs + hres;
  const float wlx = 4 * wx * vs - hres, whx = (4 * wx + 3) * vs + hres;
  const float wly = 4 * wy * vs - hres, why = (4 * wy + 3) * vs + hres;
  const float wlz = 2 * wz * vs - hres, whz = (2 * wz + 1) * vs + hres;
  const float4* la = list_xyzr + (size_t)p * cap;
  const int* lc = list_ch + (size_t)p * cap;
  const int n = list_n[p];
  const f2 C0 = pk(kC0, kC0), C1 = pk(kC1, kC1), C2 = pk(kC2, kC2), C3 = pk(kC3, kC3), C4 = pk(kC4, kC4);
  int cur = -1;
  f2 acc[kMax ? 4 : 1];  // kMax: the 8 sub-voxel sums are kept apart until the channel is flushed
#pragma unroll
  for (int q = 0; q < (kMax ? 4 : 1); q++) acc[q] = 0ull;
  auto flush = [&]() {
    float v;
    if constexpr (kMax) {
      float m = 0.f;  // sums of non-negative densities
#pragma unroll
      for (int q = 0; q < 4; q++) {
        float a, b;
        upk(acc[q], a, b);
        m = fmaxf(m, fmaxf(a, b));
        acc[q] = 0ull;
      }
      v = m;
    } else {
      float a, b;
      upk(acc[0], a, b);
      v = (a + b) * 0.125f;
      acc[0] = 0ull;
    }
    reinterpret_cast<__half*>(S.out)[((cur >> 1) * 512 + pv) * 2 + (cur & 1)] = __float2half(v);
  };
  for (int base = 0; base < n; base += kVoxChunk) {
    bool keep = false;
    float4 a = make_float4(0, 0, 0, 1);
    float bx = 0, by = 0, bz = 0, reach = 0;
    int ch = 0;
    const int ai = base + tid;
    if (ai < n) {
      a = la[ai];
      ch = lc[ai];
      bx = a.x - t0x; by = a.y - t0y; bz = a.z - t0z;
      reach = 1.5f * a.w + 1e-4f;
      keep = gap2(bx, tlo, thi) + gap2(by, tlo, thi) + gap2(bz, tlo, thi) <= reach * reach;
    }
    int tot;
    const int slot = block_ordered_slot512(keep, S.counts, tot);
    if (keep) {
      const float w = kSqrtKappa / a.w;
      const float thr = 1.5f + 1.7320508f * hres / a.w + 1e-3f;  // sub-voxel centres are +-res/2 per axis off the pooled centre
      S.A[slot] = make_float4(-bx * w, -by * w, -bz * w, vs * w);
      S.B[slot] = make_float4(bx, by, bz, reach);
      S.thr2[slot] = kKappa * thr * thr;
      S.ch[slot] = ch;
    }
    __syncthreads();
    for (int wb = 0; wb < tot; wb += 32) {
      const int m = wb + lane;
      bool hit = false;
      if (m < tot) {
        const float4 b = S.B[m];
        hit = gap2(b.x, wlx, whx) + gap2(b.y, wly, why) + gap2(b.z, wlz, whz) <= b.w * b.w;
      }
      // (1) warp level: atoms whose support sphere touches the warp's voxel box.  (2) lane level: the subset this
      // lane's own pooled voxel can see -- a sphere fills ~1/4 of the boxes it touches, so every lane walks ITS bit
      // mask (divergent loop, trip count = the busiest lane's) in list order: the sums and channel flushes are those of
      // the all-lanes loop minus terms that are identically zero.
      const unsigned wmask = __ballot_sync(0xffffffffu, hit);
      unsigned lmask = 0u;
      for (unsigned mk = wmask; mk; mk &= mk - 1) {
        const int bit = __ffs(mk) - 1;
        const float4 A = S.A[wb + bit];
        const float dx = fmaf(fx, A.w, A.x), dy = fmaf(fy, A.w, A.y), dz = fmaf(fz, A.w, A.z);
        if (fmaf(dx, dx, fmaf(dy, dy, dz * dz)) < S.thr2[wb + bit]) lmask |= 1u << bit;
      }
      while (lmask) {
        const int mm = wb + __ffs(lmask) - 1;
        lmask &= lmask - 1;
        const int chm = S.ch[mm];
        if (chm != cur) {
          if (cur >= 0) flush();
          cur = chm;
        }
        const float4 A = S.A[mm];
        const float dx = fmaf(fx, A.w, A.x), dy = fmaf(fy, A.w, A.y), dz = fmaf(fz, A.w, A.z);
        const float h2 = 0.5f * A.w;                       // 2 h, h = quarter pooled voxel in scaled units
        const float s0 = fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, 0.1875f * A.w * A.w)));  // |delta|^2 + 3 h^2
        const float ex = h2 * dx, ey = h2 * dy, ez = h2 * dz;
        const float qa = s0 - ex, qb = s0 + ex;
        const float q00 = qa - ey, q01 = qa + ey, q10 = qb - ey, q11 = qb + ey;
        const f2 r0 = density_pair(q00 - ez, q00 + ez, C0, C1, C2, C3, C4);
        const f2 r1 = density_pair(q01 - ez, q01 + ez, C0, C1, C2, C3, C4);
        const f2 r2 = density_pair(q10 - ez, q10 + ez, C0, C1, C2, C3, C4);
        const f2 r3 = density_pair(q11 - ez, q11 + ez, C0, C1, C2, C3, C4);
        if constexpr (kMax) {
          acc[0] = add2(acc[0], r0); acc[1] = add2(acc[1], r1); acc[2] = add2(acc[2], r2); acc[3] = add2(acc[3], r3);
        } else {
          acc[0] = add2(acc[0], add2(add2(r0, r1), add2(r2, r3)));
        }
      }
    }
    if (base + kVoxChunk < n) __syncthreads();  // the next chunk overwrites the staged atoms
  }
  if (cur >= 0) flush();
  // Output.  Global rows are z-contiguous (8 voxels = 128 B per chunk), but a warp owns only 2 z: the four warps that share
  // (wx, wy) -- a 4 x 4 x 8 sub-block, 16 full z rows -- synchronise among themselves (named barrier, 128 threads) and
  // each thread emits one voxel of the sub-block, z fastest.  (r2a: the block-wide barrier in front of a tile-wide
  // transposed output held 18 % of the stall samples; r2b: per-thread own-voxel stores without any barrier were slower
  // still -- 32-byte runs.)
  ptx::named_bar_sync(1 + (warp & 3), 128);
  {
    const int gt = (warp >> 2) * 32 + lane;
    const int qz = gt & 7, qy = wy * 4 + ((gt >> 3) & 3), qx = wx * 4 + (gt >> 5);
    const int qv = (qx * 8 + qy) * 8 + qz;
    const int x = tx * 8 + qx, y = ty * 8 + qy, z = tz * 8 + qz;
    // group of G poses: [group][x][chunk][q P^2 + (y+1) P + (z+1)]
    uint4* dst = x0 + (((size_t)(p / G) * D + x) * C8) * Lp + (size_t)(p % G) * P * P + (size_t)(y + 1) * P + (z + 1);
#pragma unroll
    for (int c8 = 0; c8 < C8; c8++) {
      const uint32_t* w = S.out + (c8 * 4) * 512 + qv;
      dst[(size_t)c8 * Lp] = make_uint4(w[0], w[512], w[1024], w[1536]);
    }
  }
}

}  // namespace

TcGridWorkspace::~TcGridWorkspace() {
  for (auto& k : x0)
    for (auto p : k)
      if (p) cudaFree(p);
  for (auto e : ready)
    if (e) cudaEventDestroy(e);
  for (auto e : consumed)
    if (e) cudaEventDestroy(e);
  for (auto e : started)
    if (e) cudaEventDestroy(e);
  if (list_xyzr) cudaFree(list_xyzr);
  if (list_ch) cudaFree(list_ch);
  if (list_n) cudaFree(list_n);
}

int tc_prepare_grid(const TcPoseBatch& pb, TcGridWorkspace& gw, int buf, int kinds_mask, cudaStream_t s, Profiler* prof) {
  const int nb = pb.n_poses;
  const ActLayout L1 = make_layout(24, 1, 32), LF = make_fused_x0_layout(), L48 = make_layout(24, 1, 48);
  const int cap = std::max(1, pb.n_rec + pb.max_pose_atoms);
  // allocation sizes have a floor (64 poses, 128 ligand atoms) so that small batches of varying size -- the kept
  // poses of one docked ligand -- never re-allocate: cudaFree / cudaMemset synchronise the whole device and would
  // stall the kernels of other handles (DockingPool keeps one handle per host thread)
  const int nb_alloc = std::max(nb, 64);
  const size_t need = (size_t)nb_alloc * std::max(cap, pb.n_rec + 128);
  if (gw.list_cap < need) {
    GB_CUDA(cudaStreamSynchronize(s));
    if (gw.list_xyzr) cudaFree(gw.list_xyzr);
    if (gw.list_ch) cudaFree(gw.list_ch);
    gw.list_xyzr = nullptr; gw.list_ch = nullptr; gw.list_cap = 0;
    GB_CUDA(cudaMalloc(&gw.list_xyzr, need * sizeof(float4)));
    GB_CUDA(cudaMalloc(&gw.list_ch, need * sizeof(int)));
    gw.list_cap = need;
  }
  if (gw.listn_cap < (size_t)nb_alloc) {
    GB_CUDA(cudaStreamSynchronize(s));
    if (gw.list_n) cudaFree(gw.list_n);
    gw.list_n = nullptr; gw.listn_cap = 0;
    GB_CUDA(cudaMalloc(&gw.list_n, (size_t)nb_alloc * sizeof(int)));
    gw.listn_cap = nb_alloc;
  }
  for (int kind = 0; kind < 4; kind++) {
    const size_t need0 = act_bytes(kind == 2 ? LF : kind == 3 ? L48 : L1, nb_alloc);
    if (!(kinds_mask & (1 << kind)) || gw.cap[kind][buf] >= need0) continue;
    // growing the pooled-grid buffer: only kernels of THIS handle (its two streams) can still be using the old one
    GB_CUDA(cudaStreamSynchronize(s));
    if (gw.ready[buf ^ 1]) GB_CUDA(cudaEventSynchronize(gw.ready[buf ^ 1]));
    if (gw.consumed_valid[buf]) GB_CUDA(cudaEventSynchronize(gw.consumed[buf]));
    if (gw.x0[kind][buf]) cudaFree(gw.x0[kind][buf]);
    gw.x0[kind][buf] = nullptr; gw.cap[kind][buf] = 0;
    GB_CUDA(cudaMalloc(&gw.x0[kind][buf], need0));
    GB_CUDA(cudaMemsetAsync(gw.x0[kind][buf], 0, need0, s));
    gw.cap[kind][buf] = need0;
  }
  for (int i = 0; i < 2; i++) {
    if (!gw.ready[i]) GB_CUDA(cudaEventCreateWithFlags(&gw.ready[i], cudaEventDisableTiming));
    if (!gw.consumed[i]) GB_CUDA(cudaEventCreateWithFlags(&gw.consumed[i], cudaEventDisableTiming));
    if (!gw.started[i]) GB_CUDA(cudaEventCreateWithFlags(&gw.started[i], cudaEventDisableTiming));
  }
  {
    ProfScope ps(prof, "tc_build_pose_lists", s);
    launch_build_pose_lists(pb.rec_xyzr, pb.rec_ch, pb.n_rec, pb.lig_xyzr, pb.lig_ch, pb.lig_off, pb.centers, nb,
                            pb.dimension / 2.f, cap, gw.list_xyzr, gw.list_ch, gw.list_n, s, pb.rot);
  }
  static const int occ = getenv("GB_VOX_OCC") ? atoi(getenv("GB_VOX_OCC")) : 4;  // experiment: resident CTAs per SM
  auto kern = [&](bool kmax) {
    if (kmax) return occ >= 3 ? voxelize_pool_f16_kernel<true, 3> : voxelize_pool_f16_kernel<true, 2>;
    return occ >= 4 ? voxelize_pool_f16_kernel<false, 4> : occ == 3 ? voxelize_pool_f16_kernel<false, 3> : voxelize_pool_f16_kernel<false, 2>;
  };
  {
    // > 48 KB of dynamic shared memory: opt in once per device (the attribute is per device, not per process)
    std::lock_guard<std::mutex> lk(tc_init_mutex());
    static bool attr_set[64] = {};
    int dev = 0;
    GB_CUDA(cudaGetDevice(&dev));
    if (dev < 64 && !attr_set[dev]) {
      GB_CUDA(cudaFuncSetAttribute(kern(false), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(VoxSmem)));
      GB_CUDA(cudaFuncSetAttribute(kern(true), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(VoxSmem)));
      GB_CUDA(cudaFuncSetAttribute(voxelize_pool_f16_kernel<true, 3, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(VoxSmemT<6>)));
      attr_set[dev] = true;
    }
  }
  int launches = 1;
  // (r2m: a load-balanced variant -- every in-reach (atom, voxel) pair pushed into a per-warp queue, popped 32 at a time,
  // sums accumulated with integer shared-memory atomics in fixed point -- was correct and bit-reproducible but SLOWER, 6.18 vs
  // 5.15 ms per 10 k poses: 549 k instead of 464 k warp instructions per pose.  The 32 KB accumulator forced two channel halves,
  // each with its own partial queue drain, and the int -> half conversion of the output; profiles/README.md has the numbers.)
  auto launch_avg = [&](void* dst, int lp, int grp) {
    kern(false)<<<dim3(27, nb), 512, sizeof(VoxSmem), s>>>(gw.list_xyzr, gw.list_ch, gw.list_n, cap, pb.centers, pb.resolution,
                                                          pb.dimension, reinterpret_cast<uint4*>(dst), lp, grp);
  };
  if (kinds_mask & 1) {
    ProfScope ps(prof, "tc_voxelize_pool", s);
    launch_avg(gw.x0[0][buf], L1.Lp, 1);
    launches++;
  }
  if (kinds_mask & 4) {
    ProfScope ps(prof, "tc_voxelize_pool", s);
    launch_avg(gw.x0[2][buf], LF.Lp, LF.G);
    launches++;
  }
  if (kinds_mask & 2) {
    ProfScope ps(prof, "tc_voxelize_maxpool", s);
    kern(true)<<<dim3(27, nb), 512, sizeof(VoxSmem), s>>>(gw.list_xyzr, gw.list_ch, gw.list_n, cap, pb.centers, pb.resolution,
                                                         pb.dimension, reinterpret_cast<uint4*>(gw.x0[1][buf]), L1.Lp, 1);
    launches++;
  }
  if (kinds_mask & 8) {   // default2017: max pool, 35 channels in 6 chunks
    ProfScope ps(prof, "tc_voxelize_maxpool48", s);
    voxelize_pool_f16_kernel<true, 3, 6><<<dim3(27, nb), 512, sizeof(VoxSmemT<6>), s>>>(
        gw.list_xyzr, gw.list_ch, gw.list_n, cap, pb.centers, pb.resolution, pb.dimension, reinterpret_cast<uint4*>(gw.x0[3][buf]), L48.Lp, 1);
    launches++;
  }
  return launches;
}

}  // namespace gb
