// Fast path of the CNN forward for the default2018 family (N1 of SURVEY.md §8a), sm_100a only.
//
//   voxelise + avgpool (CUDA cores)  ->  X0   fp16, "chunk-planar padded" layout (below)
//   conv 3^3 28(32)->32 @24^3        ->  Y1   tcgen05 implicit GEMM (this file, conv3_tc_kernel)
//   conv 1^3 32->32 + ReLU + avgpool ->  X2   mma.sync pointwise + pooling + re-layout
//   conv 3^3 32->64 @12^3            ->  Y3   tcgen05
//   conv 1^3 64->64 + ReLU + avgpool ->  X4
//   conv 3^3 64->128 @6^3            ->  Y5   tcgen05
//   FC heads 27648 -> 3              ->  out3
//
// Activation layout consumed by the tcgen05 convolutions ("chunk-planar padded, grouped"):
//   X[group][x: D][c8: C/8][pos: Lp][8 channels]  fp16,   pos = q*P*P + yp*P + zp,  P = D+2,
//   q = pose within the group (G poses per group), (yp,zp) in [0,P) with a zero border, x planes unpadded.
// A 3x3x3 tap (dx,dy,dz) is then a PLANE shift (dx) plus a constant offset dy*P+dz of the flat position index, so
// the A operand of every tap is the same shared-memory slab addressed with a different descriptor start address
// (no-swizzle K-major canonical layout: 16-byte rows, SBO = 128 B, LBO = slab row pitch) — im2col costs nothing.
//
// conv3_tc_kernel: one CTA owns 128 consecutive flat positions (M = 128) of one pose group and one block of 32
// output channels, and marches over the D input planes.  For input plane xi it issues, per (dy,dz) tap and
// 16-channel K step, ONE tcgen05.mma with N = 96 whose B operand stacks the three dx taps: the 96 accumulator
// columns are the TMEM slots of output planes xi-1, xi, xi+1, i.e. the dx shift is done by WHERE the MMA
// accumulates, not by moving data.  (Cout = 32 alone would make the MMA shared-memory-bound on A: N = 96 cuts the
// A traffic per FLOP 3x.)  TMEM holds a ring of R plane slots x 32 fp32 columns; a slot is drained by the epilogue
// warps (tcgen05.ld -> bias -> ReLU -> fp16 -> global) as soon as its third input plane has been accumulated.
// Warp roles: warp 0 = bulk-copy (TMA unit) producer, warp 1 = MMA issuer, warps 2..5 = epilogue.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_fp16.h>
#include <vector>
#ifndef GB_TC_STAGES32
#define GB_TC_STAGES32 3
#endif
#include "gb_ptx.cuh"
#include "gb_tc.h"

namespace gb {

// ------------------------------------------------------------------------------------------------------------
ActLayout make_layout(int D, int G, int C) {
  ActLayout L;
  L.D = D; L.P = D + 2; L.G = G; L.C8 = C / 8;
  const int span = (G - 1) * L.P * L.P + (D - 1) * L.P + D;
  L.T = (span + 127) / 128;
  L.Lp = 128 * L.T + 2 * (L.P + 1);
  L.Lp = (L.Lp + 7) & ~7;
  return L;
}

TcWeights::~TcWeights() { for (void* p : allocs) cudaFree(p); }

template <typename T>
static T* tc_upload(std::vector<void*>& allocs, const std::vector<T>& h) {
  T* d = nullptr;
  GB_CUDA(cudaMalloc(&d, h.size() * sizeof(T)));
  allocs.push_back(d);
  GB_CUDA(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return d;
}
template <typename T>
static T* tc_upload(TcWeights& w, const std::vector<T>& h) { return tc_upload(w.allocs, h); }

// Pack a 3x3x3 kernel wfn(co, ci, kx, ky, kz) into the stacked-dx B operand layout [cout/32][9][cin/8][96] x 16 B.
ConvTc make_conv_tc(std::vector<void*>& allocs, int cout, int cin, const std::function<float(int, int, int, int, int)>& wfn,
                    const float* bias) {
  ConvTc c;
  GB_CHECK(cout % 32 == 0, "tc conv shape");
  c.cout = cout;
  c.cin = (cin + 15) / 16 * 16;
  const int C8 = c.cin / 8, NB = cout / 32;
  std::vector<__half> h((size_t)NB * 9 * C8 * 96 * 8, __float2half(0.f));
  for (int nb = 0; nb < NB; nb++)
    for (int ky = 0; ky < 3; ky++)
      for (int kz = 0; kz < 3; kz++)
        for (int c8 = 0; c8 < C8; c8++)
          for (int blk = 0; blk < 3; blk++) {
            const int kx = 2 - blk;  // blk 0 <-> dx=+1 (output plane xi-1), blk 2 <-> dx=-1 (output plane xi+1)
            for (int co = 0; co < 32; co++)
              for (int e = 0; e < 8; e++) {
                const int ci = c8 * 8 + e;
                if (ci >= cin) continue;
                h[((((size_t)nb * 9 + ky * 3 + kz) * C8 + c8) * 96 + blk * 32 + co) * 8 + e] =
                    __float2half(wfn(nb * 32 + co, ci, kx, ky, kz));
              }
          }
  c.wp = reinterpret_cast<uint4*>(tc_upload(allocs, h));
  std::vector<float> b(cout, 0.f);
  if (bias) b.assign(bias, bias + cout);
  c.bias = tc_upload(allocs, b);
  return c;
}

static ConvTc prep_conv(TcWeights& tw, const Model& m, const std::string& key) {
  const HostTensor& w = m.t(key + ".weight");
  const HostTensor& b = m.t(key + ".bias");
  const int cout = w.shape[0], cin = w.shape[1];
  GB_CHECK(w.shape[2] == 3, "tc conv shape");
  return make_conv_tc(tw.allocs, cout, cin, [&](int co, int ci, int kx, int ky, int kz) {
    return w.data[((((size_t)co * cin + ci) * 3 + kx) * 3 + ky) * 3 + kz];
  }, b.data);
}

static PointwiseTc prep_pw(TcWeights& tw, const Model& m, const std::string& key) {
  const HostTensor& w = m.t(key + ".weight");
  const HostTensor& b = m.t(key + ".bias");
  PointwiseTc p;
  p.c = w.shape[0];
  GB_CHECK(w.shape[0] == w.shape[1] && w.shape[2] == 1, "pointwise shape");
  std::vector<__half> h((size_t)p.c * p.c);
  for (size_t i = 0; i < h.size(); i++) h[i] = __float2half(w.data[i]);
  p.w = tc_upload(tw, h);
  p.bias = tc_upload(tw, std::vector<float>(b.data, b.data + b.nelem));
  return p;
}

bool tc_supported(const Model& m) {
  if (m.arch == GB_ARCH_DEFAULT2017) return m.n_channels > 32 && m.n_channels <= 48 && m.npts == 48;
  return (m.arch == GB_ARCH_DEFAULT2018 || m.arch == GB_ARCH_DENSE) && m.n_channels == 28 && m.npts == 48;
}
int tc_pool_kind(const Model& m) { return m.arch == GB_ARCH_DEFAULT2018 ? 0 : m.arch == GB_ARCH_DEFAULT2017 ? 3 : 1; }
bool tc_fused_enabled() {
  static const bool on = !(getenv("GB_TC_FUSED") && atoi(getenv("GB_TC_FUSED")) == 0);
  return on;
}
int tc_grid_kind(const Model& m, bool keep_activations) {
  if (m.arch == GB_ARCH_DEFAULT2018 && !keep_activations && tc_fused_enabled()) return 2;
  return tc_pool_kind(m);
}

std::mutex& tc_init_mutex() {
  static std::mutex mu;
  return mu;
}

std::shared_ptr<TcWeights> get_tc_weights(const Model& m) {
  std::lock_guard<std::mutex> lk(tc_init_mutex());
  Model& mm = const_cast<Model&>(m);
  if (mm.tc) return mm.tc;
  auto tw = std::make_shared<TcWeights>();
  if (m.arch == GB_ARCH_DEFAULT2017) {   // maxpool -> conv -> ReLU, three times; no 1x1x1 convolutions
    tw->conv1 = prep_conv(*tw, m, "unit1_conv1");
    tw->conv3 = prep_conv(*tw, m, "unit2_conv1");
    tw->conv5 = prep_conv(*tw, m, "unit3_conv1");
  } else {
  tw->conv1 = prep_conv(*tw, m, "unit1_conv");
  tw->pw2 = prep_pw(*tw, m, "unit2_conv");
  tw->pw2_packed = pack_pointwise_tc(tw->allocs, m.t("unit2_conv.weight").data, 32);
  pack_pair_weights(tw->allocs, tw->conv1.wp, tw->pw2_packed, &tw->pair_w, &tw->pair_w2);
  tw->conv3 = prep_conv(*tw, m, "unit3_conv");
  tw->pw4 = prep_pw(*tw, m, "unit4_conv");
  tw->conv5 = prep_conv(*tw, m, "unit5_conv");
  }
  // FC heads read the NCDHW flatten idx = c*216 + pos (view(-1, 27648)); the device keeps conv5's output
  // channels-last ([pos][c]), so permute the weights once.
  const HostTensor &pw = m.t("pose_output.weight"), &pb = m.t("pose_output.bias"), &aw = m.t("affinity_output.weight"),
                   &ab = m.t("affinity_output.bias");
  const int F = 27648;
  GB_CHECK(pw.shape[1] == F, "fc features");
  std::vector<float> fw((size_t)3 * F), fb(3);
  for (int r = 0; r < 3; r++) {
    const float* src = r < 2 ? pw.data + (size_t)r * F : aw.data;
    for (int c = 0; c < 128; c++)
      for (int pos = 0; pos < 216; pos++) fw[(size_t)r * F + pos * 128 + c] = src[c * 216 + pos];
  }
  fb[0] = pb.data[0]; fb[1] = pb.data[1]; fb[2] = ab.data[0];
  tw->fcw = tc_upload(*tw, fw);
  tw->fcb = tc_upload(*tw, fb);
  mm.tc = tw;
  return tw;
}

void TcWorkspace::ensure(int i, size_t bytes) {
  if (cap[i] >= bytes) return;
  if (buf[i]) cudaFree(buf[i]);
  buf[i] = nullptr;
  GB_CUDA(cudaMalloc(&buf[i], bytes));
  // zero borders of the padded layouts; interiors are rewritten per chunk.  The memset runs on the legacy default stream,
  // the kernels on the handle's NON-BLOCKING stream, which does not wait for it: without the synchronise below the memset
  // could land after the first kernels had written the buffer (seen as a rare wrong score on a handle's first call
  // when the allocator returned recycled memory, r2d).  Growth is rare; cudaMalloc / cudaFree synchronise anyway.
  GB_CUDA(cudaMemset(buf[i], 0, bytes));
  GB_CUDA(cudaStreamSynchronize(cudaStreamLegacy));
  cap[i] = bytes;
}
TcWorkspace::~TcWorkspace() {
  for (auto p : buf)
    if (p) cudaFree(p);
}

// ------------------------------------------------------------------------------------------------------------
// Kernel B: 3x3x3 convolution as a tcgen05 implicit GEMM (see the header comment).
struct ConvTcParams {
  const uint4* xin;   // [group][x][c8][Lp]
  const uint4* wp;    // [NB][9][C8][96]
  const float* bias;  // [Cout]
  __half* out;        // [pose][D][D][D][Cout]
  int D, P, G, T, NB, Lp, Cout, n_poses, relu, n_groups;
  int out_mode, out_c8tot, out_c8off, out_lp;  // out_mode 1: write chunk-planar into a block buffer (dense family)
  uint4* out_planar;
  int dbg;  // experiment switches (GB_TC_DBG): 1 = no global stores, 2 = no slab loads, 4 = no MMAs
};

// TMEM ring of plane slots x 32 fp32 columns (256 columns allocated): 8 slots, except D = 6 (conv5): there a 6-slot ring wraps
// exactly on item boundaries, where the window is truncated anyway, so no N = 96 MMA ever has to be split at a wrap (two MMAs
// that both read the 4 KB A tile; 2 of every 8 planes otherwise): 1.31 -> 1.25 ms per 10 k poses.  For D = 12 the same idea
// (2 of 12 planes split instead of 3) loses more to the shallower ring than it gains: 2.0 -> 2.4 ms (r3f).
template <int DD> struct TcRing { static constexpr int kSlots = (DD == 6) ? 6 : 8; };
constexpr int kSlabMax = 128 + 2 * 27;

template <int CIN>
struct ConvTcSmem {
  static constexpr int C8 = CIN / 8;
  static constexpr int kWBytes = 9 * C8 * 96 * 16;
  static constexpr int kStageBytes = ((C8 * kSlabMax * 16) + 127) / 128 * 128;
  // 3 stages for CIN = 32: two conv CTAs (2 x 91 KB) leave room for one voxeliser CTA (40 KB) on the same SM, so
  // the CUDA-core voxeliser of the next chunk can overlap the tensor-core network of the current one
  static constexpr int kStages = (CIN == 32) ? (GB_TC_STAGES32) : 4;
  static constexpr int kBarOff = kWBytes + kStages * kStageBytes;
  static constexpr int kTotal = kBarOff + 512;
};

// Per-MMA start-address offsets (16-byte units) of the A slab and the B weight block, one table per kernel
// configuration, in constant memory (uniform loads straight into the uniform registers UTCHMMA consumes).
struct MmaOff { uint32_t a, b; };
__constant__ MmaOff c_mma_off[5][36];
template <int CIN, int DD> struct ConvCfg {
  static constexpr int id = (CIN == 32 && DD == 24) ? 0 : (CIN == 32 && DD == 12) ? 1 : (CIN == 64 && DD == 6) ? 2 : (CIN == 48) ? 4 : 3;
};

template <int CIN, int DD>
__global__ void __launch_bounds__(192) conv3_tc_kernel(const ConvTcParams p) {
  using S = ConvTcSmem<CIN>;
  constexpr int C8 = S::C8;
  constexpr int R = TcRing<DD>::kSlots;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_w = smem;
  uint8_t* s_stage = smem + S::kWBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kBarOff);
  uint64_t* full = bars;                        // [S::kStages]
  uint64_t* empty = bars + S::kStages;           // [S::kStages]
  uint64_t* accf = bars + 2 * S::kStages;        // [R]
  uint64_t* acce = bars + 2 * S::kStages + R;    // [R]
  uint64_t* wbar = bars + 2 * S::kStages + 2 * R;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * S::kStages + 2 * R + 1);
  float* s_bias = reinterpret_cast<float*>(bars + 2 * S::kStages + 2 * R + 2);  // 32 floats

  // Persistent CTA: work items (pose group, tile, Cout block) are dealt round-robin; gridDim.x is a multiple of NB,
  // so a CTA keeps one Cout block (weights stay resident).  All pipelines (slab ring, TMEM slot ring) run on
  // counters that continue across items, so the next item's slabs stream in while the current one drains.
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nb = blockIdx.x % p.NB;
  const int n_items = p.n_groups * p.T * p.NB;
  constexpr int D = DD, P = DD + 2;
  constexpr int SL = 128 + 2 * (P + 1);
  const uint32_t slab_row = (uint32_t)SL * 16u;  // bytes between K chunks of the A slab

  if (threadIdx.x == 0) {
    for (int s = 0; s < S::kStages; s++) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
    for (int s = 0; s < R; s++) { ptx::mbar_init(&accf[s], 1); ptx::mbar_init(&acce[s], 128); }
    ptx::mbar_init(wbar, 1);
    ptx::fence_mbar_init();
  }
  if (threadIdx.x < 32) s_bias[threadIdx.x] = p.bias[nb * 32 + threadIdx.x];
  if (warp == 1) {
    ptx::tmem_alloc(s_tmem, 256);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp == 0) {
    // ===== producer: weights once, then one A slab per input plane (whole warp converged, one elected lane issues) =====
    {
      if (ptx::elect_one()) {
        ptx::mbar_expect_tx(wbar, S::kWBytes);
        const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.wp) + (size_t)nb * S::kWBytes;
        for (int t9 = 0; t9 < 9; t9++)
          ptx::bulk_g2s(s_w + t9 * (S::kWBytes / 9), wsrc + t9 * (S::kWBytes / 9), S::kWBytes / 9, wbar);
      }
      __syncwarp();
      uint32_t gp = 0;  // global plane counter of this CTA
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int j = (item / p.NB) % p.T, g = item / (p.NB * p.T);
        const uint4* xg = p.xin + (size_t)g * D * C8 * p.Lp + (size_t)128 * j;
        for (int it = 0; it < D; it++, gp++) {
          const uint32_t st = gp % S::kStages, ph = (gp / S::kStages) & 1;
          ptx::mbar_wait(&empty[st], ph ^ 1);
          if (ptx::elect_one()) {
            if (p.dbg & 2) {
              ptx::mbar_arrive(&full[st]);
            } else {
              ptx::mbar_expect_tx(&full[st], (uint32_t)C8 * slab_row);
              uint8_t* dst = s_stage + (size_t)st * S::kStageBytes;
#pragma unroll
              for (int c8 = 0; c8 < C8; c8++)
                ptx::bulk_g2s(dst + (size_t)c8 * slab_row, xg + ((size_t)it * C8 + c8) * p.Lp, slab_row, &full[st]);
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (whole warp converged; the tcgen05 instructions are issued by one elected lane) =====
    {
      ptx::mbar_wait(wbar, 0);
      // Descriptor = (lo, hi) words; hi is constant (SBO = 128 B, version 1); lo = start address | LBO << 16.  Per
      // MMA only the 14-bit start-address field changes, by COMPILE-TIME offsets (P and the slab pitch are template
      // constants), so the issue loop is one 32-bit add per operand plus the UTCHMMA.
      constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);
      const uint32_t a_lo_fixed = ((uint32_t)SL & 0x3FFFu) << 16;  // LBO = slab row (SL x 16 B)
      const uint32_t b_lo_base = (96u << 16) | (ptx::smem_u32(s_w) >> 4);  // LBO = 96 rows x 16 B
      uint32_t gp = 0, go_base = 0;  // global input-plane / output-plane counters (output plane xo <-> go_base+xo-1)
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, go_base += D) {
        for (int it = 0; it < D; it++, gp++) {
          const int xi = it + 1;
          const uint32_t st = gp % S::kStages, ph = (gp / S::kStages) & 1;
          const int lo = xi > 1 ? xi - 1 : 1, hi = xi < D ? xi + 1 : D;
          const int fresh_lo = xi == 1 ? 1 : xi + 1;  // output planes >= fresh_lo get their first contribution now
          for (int xo = fresh_lo; xo <= hi; xo++) {
            const uint32_t go = go_base + xo - 1, u = go / R;
            if (u > 0) ptx::mbar_wait(&acce[go % R], (u - 1) & 1);
          }
          // runs of output planes with consecutive TMEM slots: {tmem column, B row offset (16 B units), idesc}
          uint32_t r_tm[2], r_boff[2], r_idesc[2];
          int nr = 0;
          {
            int rs = lo;
            for (int xo = lo; xo <= hi; xo++) {
              const uint32_t sl = (go_base + xo - 1) % R;
              if (xo == hi || sl == R - 1) {
                r_tm[nr] = tmem_base + ((go_base + rs - 1) % R) * 32u;
                r_boff[nr] = (uint32_t)(rs - (xi - 1)) * 32u;
                r_idesc[nr] = ptx::idesc_f16(128, 32 * (xo - rs + 1));
                nr++;
                rs = xo + 1;
              }
            }
          }
          ptx::mbar_wait(&full[st], ph);
          ptx::tc_fence_after();
          const uint32_t a_lo_base = a_lo_fixed | (ptx::smem_u32(s_stage + (size_t)st * S::kStageBytes) >> 4);
          if (ptx::elect_one()) {
            if (!(p.dbg & 4)) {
            // very first MMA of the plane: fresh output planes are overwritten (accumulate = 0): one N=32 MMA per plane
            for (int xo = lo; xo <= hi; xo++) {
              const uint32_t tm = tmem_base + ((go_base + xo - 1) % R) * 32u;
              const uint32_t bl = b_lo_base + (uint32_t)(xo - (xi - 1)) * 32u;
              constexpr uint32_t a0 = (uint32_t)((P + 1) - P - 1);
              if (xo >= fresh_lo) ptx::mma_f16_ss_lohi<0>(tm, a_lo_base + a0, kDescHi, bl, kDescHi, ptx::idesc_f16(128, 32));
              else ptx::mma_f16_ss_lohi<1>(tm, a_lo_base + a0, kDescHi, bl, kDescHi, ptx::idesc_f16(128, 32));
            }
            // Operand offsets from a __constant__ table, rolled loop.  (r2i: a fully unrolled loop with immediate
            // offsets made conv3 / conv5 17 % SLOWER -- the issuing thread blocks in UTCHMMA while the MMA queue is full, so
            // instructions per MMA do not matter, code size does.  The clock64 timeline of r2j shows what bounds these
            // kernels: the tensor pipe's shared-memory operand reads -- 4 KB of A + 3 KB of B per N = 96 MMA at 128 B/clk.)
            constexpr int kCfg = ConvCfg<CIN, DD>::id;
            constexpr int kNumMma = 9 * (CIN / 16);
            if (nr == 1) {
              const uint32_t tm0 = r_tm[0], id0 = r_idesc[0], bl0 = b_lo_base + r_boff[0];
#pragma unroll 1
              for (int m = 1; m < kNumMma; m++) {
                const MmaOff o = c_mma_off[kCfg][m];
                ptx::mma_f16_ss_lohi<1>(tm0, a_lo_base + o.a, kDescHi, bl0 + o.b, kDescHi, id0);
              }
            } else {
              // TMEM ring wrap inside the window (2 planes in 8): two MMAs per tap
              const uint32_t tm0 = r_tm[0], id0 = r_idesc[0], bl0 = b_lo_base + r_boff[0];
              const uint32_t tm1 = r_tm[1], id1 = r_idesc[1], bl1 = b_lo_base + r_boff[1];
#pragma unroll 1
              for (int m = 1; m < kNumMma; m++) {
                const MmaOff o = c_mma_off[kCfg][m];
                ptx::mma_f16_ss_lohi<1>(tm0, a_lo_base + o.a, kDescHi, bl0 + o.b, kDescHi, id0);
                ptx::mma_f16_ss_lohi<1>(tm1, a_lo_base + o.a, kDescHi, bl1 + o.b, kDescHi, id1);
              }
            }
            }
            ptx::tc_commit(&empty[st]);                                    // slab consumed
            if (xi >= 2) ptx::tc_commit(&accf[(go_base + xi - 2) % R]);     // output plane xi-1 is complete
            if (xi == D) ptx::tc_commit(&accf[(go_base + D - 1) % R]);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ===== epilogue: TMEM -> registers -> bias/ReLU -> fp16 -> global =====
    const int q4 = warp & 3;  // TMEM lane quarter this warp may access
    const int row = q4 * 32 + lane;
    uint32_t go = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int j = (item / p.NB) % p.T, g = item / (p.NB * p.T);
      const int m = (P + 1) + 128 * j + row;
      const int qpose = m / (P * P), rem = m % (P * P);
      const int y = rem / P, z = rem % P;
      const int pose = g * p.G + qpose;
      const bool valid = qpose < p.G && pose < p.n_poses && y >= 1 && y <= D && z >= 1 && z <= D;
      __half* obase = p.out + (((size_t)pose * D * D + (size_t)(y - 1)) * D + (z - 1)) * p.Cout + nb * 32;
      const size_t plane_stride = (size_t)D * D * p.Cout;
      for (int xo = 1; xo <= D; xo++, go++) {
        const uint32_t slot = go % R, u = go / R;
        ptx::mbar_wait(&accf[slot], u & 1);
        ptx::tc_fence_after();
        uint32_t v[32];
        ptx::tmem_ld32(tmem_base + ((uint32_t)(q4 * 32) << 16) + slot * 32u, v);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        ptx::mbar_arrive(&acce[slot]);
        if (valid && p.out_mode == 1) {
          uint4 o[4];
          uint32_t* ow = reinterpret_cast<uint32_t*>(o);
#pragma unroll
          for (int c = 0; c < 16; c++) {
            float f0 = __uint_as_float(v[2 * c]) + s_bias[2 * c];
            float f1 = __uint_as_float(v[2 * c + 1]) + s_bias[2 * c + 1];
            if (p.relu) { f0 = fmaxf(f0, 0.f); f1 = fmaxf(f1, 0.f); }
            const __half2 h = __floats2half2_rn(f0, f1);
            ow[c] = *reinterpret_cast<const uint32_t*>(&h);
          }
          // same D, P, G as the input: the flat position index m is also the output position
          uint4* dst = p.out_planar + (((size_t)g * D + (xo - 1)) * p.out_c8tot + p.out_c8off + nb * 4) * p.out_lp + m;
#pragma unroll
          for (int c = 0; c < 4; c++) dst[(size_t)c * p.out_lp] = o[c];
        } else if (valid && !(p.dbg & 1)) {
          uint4 o[4];
          uint32_t* ow = reinterpret_cast<uint32_t*>(o);
#pragma unroll
          for (int c = 0; c < 16; c++) {
            float f0 = __uint_as_float(v[2 * c]) + s_bias[2 * c];
            float f1 = __uint_as_float(v[2 * c + 1]) + s_bias[2 * c + 1];
            if (p.relu) { f0 = fmaxf(f0, 0.f); f1 = fmaxf(f1, 0.f); }
            const __half2 h = __floats2half2_rn(f0, f1);
            ow[c] = *reinterpret_cast<const uint32_t*>(&h);
          }
          uint4* dst = reinterpret_cast<uint4*>(obase + (size_t)(xo - 1) * plane_stride);
          dst[0] = o[0]; dst[1] = o[1]; dst[2] = o[2]; dst[3] = o[3];
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 256);
  }
}

// ------------------------------------------------------------------------------------------------------------
// Kernel B, "v2" organisation (what r2y found for the fused conv1 kernel, applied to conv3 / conv5 of the scoring path):
// ONE CTA per SM with all 512 TMEM columns.  The ring has as many slots as an item has planes (12 or 6), so slot = plane
// and no window ever wraps: no N = 96 MMA is split in two.  The issuer is ONE thread inside one elected region, its
// per-window MMAs unrolled with compile-time operand offsets, and it takes the NEXT window's barrier waits after two
// thirds of the current window's MMAs (with a single CTA nobody else fills the tensor-core queue while it waits).
// Channels-last output with bias + ReLU only (the gradient / dense paths keep conv3_tc_kernel).
template <int CIN, int DD>
struct ConvTcV2Smem {
  static constexpr int C8 = CIN / 8;
  static constexpr int kWBytes = 9 * C8 * 96 * 16;
  static constexpr int kSL = 128 + 2 * (DD + 3);
  static constexpr int kStageBytes = ((C8 * kSL * 16) + 127) / 128 * 128;
  static constexpr int kStages = (CIN == 32) ? 6 : 4;
  static constexpr int kBarOff = kWBytes + kStages * kStageBytes;
  static constexpr int kTotal = kBarOff + 512;
};

template <int CIN, int DD>
__global__ void __launch_bounds__(192, 1) conv3_tc_v2_kernel(const ConvTcParams p) {
  using S = ConvTcV2Smem<CIN, DD>;
  constexpr int C8 = S::C8, R = DD, D = DD, P = DD + 2, SL = S::kSL;
  constexpr int kNumMma = 9 * (CIN / 16);
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_w = smem;
  uint8_t* s_stage = smem + S::kWBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kBarOff);
  uint64_t* full = bars;                          // [kStages]
  uint64_t* empty = full + S::kStages;            // [kStages]
  uint64_t* accf = empty + S::kStages;            // [R]
  uint64_t* acce = accf + R;                      // [R] one arrival per epilogue warp
  uint64_t* wbar = acce + R;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(wbar + 1);
  float* s_bias = reinterpret_cast<float*>(s_tmem + 2);   // 32 floats

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nb = blockIdx.x % p.NB;
  const int n_items = p.n_groups * p.T * p.NB;
  constexpr uint32_t slab_row = (uint32_t)SL * 16u;
  int n_my = 0;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) n_my++;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S::kStages; s++) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
    for (int s = 0; s < R; s++) { ptx::mbar_init(&accf[s], 1); ptx::mbar_init(&acce[s], 4); }
    ptx::mbar_init(wbar, 1);
    ptx::fence_mbar_init();
  }
  if (threadIdx.x < 32) s_bias[threadIdx.x] = p.bias[nb * 32 + threadIdx.x];
  if (warp == 1) {
    ptx::tmem_alloc(s_tmem, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp == 0) {
    // ===== producer =====
    if (ptx::elect_one()) {
      ptx::mbar_expect_tx(wbar, S::kWBytes);
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.wp) + (size_t)nb * S::kWBytes;
      for (int t9 = 0; t9 < 9; t9++) ptx::bulk_g2s(s_w + t9 * (S::kWBytes / 9), wsrc + t9 * (S::kWBytes / 9), S::kWBytes / 9, wbar);
    }
    __syncwarp();
    uint32_t gp = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int j = (item / p.NB) % p.T, g = item / (p.NB * p.T);
      const uint4* xg = p.xin + (size_t)g * D * C8 * p.Lp + (size_t)128 * j;
      for (int it = 0; it < D; it++, gp++) {
        const uint32_t st = gp % S::kStages, ph = (gp / S::kStages) & 1;
        ptx::mbar_wait(&empty[st], ph ^ 1);
        if (ptx::elect_one()) {
          ptx::mbar_expect_tx(&full[st], (uint32_t)C8 * slab_row);
          uint8_t* dst = s_stage + (size_t)st * S::kStageBytes;
#pragma unroll
          for (int c8 = 0; c8 < C8; c8++) ptx::bulk_g2s(dst + (size_t)c8 * slab_row, xg + ((size_t)it * C8 + c8) * p.Lp, slab_row, &full[st]);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: one thread =====
    if (ptx::elect_one()) {
      constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);
      const uint32_t a_lo_fixed = ((uint32_t)SL & 0x3FFFu) << 16;
      const uint32_t b_lo_base = (96u << 16) | (ptx::smem_u32(s_w) >> 4);
      ptx::mbar_wait(wbar, 0);
      // waits of window (item sequence number iq, input plane xi): the slab, and the slot of every output plane that gets
      // its first contribution there -- slot = plane, previous occupant = the same plane of the previous item
      auto wait_window = [&](const int xi, const uint32_t iq, const uint32_t gpw) {
        if (iq > 0) {
          if (xi < D) ptx::mbar_wait(&acce[xi], (iq - 1) & 1);       // plane xi + 1
          if (xi == 1) ptx::mbar_wait(&acce[0], (iq - 1) & 1);       // plane 1
        }
        ptx::mbar_wait(&full[gpw % S::kStages], (gpw / S::kStages) & 1);
      };
      uint32_t gp = 0;
      if (n_my > 0) wait_window(1, 0, 0);
      for (uint32_t iq = 0; iq < (uint32_t)n_my; iq++) {
#pragma unroll 1
        for (int xi = 1; xi <= D; xi++, gp++) {
          const uint32_t st = gp % S::kStages;
          ptx::tc_fence_after();
          const uint32_t a_lo_base = a_lo_fixed | (ptx::smem_u32(s_stage + (size_t)st * S::kStageBytes) >> 4);
          uint32_t tm, bl, idn;
          if (xi == 1) { tm = tmem_base; bl = b_lo_base + 32u; idn = ptx::idesc_f16(128, 64); }
          else if (xi == D) { tm = tmem_base + (uint32_t)(D - 2) * 32u; bl = b_lo_base; idn = ptx::idesc_f16(128, 64); }
          else { tm = tmem_base + (uint32_t)(xi - 2) * 32u; bl = b_lo_base; idn = ptx::idesc_f16(128, 96); }
          // per MMA m = tap * (CIN / 16) + k step: A start = flat offset of the tap + k step * 2 slab rows, B = weight block
          constexpr uint32_t a0 = 0;   // tap (dy, dz) = (-1, -1), k step 0: (P + 1) - P - 1
          if (xi == 1) ptx::mma_f16_ss_lohi<0>(tm, a_lo_base + a0, kDescHi, bl, kDescHi, idn);                 // planes 1, 2: fresh
          else if (xi == D) ptx::mma_f16_ss_lohi<1>(tm, a_lo_base + a0, kDescHi, bl, kDescHi, idn);
          else {
            ptx::mma_f16_ss_lohi<1>(tm, a_lo_base + a0, kDescHi, bl, kDescHi, ptx::idesc_f16(128, 64));         // planes xi - 1, xi
            ptx::mma_f16_ss_lohi<0>(tm + 64u, a_lo_base + a0, kDescHi, bl + 64u, kDescHi, ptx::idesc_f16(128, 32));   // plane xi + 1: fresh
          }
          constexpr int kSplit = (2 * kNumMma) / 3;
#pragma unroll
          for (int m = 1; m < kNumMma; m++) {
            if (m == kSplit) {
              // the next window's waits, behind a third of this window's MMAs still queued
              const bool last = xi == D && iq + 1 == (uint32_t)n_my;
              if (!last) {
                if (xi == D) wait_window(1, iq + 1, gp + 1);
                else wait_window(xi + 1, iq, gp + 1);
                ptx::tc_fence_after();
              }
            }
            constexpr int kKs = CIN / 16;
            const int t9 = m / kKs, ks = m % kKs;
            const uint32_t oa = (uint32_t)((P + 1) + (t9 / 3 - 1) * P + (t9 % 3 - 1) + 2 * ks * SL);
            const uint32_t ob = (uint32_t)((t9 * C8 + 2 * ks) * 96);
            ptx::mma_f16_ss_lohi<1>(tm, a_lo_base + oa, kDescHi, bl + ob, kDescHi, idn);
          }
          ptx::tc_commit(&empty[st]);
          if (xi >= 2) ptx::tc_commit(&accf[xi - 2]);     // output plane xi - 1 is complete
          if (xi == D) ptx::tc_commit(&accf[D - 1]);
        }
      }
    }
  } else {
    // ===== epilogue: TMEM -> registers -> bias / ReLU -> fp16 -> global (channels-last) =====
    const int q4 = warp & 3;
    const int row = q4 * 32 + lane;
    uint32_t iq = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, iq++) {
      const int j = (item / p.NB) % p.T, g = item / (p.NB * p.T);
      const int m = (P + 1) + 128 * j + row;
      const int qpose = m / (P * P), rem = m % (P * P);
      const int y = rem / P, z = rem % P;
      const int pose = g * p.G + qpose;
      const bool valid = qpose < p.G && pose < p.n_poses && y >= 1 && y <= D && z >= 1 && z <= D;
      __half* obase = p.out + (((size_t)pose * D * D + (size_t)(y - 1)) * D + (z - 1)) * p.Cout + nb * 32;
      const size_t plane_stride = (size_t)D * D * p.Cout;
      for (int xo = 1; xo <= D; xo++) {
        const uint32_t slot = (uint32_t)(xo - 1);
        ptx::mbar_wait(&accf[slot], iq & 1);
        ptx::tc_fence_after();
        uint32_t v[32];
        ptx::tmem_ld32(tmem_base + ((uint32_t)(q4 * 32) << 16) + slot * 32u, v);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&acce[slot]);
        if (valid) {
          uint4 o[4];
          uint32_t* ow = reinterpret_cast<uint32_t*>(o);
#pragma unroll
          for (int c = 0; c < 16; c++) {
            const float f0 = fmaxf(__uint_as_float(v[2 * c]) + s_bias[2 * c], 0.f);
            const float f1 = fmaxf(__uint_as_float(v[2 * c + 1]) + s_bias[2 * c + 1], 0.f);
            const __half2 h = __floats2half2_rn(f0, f1);
            ow[c] = *reinterpret_cast<const uint32_t*>(&h);
          }
          uint4* dst = reinterpret_cast<uint4*>(obase + (size_t)(xo - 1) * plane_stride);
          dst[0] = o[0]; dst[1] = o[1]; dst[2] = o[2]; dst[3] = o[3];
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

// GB_TC_CONV_V2 (read per call): bit 0 conv3, bit 1 conv5 through conv3_tc_v2_kernel (default both: 2.00 -> 1.81 and
// 1.25 -> 1.07 ms per 10 k poses, bit-identical results; 0 = conv3_tc_kernel for both)
static int tc_conv_v2() {
  const char* e = getenv("GB_TC_CONV_V2");
  return e ? atoi(e) : 3;
}

template <int CIN, int DD>
static void launch_conv_tc_v2(const ConvTc& c, const ActLayout& L, const uint4* xin, __half* out, int n_poses, cudaStream_t s) {
  using S = ConvTcV2Smem<CIN, DD>;
  GB_CHECK(c.cin == CIN && L.D == DD, "conv shape");
  static bool attr_set[64] = {};
  static int n_sm_dev[64] = {};
  int dev = 0;
  GB_CUDA(cudaGetDevice(&dev));
  GB_CHECK(dev >= 0 && dev < 64, "device index");
  {
    std::lock_guard<std::mutex> init_lock(tc_init_mutex());
    if (!attr_set[dev]) {
      GB_CUDA(cudaFuncSetAttribute(conv3_tc_v2_kernel<CIN, DD>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
      GB_CUDA(cudaDeviceGetAttribute(&n_sm_dev[dev], cudaDevAttrMultiProcessorCount, dev));
      attr_set[dev] = true;
    }
  }
  ConvTcParams p = {};
  p.xin = xin; p.wp = c.wp; p.bias = c.bias; p.out = out;
  p.D = L.D; p.P = L.P; p.G = L.G; p.T = L.T; p.NB = c.cout / 32; p.Lp = L.Lp; p.Cout = c.cout; p.n_poses = n_poses;
  p.relu = 1;
  p.n_groups = (n_poses + L.G - 1) / L.G;
  const int n_items = p.n_groups * L.T * p.NB;
  int grid = n_sm_dev[dev];
  grid -= grid % p.NB;                  // a CTA keeps one Cout block
  if (grid > n_items) grid = n_items;
  conv3_tc_v2_kernel<CIN, DD><<<grid, 192, S::kTotal, s>>>(p);
}

// ------------------------------------------------------------------------------------------------------------
// Kernel C: pointwise (1x1x1) convolution + bias + ReLU + 2x2x2 average pool + re-layout for the next conv.
// in : Y [pose][D][D][D][C] fp16 (already ReLU'd output of the preceding 3^3 conv)
// out: chunk-planar padded grouped layout with Dn = D/2.
// One warp handles 2 pooled voxels = 16 fine voxels as the M dimension of mma.sync.m16n8k16 (fp32 accumulate).
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Variant A (rows = voxels): one warp = 2 pooled voxels as the 16 rows of the MMA; faster for C = 32.
template <int C>
__global__ void __launch_bounds__(256) pointwise_pool_rows_kernel(const __half* __restrict__ yin, const __half* __restrict__ w,
                                                             const float* __restrict__ bias, uint4* __restrict__ xout,
                                                             int D, int n_poses, int Gn, int Lpn) {
  constexpr int NT = C / 8, KS = C / 16;
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int Dn = D / 2, Pn = Dn + 2, C8n = C / 8;
  uint32_t bf[NT][KS][2];
#pragma unroll
  for (int nt = 0; nt < NT; nt++)
#pragma unroll
    for (int ks = 0; ks < KS; ks++) {
      const __half* wr = w + (size_t)(nt * 8 + g) * C + ks * 16 + 2 * t;
      bf[nt][ks][0] = *reinterpret_cast<const uint32_t*>(wr);
      bf[nt][ks][1] = *reinterpret_cast<const uint32_t*>(wr + 8);
    }
  float bs[NT][2];
#pragma unroll
  for (int nt = 0; nt < NT; nt++) { bs[nt][0] = bias[nt * 8 + 2 * t]; bs[nt][1] = bias[nt * 8 + 2 * t + 1]; }

  const int n_pairs = n_poses * Dn * Dn * Dn / 2;  // < 2^31 for any chunk the handle allows
  const int warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int n_warps = gridDim.x * (blockDim.x >> 5);
  const int di = (g >> 2) & 1, dj = (g >> 1) & 1, dk = g & 1;
  for (int pair = warp_global; pair < n_pairs; pair += n_warps) {
    // pooled voxels 2*pair (rows 0-7) and 2*pair+1 (rows 8-15); Dn is even so both share pose, x, y
    const int pvA = 2 * pair;
    const int z0 = pvA % Dn;
    int r = pvA / Dn;
    const int y0 = r % Dn; r /= Dn;
    const int x0 = r % Dn;
    const int pose = r / Dn;
    const __half* rowA = yin + ((((size_t)pose * D + (2 * x0 + di)) * D + (2 * y0 + dj)) * D + (2 * z0 + dk)) * C;
    const __half* rowB = rowA + (size_t)2 * C;  // next pooled voxel in z: fine z + 2
    float acc[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; nt++) { acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < KS; ks++) {
      uint32_t a[4];
      a[0] = *reinterpret_cast<const uint32_t*>(rowA + ks * 16 + 2 * t);
      a[1] = *reinterpret_cast<const uint32_t*>(rowB + ks * 16 + 2 * t);
      a[2] = *reinterpret_cast<const uint32_t*>(rowA + ks * 16 + 8 + 2 * t);
      a[3] = *reinterpret_cast<const uint32_t*>(rowB + ks * 16 + 8 + 2 * t);
#pragma unroll
      for (int nt = 0; nt < NT; nt++) mma_16816(acc[nt], a, bf[nt][ks][0], bf[nt][ks][1]);
    }
    const int grp = pose / Gn, q = pose % Gn;
    const size_t posA = (size_t)q * Pn * Pn + (size_t)(y0 + 1) * Pn + (z0 + 1);
#pragma unroll
    for (int nt = 0; nt < NT; nt++) {
      float v[4];
#pragma unroll
      for (int e = 0; e < 4; e++) {
        float x = fmaxf(acc[nt][e] + bs[nt][e & 1], 0.f);
        x += __shfl_xor_sync(0xffffffffu, x, 4);
        x += __shfl_xor_sync(0xffffffffu, x, 8);
        x += __shfl_xor_sync(0xffffffffu, x, 16);
        v[e] = x * 0.125f;
      }
      if (g == 0) {
        uint32_t* base = reinterpret_cast<uint32_t*>(xout + (((size_t)grp * Dn + x0) * C8n + nt) * Lpn + posA);
        const __half2 hA = __floats2half2_rn(v[0], v[1]);
        const __half2 hB = __floats2half2_rn(v[2], v[3]);
        base[t] = *reinterpret_cast<const uint32_t*>(&hA);      // pooled voxel A
        base[4 + t] = *reinterpret_cast<const uint32_t*>(&hB);  // pooled voxel B = next position
      }
    }
  }
}

// Variant B, faster for C = 64.
// Transposed formulation: Y2^T[co][voxel] = W[co][ci] * Y1^T[ci][voxel]  (A = weights, resident in registers;
// B = 8 fine voxels of ONE pooled voxel, loaded straight from the channels-last input), so the 2x2x2 average is a
// sum over the N dimension of the accumulator fragment: c0+c1 in-thread, then two xor-shuffles over the 4 lanes of
// a quad — instead of three shuffles on every accumulator register.
template <int C>
__global__ void __launch_bounds__(256) pointwise_pool_kernel(const __half* __restrict__ yin, const __half* __restrict__ w,
                                                             const float* __restrict__ bias, __half* __restrict__ xout,
                                                             int D, int n_poses, int Gn, int Lpn) {
  constexpr int MT = C / 16, KS = C / 16;
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int Dn = D / 2, Pn = Dn + 2, C8n = C / 8;
  uint32_t af[MT][KS][4];  // A fragments of W: rows = output channels
#pragma unroll
  for (int mt = 0; mt < MT; mt++)
#pragma unroll
    for (int ks = 0; ks < KS; ks++) {
      const __half* w0 = w + (size_t)(mt * 16 + g) * C + ks * 16 + 2 * t;
      const __half* w1 = w0 + (size_t)8 * C;
      af[mt][ks][0] = *reinterpret_cast<const uint32_t*>(w0);
      af[mt][ks][1] = *reinterpret_cast<const uint32_t*>(w1);
      af[mt][ks][2] = *reinterpret_cast<const uint32_t*>(w0 + 8);
      af[mt][ks][3] = *reinterpret_cast<const uint32_t*>(w1 + 8);
    }
  float bs[MT][2];
#pragma unroll
  for (int mt = 0; mt < MT; mt++) { bs[mt][0] = bias[mt * 16 + g]; bs[mt][1] = bias[mt * 16 + g + 8]; }

  const int n_pv = n_poses * Dn * Dn * Dn;  // < 2^31 for any chunk size the handle allows
  const int warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int n_warps = gridDim.x * (blockDim.x >> 5);
  // B fragment column n = g is fine voxel g of the pooled voxel: (di,dj,dk) = bits of g
  const int di = (g >> 2) & 1, dj = (g >> 1) & 1, dk = g & 1;
  for (int pv = warp_global; pv < n_pv; pv += n_warps) {
    const int z0 = pv % Dn;
    int r = pv / Dn;
    const int y0 = r % Dn; r /= Dn;
    const int x0 = r % Dn;
    const int pose = r / Dn;
    const __half* row = yin + ((((size_t)pose * D + (2 * x0 + di)) * D + (2 * y0 + dj)) * D + (2 * z0 + dk)) * C;
    uint32_t bfr[KS][2];
#pragma unroll
    for (int ks = 0; ks < KS; ks++) {
      bfr[ks][0] = *reinterpret_cast<const uint32_t*>(row + ks * 16 + 2 * t);
      bfr[ks][1] = *reinterpret_cast<const uint32_t*>(row + ks * 16 + 8 + 2 * t);
    }
    const int grp = pose / Gn, q = pose % Gn;
    const size_t pos = (size_t)q * Pn * Pn + (size_t)(y0 + 1) * Pn + (z0 + 1);
#pragma unroll
    for (int mt = 0; mt < MT; mt++) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int ks = 0; ks < KS; ks++) mma_16816(acc, af[mt][ks], bfr[ks][0], bfr[ks][1]);
      // acc[0],acc[1]: channel mt*16+g, fine voxels 2t,2t+1 ; acc[2],acc[3]: channel mt*16+g+8
      float s0 = fmaxf(acc[0] + bs[mt][0], 0.f) + fmaxf(acc[1] + bs[mt][0], 0.f);
      float s1 = fmaxf(acc[2] + bs[mt][1], 0.f) + fmaxf(acc[3] + bs[mt][1], 0.f);
      s0 += __shfl_xor_sync(0xffffffffu, s0, 1);
      s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
      s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
      s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
      if (t == 0) {
        // chunk 2*mt holds channels mt*16 .. +7 (lane g writes channel g), chunk 2*mt+1 the next eight
        __half* o0 = xout + ((((size_t)grp * Dn + x0) * C8n + 2 * mt) * Lpn + pos) * 8 + g;
        o0[0] = __float2half(s0 * 0.125f);
        o0[(size_t)Lpn * 8] = __float2half(s1 * 0.125f);
      }
    }
  }
}

// Variant C (current): transposed formulation with 16-byte loads and the 2x2x2 average done by a second MMA.
//   * B operand = 16 fine voxels (two z-adjacent pooled voxels) straight from the channels-last input with ONE
//     16-byte load per lane and 32 channels: lane (g, t) reads channels [8t, 8t+8) of voxel g.  The reduction index
//     K may be permuted freely as long as A uses the same permutation: MMA k-slots (2t, 2t+1, 2t+8, 2t+9) of k-step
//     s carry channels 8t + 4s + (0, 1, 2, 3) (+32 for the second 16-byte load when C = 64); the weight fragments
//     are gathered with that permutation once, into registers.
//   * bias + ReLU on the accumulator fragments, which re-packed to fp16 ARE the A fragments (channels x 16 voxels) of
//     a second MMA against the constant pooling matrix S[voxel][pooled] = 1/8: no shuffles.
//   * D is a template parameter, so the index arithmetic is constant division.
template <int C, int D>
__global__ void __launch_bounds__(256) pointwise_pool_mma_kernel(const __half* __restrict__ yin, const __half* __restrict__ w,
                                                                 const float* __restrict__ bias, __half* __restrict__ xout,
                                                                 int n_poses, int Gn, int Lpn) {
  constexpr int MT = C / 16, KS = C / 16, LD = C / 32;  // LD: 16-byte loads per lane and voxel
  constexpr int Dn = D / 2, Pn = Dn + 2, C8n = C / 8;
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  uint32_t af[MT][KS][4];
#pragma unroll
  for (int mt = 0; mt < MT; mt++)
#pragma unroll
    for (int ks = 0; ks < KS; ks++) {
      const int ch = 32 * (ks >> 1) + 8 * t + 4 * (ks & 1);
      const __half* w0 = w + (size_t)(mt * 16 + g) * C + ch;
      const __half* w1 = w0 + (size_t)8 * C;
      af[mt][ks][0] = *reinterpret_cast<const uint32_t*>(w0);
      af[mt][ks][1] = *reinterpret_cast<const uint32_t*>(w1);
      af[mt][ks][2] = *reinterpret_cast<const uint32_t*>(w0 + 2);
      af[mt][ks][3] = *reinterpret_cast<const uint32_t*>(w1 + 2);
    }
  float bs[MT][2];
#pragma unroll
  for (int mt = 0; mt < MT; mt++) { bs[mt][0] = bias[mt * 16 + g]; bs[mt][1] = bias[mt * 16 + g + 8]; }
  // pooling matrix fragment: column n = g of S; rows k < 8 -> pooled voxel 0 (A), k >= 8 -> pooled voxel 1 (B)
  const uint32_t eighth2 = 0x30003000u;  // half2(0.125, 0.125)
  const uint32_t sb0 = g == 0 ? eighth2 : 0u, sb1 = g == 1 ? eighth2 : 0u;

  const int n_pairs = n_poses * (Dn * Dn * Dn / 2);
  const int warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int n_warps = gridDim.x * (blockDim.x >> 5);
  const int di = (g >> 2) & 1, dj = (g >> 1) & 1, dk = g & 1;
  const int voff = ((di * D + dj) * D + dk) * C + 8 * t;  // this lane's fine voxel inside the 2x2x2 window, channel chunk
  // software pipeline: the 16-byte loads of the NEXT pair are issued before the MMAs of the current one
  auto pair_base = [&](int pair) -> const __half* {
    const int pvA = 2 * pair;
    const int z0 = pvA % Dn;
    int r = pvA / Dn;
    const int y0 = r % Dn; r /= Dn;
    const int x0 = r % Dn;
    const int pose = r / Dn;
    return yin + ((((size_t)pose * D + 2 * x0) * D + 2 * y0) * D + 2 * z0) * C + voff;
  };
  constexpr bool kPrefetch = false;  // measured: prefetching the next pair costs occupancy and is slower (r1i)
  uint4 va[LD], vb[LD];
  if (kPrefetch && warp_global < n_pairs) {
    const __half* base = pair_base(warp_global);
#pragma unroll
    for (int l = 0; l < LD; l++) {
      va[l] = *reinterpret_cast<const uint4*>(base + 32 * l);           // pooled voxel A
      vb[l] = *reinterpret_cast<const uint4*>(base + 2 * C + 32 * l);   // pooled voxel B = fine z + 2
    }
  }
  for (int pair = warp_global; pair < n_pairs; pair += n_warps) {
    const int pvA = 2 * pair;
    const int z0 = pvA % Dn;
    int r = pvA / Dn;
    const int y0 = r % Dn; r /= Dn;
    const int x0 = r % Dn;
    const int pose = r / Dn;
    if (!kPrefetch) {
      const __half* base = pair_base(pair);
#pragma unroll
      for (int l = 0; l < LD; l++) {
        va[l] = *reinterpret_cast<const uint4*>(base + 32 * l);
        vb[l] = *reinterpret_cast<const uint4*>(base + 2 * C + 32 * l);
      }
    }
    uint4 na[LD], nb2[LD];
    const bool more = kPrefetch && pair + n_warps < n_pairs;
    if (more) {
      const __half* base = pair_base(pair + n_warps);
#pragma unroll
      for (int l = 0; l < LD; l++) {
        na[l] = *reinterpret_cast<const uint4*>(base + 32 * l);
        nb2[l] = *reinterpret_cast<const uint4*>(base + 2 * C + 32 * l);
      }
    }
    const int grp = pose / Gn, q = pose % Gn;
    const size_t pos = (size_t)q * Pn * Pn + (size_t)(y0 + 1) * Pn + (z0 + 1);
#pragma unroll
    for (int mt = 0; mt < MT; mt++) {
      float ca[4] = {0.f, 0.f, 0.f, 0.f}, cb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int ks = 0; ks < KS; ks++) {
        const uint32_t* pa = reinterpret_cast<const uint32_t*>(&va[ks >> 1]);
        const uint32_t* pb = reinterpret_cast<const uint32_t*>(&vb[ks >> 1]);
        mma_16816(ca, af[mt][ks], pa[2 * (ks & 1)], pa[2 * (ks & 1) + 1]);
        mma_16816(cb, af[mt][ks], pb[2 * (ks & 1)], pb[2 * (ks & 1) + 1]);
      }
      // ca: channels (mt*16 + g | + 8) x fine voxels (2t, 2t+1) of A; cb: same for B
      uint32_t a2[4];
      {
        const __half2 h0 = __floats2half2_rn(fmaxf(ca[0] + bs[mt][0], 0.f), fmaxf(ca[1] + bs[mt][0], 0.f));
        const __half2 h1 = __floats2half2_rn(fmaxf(ca[2] + bs[mt][1], 0.f), fmaxf(ca[3] + bs[mt][1], 0.f));
        const __half2 h2 = __floats2half2_rn(fmaxf(cb[0] + bs[mt][0], 0.f), fmaxf(cb[1] + bs[mt][0], 0.f));
        const __half2 h3 = __floats2half2_rn(fmaxf(cb[2] + bs[mt][1], 0.f), fmaxf(cb[3] + bs[mt][1], 0.f));
        a2[0] = *reinterpret_cast<const uint32_t*>(&h0); a2[1] = *reinterpret_cast<const uint32_t*>(&h1);
        a2[2] = *reinterpret_cast<const uint32_t*>(&h2); a2[3] = *reinterpret_cast<const uint32_t*>(&h3);
      }
      float pz[4] = {0.f, 0.f, 0.f, 0.f};
      mma_16816(pz, a2, sb0, sb1);
      // lanes t == 0: pz[0], pz[1] = channel mt*16+g of pooled A, B; pz[2], pz[3] = channel mt*16+8+g
      if (t == 0) {
        __half* o0 = xout + ((((size_t)grp * Dn + x0) * C8n + 2 * mt) * Lpn + pos) * 8 + g;
        o0[0] = __float2half(pz[0]);
        o0[8] = __float2half(pz[1]);
        o0[(size_t)Lpn * 8] = __float2half(pz[2]);
        o0[(size_t)Lpn * 8 + 8] = __float2half(pz[3]);
      }
    }
    if (more) {
#pragma unroll
      for (int l = 0; l < LD; l++) { va[l] = na[l]; vb[l] = nb2[l]; }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// Kernel D: FC heads on conv5's channels-last fp16 output [pose][216][128].
__global__ void __launch_bounds__(256) fc_heads_f16_kernel(const __half* __restrict__ y5, const float* __restrict__ w,
                                                           const float* __restrict__ bias, float* __restrict__ out3) {
  constexpr int F = 27648;
  __shared__ float red[3][8];
  const __half2* f = reinterpret_cast<const __half2*>(y5 + (size_t)blockIdx.x * F);
  float a0 = 0, a1 = 0, a2 = 0;
  for (int e = threadIdx.x; e < F / 2; e += 256) {
    const float2 x = __half22float2(f[e]);
    const float2 w0 = *reinterpret_cast<const float2*>(w + 2 * e);
    const float2 w1 = *reinterpret_cast<const float2*>(w + F + 2 * e);
    const float2 w2 = *reinterpret_cast<const float2*>(w + 2 * (size_t)F + 2 * e);
    a0 = fmaf(x.x, w0.x, fmaf(x.y, w0.y, a0));
    a1 = fmaf(x.x, w1.x, fmaf(x.y, w1.y, a1));
    a2 = fmaf(x.x, w2.x, fmaf(x.y, w2.y, a2));
  }
  for (int o = 16; o; o >>= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, o);
    a1 += __shfl_xor_sync(0xffffffffu, a1, o);
    a2 += __shfl_xor_sync(0xffffffffu, a2, o);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { red[0][warp] = a0; red[1][warp] = a1; red[2][warp] = a2; }
  __syncthreads();
  if (threadIdx.x < 3) {
    float v = bias[threadIdx.x];
    for (int q = 0; q < 8; q++) v += red[threadIdx.x][q];
    out3[(size_t)blockIdx.x * 3 + threadIdx.x] = v;
  }
}

// ------------------------------------------------------------------------------------------------------------
template <int CIN, int DD>
static void launch_conv_tc(const ConvTc& c, const ActLayout& L, const uint4* xin, __half* out, int n_poses, cudaStream_t s,
                           uint4* out_planar = nullptr, int out_c8tot = 0, int out_c8off = 0, int out_lp = 0, int relu = 1) {
  using S = ConvTcSmem<CIN>;
  GB_CHECK(c.cin == CIN && L.D == DD, "conv shape");
  // function attributes and occupancy are per DEVICE (one process may own handles on several GPUs)
  static bool attr_set[64] = {};
  static int ctas_per_sm_dev[64] = {}, n_sm_dev[64] = {};
  int dev = 0;
  GB_CUDA(cudaGetDevice(&dev));
  GB_CHECK(dev >= 0 && dev < 64, "device index");
  {
    std::lock_guard<std::mutex> init_lock(tc_init_mutex());
    if (!attr_set[dev]) {
      GB_CUDA(cudaFuncSetAttribute(conv3_tc_kernel<CIN, DD>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
      GB_CUDA(cudaDeviceGetAttribute(&n_sm_dev[dev], cudaDevAttrMultiProcessorCount, dev));
      GB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm_dev[dev], conv3_tc_kernel<CIN, DD>, 192, S::kTotal));
      MmaOff h[36] = {};
      const int P = DD + 2, SL = 128 + 2 * (P + 1), C8 = CIN / 8;
      int m = 0;
      for (int t9 = 0; t9 < 9; t9++)
        for (int ks = 0; ks < CIN / 16; ks++, m++) {
          h[m].a = (uint32_t)((P + 1) + (t9 / 3 - 1) * P + (t9 % 3 - 1) + 2 * ks * SL);
          h[m].b = (uint32_t)((t9 * C8 + 2 * ks) * 96);
        }
      GB_CUDA(cudaMemcpyToSymbol(c_mma_off, h, sizeof(h), sizeof(MmaOff) * 36 * ConvCfg<CIN, DD>::id));
      if (ctas_per_sm_dev[dev] < 1) ctas_per_sm_dev[dev] = 1;
      attr_set[dev] = true;
    }
  }
  const int ctas_per_sm = ctas_per_sm_dev[dev], n_sm = n_sm_dev[dev];
  ConvTcParams p;
  p.xin = xin; p.wp = c.wp; p.bias = c.bias; p.out = out;
  p.D = L.D; p.P = L.P; p.G = L.G; p.T = L.T; p.NB = c.cout / 32; p.Lp = L.Lp; p.Cout = c.cout; p.n_poses = n_poses;
  p.relu = relu;
  p.out_mode = out_planar ? 1 : 0; p.out_planar = out_planar; p.out_c8tot = out_c8tot; p.out_c8off = out_c8off; p.out_lp = out_lp;
  static const int dbg = getenv("GB_TC_DBG") ? atoi(getenv("GB_TC_DBG")) : 0;
  p.dbg = dbg;
  p.n_groups = (n_poses + L.G - 1) / L.G;
  const int n_items = p.n_groups * L.T * p.NB;
  int grid = n_sm * ctas_per_sm;
  static const int persist = getenv("GB_TC_PERSIST") ? atoi(getenv("GB_TC_PERSIST")) : (CIN == 64 ? 1 : 0);
  if (persist == 0) grid = n_items;            // experiment: one item per CTA
  else if (persist > 1) grid = n_sm * persist;  // experiment: force CTAs per SM
  grid -= grid % p.NB;                  // a CTA keeps one Cout block
  if (grid > n_items) grid = n_items;  // n_items is a multiple of NB
  conv3_tc_kernel<CIN, DD><<<grid, 192, S::kTotal, s>>>(p);
}

size_t act_bytes(const ActLayout& L, int n_poses) {
  const size_t n_groups = (n_poses + L.G - 1) / L.G;
  return (n_groups * L.group_u4() + 256) * sizeof(uint4);
}

struct TcDebug {
  const void* ptr[8];
  size_t bytes[8];
};
static thread_local TcDebug t_debug;
void tc_debug_set(int i, const void* p, size_t bytes) { t_debug.ptr[i] = p; t_debug.bytes[i] = bytes; }
void launch_conv_tc_32_24_planar(const ConvTc& c, const uint4* xin, uint4* xout, int out_c8tot, int out_c8off, int out_lp,
                                 int n_poses, cudaStream_t s) {
  launch_conv_tc<32, 24>(c, make_layout(24, 1, 32), xin, nullptr, n_poses, s, xout, out_c8tot, out_c8off, out_lp);
}
void launch_conv_tc_any(int cin, int D, const ConvTc& c, const ActLayout& L, const uint4* xin, __half* out, int n_poses,
                        cudaStream_t s, uint4* out_planar, int out_c8tot, int out_c8off, int out_lp, int relu) {
  if (cin == 48 && D == 24) launch_conv_tc<48, 24>(c, L, xin, out, n_poses, s, out_planar, out_c8tot, out_c8off, out_lp, relu);
  else if (cin == 32 && D == 24) launch_conv_tc<32, 24>(c, L, xin, out, n_poses, s, out_planar, out_c8tot, out_c8off, out_lp, relu);
  else if (cin == 32 && D == 12) launch_conv_tc<32, 12>(c, L, xin, out, n_poses, s, out_planar, out_c8tot, out_c8off, out_lp, relu);
  else if (cin == 64 && D == 6) launch_conv_tc<64, 6>(c, L, xin, out, n_poses, s, out_planar, out_c8tot, out_c8off, out_lp, relu);
  else if (cin == 64 && D == 12) launch_conv_tc<64, 12>(c, L, xin, out, n_poses, s, out_planar, out_c8tot, out_c8off, out_lp, relu);
  else throw Error(GB_ERR_INTERNAL, "no tensor-core conv instantiation for this shape");
}
const void* tc_debug_buffer(int i, size_t* bytes) {
  if (i < 0 || i >= 8) return nullptr;
  if (bytes) *bytes = t_debug.bytes[i];
  return t_debug.ptr[i];
}

// default2017: 2x2x2 MAX pool of a channels-last conv output (already ReLU'd) + re-layout into the chunk-planar padded
// input of the next convolution.  One thread per (pooled voxel, 8-channel chunk): eight 16-byte loads, one 16-byte store.
template <int C>
__global__ void __launch_bounds__(256) maxpool_relayout_kernel(const __half* __restrict__ yin, uint4* __restrict__ xout, int D,
                                                               int n_poses, int Gn, int Lpn) {
  constexpr int C8 = C / 8;
  const int Dn = D / 2, Pn = Dn + 2;
  const long long total = (long long)n_poses * Dn * Dn * Dn * C8;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(e % C8);
    long long r = e / C8;
    const int z0 = (int)(r % Dn); r /= Dn;
    const int y0 = (int)(r % Dn); r /= Dn;
    const int x0 = (int)(r % Dn);
    const int pose = (int)(r / Dn);
    __half2 m[4];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int xi = 2 * x0 + (k >> 2), yi = 2 * y0 + ((k >> 1) & 1), zi = 2 * z0 + (k & 1);
      const uint4 v = *reinterpret_cast<const uint4*>(yin + ((((size_t)pose * D + xi) * D + yi) * D + zi) * C + c8 * 8);
      const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
      for (int q = 0; q < 4; q++) m[q] = k == 0 ? h[q] : __hmax2(m[q], h[q]);
    }
    uint4 o;
    o.x = *reinterpret_cast<uint32_t*>(&m[0]); o.y = *reinterpret_cast<uint32_t*>(&m[1]);
    o.z = *reinterpret_cast<uint32_t*>(&m[2]); o.w = *reinterpret_cast<uint32_t*>(&m[3]);
    xout[(((size_t)(pose / Gn) * Dn + x0) * C8 + c8) * Lpn + (size_t)(pose % Gn) * Pn * Pn + (size_t)(y0 + 1) * Pn + (z0 + 1)] = o;
  }
}

// default2017 (N3 of SURVEY.md §8a): maxpool2 (in the voxeliser) -> conv3^3(35(48)->32)+ReLU -> maxpool2 -> conv3^3(32->64)+ReLU
// -> maxpool2 -> conv3^3(64->128)+ReLU -> FC heads, the three convolutions on tcgen05
static int tc_forward_2017(const Model& m, const TcPoseBatch& pb, const void* x0v, TcWorkspace& ws, float* out3, cudaStream_t s,
                           Profiler* prof, cudaEvent_t x0_consumed) {
  auto tw = get_tc_weights(m);
  const int nb = pb.n_poses, nb_alloc = std::max(nb, 64);
  const ActLayout L1 = make_layout(24, 1, 48), L3 = make_layout(12, 2, 32), L5 = make_layout(6, 2, 64);
  ws.ensure(0, (size_t)nb_alloc * 24 * 24 * 24 * 32 * sizeof(__half) + 1024);
  ws.ensure(1, act_bytes(L3, nb_alloc));
  ws.ensure(2, act_bytes(L5, nb_alloc));
  ws.ensure(3, (size_t)nb_alloc * 216 * 128 * sizeof(__half) + 1024);
  __half* Y = reinterpret_cast<__half*>(ws.buf[0]);
  uint4* X2 = reinterpret_cast<uint4*>(ws.buf[1]);
  uint4* X4 = reinterpret_cast<uint4*>(ws.buf[2]);
  __half* Y5 = reinterpret_cast<__half*>(ws.buf[3]);
  { ProfScope ps(prof, "tc17_conv1_3x3x3_35x32_d24", s); launch_conv_tc<48, 24>(tw->conv1, L1, reinterpret_cast<const uint4*>(x0v), Y, nb, s); }
  if (x0_consumed) GB_CUDA(cudaEventRecord(x0_consumed, s));
  { ProfScope ps(prof, "tc17_maxpool_d24", s); maxpool_relayout_kernel<32><<<148 * 8, 256, 0, s>>>(Y, X2, 24, nb, L3.G, L3.Lp); }
  { ProfScope ps(prof, "tc17_conv2_3x3x3_32x64_d12", s); launch_conv_tc<32, 12>(tw->conv3, L3, X2, Y, nb, s); }
  { ProfScope ps(prof, "tc17_maxpool_d12", s); maxpool_relayout_kernel<64><<<148 * 8, 256, 0, s>>>(Y, X4, 12, nb, L5.G, L5.Lp); }
  { ProfScope ps(prof, "tc17_conv3_3x3x3_64x128_d6", s); launch_conv_tc<64, 6>(tw->conv5, L5, X4, Y5, nb, s); }
  { ProfScope ps(prof, "tc_fc_heads", s); fc_heads_f16_kernel<<<nb, 256, 0, s>>>(Y5, tw->fcw, tw->fcb, out3); }
  return 6;
}

int tc_forward(const Model& m, const TcPoseBatch& pb, const void* x0v, TcWorkspace& ws, float* out3, cudaStream_t s,
               Profiler* prof, cudaEvent_t x0_consumed, bool keep_activations) {
  GB_CHECK(tc_supported(m), "model has no tensor-core path");
  if (m.arch == GB_ARCH_DENSE) return tc_forward_dense(m, pb, x0v, ws, out3, s, prof, x0_consumed);
  if (m.arch == GB_ARCH_DEFAULT2017) {
    GB_CHECK(!keep_activations, "the fast gradient path covers the default2018 family");
    return tc_forward_2017(m, pb, x0v, ws, out3, s, prof, x0_consumed);
  }
  auto tw = get_tc_weights(m);
  int launches = 0;
  const int nb = pb.n_poses;
  const int nb_alloc = std::max(nb, 64);  // see tc_prepare_grid
  const ActLayout L1 = make_layout(24, 1, 32), L3 = make_layout(12, 2, 32), L5 = make_layout(6, 2, 64);
  const uint4* x0 = reinterpret_cast<const uint4*>(x0v);
  // --- workspaces: 0 = Y (conv outputs, reused), 1 = X2, 2 = X4 ---
  ws.ensure(0, (size_t)nb_alloc * 24 * 24 * 24 * 32 * sizeof(__half) + 1024);
  ws.ensure(1, act_bytes(L3, nb_alloc));
  ws.ensure(2, act_bytes(L5, nb_alloc));
  ws.ensure(3, (size_t)nb_alloc * 216 * 128 * sizeof(__half) + 1024);
  __half* Y = reinterpret_cast<__half*>(ws.buf[0]);
  uint4* X2 = reinterpret_cast<uint4*>(ws.buf[1]);
  uint4* X4 = reinterpret_cast<uint4*>(ws.buf[2]);
  __half* Y5 = reinterpret_cast<__half*>(ws.buf[3]);
  __half* Y3 = Y;  // conv3's output reuses conv1's buffer unless the backward pass needs both
  if (keep_activations) {
    GB_CHECK(m.arch == GB_ARCH_DEFAULT2018, "activations are kept for the default2018 family only");
    ws.ensure(7, (size_t)nb_alloc * 12 * 12 * 12 * 64 * sizeof(__half) + 1024);
    Y3 = reinterpret_cast<__half*>(ws.buf[7]);
  }
  const int pw_blocks = 148 * 8;
  const bool fused = tc_grid_kind(m, keep_activations) == 2;
  if (fused) {
    // unit1_conv + ReLU + unit2_conv + ReLU + avg-pool in one kernel (x0 is in the row-group layout): straight to X2
    ProfScope ps(prof, "tc_conv1_pw2_pool_fused", s);
    launch_conv1_pw2_pool(tw->conv1, tw->pw2.w, tw->pw2_packed, tw->pw2.bias, tw->pair_w, tw->pair_w2, x0, make_fused_x0_layout(), X2, L3, nb, s);
  } else {
    ProfScope ps(prof, "tc_conv1_3x3x3_28x32_d24", s);
    launch_conv_tc<32, 24>(tw->conv1, L1, x0, Y, nb, s);
  }
  if (x0_consumed) GB_CUDA(cudaEventRecord(x0_consumed, s));
  if (!fused) {
    ProfScope ps(prof, "tc_pw2_pool", s);
    static const int pw_variant = getenv("GB_TC_PW") ? atoi(getenv("GB_TC_PW")) : 2;
    if (pw_variant == 2)
      pointwise_pool_mma_kernel<32, 24><<<pw_blocks, 256, 0, s>>>(Y, tw->pw2.w, tw->pw2.bias, reinterpret_cast<__half*>(X2), nb,
                                                                  L3.G, L3.Lp);
    else
      pointwise_pool_rows_kernel<32><<<pw_blocks, 256, 0, s>>>(Y, tw->pw2.w, tw->pw2.bias, X2, 24, nb, L3.G, L3.Lp);
  }
  {
    ProfScope ps(prof, "tc_conv3_3x3x3_32x64_d12", s);
    if (tc_conv_v2() & 1) launch_conv_tc_v2<32, 12>(tw->conv3, L3, X2, Y3, nb, s);
    else launch_conv_tc<32, 12>(tw->conv3, L3, X2, Y3, nb, s);
  }
  {
    ProfScope ps(prof, "tc_pw4_pool", s);
    static const int pw_variant = getenv("GB_TC_PW") ? atoi(getenv("GB_TC_PW")) : 2;
    if (pw_variant == 2)
      pointwise_pool_mma_kernel<64, 12><<<pw_blocks, 256, 0, s>>>(Y3, tw->pw4.w, tw->pw4.bias, reinterpret_cast<__half*>(X4), nb,
                                                                  L5.G, L5.Lp);
    else
      pointwise_pool_kernel<64><<<pw_blocks, 256, 0, s>>>(Y3, tw->pw4.w, tw->pw4.bias, reinterpret_cast<__half*>(X4), 12, nb, L5.G, L5.Lp);
  }
  {
    ProfScope ps(prof, "tc_conv5_3x3x3_64x128_d6", s);
    if (tc_conv_v2() & 2) launch_conv_tc_v2<64, 6>(tw->conv5, L5, X4, Y5, nb, s);
    else launch_conv_tc<64, 6>(tw->conv5, L5, X4, Y5, nb, s);
  }
  {
    ProfScope ps(prof, "tc_fc_heads", s);
    fc_heads_f16_kernel<<<nb, 256, 0, s>>>(Y5, tw->fcw, tw->fcb, out3);
  }
  launches += fused ? 5 : 6;
  t_debug.ptr[0] = x0v;      t_debug.bytes[0] = act_bytes(fused ? make_fused_x0_layout() : L1, nb);
  t_debug.ptr[1] = Y3;       t_debug.bytes[1] = (size_t)nb * 1728 * 64 * sizeof(__half);  // holds Y3 after the pass
  t_debug.ptr[2] = X2;       t_debug.bytes[2] = act_bytes(L3, nb);
  t_debug.ptr[3] = X4;       t_debug.bytes[3] = act_bytes(L5, nb);
  t_debug.ptr[4] = Y5;       t_debug.bytes[4] = (size_t)nb * 216 * 128 * sizeof(__half);
  return launches;
}

}  // namespace gb
