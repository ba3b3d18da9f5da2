// Scoring path of the default2018 family: unit1_conv (3x3x3, 28(32) -> 32 @ 24^3) + ReLU + unit2_conv (1x1x1, 32 -> 32)
// + ReLU + 2x2x2 average pool in ONE tcgen05 kernel (N1 of SURVEY.md §8a; the TorchScript graph executed at
// lib/torch_model.cpp:185).  The 3x3x3 convolution is the implicit GEMM of gb_cnn_tc.cu (three dx taps stacked into
// N = 96, plane-marching TMEM ring); what is new is the tile shape and everything after the accumulator:
//
//   M tile = 16 rows (y) x 8 (z) of one x plane, instead of 128 consecutive flat positions.  The rows of a group of
//   8 poses are stacked (row R = q * 26 + yp, 26 = 24 + 2 zero rows per pose) and tiles start at ODD rows, so the two
//   rows and the two columns of every 2x2 pooling window always lie in the same tile -- in fact in the same warp of
//   the epilogue (partners are lanes ^1 and ^8).  In shared memory the A slab is a dense [chunk][18 rows][10 z][8 ch]
//   box: MMA row group g (8 consecutive z) sits at g * 160 B, so the UMMA descriptor simply uses SBO = 160 B and a tap
//   (dy, dz) is again a constant start-address offset, (1+dy) * 10 + (1+dz) elements.  The box is fetched by ONE
//   tensor-map TMA instruction per input plane (cp.async.bulk.tensor.4d, zero fill outside the tensor): this is the
//   "TMA im2col" of the north star -- the halo is part of the box, nothing is materialised.  Valid rows: 24 of every 26
//   (92 %; the flat tiling had 576 / 640 = 90 %).
//
//   After the accumulator (default, variant B of the template parameter below): unit2_conv is a SECOND tcgen05.mma.  The
//   epilogue warps turn an accumulator plane into relu(conv + bias) fp16 and store it as a K-major A tile in shared memory
//   (double buffered); the MMA warp issues D2 = A2 x W2^T (M = 128, N = 32, K = 32) on a fixed schedule -- the pointwise
//   MMA of output plane j enters the queue right before the convolution MMAs of the input plane two commits later --
//   and the epilogue runs one plane behind itself: stage plane j, then finish plane j - 1 (bias, ReLU, keep; after the
//   second plane of an (odd, even) x pair the 2x2x2 average = two shfl.xor levels, lanes ^1 and ^8, plus the kept plane)
//   and write the 12^3 input of unit3_conv with 16-byte stores.  (r2c: issuing the pointwise MMA opportunistically
//   queued it behind up to three planes of convolution MMAs -- 62 ms instead of 7.7 per 10 k poses.)
//   Variant A keeps the pointwise stage inside the epilogue warps on mma.sync (the formulation of
//   pointwise_pool_mma_kernel, operands from an XOR-swizzled shared-memory tile); bit-identical scores, 9.9 ms: HMMA
//   contends with tcgen05 for the tensor pipe.  It stays as the cross-check of variant B (GB_TC_FUSED_PW=0).
//   Y1 (0.88 MB per pose) never exists in HBM and the pointwise kernel is gone (round 1: 11 % of the step).
//
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..5 = epilogue.
//
// This file holds two organisations of that arithmetic: conv1_pw2_pool_kernel (two CTAs per SM, 256 TMEM columns each; round 2's
// first version, 6.9 ms per 10 k poses) and conv1_pw2_pool_v2_kernel further down (one CTA per SM, all 512 TMEM columns, ghost
// slots, single-thread issuers, two epilogue teams: 5.4 ms; the default -- and, as a template variant, the same on CTA pairs with
// tcgen05 cta_group::2, which is slower).  launch_conv1_pw2_pool() picks by GB_TC_FUSED_V2; the tests run all of them.
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_fp16.h>
#include <map>
#include <mutex>
#include <type_traits>
#include "gb_ptx.cuh"
#include "gb_tc.h"

namespace gb {

namespace {

constexpr int kG = kFusedGroup;       // poses per row group
constexpr int kD = 24, kP = 26;
constexpr int kRows = kG * kP;        // 208 stacked rows per group
constexpr int kRowTiles = (kRows + 15) / 16;   // 13
constexpr int kZBlocks = 3;
constexpr int kSlabRows = 18, kSlabZ = 10;
constexpr int kChunkBytes = kSlabRows * kSlabZ * 16;   // 2880: one 8-channel chunk of the slab = LBO
constexpr int kStageBytes = 4 * kChunkBytes;           // 11520
constexpr int kStages = 3;
constexpr int kWBytes = 9 * 4 * 96 * 16;               // 55296
constexpr int kYBytes = 2 * 128 * 64;                  // 16384: variant A [plane parity][row][32 ch fp16] (chunks swizzled),
                                                       //        variant B 2 x [chunk][row][8 ch] = double-buffered A operand
constexpr int kW2Bytes = 4 * 32 * 16;                  // 2048  [chunk][co][8 ci] (variant B)
constexpr int kOffStage = kWBytes;
constexpr int kOffY = kOffStage + kStages * kStageBytes;
constexpr int kOffW2 = kOffY + kYBytes;
constexpr int kOffBar = kOffW2 + kW2Bytes;
constexpr int kSmemTotal = kOffBar + 512;
// The 1x1x1 convolution after the accumulator comes in two variants (template parameter kTcPw):
//   A (false): mma.sync from shared memory inside the epilogue warps; TMEM ring of 8 plane slots
//   B (true) : a second tcgen05.mma (D2 = relu(conv) x W2^T, 32 more TMEM columns, double buffered) issued by the MMA warp on
//              a FIXED schedule -- the pointwise MMA of output plane j goes into the queue right before the convolution
//              MMAs of the input plane two commits later, when plane j's accumulator has long been complete and staged --
//              and an epilogue that runs one plane behind itself (stage plane j, then finish plane j - 1), so neither side
//              ever waits for the other's latency.  The pointwise accumulator of plane j is written INTO plane j's own ring
//              slot (the epilogue has drained it by then), so all 256 TMEM columns form a ring of 8 plane slots: with 24
//              planes per item the ring wraps at item boundaries and an N = 96 MMA has to be split at a wrap (two MMAs that
//              both read the 4 KB A tile) for 4 of 24 planes instead of 8 of 24 with a 6-slot ring + separate D2 columns.
template <bool kTcPw> struct Var { static constexpr int kR = 8; };

struct FusedParams {
  const uint4* wp;     // conv1: [9][4][96] x 16 B (make_conv_tc)
  const float* bias1;  // [32]
  const __half* w2;    // pointwise: [co][ci] row-major fp16 (variant A)
  const uint4* w2p;    // pointwise: [ci / 8][co][ci % 8] fp16 = K-major B operand (variant B)
  const float* bias2;  // [32]
  __half* xout;        // X2: chunk-planar D = 12 (make_layout(12, Gn, 32))
  int out_lp, out_G, n_poses, n_groups;
  // timeline instrumentation (GB_TC_FUSED_TRACE=<file>): clock64 stamps of CTA 0's producer / MMA / epilogue warps for
  // the first kTracePlanes planes: [role 3][plane][8]
  unsigned long long* trace;
};
constexpr int kTracePlanes = 96;
__device__ __forceinline__ void tr(unsigned long long* t, int role, uint32_t plane, int k) {
  if (t && blockIdx.x == 0 && plane < (uint32_t)kTracePlanes && (threadIdx.x & 31) == 0) t[(role * kTracePlanes + plane) * 8 + k] = clock64();
}

__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// per MMA m = tap * 2 + k step (tap = (dy+1) * 3 + (dz+1)): start-address offsets of the A slab and of the B weight block,
// in 16-byte units -- compile-time constants of the tile geometry
__host__ __device__ constexpr uint32_t off_a(int m) { return (uint32_t)(((m >> 1) / 3) * kSlabZ + ((m >> 1) % 3) + 2 * (m & 1) * (kChunkBytes >> 4)); }
__host__ __device__ constexpr uint32_t off_b(int m) { return (uint32_t)(((m >> 1) * 4 + 2 * (m & 1)) * 96); }

template <bool kTcPw>
__global__ void __launch_bounds__(192) conv1_pw2_pool_kernel(const __grid_constant__ CUtensorMap tmap, const FusedParams p) {
  constexpr int kR = Var<kTcPw>::kR;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_w = smem;
  uint8_t* s_stage = smem + kOffStage;
  uint8_t* s_y = smem + kOffY;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* full = bars;                  // [kStages]
  uint64_t* empty = bars + kStages;       // [kStages]
  uint64_t* accf = bars + 2 * kStages;    // [kR]  conv plane complete
  uint64_t* acce = accf + kR;             // [kR]  conv plane drained (128 arrivals)
  uint64_t* wbar = acce + kR;             // weights landed
  uint64_t* a2_full = wbar + 1;           // [2] variant B: a plane is staged for the pointwise MMA (128 arrivals)
  uint64_t* d2_full = a2_full + 2;        // [2] variant B: pointwise accumulator complete
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(d2_full + 2);
  float* s_bias = reinterpret_cast<float*>(s_tmem + 2);   // bias1[32], bias2[32]
  uint8_t* s_w2 = smem + kOffW2;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = p.n_groups * kRowTiles * kZBlocks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; s++) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
    for (int s = 0; s < kR; s++) { ptx::mbar_init(&accf[s], 1); ptx::mbar_init(&acce[s], 128); }
    ptx::mbar_init(wbar, 1);
    for (int b = 0; b < 2; b++) { ptx::mbar_init(&a2_full[b], 128); ptx::mbar_init(&d2_full[b], 1); }
    ptx::fence_mbar_init();
  }
  if (threadIdx.x < 32) s_bias[threadIdx.x] = p.bias1[threadIdx.x];
  else if (threadIdx.x < 64) s_bias[threadIdx.x] = p.bias2[threadIdx.x - 32];
  if (warp == 1) {
    ptx::tmem_alloc(s_tmem, 256);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp == 0) {
    // ===== producer: weights once (bulk copies), then one tensor-map TMA box per input plane =====
    if (ptx::elect_one()) {
      ptx::prefetch_tmap(&tmap);
      ptx::mbar_expect_tx(wbar, kWBytes + (kTcPw ? kW2Bytes : 0));
      for (int t9 = 0; t9 < 9; t9++)
        ptx::bulk_g2s(s_w + t9 * (kWBytes / 9), reinterpret_cast<const uint8_t*>(p.wp) + t9 * (kWBytes / 9), kWBytes / 9, wbar);
      if (kTcPw) ptx::bulk_g2s(s_w2, p.w2p, kW2Bytes, wbar);
    }
    __syncwarp();
    uint32_t gp = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int zb = item % kZBlocks, k = (item / kZBlocks) % kRowTiles, g = item / (kZBlocks * kRowTiles);
      for (int it = 0; it < kD; it++, gp++) {
        const uint32_t st = gp % kStages, ph = (gp / kStages) & 1;
        tr(p.trace, 0, gp, 0);
        ptx::mbar_wait(&empty[st], ph ^ 1);
        tr(p.trace, 0, gp, 1);
        if (ptx::elect_one()) {
          ptx::mbar_expect_tx(&full[st], kStageBytes);
          // box = {10 z x 8 ch, 18 rows, 4 chunks, 1 plane} at (z0 - 1, R0 - 1) = (8 zb, 16 k); rows past the group are zero-filled
          ptx::tma_load_4d(s_stage + (size_t)st * kStageBytes, &tmap, 64 * zb, 16 * k, 0, g * kD + it, &full[st]);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    constexpr uint32_t kDescHiA = (uint32_t)(kSlabZ * 16 >> 4) | (1u << 14);   // SBO = 160 B between 8-row groups
    constexpr uint32_t kDescHiB = (128u >> 4) | (1u << 14);
    const uint32_t a_lo_fixed = ((uint32_t)(kChunkBytes >> 4)) << 16;         // LBO = 2880 B between K chunks
    const uint32_t b_lo_base = (96u << 16) | (ptx::smem_u32(s_w) >> 4);
    auto wait_service = [&](uint64_t* bar, uint32_t parity) { ptx::mbar_wait(bar, parity); };
    wait_service(wbar, 0);
    // variant B: pointwise MMA of output plane j: D2[j & 1] = A2[j & 1] (128 x 32) x W2^T, two K = 16 steps
    const uint32_t w2_lo = ((uint32_t)(512 >> 4) << 16) | (ptx::smem_u32(s_w2) >> 4);   // LBO = 32 rows x 16 B
    uint32_t pw_issued = 0, committed = 0;
    auto issue_pw_until = [&](uint32_t upto) {
      for (; pw_issued < upto; pw_issued++) {
        const uint32_t b = pw_issued & 1;
        ptx::mbar_wait(&a2_full[b], (pw_issued >> 1) & 1);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint32_t a2_lo = ((uint32_t)(2048 >> 4) << 16) | (ptx::smem_u32(s_y + b * 8192) >> 4);   // LBO = 128 rows x 16 B
          const uint32_t tm_d2 = tmem_base + (pw_issued % kR) * 32u;   // plane j's own slot, drained before a2_full[b] completed
          ptx::mma_f16_ss_lohi<0>(tm_d2, a2_lo, kDescHiB, w2_lo, kDescHiB, ptx::idesc_f16(128, 32));
          ptx::mma_f16_ss_lohi<1>(tm_d2, a2_lo + 2 * (2048 >> 4), kDescHiB, w2_lo + 2 * (512 >> 4), kDescHiB, ptx::idesc_f16(128, 32));
          ptx::tc_commit(&d2_full[b]);
        }
        __syncwarp();
      }
    };
    uint32_t gp = 0, go_base = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, go_base += kD) {
      for (int it = 0; it < kD; it++, gp++) {
        // planes committed at least one input plane ago are complete (or about to be) and staged: their pointwise MMAs go
        // into the queue now, ahead of this input plane's convolution MMAs; the most recent commit is left for next time
        tr(p.trace, 1, gp, 0);
        if (kTcPw && committed > 0) issue_pw_until(committed - 1);
        tr(p.trace, 1, gp, 1);
        const int xi = it + 1;
        const uint32_t st = gp % kStages, ph = (gp / kStages) & 1;
        const int lo = xi > 1 ? xi - 1 : 1, hi = xi < kD ? xi + 1 : kD;
        const int fresh_lo = xi == 1 ? 1 : xi + 1;  // output planes >= fresh_lo get their first contribution now
        for (int xo = fresh_lo; xo <= hi; xo++) {
          const uint32_t go = go_base + xo - 1, u = go / kR;
          if (u > 0) wait_service(&acce[go % kR], (u - 1) & 1);
        }
        uint32_t r_tm[2], r_boff[2], r_idesc[2];
        int nr = 0;
        {
          int rs = lo;
          for (int xo = lo; xo <= hi; xo++) {
            const uint32_t sl = (go_base + xo - 1) % kR;
            if (xo == hi || sl == kR - 1) {
              r_tm[nr] = tmem_base + ((go_base + rs - 1) % kR) * 32u;
              r_boff[nr] = (uint32_t)(rs - (xi - 1)) * 32u;
              r_idesc[nr] = ptx::idesc_f16(128, 32 * (xo - rs + 1));
              nr++;
              rs = xo + 1;
            }
          }
        }
        tr(p.trace, 1, gp, 2);
        wait_service(&full[st], ph);
        tr(p.trace, 1, gp, 3);
        ptx::tc_fence_after();
        const uint32_t a_lo_base = a_lo_fixed | (ptx::smem_u32(s_stage + (size_t)st * kStageBytes) >> 4);
        if (ptx::elect_one()) {
          // first MMA of the plane (tap 0, k step 0): fresh output planes are overwritten (accumulate = 0), one N = 32 MMA each
          for (int xo = lo; xo <= hi; xo++) {
            const uint32_t tm = tmem_base + ((go_base + xo - 1) % kR) * 32u;
            const uint32_t bl = b_lo_base + (uint32_t)(xo - (xi - 1)) * 32u;
            if (xo >= fresh_lo) ptx::mma_f16_ss_lohi<0>(tm, a_lo_base, kDescHiA, bl, kDescHiB, ptx::idesc_f16(128, 32));
            else ptx::mma_f16_ss_lohi<1>(tm, a_lo_base, kDescHiA, bl, kDescHiB, ptx::idesc_f16(128, 32));
          }
          // the issue loop stays inside ONE elected region (ptxas then keeps UTCHMMA on the uniform datapath without
          // re-electing per instruction; r2d: electing per MMA cost ~1 ms per 10 k poses)
          // Fully unrolled with compile-time operand offsets: per MMA the uniform datapath executes two adds and the
          // UTCHMMA.  (r2h: with the offsets in a __constant__ table the loop cost ~14 uniform instructions per MMA, a
          // dependent chain through a constant load, and the ISSUE of the MMAs -- not the tensor pipe, not memory -- bounded
          // the kernel: two extra instructions per MMA cost 25 %, issuing every MMA twice only 24 %.)
          if (nr == 1) {
            const uint32_t tm0 = r_tm[0], id0 = r_idesc[0], bl0 = b_lo_base + r_boff[0];
#pragma unroll
            for (int m = 1; m < 18; m++) ptx::mma_f16_ss_lohi<1>(tm0, a_lo_base + off_a(m), kDescHiA, bl0 + off_b(m), kDescHiB, id0);
          } else {
            const uint32_t tm0 = r_tm[0], id0 = r_idesc[0], bl0 = b_lo_base + r_boff[0];
            const uint32_t tm1 = r_tm[1], id1 = r_idesc[1], bl1 = b_lo_base + r_boff[1];
#pragma unroll
            for (int m = 1; m < 18; m++) {
              ptx::mma_f16_ss_lohi<1>(tm0, a_lo_base + off_a(m), kDescHiA, bl0 + off_b(m), kDescHiB, id0);
              ptx::mma_f16_ss_lohi<1>(tm1, a_lo_base + off_a(m), kDescHiA, bl1 + off_b(m), kDescHiB, id1);
            }
          }
          ptx::tc_commit(&empty[st]);                                     // slab consumed
          if (xi >= 2) ptx::tc_commit(&accf[(go_base + xi - 2) % kR]);      // output plane xi - 1 is complete
          if (xi == kD) ptx::tc_commit(&accf[(go_base + kD - 1) % kR]);
        }
        __syncwarp();
        tr(p.trace, 1, gp, 4);
        committed += (xi >= 2) + (xi == kD);
      }
    }
    if (kTcPw) issue_pw_until(committed);   // the last planes
  } else if constexpr (kTcPw) {
    // ===== epilogue, variant B: stage plane j for the pointwise tcgen05.mma, then finish plane j - 1 =====
    const int q4 = warp & 3;
    const int row = q4 * 32 + lane;
    const int yrow = row >> 3, zz = row & 7;
    const uint32_t tm_lane = (uint32_t)(q4 * 32) << 16;
    constexpr int Dn = 12, Pn = 14;
    int n_my = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) n_my++;
    const uint32_t n_planes = (uint32_t)n_my * kD;
    float keep[32];
    uint4* xo4 = reinterpret_cast<uint4*>(p.xout);
    for (uint32_t j = 0; j <= n_planes; j++) {
      if (j < n_planes) {
        // ---- step 1 (plane j): conv accumulator -> bias, ReLU, fp16 -> A operand buffer j & 1 ----
        const uint32_t slot = j % kR, u = j / kR, b = j & 1;
        if (warp == 2) tr(p.trace, 2, j, 0);
        ptx::mbar_wait(&accf[slot], u & 1);
        if (warp == 2) tr(p.trace, 2, j, 1);
        ptx::tc_fence_after();
        uint32_t v[32];
        ptx::tmem_ld32(tmem_base + tm_lane + slot * 32u, v);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();   // ordered before the a2_full arrival below: the pointwise MMA may then overwrite the slot
#pragma unroll
        for (int c8 = 0; c8 < 4; c8++) {
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; e++) {
            const int c = c8 * 8 + 2 * e;
            const float f0 = fmaxf(__uint_as_float(v[c]) + s_bias[c], 0.f);
            const float f1 = fmaxf(__uint_as_float(v[c + 1]) + s_bias[c + 1], 0.f);
            const __half2 h = __floats2half2_rn(f0, f1);
            w[e] = *reinterpret_cast<const uint32_t*>(&h);
          }
          *reinterpret_cast<uint4*>(s_y + b * 8192 + c8 * 2048 + row * 16) = make_uint4(w[0], w[1], w[2], w[3]);
        }
        ptx::fence_proxy_async();   // generic-proxy stores -> visible to the tensor core
        ptx::mbar_arrive(&a2_full[b]);
        if (warp == 2) tr(p.trace, 2, j, 2);
      }
      if (j == 0) continue;
      // ---- step 2 (plane j - 1): pointwise accumulator -> bias, ReLU, 2x2x2 average ----
      const uint32_t jj = j - 1, b = jj & 1;
      ptx::mbar_wait(&d2_full[b], (jj >> 1) & 1);
      if (warp == 2) tr(p.trace, 2, j, 3);
      ptx::tc_fence_after();
      uint32_t v[32];
      ptx::tmem_ld32(tmem_base + tm_lane + (jj % kR) * 32u, v);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&acce[jj % kR]);   // the slot is free for the convolution of plane jj + kR
      const int xo = (int)(jj % kD) + 1;
      if (xo & 1) {
#pragma unroll
        for (int c = 0; c < 32; c++) keep[c] = fmaxf(__uint_as_float(v[c]) + s_bias[32 + c], 0.f);
      } else {
        uint32_t o[16];
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          float s0 = keep[c] + fmaxf(__uint_as_float(v[c]) + s_bias[32 + c], 0.f);
          float s1 = keep[c + 1] + fmaxf(__uint_as_float(v[c + 1]) + s_bias[32 + c + 1], 0.f);
          s0 += __shfl_xor_sync(0xffffffffu, s0, 1);
          s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
          s0 += __shfl_xor_sync(0xffffffffu, s0, 8);
          s1 += __shfl_xor_sync(0xffffffffu, s1, 8);
          const __half2 h = __floats2half2_rn(s0 * 0.125f, s1 * 0.125f);
          o[c >> 1] = *reinterpret_cast<const uint32_t*>(&h);
        }
        const int item = blockIdx.x + (int)(jj / kD) * gridDim.x;
        const int zb = item % kZBlocks, k = (item / kZBlocks) % kRowTiles, grp = item / (kZBlocks * kRowTiles);
        const int R = 16 * k + 1 + yrow;
        const int q = R / kP, yp = R - q * kP;
        const int pose = grp * kG + q;
        // the lane that stores a pooled voxel: even tile row (= odd yp) and even z; its window is rows R, R+1, columns zz, zz+1
        if (((lane & 9) == 0) && q < kG && pose < p.n_poses && yp >= 1 && yp <= kD) {
          const int yo = (yp - 1) >> 1, zo = 4 * zb + (zz >> 1);
          uint4* dst = xo4 + (((size_t)(pose / p.out_G) * Dn + ((xo >> 1) - 1)) * 4) * p.out_lp + (size_t)(pose % p.out_G) * Pn * Pn +
                       (size_t)(yo + 1) * Pn + (zo + 1);
#pragma unroll
          for (int c8 = 0; c8 < 4; c8++) dst[(size_t)c8 * p.out_lp] = make_uint4(o[4 * c8], o[4 * c8 + 1], o[4 * c8 + 2], o[4 * c8 + 3]);
        }
      }
    }
  } else {
    // ===== epilogue =====
    const int q4 = warp & 3;
    const int row = q4 * 32 + lane;
    const uint32_t tm_lane = (uint32_t)(q4 * 32) << 16;
    constexpr int Dn = 12, Pn = 14;
    // --- pointwise weights as mma.sync A fragments (rows = output channels), reduction index permuted: k-slots
    // (2t, 2t+1, 2t+8, 2t+9) of k step s carry input channels 8t + 4s + (0..3); see pointwise_pool_mma_kernel ---
    const int g = lane >> 2, t = lane & 3;
    uint32_t af[2][2][4];
#pragma unroll
    for (int mt = 0; mt < 2; mt++)
#pragma unroll
      for (int ks = 0; ks < 2; ks++) {
        const __half* w0 = p.w2 + (size_t)(mt * 16 + g) * 32 + 8 * t + 4 * ks;
        const __half* w1 = w0 + (size_t)8 * 32;
        af[mt][ks][0] = *reinterpret_cast<const uint32_t*>(w0);
        af[mt][ks][1] = *reinterpret_cast<const uint32_t*>(w1);
        af[mt][ks][2] = *reinterpret_cast<const uint32_t*>(w0 + 2);
        af[mt][ks][3] = *reinterpret_cast<const uint32_t*>(w1 + 2);
      }
    float bs[2][2];
#pragma unroll
    for (int mt = 0; mt < 2; mt++) { bs[mt][0] = p.bias2[mt * 16 + g]; bs[mt][1] = p.bias2[mt * 16 + g + 8]; }
    const uint32_t eighth2 = 0x30003000u;  // half2(0.125, 0.125): pooling matrix, column n = g
    const uint32_t sb0 = g == 0 ? eighth2 : 0u, sb1 = g == 1 ? eighth2 : 0u;
    // fine voxel of this lane inside a 2x2x2 window: (di, dj, dk) = bits of g (plane, tile row, z)
    const int di = (g >> 2) & 1, dj = (g >> 1) & 1, dk = g & 1;
    uint32_t go = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int zb = item % kZBlocks, k = (item / kZBlocks) % kRowTiles, grp = item / (kZBlocks * kRowTiles);
      for (int xo = 1; xo <= kD; xo++, go++) {
        const uint32_t slot = go % kR, u = go / kR;
        // ---- conv accumulator -> bias, ReLU, fp16 -> this thread's row of the shared tile of plane parity (xo - 1) & 1 ----
        ptx::mbar_wait(&accf[slot], u & 1);
        ptx::tc_fence_after();
        uint32_t v[32];
        ptx::tmem_ld32(tmem_base + tm_lane + slot * 32u, v);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        ptx::mbar_arrive(&acce[slot]);
        uint8_t* yrow_ptr = s_y + ((xo - 1) & 1) * 8192 + row * 64;
        const int sw = (row >> 1) & 3;
#pragma unroll
        for (int c8 = 0; c8 < 4; c8++) {
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; e++) {
            const int c = c8 * 8 + 2 * e;
            const float f0 = fmaxf(__uint_as_float(v[c]) + s_bias[c], 0.f);
            const float f1 = fmaxf(__uint_as_float(v[c + 1]) + s_bias[c + 1], 0.f);
            const __half2 h = __floats2half2_rn(f0, f1);
            w[e] = *reinterpret_cast<const uint32_t*>(&h);
          }
          *reinterpret_cast<uint4*>(yrow_ptr + ((c8 ^ sw) * 16)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
        if (xo & 1) continue;
        __syncwarp();
        // ---- both planes of the pair are staged: unit2_conv + ReLU + 2x2x2 average on this warp's 8 pooling windows ----
        const int xq = (xo >> 1) - 1;
#pragma unroll
        for (int jy = 0; jy < 2; jy++) {
          // tile rows of the window pair (jy): R, R + 1 with R = 16 k + 1 + 4 q4 + 2 jy (odd)
          const int R = 16 * k + 1 + 4 * q4 + 2 * jy;
          const int q = R / kP, yp = R - q * kP;
          const int pose = grp * kG + q;
          const bool ok = q < kG && pose < p.n_poses && yp >= 1 && yp <= kD;   // uniform across the warp
          const int yo = (yp - 1) >> 1;
#pragma unroll
          for (int jz = 0; jz < 2; jz++) {
            // pooled voxels A = (jy, 2 jz), B = (jy, 2 jz + 1); this lane's fine voxel of A: tile row 2 jy + dj, z 4 jz + dk
            const int ra = q4 * 32 + (2 * jy + dj) * 8 + 4 * jz + dk;
            const uint8_t* pa_ = s_y + di * 8192 + ra * 64;
            const uint4 va = *reinterpret_cast<const uint4*>(pa_ + ((t ^ ((ra >> 1) & 3)) * 16));
            const int rb = ra + 2;
            const uint4 vb = *reinterpret_cast<const uint4*>(pa_ + 128 + ((t ^ ((rb >> 1) & 3)) * 16));
            const uint32_t* pa = reinterpret_cast<const uint32_t*>(&va);
            const uint32_t* pb = reinterpret_cast<const uint32_t*>(&vb);
#pragma unroll
            for (int mt = 0; mt < 2; mt++) {
              float ca[4] = {0.f, 0.f, 0.f, 0.f}, cb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
              for (int ks = 0; ks < 2; ks++) {
                mma_16816(ca, af[mt][ks], pa[2 * ks], pa[2 * ks + 1]);
                mma_16816(cb, af[mt][ks], pb[2 * ks], pb[2 * ks + 1]);
              }
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
              if (t == 0 && ok) {
                const int zo = 4 * zb + 2 * jz;
                const size_t pos = (size_t)(pose % p.out_G) * Pn * Pn + (size_t)(yo + 1) * Pn + (zo + 1);
                __half* o0 = p.xout + ((((size_t)(pose / p.out_G) * Dn + xq) * 4 + 2 * mt) * p.out_lp + pos) * 8 + g;
                o0[0] = __float2half(pz[0]);
                o0[8] = __float2half(pz[1]);
                o0[(size_t)p.out_lp * 8] = __float2half(pz[2]);
                o0[(size_t)p.out_lp * 8 + 8] = __float2half(pz[3]);
              }
            }
          }
        }
        __syncwarp();   // the next odd plane overwrites this warp's rows of parity 0
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

// ======================================================================================================================
// "v2" of the fused kernel (GB_TC_FUSED_V2=1): ONE CTA per SM with all 512 TMEM columns.
//   * ring of 8 plane slots + 2 GHOST slots (columns 256..319): the window of input plane xi always writes three CONSECUTIVE
//     slots starting at slot((xi - 2) & 7); when that runs past slot 7 the planes that belong in slots 0 / 1 receive their
//     first contributions in the ghosts, their remaining ones in their own slot, and the epilogue adds the two.  No N = 96 MMA
//     is ever split at a ring wrap (a split pair reads the 4 KB A tile twice: 88 instead of 56 cycles of operand fetch)
//   * 16 epilogue warps in two TEAMS that own alternate plane pairs (a 2x2x2 pooling window needs both planes of a pair);
//     inside a team two warps per TMEM lane quarter, 16 of the 32 channels each.  One plane is a ~1250-cycle chain of
//     dependent latencies for a warp (barrier, tcgen05.ld, shared stores, proxy fence, barrier, tcgen05.ld, shuffles): r2y, with
//     one team the epilogue bounded the kernel however the columns were split; two planes in flight remove that
//   * 6 TMA stages, the issue loop unrolled over the ring period (slot numbers and accumulate patterns are compile-time)
// Same arithmetic as variant B above (pointwise tcgen05.mma into the plane's own slot on the fixed schedule).
constexpr int k2Stages = 6;
constexpr int k2Ncg = 2;                  // column groups: epilogue warps per TMEM lane quarter and team
constexpr int k2Cw = 32 / k2Ncg;          // channels per epilogue warp
constexpr int k2Teams = 2;                // epilogue teams; team t owns the plane PAIRS t, t + 2, t + 4, ...
constexpr int k2TeamWarps = 4 * k2Ncg;
constexpr int k2EpiWarps = k2Teams * k2TeamWarps;
constexpr int k2Bufs = 2 * k2Teams;       // A-operand buffers of the pointwise MMA (one per plane in flight)
constexpr int k2Threads = 32 * (3 + k2EpiWarps);   // producer, conv issuer, 8 epilogue warps, pointwise issuer
// pair mode: per-rank B rows.  [0,96): 48 rows per (tap, chunk); [0,64) and [32,96): 32 rows; [64,96) of tap 0, chunks 0-1: 16 rows
constexpr int kPairOffA64 = 9 * 4 * 48 * 16;                       // 27648
constexpr int kPairOffB64 = kPairOffA64 + 9 * 4 * 32 * 16;         // 46080
constexpr int kPairOffC32 = kPairOffB64 + 9 * 4 * 32 * 16;         // 64512
constexpr int kPairWBytes = kPairOffC32 + 2 * 16 * 16 + 512;       // 65536 (padded: eight equal bulk copies)
constexpr int kPairW2Bytes = 4 * 16 * 16;                          // 1024
constexpr int k2OffStage = 65536;
constexpr int k2OffY = k2OffStage + k2Stages * kStageBytes;
constexpr int k2OffW2 = k2OffY + k2Bufs * 8192;
constexpr int k2OffBar = k2OffW2 + kW2Bytes;
constexpr int k2SmemTotal = k2OffBar + 1024;

// kPair: a cluster of two CTAs (the two SMs of a TPC) works on two items in lockstep with tcgen05 ...cta_group::2: every MMA is
// M = 256 (128 rows of each CTA's own A slab and TMEM) and reads only HALF of the B rows from each CTA's shared memory -- 5.5 KB
// of operands per N = 96 MMA instead of 7, under the 48 cycles of math.  Only the leader CTA (rank 0) issues MMAs; its barriers
// collect the arrivals of both CTAs' epilogue warps, tcgen05.commit multicasts completion to both.  The B rows are packed per
// rank by pack_pair_weights(): the sub-ranges of N used at item edges and for the first-touch MMA need their own half-split
// copies (a sub-range of a half-split array is not a half-split of the sub-range).
template <bool kPair>
__global__ void __launch_bounds__(k2Threads, 1) conv1_pw2_pool_v2_kernel(const __grid_constant__ CUtensorMap tmap, const FusedParams p) {
  constexpr int kR = 8;
  constexpr uint32_t kM = kPair ? 256u : 128u;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_w = smem;
  uint8_t* s_stage = smem + k2OffStage;
  uint8_t* s_y = smem + k2OffY;
  uint8_t* s_w2 = smem + k2OffW2;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + k2OffBar);
  uint64_t* full = bars;                   // [k2Stages]
  uint64_t* empty = full + k2Stages;       // [k2Stages]
  uint64_t* accf = empty + k2Stages;       // [kR] conv plane complete
  uint64_t* acce = accf + kR;              // [kR] slot handed back (pointwise result read): one arrival per epilogue warp
  uint64_t* wbar = acce + kR;
  uint64_t* a2_full = wbar + 1;            // [k2Bufs]
  uint64_t* d2_full = a2_full + k2Bufs;    // [k2Bufs]
  uint64_t* pfull = d2_full + k2Bufs;      // [k2Stages] pair mode, leader: the peer's TMA box has landed (relayed)
  uint64_t* pwbar = pfull + k2Stages;      // pair mode, leader: the peer's weights have landed
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(pwbar + 1);
  float* s_bias = reinterpret_cast<float*>(s_tmem + 2);   // bias1[32], bias2[32]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = p.n_groups * kRowTiles * kZBlocks;
  // work units: items, or pairs of items (unit u = items 2u, 2u + 1; an odd tail lets rank 1 redo item 2u without storing)
  const uint32_t rank = kPair ? ptx::cluster_ctarank() : 0u;
  const int n_units = kPair ? (n_items + 1) / 2 : n_items;
  const int u0 = kPair ? (int)ptx::cluster_id_x() : (int)blockIdx.x, ustride = kPair ? (int)ptx::cluster_nctaid_x() : (int)gridDim.x;
  int n_my = 0;
  for (int u = u0; u < n_units; u += ustride) n_my++;
  auto item_of = [&](int i, bool& valid) {   // i-th unit of this CTA -> its item
    const int u = u0 + i * ustride;
    if (!kPair) { valid = true; return u; }
    const int it = 2 * u + (int)rank;
    valid = it < n_items;
    return valid ? it : 2 * u;
  };
  constexpr uint32_t kArr = kPair ? 2u * k2TeamWarps : (uint32_t)k2TeamWarps;   // arrivals on the leader's acce / a2_full barriers

  if (threadIdx.x == 0) {
    for (int s = 0; s < k2Stages; s++) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
    for (int s = 0; s < kR; s++) { ptx::mbar_init(&accf[s], 1); ptx::mbar_init(&acce[s], kArr); }   // one arrival per warp of the owning team(s)
    ptx::mbar_init(wbar, 1);
    ptx::mbar_init(pwbar, 1);
    for (int s = 0; s < k2Stages; s++) ptx::mbar_init(&pfull[s], 1);
    for (int b = 0; b < k2Bufs; b++) { ptx::mbar_init(&a2_full[b], kArr); ptx::mbar_init(&d2_full[b], 1); }
    ptx::fence_mbar_init();
  }
  if (threadIdx.x < 32) s_bias[threadIdx.x] = p.bias1[threadIdx.x];
  else if (threadIdx.x < 64) s_bias[threadIdx.x] = p.bias2[threadIdx.x - 32];
  if (warp == 1) {
    if constexpr (kPair) { ptx::tmem_alloc2(s_tmem, 512); ptx::tmem_relinquish2(); }
    else { ptx::tmem_alloc(s_tmem, 512); ptx::tmem_relinquish(); }
  }
  ptx::tc_fence_before();
  if constexpr (kPair) ptx::cluster_sync_all();   // both CTAs' barriers are initialised before anyone signals across
  else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  // leader-side barrier wait / peer-side arrival that work in both modes
  auto wait_x = [&](uint64_t* bar, uint32_t parity) {
    if constexpr (kPair) ptx::mbar_wait_cluster(bar, parity);
    else ptx::mbar_wait(bar, parity);
  };
  auto arrive_x = [&](uint64_t* bar) {   // on the LEADER's barrier
    if (kPair && rank != 0) ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(bar), 0));
    else ptx::mbar_arrive(bar);
  };
  auto commit_x = [&](uint64_t* bar) {
    if constexpr (kPair) ptx::tc_commit2(bar);
    else ptx::tc_commit(bar);
  };

  if (warp == 0) {
    // ===== producer =====
    if (ptx::elect_one()) {
      ptx::prefetch_tmap(&tmap);
      if constexpr (kPair) {
        ptx::mbar_expect_tx(wbar, kPairWBytes + kPairW2Bytes);
        const uint8_t* src = reinterpret_cast<const uint8_t*>(p.wp) + (size_t)rank * kPairWBytes;
        for (int t8 = 0; t8 < 8; t8++) ptx::bulk_g2s(s_w + t8 * (kPairWBytes / 8), src + t8 * (kPairWBytes / 8), kPairWBytes / 8, wbar);
        ptx::bulk_g2s(s_w2, reinterpret_cast<const uint8_t*>(p.w2p) + (size_t)rank * kPairW2Bytes, kPairW2Bytes, wbar);
      } else {
        ptx::mbar_expect_tx(wbar, kWBytes + kW2Bytes);
        for (int t9 = 0; t9 < 9; t9++)
          ptx::bulk_g2s(s_w + t9 * (kWBytes / 9), reinterpret_cast<const uint8_t*>(p.wp) + t9 * (kWBytes / 9), kWBytes / 9, wbar);
        ptx::bulk_g2s(s_w2, p.w2p, kW2Bytes, wbar);
      }
    }
    __syncwarp();
    uint32_t gp = 0;
    for (int im = 0; im < n_my; im++) {
      bool valid;
      const int item = item_of(im, valid);
      const int zb = item % kZBlocks, k = (item / kZBlocks) % kRowTiles, g = item / (kZBlocks * kRowTiles);
      for (int it = 0; it < kD; it++, gp++) {
        const uint32_t st = gp % k2Stages, ph = (gp / k2Stages) & 1;
        ptx::mbar_wait(&empty[st], ph ^ 1);
        if (ptx::elect_one()) {
          ptx::mbar_expect_tx(&full[st], kStageBytes);
          ptx::tma_load_4d(s_stage + (size_t)st * kStageBytes, &tmap, 64 * zb, 16 * k, 0, g * kD + it, &full[st]);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===== convolution MMA issuer: ONE thread does everything (waits included).  With a single CTA per SM nobody else
    // fills the tensor-core queue while this thread is away, and the queue is only a few MMAs deep (r2y: every cycle this
    // thread spent waiting was an idle cycle of the unit), so the barrier waits of the NEXT window are taken in the middle
    // of the current window's MMAs, where a few queued MMAs cover their latency. =====
    if (kPair && rank != 0) {
      // peer CTA: this warp only relays "my weights / my TMA box have landed" to the leader, which issues the MMAs that read them
      if (ptx::elect_one()) {
        ptx::mbar_wait(wbar, 0);
        ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(pwbar), 0));
        const uint32_t n_planes = (uint32_t)n_my * kD;
        for (uint32_t gp = 0; gp < n_planes; gp++) {
          const uint32_t st = gp % k2Stages;
          ptx::mbar_wait(&full[st], (gp / k2Stages) & 1);
          ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(&pfull[st]), 0));
        }
      }
    } else if (ptx::elect_one()) {
      constexpr uint32_t kDescHiA = (uint32_t)(kSlabZ * 16 >> 4) | (1u << 14);
      constexpr uint32_t kDescHiB = (128u >> 4) | (1u << 14);
      const uint32_t a_lo_fixed = ((uint32_t)(kChunkBytes >> 4)) << 16;
      // B descriptors (lo words: LBO << 16 | address >> 4) and row strides of the N ranges [0,96) [0,64) [32,96) [64,96)
      const uint32_t wbase = ptx::smem_u32(s_w) >> 4;
      const uint32_t b_main = kPair ? ((48u << 16) | wbase) : ((96u << 16) | wbase);
      const uint32_t b_a64 = kPair ? ((32u << 16) | (wbase + (kPairOffA64 >> 4))) : b_main;
      const uint32_t b_b64 = kPair ? ((32u << 16) | (wbase + (kPairOffB64 >> 4))) : b_main + 32u;
      const uint32_t b_c32 = kPair ? ((16u << 16) | (wbase + (kPairOffC32 >> 4))) : b_main + 64u;
      constexpr uint32_t kS96 = kPair ? 48u : 96u, kS64 = kPair ? 32u : 96u;
      ptx::mbar_wait(wbar, 0);
      if constexpr (kPair) ptx::mbar_wait_cluster(pwbar, 0);
      uint32_t gp = 0, pbase = 0;   // pbase: global index of this item's plane 1 (a multiple of 24, hence of the ring period)
      // barrier waits of window (it8, r): the TMA box, and the slot of every output plane that receives its FIRST
      // contribution there (handed back by the plane 8 earlier; a first contribution that lands in a ghost slot is covered
      // by the same wait)
      auto wait_window = [&](const int r, const int it8, const uint32_t gpw, const uint32_t pb) {
        const int xi = 8 * it8 + r + 1;
        if (xi != kD) {
          const uint32_t gpl = pb + (uint32_t)xi, u = gpl >> 3;      // plane xi + 1
          if (u > 0) wait_x(&acce[(r + 1) & 7], (u - 1) & 1);
        }
        if (xi == 1) {
          const uint32_t u = pb >> 3;                                  // plane 1
          if (u > 0) wait_x(&acce[0], (u - 1) & 1);
        }
        ptx::mbar_wait(&full[gpw % k2Stages], (gpw / k2Stages) & 1);
        if constexpr (kPair) ptx::mbar_wait_cluster(&pfull[gpw % k2Stages], (gpw / k2Stages) & 1);
      };
      // one window = one input plane; r = (xi - 1) & 7 is a compile-time constant of the unrolled body
      auto window = [&](auto rc, const int it8, const bool more) {
        constexpr int r = decltype(rc)::value;
        const bool first_w = (r == 0) && it8 == 0;   // xi == 1: output planes 1, 2 only
        const bool last_w = (r == 7) && it8 == 2;    // xi == 24: output planes 23, 24 only
        tr(p.trace, 1, gp, 0);
        const uint32_t st = gp % k2Stages;
        ptx::tc_fence_after();
        const uint32_t a_lo_base = a_lo_fixed | (ptx::smem_u32(s_stage + (size_t)st * kStageBytes) >> 4);
        constexpr uint32_t s0 = (uint32_t)((r + 7) & 7);   // slot of output plane xi - 1; s0 = 6, 7 run on into the ghosts
        uint32_t tm, bl, bs, idn;   // D columns, B descriptor, its row stride, instruction descriptor of the window's MMAs
        if (first_w) { tm = tmem_base; bl = b_b64; bs = kS64; idn = ptx::idesc_f16(kM, 64); }
        else if (last_w) { tm = tmem_base + 6 * 32u; bl = b_a64; bs = kS64; idn = ptx::idesc_f16(kM, 64); }
        else { tm = tmem_base + s0 * 32u; bl = b_main; bs = kS96; idn = ptx::idesc_f16(kM, 96); }
        auto mma = [&](auto acc, uint32_t d, uint32_t alo, uint32_t blo, uint32_t idesc) {
          if constexpr (kPair) ptx::mma2_f16_ss_lohi<decltype(acc)::value>(d, alo, kDescHiA, blo, kDescHiB, idesc);
          else ptx::mma_f16_ss_lohi<decltype(acc)::value>(d, alo, kDescHiA, blo, kDescHiB, idesc);
        };
        constexpr std::integral_constant<int, 0> kFresh{};
        constexpr std::integral_constant<int, 1> kAcc{};
        // first MMA of the window (tap 0, k step 0): who is fresh
        if (first_w || (r == 1 && it8 > 0)) {
          mma(kFresh, tm, a_lo_base, bl, idn);                                   // every plane of the window is fresh
        } else if (last_w) {
          mma(kAcc, tm, a_lo_base, bl, idn);
        } else {
          mma(kAcc, tm, a_lo_base, b_a64, ptx::idesc_f16(kM, 64));             // planes xi - 1, xi
          mma(kFresh, tm + 64u, a_lo_base, b_c32, ptx::idesc_f16(kM, 32));     // plane xi + 1: fresh
        }
#pragma unroll
        for (int m = 1; m < 12; m++) mma(kAcc, tm, a_lo_base + off_a(m), bl + (uint32_t)(((m >> 1) * 4 + 2 * (m & 1))) * bs, idn);
        tr(p.trace, 1, gp, 1);
        // the next window's waits, behind a dozen queued MMAs (the item loop passes the next item's plane base when r = 7 of
        // the last period wraps around)
        if (more) {
          constexpr int rn = (r + 1) & 7;
          const int it8n = (r == 7) ? (it8 == 2 ? 0 : it8 + 1) : it8;
          wait_window(rn, it8n, gp + 1, (r == 7 && it8 == 2) ? pbase + kD : pbase);
          ptx::tc_fence_after();
        }
        tr(p.trace, 1, gp, 2);
#pragma unroll
        for (int m = 12; m < 18; m++) mma(kAcc, tm, a_lo_base + off_a(m), bl + (uint32_t)(((m >> 1) * 4 + 2 * (m & 1))) * bs, idn);
        commit_x(&empty[st]);
        if (!first_w) commit_x(&accf[s0]);        // output plane xi - 1 is complete
        if (last_w) commit_x(&accf[7]);           // ... and so is plane 24
        tr(p.trace, 1, gp, 4);
        gp++;
      };
      if (n_my > 0) wait_window(0, 0, 0, 0);
      for (int im = 0; im < n_my; im++, pbase += kD) {
        const bool last_item = im + 1 == n_my;
#pragma unroll 1
        for (int it8 = 0; it8 < 3; it8++) {
          window(std::integral_constant<int, 0>{}, it8, true); window(std::integral_constant<int, 1>{}, it8, true);
          window(std::integral_constant<int, 2>{}, it8, true); window(std::integral_constant<int, 3>{}, it8, true);
          window(std::integral_constant<int, 4>{}, it8, true); window(std::integral_constant<int, 5>{}, it8, true);
          window(std::integral_constant<int, 6>{}, it8, true); window(std::integral_constant<int, 7>{}, it8, !(last_item && it8 == 2));
        }
      }
    }
  } else if (warp == 2 + k2EpiWarps) {
    // ===== pointwise MMA issuer (one thread): D2(j) = A2[j & 1] (128 x 32, staged by the epilogue) x W2^T into plane j's own
    // slot.  The two issuers never touch the same TMEM columns at the same time, so their relative order in the queue does
    // not matter. =====
    if ((!kPair || rank == 0) && ptx::elect_one()) {
      constexpr uint32_t kDescHiB = (128u >> 4) | (1u << 14);
      constexpr uint32_t kW2Lbo = kPair ? 16u : 32u;   // rows of W2 in this CTA = chunk pitch in 16-byte units
      ptx::mbar_wait(wbar, 0);
      if constexpr (kPair) ptx::mbar_wait_cluster(pwbar, 0);
      const uint32_t w2_lo = (kW2Lbo << 16) | (ptx::smem_u32(s_w2) >> 4);
      const uint32_t n_planes = (uint32_t)n_my * kD;
      for (uint32_t j = 0; j < n_planes; j++) {
        const uint32_t b = j % k2Bufs;
        wait_x(&a2_full[b], (j / k2Bufs) & 1);
        ptx::tc_fence_after();
        const uint32_t a2_lo = ((uint32_t)(2048 >> 4) << 16) | (ptx::smem_u32(s_y + b * 8192) >> 4);
        const uint32_t tm_d2 = tmem_base + (j & 7) * 32u;
        if constexpr (kPair) {
          ptx::mma2_f16_ss_lohi<0>(tm_d2, a2_lo, kDescHiB, w2_lo, kDescHiB, ptx::idesc_f16(kM, 32));
          ptx::mma2_f16_ss_lohi<1>(tm_d2, a2_lo + 2 * (2048 >> 4), kDescHiB, w2_lo + 2 * kW2Lbo, kDescHiB, ptx::idesc_f16(kM, 32));
        } else {
          ptx::mma_f16_ss_lohi<0>(tm_d2, a2_lo, kDescHiB, w2_lo, kDescHiB, ptx::idesc_f16(kM, 32));
          ptx::mma_f16_ss_lohi<1>(tm_d2, a2_lo + 2 * (2048 >> 4), kDescHiB, w2_lo + 2 * kW2Lbo, kDescHiB, ptx::idesc_f16(kM, 32));
        }
        commit_x(&d2_full[b]);
      }
    }
  } else {
    // ===== epilogue: k2Teams teams x (4 TMEM lane quarters x k2Ncg groups of k2Cw channels) =====
    const int ew = warp - 2, team = ew / k2TeamWarps;
    const int q4 = warp & 3, cg = (ew % k2TeamWarps) >> 2;
    const int row = q4 * 32 + lane;
    const int yrow = row >> 3, zz = row & 7;
    const uint32_t tm_mine = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(cg * k2Cw);
    constexpr int Dn = 12, Pn = 14;
    const uint32_t n_planes = (uint32_t)n_my * kD;
    float keep[k2Cw];
    uint4* xo4 = reinterpret_cast<uint4*>(p.xout);
    const float* b1 = s_bias + cg * k2Cw;
    const float* b2 = s_bias + 32 + cg * k2Cw;
    const bool tracer = warp == 2;
    // this team's planes in order: pairs team, team + k2Teams, ... ; the loop stages plane cur, then finishes plane prev
    uint32_t prev = 0xffffffffu;
    uint4* dst_base = nullptr;   // per item: this lane's pooled output voxel (nullptr: the lane stores nothing)
    int dst_item = -1;
    for (uint32_t t = 0;; t++) {
      const uint32_t cur = 2u * (uint32_t)(team + (int)(t >> 1) * k2Teams) + (t & 1);
      const bool have_cur = cur < n_planes;
      if (have_cur) {
        // ---- step 1 (plane cur): conv accumulator (+ ghost part) -> bias, ReLU, fp16 -> A operand buffer ----
        const uint32_t j = cur, slot = j & 7, b = j % k2Bufs, pj = j % kD;
        if (tracer) tr(p.trace, 2, j, 0);
        ptx::mbar_wait(&accf[slot], (j >> 3) & 1);
        if (tracer) tr(p.trace, 2, j, 1);
        ptx::tc_fence_after();
        uint32_t v[k2Cw];
        ptx::tmem_ld_cols(tm_mine + slot * 32u, v);
        if (pj >= 8 && (pj & 7) < 2) {   // planes 9, 10, 17, 18 of the item: first contributions sit in ghost (pj & 1)
          uint32_t g[k2Cw];
          ptx::tmem_ld_cols(tm_mine + 256u + (pj & 1) * 32u, g);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < k2Cw; c++) v[c] = __float_as_uint(__uint_as_float(g[c]) + __uint_as_float(v[c]));
        } else {
          ptx::tmem_ld_wait();
        }
        ptx::tc_fence_before();
#pragma unroll
        for (int c8 = 0; c8 < k2Cw / 8; c8++) {
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; e++) {
            const int c = c8 * 8 + 2 * e;
            const float f0 = fmaxf(__uint_as_float(v[c]) + b1[c], 0.f);
            const float f1 = fmaxf(__uint_as_float(v[c + 1]) + b1[c + 1], 0.f);
            const __half2 h = __floats2half2_rn(f0, f1);
            w[e] = *reinterpret_cast<const uint32_t*>(&h);
          }
          *reinterpret_cast<uint4*>(s_y + b * 8192 + (cg * (k2Cw / 8) + c8) * 2048 + row * 16) = make_uint4(w[0], w[1], w[2], w[3]);
        }
        ptx::fence_proxy_async();   // every lane: its shared-memory stores -> visible to the tensor core (r3b: the fence over all
                                    // state spaces waits for this warp's global stores too: 1400 instead of 350 cycles)
        __syncwarp();
        if (lane == 0) arrive_x(&a2_full[b]);                 // one arrival for the warp
        if (tracer) tr(p.trace, 2, j, 2);
      }
      if (prev != 0xffffffffu) {
        // ---- step 2 (plane prev): pointwise accumulator -> bias, ReLU, 2x2x2 average ----
        const uint32_t jj = prev, b = jj % k2Bufs, slot = jj & 7;
        ptx::mbar_wait(&d2_full[b], (jj / k2Bufs) & 1);
        if (tracer) tr(p.trace, 2, jj, 3);
        ptx::tc_fence_after();
        uint32_t v[k2Cw];
        ptx::tmem_ld_cols(tm_mine + slot * 32u, v);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_x(&acce[slot]);
        const uint32_t pj = jj % kD;
        const int xo = (int)pj + 1;
        if (xo & 1) {
          if ((int)(jj / kD) != dst_item) {   // first plane of an item this team sees: where this lane's pooled voxels of the item go
            dst_item = (int)(jj / kD);
            bool valid;
            const int item = item_of(dst_item, valid);
            const int zb = item % kZBlocks, k = (item / kZBlocks) % kRowTiles, grp = item / (kZBlocks * kRowTiles);
            const int R = 16 * k + 1 + yrow;
            const int q = R / kP, yp = R - q * kP;
            const int pose = grp * kG + q;
            dst_base = nullptr;
            if (valid && ((lane & 9) == 0) && q < kG && pose < p.n_poses && yp >= 1 && yp <= kD) {
              const int yo = (yp - 1) >> 1, zo = 4 * zb + (zz >> 1);
              dst_base = xo4 + ((size_t)(pose / p.out_G) * Dn * 4 + cg * (k2Cw / 8)) * p.out_lp + (size_t)(pose % p.out_G) * Pn * Pn +
                         (size_t)(yo + 1) * Pn + (zo + 1);
            }
          }
#pragma unroll
          for (int c = 0; c < k2Cw; c++) keep[c] = fmaxf(__uint_as_float(v[c]) + b2[c], 0.f);
        } else {
          uint32_t o[k2Cw / 2];
#pragma unroll
          for (int c = 0; c < k2Cw; c += 2) {
            float s0 = keep[c] + fmaxf(__uint_as_float(v[c]) + b2[c], 0.f);
            float s1 = keep[c + 1] + fmaxf(__uint_as_float(v[c + 1]) + b2[c + 1], 0.f);
            s0 += __shfl_xor_sync(0xffffffffu, s0, 1);
            s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
            s0 += __shfl_xor_sync(0xffffffffu, s0, 8);
            s1 += __shfl_xor_sync(0xffffffffu, s1, 8);
            const __half2 h = __floats2half2_rn(s0 * 0.125f, s1 * 0.125f);
            o[c >> 1] = *reinterpret_cast<const uint32_t*>(&h);
          }
          if (dst_base) {
            uint4* dst = dst_base + (size_t)((xo >> 1) - 1) * 4 * p.out_lp;
#pragma unroll
            for (int c8 = 0; c8 < k2Cw / 8; c8++) dst[(size_t)c8 * p.out_lp] = make_uint4(o[4 * c8], o[4 * c8 + 1], o[4 * c8 + 2], o[4 * c8 + 3]);
          }
        }
      }
      if (!have_cur) break;
      prev = cur;
    }
  }
  ptx::tc_fence_before();
  if constexpr (kPair) ptx::cluster_sync_all();   // the leader's MMAs read the peer's shared memory and write its TMEM until the end
  else __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    if constexpr (kPair) ptx::tmem_dealloc2(tmem_base, 512);
    else ptx::tmem_dealloc(tmem_base, 512);
  }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no direct libcuda symbol dependency)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  });
  return fn;
}

}  // namespace

ActLayout make_fused_x0_layout() { return make_layout(kD, kG, 32); }

// unit2_conv weight [co][ci] fp32 -> [ci / 8][co][ci % 8] fp16: the K-major B operand (LBO = 32 rows x 16 B)
uint4* pack_pointwise_tc(std::vector<void*>& allocs, const float* w, int c) {
  std::vector<__half> h((size_t)c * c);
  for (int co = 0; co < c; co++)
    for (int ci = 0; ci < c; ci++) h[((size_t)(ci / 8) * c + co) * 8 + (ci % 8)] = __float2half(w[(size_t)co * c + ci]);
  __half* d = nullptr;
  GB_CUDA(cudaMalloc(&d, h.size() * sizeof(__half)));
  allocs.push_back(d);
  GB_CUDA(cudaMemcpy(d, h.data(), h.size() * sizeof(__half), cudaMemcpyHostToDevice));
  return reinterpret_cast<uint4*>(d);
}

// pair mode (conv1_pw2_pool_v2_kernel<true>): the B rows each CTA of a pair holds, see the kernel's comment.  Built with the
// weight set and freed with it (allocs).
void pack_pair_weights(std::vector<void*>& allocs, const uint4* wp, const uint4* w2p, uint4** pair_w, uint4** pair_w2) {
  std::vector<uint4> src(9 * 4 * 96), src2(4 * 32);
  GB_CUDA(cudaMemcpy(src.data(), wp, src.size() * 16, cudaMemcpyDeviceToHost));
  GB_CUDA(cudaMemcpy(src2.data(), w2p, src2.size() * 16, cudaMemcpyDeviceToHost));
  std::vector<uint4> dst(2 * kPairWBytes / 16, make_uint4(0, 0, 0, 0)), dst2(2 * kPairW2Bytes / 16);
  for (int r = 0; r < 2; r++) {
    uint4* d = dst.data() + (size_t)r * kPairWBytes / 16;
    for (int tc = 0; tc < 36; tc++) {   // (tap, chunk)
      for (int j = 0; j < 48; j++) d[tc * 48 + j] = src[tc * 96 + 48 * r + j];
      for (int j = 0; j < 32; j++) d[kPairOffA64 / 16 + tc * 32 + j] = src[tc * 96 + 32 * r + j];
      for (int j = 0; j < 32; j++) d[kPairOffB64 / 16 + tc * 32 + j] = src[tc * 96 + 32 + 32 * r + j];
    }
    for (int c8 = 0; c8 < 2; c8++)
      for (int j = 0; j < 16; j++) d[kPairOffC32 / 16 + c8 * 16 + j] = src[c8 * 96 + 64 + 16 * r + j];
    for (int c8 = 0; c8 < 4; c8++)
      for (int j = 0; j < 16; j++) dst2[(size_t)r * kPairW2Bytes / 16 + c8 * 16 + j] = src2[c8 * 32 + 16 * r + j];
  }
  void *dw = nullptr, *dw2 = nullptr;
  GB_CUDA(cudaMalloc(&dw, dst.size() * 16));
  allocs.push_back(dw);
  GB_CUDA(cudaMalloc(&dw2, dst2.size() * 16));
  allocs.push_back(dw2);
  GB_CUDA(cudaMemcpy(dw, dst.data(), dst.size() * 16, cudaMemcpyHostToDevice));
  GB_CUDA(cudaMemcpy(dw2, dst2.data(), dst2.size() * 16, cudaMemcpyHostToDevice));
  *pair_w = reinterpret_cast<uint4*>(dw);
  *pair_w2 = reinterpret_cast<uint4*>(dw2);
}

void launch_conv1_pw2_pool(const ConvTc& conv1, const __half* w2, const uint4* w2p, const float* bias2, const uint4* pair_w, const uint4* pair_w2,
                           const uint4* x0, const ActLayout& L0, uint4* x2, const ActLayout& L2, int n_poses, cudaStream_t s) {
  GB_CHECK(conv1.cin == 32 && conv1.cout == 32 && L0.D == kD && L0.G == kG && L2.D == 12, "fused conv1 shape");
  EncodeTiledFn enc = encode_tiled_fn();
  GB_CHECK(enc != nullptr, "cuTensorMapEncodeTiled is not available in this driver");
  const int n_groups = (n_poses + kG - 1) / kG;
  int dev = 0;
  GB_CUDA(cudaGetDevice(&dev));
  static std::mutex mu;
  static bool attr_set[64] = {};
  static int n_sm[64] = {};
  {
    std::lock_guard<std::mutex> lk(mu);
    GB_CHECK(dev < 64, "device index");
    if (!attr_set[dev]) {
      GB_CUDA(cudaFuncSetAttribute(conv1_pw2_pool_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
      GB_CUDA(cudaFuncSetAttribute(conv1_pw2_pool_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
      GB_CUDA(cudaFuncSetAttribute(conv1_pw2_pool_v2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, k2SmemTotal));
      GB_CUDA(cudaFuncSetAttribute(conv1_pw2_pool_v2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, k2SmemTotal));
      GB_CUDA(cudaDeviceGetAttribute(&n_sm[dev], cudaDevAttrMultiProcessorCount, dev));
      attr_set[dev] = true;
    }
  }
  // x0 as a 4-D tensor {26 z x 8 channels (contiguous: one 416-byte row), 208 rows, 4 chunks, planes x groups}; a box row is
  // 10 z x 8 channels = 160 contiguous bytes (a 5-D map with the 16-byte channel group as its own dimension made the TMA
  // unit fetch 720 16-byte pieces per plane)
  CUtensorMap tmap;
  const cuuint64_t gdim[4] = {(cuuint64_t)kP * 8, (cuuint64_t)kRows, 4, (cuuint64_t)kD * n_groups};
  const cuuint64_t gstr[3] = {(cuuint64_t)kP * 16, (cuuint64_t)L0.Lp * 16, (cuuint64_t)4 * L0.Lp * 16};
  const cuuint32_t box[4] = {kSlabZ * 8, kSlabRows, 4, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<uint4*>(x0), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GB_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed");
  FusedParams p;
  p.wp = conv1.wp; p.bias1 = conv1.bias; p.w2 = w2; p.w2p = w2p; p.bias2 = bias2; p.xout = reinterpret_cast<__half*>(x2); p.out_lp = L2.Lp; p.out_G = L2.G;
  p.n_poses = n_poses; p.n_groups = n_groups;
  p.trace = nullptr;
  static const char* trace_path = getenv("GB_TC_FUSED_TRACE");
  static unsigned long long* d_trace = nullptr;
  static int trace_left = 2;   // the second launch of the process is traced (warm caches)
  if (trace_path && trace_left > 0 && --trace_left == 0) {
    GB_CUDA(cudaMalloc(&d_trace, sizeof(unsigned long long) * 3 * kTracePlanes * 8));
    GB_CUDA(cudaMemsetAsync(d_trace, 0, sizeof(unsigned long long) * 3 * kTracePlanes * 8, s));
    p.trace = d_trace;
  }

  const int n_items = n_groups * kRowTiles * kZBlocks;
  static const int persist = getenv("GB_TC_FUSED_PERSIST") ? atoi(getenv("GB_TC_FUSED_PERSIST")) : 2;
  static const int tc_pw = getenv("GB_TC_FUSED_PW") ? atoi(getenv("GB_TC_FUSED_PW")) : 1;   // 1 (default): pointwise conv as a second tcgen05.mma;
                                                                                                  // 0: mma.sync in the epilogue (r2g: 9.9 vs 7.6 ms)
  int grid = persist > 0 ? std::min(n_items, n_sm[dev] * persist) : n_items;
  // read per launch (tests switch variants inside one process): 1 (default) v2, one CTA per SM; 0 the two-CTAs-per-SM kernel above
  // (r2x: 6.9 ms per 10 k poses against 5.5); 2 v2 on CTA pairs (cta_group::2; r3c: 8.7 ms -- an M = 256, N = 96 MMA takes ~100
  // cycles instead of 56 for M = 128: the half of B that lives in the other SM's shared memory costs more than it saves)
  const char* v2env = getenv("GB_TC_FUSED_V2");
  const int v2 = v2env ? atoi(v2env) : 1;
  if (v2 == 2) {
    // CTA pairs: per-rank B rows packed once per weight set
    GB_CHECK(pair_w && pair_w2, "pair weights");
    p.wp = pair_w;
    p.w2p = pair_w2;
    const int n_units = (n_items + 1) / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * std::min(n_units, n_sm[dev] / 2));
    cfg.blockDim = dim3(k2Threads);
    cfg.dynamicSmemBytes = k2SmemTotal;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    GB_CUDA(cudaLaunchKernelEx(&cfg, conv1_pw2_pool_v2_kernel<true>, tmap, p));
  } else if (v2) conv1_pw2_pool_v2_kernel<false><<<std::min(n_items, n_sm[dev]), k2Threads, k2SmemTotal, s>>>(tmap, p);
  else if (tc_pw) conv1_pw2_pool_kernel<true><<<grid, 192, kSmemTotal, s>>>(tmap, p);
  else conv1_pw2_pool_kernel<false><<<grid, 192, kSmemTotal, s>>>(tmap, p);
  if (p.trace) {
    std::vector<unsigned long long> h(3 * kTracePlanes * 8);
    GB_CUDA(cudaStreamSynchronize(s));
    GB_CUDA(cudaMemcpy(h.data(), d_trace, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    if (FILE* f = fopen(trace_path, "w")) {
      for (int r = 0; r < 3; r++)
        for (int pl = 0; pl < kTracePlanes; pl++) {
          fprintf(f, "%d %d", r, pl);
          for (int k = 0; k < 8; k++) fprintf(f, " %llu", h[(r * kTracePlanes + pl) * 8 + k]);
          fprintf(f, "\n");
        }
      fclose(f);
    }
  }
}

}  // namespace gb
