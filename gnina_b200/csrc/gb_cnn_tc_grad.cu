// Fast (fp16 tensor-core) backward pass of the default2018 family: d(CE loss)/d(ligand atom coordinates).
//
// Reference behaviour: TorchModel::forward(compute_gradient = true) (torch_model.cpp:100-137) back-propagates
// loss = -log softmax(pose)[1]... through the TorchScript graph to the input grid and GridMaker::backward
// (libmolgrid) turns the grid gradient into per-atom gradients; CNNTorchScorer::score accumulates them over the
// ensemble (cnn_torch_scorer.cpp:164-179).  The fp32 kernels in gb_cnn_fp32.cu are the validation path; this file
// is the production path that reuses the forward pass's machinery:
//
//   * the three 3x3x3 backward-data convolutions are ordinary 3x3x3 convolutions with flipped taps and swapped
//     channel roles, so they run on conv3_tc_kernel (tcgen05, TMEM accumulators) with repacked weights:
//       d conv5: 128 -> 64 @ 6^3  (K = 128 split in two launches of CIN = 64, partial sums added by the consumer)
//       d conv3:  64 -> 32 @ 12^3
//       d conv1:  32 -> 28 @ 24^3
//   * unpool + ReLU mask + pointwise-conv backward + ReLU mask of the producing 3x3x3 conv are ONE kernel per
//     level (mma.sync m16n8k16): it recomputes u = W y + b for the mask, multiplies the unpooled gradient, applies
//     W^T, masks with y > 0 and writes straight into the chunk-planar padded layout the next backward conv reads.
//   * the avg-pool 1/8 factors are deferred and a loss scale S = 16 is applied at the head, so fp16 gradients stay
//     in the normal range at every level (measured |g| 1e-3 .. 1 in fp32  ->  2e-2 .. 7e2 here); the final
//     per-atom kernel multiplies by 1 / (512 S).
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "gb_tc.h"

namespace gb {

constexpr float kLossScale = 16.f;

struct TcGradWeights {
  ConvTc d5[2], d3, d1;
  std::vector<void*> allocs;
  ~TcGradWeights() { for (void* p : allocs) cudaFree(p); }
};

static std::shared_ptr<TcGradWeights> get_grad_weights(const Model& m) {
  std::lock_guard<std::mutex> lk(tc_init_mutex());
  Model& mm = const_cast<Model&>(m);
  if (mm.tc_grad) return mm.tc_grad;
  auto gw = std::make_shared<TcGradWeights>();
  // backward-data kernel: W'[co' = ci][ci' = co][k] = W[co][ci][2 - k]
  auto flipped = [](const HostTensor& w, int co_lo) {
    const int cin = w.shape[1];
    const float* d = w.data;
    return [d, cin, co_lo](int co2, int ci2, int kx, int ky, int kz) -> float {
      if (co2 >= cin) return 0.f;  // padded output channels (28 -> 32)
      return d[((((size_t)(co_lo + ci2) * cin + co2) * 3 + (2 - kx)) * 3 + (2 - ky)) * 3 + (2 - kz)];
    };
  };
  const HostTensor &w1 = m.t("unit1_conv.weight"), &w3 = m.t("unit3_conv.weight"), &w5 = m.t("unit5_conv.weight");
  GB_CHECK(w1.shape[0] == 32 && w1.shape[1] <= 32 && w3.shape[0] == 64 && w3.shape[1] == 32 && w5.shape[0] == 128 &&
               w5.shape[1] == 64,
           "default2018 conv shapes");
  gw->d5[0] = make_conv_tc(gw->allocs, 64, 64, flipped(w5, 0), nullptr);
  gw->d5[1] = make_conv_tc(gw->allocs, 64, 64, flipped(w5, 64), nullptr);
  gw->d3 = make_conv_tc(gw->allocs, 32, 64, flipped(w3, 0), nullptr);
  gw->d1 = make_conv_tc(gw->allocs, 32, 32, flipped(w1, 0), nullptr);
  mm.tc_grad = gw;
  return gw;
}

// ------------------------------------------------------------------------------------------------------------
// d loss / d y5, masked by y5 > 0, scaled by S, written as the two 64-channel halves of conv5's backward input
// (chunk-planar padded, D = 6, G = 2).  fc3_backward_kernel (gb_cnn_fp32.cu) is the fp32 statement.
__global__ void __launch_bounds__(256) heads_backward_f16_kernel(const float* __restrict__ out3, const float* __restrict__ fcw,
                                                                 const __half* __restrict__ y5, uint4* __restrict__ g5a,
                                                                 uint4* __restrict__ g5b, int Lp) {
  constexpr int F = 27648, D = 6, P = 8;
  const int pose = blockIdx.x;
  const float z0 = out3[3 * pose], z1 = out3[3 * pose + 1];
  const float mx = fmaxf(z0, z1);
  const float e0 = __expf(z0 - mx), e1 = __expf(z1 - mx);
  const float sp0 = kLossScale * e0 / (e0 + e1);
  const int grp = pose >> 1, q = pose & 1;
  for (int e = threadIdx.x; e < 216 * 16; e += 256) {
    const int pos = e >> 4, c16 = e & 15;  // c16: chunk of 8 channels among 128
    const int x = pos / 36, y = (pos / 6) % 6, z = pos % 6;
    const uint4 yv = *reinterpret_cast<const uint4*>(y5 + (size_t)pose * F + pos * 128 + c16 * 8);
    const __half2* yh = reinterpret_cast<const __half2*>(&yv);
    const float4* wa = reinterpret_cast<const float4*>(fcw + pos * 128 + c16 * 8);
    const float4* wb = reinterpret_cast<const float4*>(fcw + F + pos * 128 + c16 * 8);
    const float4 a0 = wa[0], a1 = wa[1], b0 = wb[0], b1 = wb[1];
    const float d[8] = {a0.x - b0.x, a0.y - b0.y, a0.z - b0.z, a0.w - b0.w, a1.x - b1.x, a1.y - b1.y, a1.z - b1.z, a1.w - b1.w};
    uint4 o;
    uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const float2 yy = __half22float2(yh[k]);
      const __half2 h = __floats2half2_rn(yy.x > 0.f ? sp0 * d[2 * k] : 0.f, yy.y > 0.f ? sp0 * d[2 * k + 1] : 0.f);
      ow[k] = *reinterpret_cast<const uint32_t*>(&h);
    }
    uint4* dst = (c16 < 8 ? g5a : g5b) + (((size_t)grp * D + x) * 8 + (c16 & 7)) * Lp + q * P * P + (y + 1) * P + (z + 1);
    *dst = o;
  }
}

// ------------------------------------------------------------------------------------------------------------
// One level of "unpool -> ReLU' -> (1x1x1 conv)^T -> ReLU'" (see the header).  Rows of the MMA = the 16 fine voxels
// of two z-adjacent pooled voxels, exactly as pointwise_pool_rows_kernel (gb_cnn_tc.cu) lays them out.
//   yin : [pose][D][D][D][C] fp16, the 3x3x3 conv's ReLU'd output kept by the forward pass
//   gxa (+ gxb): gradient w.r.t. the pooled tensor, chunk-planar (Dn = D/2, Gn, C), 1/8 deferred
//   out : gradient w.r.t. yin's pre-activation, chunk-planar (D, Gf, C) = input of the next backward 3x3x3 conv
__device__ __forceinline__ void mma_16816_g(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_h2(uint32_t u) { return __half22float2(*reinterpret_cast<const __half2*>(&u)); }

// Data movement (same ideas as pointwise_pool_mma_kernel, gb_cnn_tc.cu):
//   * y rows are read with ONE 16-byte load per lane and 32 channels -- lane (g, t) holds channels [8t, 8t+8) of voxel
//     rows g and g+8 -- and used as A fragments under a permutation of the reduction index (k-slots 2t, 2t+1, 2t+8,
//     2t+9 of k-step s carry channels 8t + 4s + 0..3, +32 for the second load when C = 64);
//   * the output channels of the second GEMM are permuted the other way round (column pair (2t, 2t+1) of n-tile j is
//     channel pair 8t + 2j, +1), so that a lane's four accumulator pairs are exactly the 16-byte element (chunk t,
//     this voxel) of the chunk-planar output -- one 16-byte store -- and the y > 0 mask comes from the registers the
//     lane already loaded;
//   * both weight matrices live in shared memory in FRAGMENT order (one conflict-free 8-byte load per MMA).
template <int C, int D>
__global__ void __launch_bounds__(256) pw_backward_kernel(const __half* __restrict__ yin, const __half* __restrict__ w,
                                                          const float* __restrict__ bias, const uint32_t* __restrict__ gxa,
                                                          const uint32_t* __restrict__ gxb, uint4* __restrict__ out,
                                                          int n_poses, int Gn, int Lpn, int Gf, int Lpf) {
  constexpr int NT = C / 8, KS = C / 16, LD = C / 32;
  constexpr int Dn = D / 2, Pn = Dn + 2, Pf = D + 2;
  __shared__ uint2 s_wf[NT * KS * 32];   // first GEMM  u = y W^T : B fragments, k permuted
  __shared__ uint2 s_wtf[NT * KS * 32];  // second GEMM dy = d W  : B fragments, n permuted
  __shared__ float s_b[C];
  for (int e = threadIdx.x; e < NT * KS * 32; e += 256) {
    const int ln = e & 31, ks = (e >> 5) % KS, nt = e / (32 * KS);
    const int gg = ln >> 2, tt = ln & 3;
    {
      const int co = nt * 8 + gg, ch = 32 * (ks >> 1) + 8 * tt + 4 * (ks & 1);
      const uint32_t* p = reinterpret_cast<const uint32_t*>(w + (size_t)co * C + ch);
      s_wf[e] = make_uint2(p[0], p[1]);
    }
    {
      // n-tile nt, column gg <-> input channel ci; rows k = output channels 16 ks + (2tt, 2tt+1 | +8)
      const int ci = 32 * (nt >> 2) + 8 * (gg >> 1) + 2 * (nt & 3) + (gg & 1);
      const int co = 16 * ks + 2 * tt;
      const __half2 lo = __halves2half2(w[(size_t)co * C + ci], w[(size_t)(co + 1) * C + ci]);
      const __half2 hi = __halves2half2(w[(size_t)(co + 8) * C + ci], w[(size_t)(co + 9) * C + ci]);
      s_wtf[e] = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
    }
  }
  if (threadIdx.x < C) s_b[threadIdx.x] = bias[threadIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  // Work item = 16 fine voxels (the 16 MMA rows; lane (g, t) owns rows g and g+8).  The backward pass has no pooling
  // reduction -- every fine voxel just reads its pooled parent -- so any 16 voxels do.  D = 24: row g = (x, y, zb+g),
  // row g+8 = (x, y+1, zb+g) with zb a multiple of 8, so that the 8 lanes of a quad-column write 128 contiguous bytes
  // of each chunk plane (the 2x2x2-window mapping writes isolated 16-byte pieces: half-filled 32-byte sectors); both
  // rows share their pooled parent.  D = 12 (8 does not divide 12): rows = the 2x2x2 windows of two z-adjacent pooled
  // voxels.
  constexpr bool kRowZ = D % 8 == 0;
  constexpr int kItemsPerPose = D * D * D / 16;
  const int n_items = n_poses * kItemsPerPose;
  const int warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int n_warps = gridDim.x * (blockDim.x >> 5);
  for (int item = warp_global; item < n_items; item += n_warps) {
    int pose, xa, ya_, za, x0, y0a, z0a, z0b;  // fine voxel of row g; pooled parents of rows g / g+8 (x0, y0a shared)
    int rowb_off, posb_off;                    // row g+8 relative to row g: input halves, output positions
    if constexpr (kRowZ) {
      const int zb = (item % (D / 8)) * 8;
      int r = item / (D / 8);
      const int yp = r % Dn; r /= Dn;
      xa = r % D;
      pose = r / D;
      ya_ = 2 * yp; za = zb + g;
      x0 = xa >> 1; y0a = yp; z0a = za >> 1; z0b = z0a;
      rowb_off = D * C; posb_off = Pf;
    } else {
      const int pvA = 2 * item;
      const int z0 = pvA % Dn;
      int r = pvA / Dn;
      y0a = r % Dn; r /= Dn;
      x0 = r % Dn;
      pose = r / Dn;
      xa = 2 * x0 + ((g >> 2) & 1); ya_ = 2 * y0a + ((g >> 1) & 1); za = 2 * z0 + (g & 1);
      z0a = z0; z0b = z0 + 1;
      rowb_off = 2 * C; posb_off = 2;
    }
    const __half* rowA = yin + ((((size_t)pose * D + xa) * D + ya_) * D + za) * C + 8 * t;
    uint4 ya[LD], yb[LD];
#pragma unroll
    for (int l = 0; l < LD; l++) {
      ya[l] = *reinterpret_cast<const uint4*>(rowA + 32 * l);
      yb[l] = *reinterpret_cast<const uint4*>(rowA + rowb_off + 32 * l);
    }
    // gradient gathers issued before the first GEMM so their latency hides behind it
    const int grpn = pose / Gn, qn = pose % Gn;
    const size_t posn = (size_t)qn * Pn * Pn + (size_t)(y0a + 1) * Pn + (z0a + 1);
    const int gb_off = (z0b - z0a) * 4;  // uint32 units; 0 when both rows share the parent
    uint32_t gra[NT], grb[NT];
#pragma unroll
    for (int nt = 0; nt < NT; nt++) {
      const size_t gi = ((((size_t)grpn * Dn + x0) * NT + nt) * Lpn + posn) * 4 + t;
      gra[nt] = gxa[gi];
      grb[nt] = kRowZ ? gra[nt] : gxa[gi + gb_off];
    }
    uint32_t gra2[NT], grb2[NT];
    if (gxb) {
#pragma unroll
      for (int nt = 0; nt < NT; nt++) {
        const size_t gi = ((((size_t)grpn * Dn + x0) * NT + nt) * Lpn + posn) * 4 + t;
        gra2[nt] = gxb[gi];
        grb2[nt] = kRowZ ? gra2[nt] : gxb[gi + gb_off];
      }
    }
    // u = W y (+ b): pre-activation of the pointwise conv, for its ReLU mask
    float acc[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; nt++) { acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < KS; ks++) {
      const uint32_t* pa = reinterpret_cast<const uint32_t*>(&ya[ks >> 1]);
      const uint32_t* pb = reinterpret_cast<const uint32_t*>(&yb[ks >> 1]);
      const uint32_t a[4] = {pa[2 * (ks & 1)], pb[2 * (ks & 1)], pa[2 * (ks & 1) + 1], pb[2 * (ks & 1) + 1]};
#pragma unroll
      for (int nt = 0; nt < NT; nt++) {
        const uint2 b = s_wf[(nt * KS + ks) * 32 + lane];
        mma_16816_g(acc[nt], a, b.x, b.y);
      }
    }
    // d = unpool(gx) * [u > 0], re-packed from accumulator fragments into A fragments (natural channel order)
    uint32_t a2[KS][4];
#pragma unroll
    for (int nt = 0; nt < NT; nt++) {
      float2 fa = unpack_h2(gra[nt]), fb = unpack_h2(grb[nt]);
      if (gxb) {
        const float2 fa2 = unpack_h2(gra2[nt]), fb2 = unpack_h2(grb2[nt]);
        fa.x += fa2.x; fa.y += fa2.y; fb.x += fb2.x; fb.y += fb2.y;
      }
      const float b0 = s_b[nt * 8 + 2 * t], b1 = s_b[nt * 8 + 2 * t + 1];
      a2[nt >> 1][(nt & 1) * 2 + 0] = pack_h2(acc[nt][0] + b0 > 0.f ? fa.x : 0.f, acc[nt][1] + b1 > 0.f ? fa.y : 0.f);
      a2[nt >> 1][(nt & 1) * 2 + 1] = pack_h2(acc[nt][2] + b0 > 0.f ? fb.x : 0.f, acc[nt][3] + b1 > 0.f ? fb.y : 0.f);
    }
    // d_in = d W  (rows voxels, K = co natural, N = ci permuted: column pair (2t, 2t+1) of n-tile j = channels 8t+2j, +1)
#pragma unroll
    for (int nt = 0; nt < NT; nt++) { acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < KS; ks++)
#pragma unroll
      for (int nt = 0; nt < NT; nt++) {
        const uint2 b = s_wtf[(nt * KS + ks) * 32 + lane];
        mma_16816_g(acc[nt], a2[ks], b.x, b.y);
      }
    // mask with y > 0 (ReLU of the 3x3x3 conv) and store one 16-byte element per row and channel chunk
    const int grpf = pose / Gf, qf = pose % Gf;
    const size_t posf = (size_t)qf * Pf * Pf + (size_t)(ya_ + 1) * Pf + (za + 1);
#pragma unroll
    for (int l = 0; l < LD; l++) {
      uint4 oa, ob;
      uint32_t* wa = reinterpret_cast<uint32_t*>(&oa);
      uint32_t* wb = reinterpret_cast<uint32_t*>(&ob);
      const uint32_t* pa = reinterpret_cast<const uint32_t*>(&ya[l]);
      const uint32_t* pb = reinterpret_cast<const uint32_t*>(&yb[l]);
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const float2 va = unpack_h2(pa[j]), vb = unpack_h2(pb[j]);
        const float* c = acc[4 * l + j];
        wa[j] = pack_h2(va.x > 0.f ? c[0] : 0.f, va.y > 0.f ? c[1] : 0.f);
        wb[j] = pack_h2(vb.x > 0.f ? c[2] : 0.f, vb.y > 0.f ? c[3] : 0.f);
      }
      uint4* dst = out + (((size_t)grpf * D + xa) * NT + 4 * l + t) * Lpf + posf;
      dst[0] = oa;
      dst[posb_off] = ob;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// GridMaker::backward restricted to the ligand atoms, reading the POOLED grid gradient gx0 [pose][24^3][32] fp16:
// d grid[fine] = gx0[fine / 2] * scale.  Warp per atom, as grid_backward_kernel (gb_grid.cu).
__global__ void __launch_bounds__(256) grid_backward_pooled_kernel(const float4* __restrict__ atoms_xyzr,
                                                                   const int* __restrict__ atoms_ch,
                                                                   const int* __restrict__ pose_off,
                                                                   const float* __restrict__ centers, int n_channels,
                                                                   float resolution, float dimension,
                                                                   const __half* __restrict__ gx0, float scale,
                                                                   float* __restrict__ atom_grad, const float* __restrict__ rot) {
  constexpr int npts = 48, Dp = 24;
  const int p = blockIdx.y;
  const int a_i = pose_off[p] + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (a_i >= pose_off[p + 1]) return;
  float4 a = atoms_xyzr[a_i];
  float R[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f};  // G3, as grid_backward_kernel (gb_grid.cu)
  if (rot) {
    for (int k = 0; k < 9; k++) R[k] = rot[9 * p + k];
    const float cx = centers[3 * p], cy = centers[3 * p + 1], cz = centers[3 * p + 2];
    const float dx = a.x - cx, dy = a.y - cy, dz = a.z - cz;
    a.x = cx + (R[0] * dx + R[1] * dy + R[2] * dz);
    a.y = cy + (R[3] * dx + R[4] * dy + R[5] * dz);
    a.z = cz + (R[6] * dx + R[7] * dy + R[8] * dz);
  }
  const int ch = atoms_ch[a_i];
  const float half = dimension / 2.f;
  const float ox = centers[3 * p] - half, oy = centers[3 * p + 1] - half, oz = centers[3 * p + 2] - half;
  const float ar = a.w, reach = 1.5f * ar;
  int lo[3], hi[3];
  const float o[3] = {ox, oy, oz}, c[3] = {a.x, a.y, a.z};
  bool empty = false;
#pragma unroll
  for (int d = 0; d < 3; d++) {
    lo[d] = max(0, (int)floorf((c[d] - reach - o[d]) / resolution));
    hi[d] = min(npts - 1, (int)ceilf((c[d] + reach - o[d]) / resolution));
    empty |= lo[d] > hi[d];
  }
  float gx = 0.f, gy = 0.f, gz = 0.f;
  if (!empty && ch >= 0 && ch < n_channels) {
    const int ni = hi[0] - lo[0] + 1, nj = hi[1] - lo[1] + 1, nk = hi[2] - lo[2] + 1;
    const __half* gp = gx0 + (size_t)p * Dp * Dp * Dp * 32 + ch;
    const float e2 = 0.13533528323661270f, A = 4.f * e2, Bq = -12.f * e2;
    const float inv_r2 = 1.f / (ar * ar);
    for (int e = lane; e < ni * nj * nk; e += 32) {
      const int k = lo[2] + e % nk, j = lo[1] + (e / nk) % nj, i = lo[0] + e / (nk * nj);
      const float dx = (ox + i * resolution) - a.x, dy = (oy + j * resolution) - a.y, dz = (oz + k * resolution) - a.z;
      const float d2 = dx * dx + dy * dy + dz * dz;
      const float dist = sqrtf(d2);
      if (dist >= reach || dist == 0.f) continue;
      float dr;
      if (dist <= ar) dr = -4.f * dist * inv_r2 * __expf(-2.f * d2 * inv_r2);
      else dr = (2.f * A * (dist / ar) + Bq) / ar;
      const float gv = __half2float(gp[((size_t)((i >> 1) * Dp + (j >> 1)) * Dp + (k >> 1)) * 32]);
      const float sc = gv * dr / dist;
      gx -= sc * dx; gy -= sc * dy; gz -= sc * dz;
    }
  }
  for (int off = 16; off; off >>= 1) {
    gx += __shfl_xor_sync(0xffffffffu, gx, off);
    gy += __shfl_xor_sync(0xffffffffu, gy, off);
    gz += __shfl_xor_sync(0xffffffffu, gz, off);
  }
  if (lane == 0) {
    float* og = atom_grad + (size_t)a_i * 3;
    gx *= scale; gy *= scale; gz *= scale;
    og[0] = R[0] * gx + R[3] * gy + R[6] * gz;
    og[1] = R[1] * gx + R[4] * gy + R[7] * gz;
    og[2] = R[2] * gx + R[5] * gy + R[8] * gz;
  }
}

// ------------------------------------------------------------------------------------------------------------
int tc_backward(const Model& m, const TcPoseBatch& pb, TcWorkspace& ws, const float* out3, float* atom_grad, cudaStream_t s,
                Profiler* prof, const int* rec_off, float* rec_grad) {
  GB_CHECK(m.arch == GB_ARCH_DEFAULT2018 && tc_supported(m), "fast backward: default2018 family only");
  auto tw = get_tc_weights(m);
  auto gw = get_grad_weights(m);
  const int nb = pb.n_poses;
  const int nb_alloc = std::max(nb, 64);  // allocation floor, see tc_prepare_grid
  const ActLayout L1 = make_layout(24, 1, 32), L3 = make_layout(12, 2, 64), L5 = make_layout(6, 2, 64);
  // forward activations kept by tc_forward(keep_activations = true)
  const __half* Y1 = reinterpret_cast<const __half*>(ws.buf[0]);
  const __half* Y3 = reinterpret_cast<const __half*>(ws.buf[7]);
  const __half* Y5 = reinterpret_cast<const __half*>(ws.buf[3]);
  GB_CHECK(Y1 && Y3 && Y5, "tc_backward needs the activations of tc_forward(keep_activations = true)");
  ws.ensure(8, act_bytes(L5, nb_alloc));
  ws.ensure(9, act_bytes(L5, nb_alloc));
  ws.ensure(10, act_bytes(L5, nb_alloc));
  ws.ensure(11, act_bytes(L5, nb_alloc));
  ws.ensure(12, act_bytes(L3, nb_alloc));
  ws.ensure(13, act_bytes(make_layout(12, 2, 32), nb_alloc));
  ws.ensure(14, act_bytes(L1, nb_alloc));
  ws.ensure(15, (size_t)nb_alloc * 24 * 24 * 24 * 32 * sizeof(__half) + 1024);
  uint4 *G5a = reinterpret_cast<uint4*>(ws.buf[8]), *G5b = reinterpret_cast<uint4*>(ws.buf[9]);
  uint4 *GX4a = reinterpret_cast<uint4*>(ws.buf[10]), *GX4b = reinterpret_cast<uint4*>(ws.buf[11]);
  uint4* GY3 = reinterpret_cast<uint4*>(ws.buf[12]);
  uint4* GX2 = reinterpret_cast<uint4*>(ws.buf[13]);
  uint4* GY1 = reinterpret_cast<uint4*>(ws.buf[14]);
  __half* GX0 = reinterpret_cast<__half*>(ws.buf[15]);
  const int pw_blocks = 148 * 4;
  {
    ProfScope ps(prof, "tcg_heads_backward", s);
    heads_backward_f16_kernel<<<nb, 256, 0, s>>>(out3, tw->fcw, Y5, G5a, G5b, L5.Lp);
  }
  {
    ProfScope ps(prof, "tcg_dconv5_3x3x3_128x64_d6", s);
    launch_conv_tc_any(64, 6, gw->d5[0], L5, G5a, nullptr, nb, s, GX4a, 8, 0, L5.Lp, 0);
    launch_conv_tc_any(64, 6, gw->d5[1], L5, G5b, nullptr, nb, s, GX4b, 8, 0, L5.Lp, 0);
  }
  {
    ProfScope ps(prof, "tcg_unpool_dpw4", s);
    pw_backward_kernel<64, 12><<<pw_blocks, 256, 0, s>>>(Y3, tw->pw4.w, tw->pw4.bias, reinterpret_cast<const uint32_t*>(GX4a),
                                                         reinterpret_cast<const uint32_t*>(GX4b), GY3, nb, L5.G, L5.Lp, L3.G,
                                                         L3.Lp);
  }
  {
    ProfScope ps(prof, "tcg_dconv3_3x3x3_64x32_d12", s);
    launch_conv_tc_any(64, 12, gw->d3, L3, GY3, nullptr, nb, s, GX2, 4, 0, L3.Lp, 0);
  }
  {
    ProfScope ps(prof, "tcg_unpool_dpw2", s);
    pw_backward_kernel<32, 24><<<pw_blocks, 256, 0, s>>>(Y1, tw->pw2.w, tw->pw2.bias, reinterpret_cast<const uint32_t*>(GX2),
                                                         nullptr, GY1, nb, L3.G, L3.Lp, L1.G, L1.Lp);
  }
  {
    ProfScope ps(prof, "tcg_dconv1_3x3x3_32x28_d24", s);
    launch_conv_tc_any(32, 24, gw->d1, L1, GY1, GX0, nb, s, nullptr, 0, 0, 0, 0);
  }
  {
    ProfScope ps(prof, "tcg_grid_backward_atoms", s);
    if (pb.max_pose_atoms > 0) {
      dim3 g((pb.max_pose_atoms + 7) / 8, nb);
      grid_backward_pooled_kernel<<<g, 256, 0, s>>>(pb.lig_xyzr, pb.lig_ch, pb.lig_off, pb.centers, pb.n_channels, pb.resolution,
                                                    pb.dimension, GX0, 1.f / (512.f * kLossScale), atom_grad, pb.rot);
    }
  }
  if (rec_grad && pb.n_rec > 0) {  // getReceptorGradient: the receptor atoms of the (single) pose
    GB_CHECK(nb == 1 && rec_off, "receptor gradients need a single-pose chunk");
    ProfScope ps(prof, "tcg_grid_backward_receptor", s);
    dim3 g((pb.n_rec + 7) / 8, 1);
    grid_backward_pooled_kernel<<<g, 256, 0, s>>>(pb.rec_xyzr, pb.rec_ch, rec_off, pb.centers, pb.n_channels, pb.resolution,
                                                  pb.dimension, GX0, 1.f / (512.f * kLossScale), rec_grad, pb.rot);
  }
  tc_debug_set(5, GX0, (size_t)nb * 24 * 24 * 24 * 32 * sizeof(__half));
  tc_debug_set(6, GY1, act_bytes(L1, nb));
  tc_debug_set(7, GX2, act_bytes(make_layout(12, 2, 32), nb));
  return 8;
}

}  // namespace gb
