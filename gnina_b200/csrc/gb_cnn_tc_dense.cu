// Fast path of the dense family (N2 of SURVEY.md §8a; graph read from dense_1.3.pt):
//   maxpool2 -> conv3(28->32)+ReLU -> DB0@24^3 -> conv1(96->96)+ReLU -> maxpool2 -> DB1@12^3 -> conv1(160->160)+ReLU
//   -> maxpool2 -> DB2@6^3 -> global maxpool -> 224 -> {Linear->2, Linear->1}
//   DBk = 4 x [BatchNorm3d(eval) -> conv3(Cin->16, pad 1) -> ReLU -> concat]
//
// Everything lives in the chunk-planar padded layout of gb_cnn_tc.cu, which makes the channel concatenation free: a
// dense block is ONE buffer [group][x][C8 total][Lp][8 ch]; a layer reads its first Cin/8 chunks and writes two more.
//
// dense_conv_tc_kernel: 3x3x3 conv Cin -> 16 as tcgen05 implicit GEMM.  Differences from conv3_tc_kernel:
//   * K is streamed in blocks of 16 channels (weights 13.8 KB + slab 2 chunks per stage), so any Cin fits;
//   * the accumulators of ALL output planes of the item stay resident in TMEM (XS planes x 16 columns) and are
//     zero-initialised by the epilogue warps (tcgen05.st), so every MMA accumulates and the three dx taps are again
//     stacked in N (N = 48 -> planes xi-1, xi, xi+1);
//   * BatchNorm is folded into the weights (scale) and a per-border-class bias correction (shift): padding is applied
//     AFTER BN in the reference, so the shift must only be counted for in-bounds taps — 27 position classes.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_fp16.h>
#include <vector>
#include "gb_ptx.cuh"
#include "gb_tc.h"

namespace gb {

namespace ptx {
__device__ __forceinline__ void tmem_st16_zero(uint32_t taddr) {  // this warp's 32 lanes x 16 columns <- 0
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr),
      "r"(0u)
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
}  // namespace ptx

// ------------------------------------------------------------------------------------------------------------
struct DenseLayerTc {
  int cin = 0;
  uint4* wp = nullptr;   // [cin/16][9][2][48] x 16 B
  float* bias = nullptr; // [16]
  float* corr = nullptr; // [27][16]
};
struct PointwiseGen {
  int c = 0;
  __half* w = nullptr;   // [co][ci] fp16
  float* bias = nullptr;
};
struct TcDenseWeights {
  ConvTc init;
  DenseLayerTc layer[3][4];
  PointwiseGen bott[2];
  float* fcw = nullptr;  // [3][224]
  float* fcb = nullptr;
  std::vector<void*> allocs;
  ~TcDenseWeights() { for (void* p : allocs) cudaFree(p); }
};

template <typename T>
static T* dn_upload(TcDenseWeights& w, const std::vector<T>& h) {
  T* d = nullptr;
  GB_CUDA(cudaMalloc(&d, h.size() * sizeof(T)));
  w.allocs.push_back(d);
  GB_CUDA(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return d;
}

static DenseLayerTc prep_dense_layer(TcDenseWeights& tw, const Model& m, int L, int i) {
  const std::string base = "dense_block_" + std::to_string(L) + ".data_enc_level" + std::to_string(L);
  const std::string cv = base + "_conv" + std::to_string(i), bn = base + "_batchnorm_conv" + std::to_string(i);
  const HostTensor &w = m.t(cv + ".weight"), &b = m.t(cv + ".bias"), &g = m.t(bn + ".weight"), &be = m.t(bn + ".bias"),
                   &mu = m.t(bn + ".running_mean"), &var = m.t(bn + ".running_var");
  DenseLayerTc d;
  d.cin = w.shape[1];
  GB_CHECK(w.shape[0] == 16 && d.cin % 16 == 0 && w.shape[2] == 3, "dense layer shape");
  std::vector<float> sc(d.cin), sh(d.cin);
  for (int c = 0; c < d.cin; c++) {  // eval-mode BatchNorm3d, eps 1e-5
    sc[c] = g.data[c] / std::sqrt(var.data[c] + 1e-5f);
    sh[c] = be.data[c] - mu.data[c] * sc[c];
  }
  const int KS = d.cin / 16;
  std::vector<__half> h((size_t)KS * 9 * 2 * 48 * 8, __float2half(0.f));
  std::vector<double> S((size_t)16 * 27, 0.0);  // S[co][tap] = sum_ci W[co][ci][tap] * shift[ci]
  for (int co = 0; co < 16; co++)
    for (int ci = 0; ci < d.cin; ci++)
      for (int kx = 0; kx < 3; kx++)
        for (int ky = 0; ky < 3; ky++)
          for (int kz = 0; kz < 3; kz++) {
            const float wv = w.data[((((size_t)co * d.cin + ci) * 3 + kx) * 3 + ky) * 3 + kz];
            S[(size_t)co * 27 + (kx * 3 + ky) * 3 + kz] += (double)wv * sh[ci];
            const int ks = ci / 16, c2 = (ci % 16) / 8, e = ci % 8, blk = 2 - kx;  // blk 0 <-> dx=+1 (output plane xi-1)
            h[((((size_t)ks * 9 + ky * 3 + kz) * 2 + c2) * 48 + blk * 16 + co) * 8 + e] = __float2half(wv * sc[ci]);
          }
  // border classes: class = (cx*3 + cy)*3 + cz with c = 0 (first index), 1 (interior), 2 (last index); a tap with
  // offset -1 is out of bounds in class 0, offset +1 in class 2
  std::vector<float> corr((size_t)27 * 16);
  for (int cx = 0; cx < 3; cx++)
    for (int cy = 0; cy < 3; cy++)
      for (int cz = 0; cz < 3; cz++)
        for (int co = 0; co < 16; co++) {
          double acc = 0;
          for (int kx = 0; kx < 3; kx++)
            for (int ky = 0; ky < 3; ky++)
              for (int kz = 0; kz < 3; kz++) {
                const bool ok = !(cx == 0 && kx == 0) && !(cx == 2 && kx == 2) && !(cy == 0 && ky == 0) && !(cy == 2 && ky == 2) &&
                                !(cz == 0 && kz == 0) && !(cz == 2 && kz == 2);
                if (ok) acc += S[(size_t)co * 27 + (kx * 3 + ky) * 3 + kz];
              }
          corr[(size_t)((cx * 3 + cy) * 3 + cz) * 16 + co] = (float)acc;
        }
  d.wp = reinterpret_cast<uint4*>(dn_upload(tw, h));
  d.bias = dn_upload(tw, std::vector<float>(b.data, b.data + 16));
  d.corr = dn_upload(tw, corr);
  return d;
}

static std::shared_ptr<TcDenseWeights> get_dense_weights(const Model& m) {
  std::lock_guard<std::mutex> lk(tc_init_mutex());
  Model& mm = const_cast<Model&>(m);
  if (mm.tc_dense) return mm.tc_dense;
  auto tw = std::make_shared<TcDenseWeights>();
  {  // init conv 28 -> 32: same packing as the default2018 convs (prep_conv in gb_cnn_tc.cu), restated here
    const HostTensor &w = m.t("data_enc_init_conv.weight"), &b = m.t("data_enc_init_conv.bias");
    const int cout = w.shape[0], cin = w.shape[1];
    GB_CHECK(cout == 32 && cin == 28, "dense init conv shape");
    std::vector<__half> h((size_t)9 * 4 * 96 * 8, __float2half(0.f));
    for (int ky = 0; ky < 3; ky++)
      for (int kz = 0; kz < 3; kz++)
        for (int ci = 0; ci < cin; ci++)
          for (int blk = 0; blk < 3; blk++)
            for (int co = 0; co < 32; co++)
              h[((((size_t)ky * 3 + kz) * 4 + ci / 8) * 96 + blk * 32 + co) * 8 + ci % 8] =
                  __float2half(w.data[((((size_t)co * cin + ci) * 3 + (2 - blk)) * 3 + ky) * 3 + kz]);
    tw->init.cin = 32; tw->init.cout = 32;
    tw->init.wp = reinterpret_cast<uint4*>(dn_upload(*tw, h));
    tw->init.bias = dn_upload(*tw, std::vector<float>(b.data, b.data + 32));
  }
  for (int L = 0; L < 3; L++)
    for (int i = 0; i < 4; i++) tw->layer[L][i] = prep_dense_layer(*tw, m, L, i);
  for (int L = 0; L < 2; L++) {
    const std::string k = "data_enc_level" + std::to_string(L) + "_bottleneck";
    const HostTensor &w = m.t(k + ".weight"), &b = m.t(k + ".bias");
    PointwiseGen p;
    p.c = w.shape[0];
    std::vector<__half> h((size_t)p.c * p.c);
    for (size_t q = 0; q < h.size(); q++) h[q] = __float2half(w.data[q]);
    p.w = dn_upload(*tw, h);
    p.bias = dn_upload(*tw, std::vector<float>(b.data, b.data + p.c));
    tw->bott[L] = p;
  }
  const HostTensor &pw = m.t("pose_output.weight"), &pb = m.t("pose_output.bias"), &aw = m.t("affinity_output.weight"),
                   &ab = m.t("affinity_output.bias");
  GB_CHECK(pw.shape[1] == 224, "dense fc features");
  std::vector<float> fw(3 * 224), fb(3);
  memcpy(fw.data(), pw.data, sizeof(float) * 2 * 224);
  memcpy(fw.data() + 2 * 224, aw.data, sizeof(float) * 224);
  fb[0] = pb.data[0]; fb[1] = pb.data[1]; fb[2] = ab.data[0];
  tw->fcw = dn_upload(*tw, fw);
  tw->fcb = dn_upload(*tw, fb);
  mm.tc_dense = tw;
  return tw;
}

// ------------------------------------------------------------------------------------------------------------
struct DenseConvParams {
  const uint4* xin;   // block buffer [group][x][C8tot][Lp]
  uint4* xout;        // same buffer
  const uint4* wp;    // [KS][9][2][48]
  const float* bias;  // [16]
  const float* corr;  // [27][16]
  int KS, C8tot, c8_off, G, T, Lp, n_poses, n_groups;
};

constexpr int kDnStages = 4;
constexpr int kDnWBytes = 9 * 2 * 48 * 16;  // 13,824 B per 16-channel K block

template <int DD> struct DenseCfg { static constexpr int id = DD == 24 ? 0 : DD == 12 ? 1 : 2; static constexpr int XS = DD == 24 ? 12 : DD; };

template <int DD>
__global__ void __launch_bounds__(192) dense_conv_tc_kernel(const DenseConvParams p) {
  constexpr int D = DD, P = DD + 2, SL = 128 + 2 * (P + 1), XS = DenseCfg<DD>::XS;
  constexpr int kTmemCols = XS * 16 <= 128 ? 128 : 256;
  constexpr int kStageBytes = ((2 * SL * 16) + 127) / 128 * 128;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_w = smem;                          // 2 x kDnWBytes
  uint8_t* s_stage = smem + 2 * kDnWBytes;      // kDnStages x kStageBytes
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * kDnWBytes + kDnStages * kStageBytes);
  uint64_t* full = bars;                   // [kDnStages]
  uint64_t* empty = bars + kDnStages;      // [kDnStages]
  uint64_t* wfull = bars + 2 * kDnStages;  // [2]
  uint64_t* wempty = wfull + 2;            // [2]
  uint64_t* zeroed = wempty + 2;           // accumulators zero-initialised
  uint64_t* accdone = zeroed + 1;          // all MMAs of the item complete
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(accdone + 1);
  float* s_bias = reinterpret_cast<float*>(s_tmem + 2);  // 16
  float* s_corr = s_bias + 16;                           // 27 * 16

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int NXH = D / XS;
  const int xh = blockIdx.x % NXH;
  const int j = (blockIdx.x / NXH) % p.T;
  const int g = blockIdx.x / (NXH * p.T);
  const int x_lo = xh * XS, x_hi = x_lo + XS - 1;            // output planes (0-based)
  const int xi_lo = x_lo > 0 ? x_lo - 1 : 0, xi_hi = x_hi < D - 1 ? x_hi + 1 : D - 1;  // input planes
  const uint32_t slab_row = (uint32_t)SL * 16u;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kDnStages; s++) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; s++) { ptx::mbar_init(&wfull[s], 1); ptx::mbar_init(&wempty[s], 1); }
    ptx::mbar_init(zeroed, 128);
    ptx::mbar_init(accdone, 1);
    ptx::fence_mbar_init();
  }
  if (threadIdx.x < 16) s_bias[threadIdx.x] = p.bias[threadIdx.x];
  for (int e = threadIdx.x; e < 27 * 16; e += blockDim.x) s_corr[e] = p.corr[e];
  if (warp == 1) {
    ptx::tmem_alloc(s_tmem, kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp == 0) {
    // ===== producer: per K block the weights, then one 2-chunk slab per input plane =====
    const uint4* xg = p.xin + (size_t)g * D * p.C8tot * p.Lp + (size_t)128 * j;
    uint32_t gp = 0;
    for (int ks = 0; ks < p.KS; ks++) {
      const uint32_t wb = ks & 1, wph = (ks >> 1) & 1;
      ptx::mbar_wait(&wempty[wb], wph ^ 1);
      if (ptx::elect_one()) {
        ptx::mbar_expect_tx(&wfull[wb], kDnWBytes);
        ptx::bulk_g2s(s_w + wb * kDnWBytes, reinterpret_cast<const uint8_t*>(p.wp) + (size_t)ks * kDnWBytes, kDnWBytes, &wfull[wb]);
      }
      __syncwarp();
      for (int xi = xi_lo; xi <= xi_hi; xi++, gp++) {
        const uint32_t st = gp % kDnStages, ph = (gp / kDnStages) & 1;
        ptx::mbar_wait(&empty[st], ph ^ 1);
        if (ptx::elect_one()) {
          ptx::mbar_expect_tx(&full[st], 2u * slab_row);
          uint8_t* dst = s_stage + (size_t)st * kStageBytes;
          const uint4* src = xg + ((size_t)xi * p.C8tot + 2 * ks) * p.Lp;
          ptx::bulk_g2s(dst, src, slab_row, &full[st]);
          ptx::bulk_g2s(dst + slab_row, src + p.Lp, slab_row, &full[st]);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);
    const uint32_t a_lo_fixed = ((uint32_t)SL & 0x3FFFu) << 16;
    ptx::mbar_wait(zeroed, 0);
    ptx::tc_fence_after();
    uint32_t gp = 0;
    for (int ks = 0; ks < p.KS; ks++) {
      const uint32_t wb = ks & 1, wph = (ks >> 1) & 1;
      ptx::mbar_wait(&wfull[wb], wph);
      const uint32_t b_lo_base = (48u << 16) | (ptx::smem_u32(s_w + wb * kDnWBytes) >> 4);  // LBO = 48 rows x 16 B
      for (int xi = xi_lo; xi <= xi_hi; xi++, gp++) {
        const uint32_t st = gp % kDnStages, ph = (gp / kDnStages) & 1;
        const int lo = (xi - 1 > x_lo) ? xi - 1 : x_lo, hi = (xi + 1 < x_hi) ? xi + 1 : x_hi;  // output planes fed by xi
        const uint32_t tm = tmem_base + (uint32_t)(lo - x_lo) * 16u;
        const uint32_t bl = b_lo_base + (uint32_t)(lo - (xi - 1)) * 16u;  // skip weight row blocks of planes outside the item
        const uint32_t idesc = ptx::idesc_f16(128, 16 * (hi - lo + 1));
        ptx::mbar_wait(&full[st], ph);
        ptx::tc_fence_after();
        const uint32_t a_lo_base = a_lo_fixed | (ptx::smem_u32(s_stage + (size_t)st * kStageBytes) >> 4);
        if (ptx::elect_one()) {
          if (hi >= lo) {
            // unrolled, compile-time operand offsets (the kernels are MMA-issue bound, see conv3_tc_kernel)
#pragma unroll
            for (int t9 = 0; t9 < 9; t9++)
              ptx::mma_f16_ss_lohi<1>(tm, a_lo_base + (uint32_t)((DD + 3) + (t9 / 3 - 1) * (DD + 2) + (t9 % 3 - 1)), kDescHi,
                                      bl + (uint32_t)(t9 * 2 * 48), kDescHi, idesc);
          }
          ptx::tc_commit(&empty[st]);
          if (xi == xi_hi) ptx::tc_commit(&wempty[wb]);
          if (xi == xi_hi && ks == p.KS - 1) ptx::tc_commit(accdone);
        }
        __syncwarp();
      }
    }
  } else {
    // ===== epilogue warps: zero the accumulators, later drain them =====
    const int q4 = warp & 3;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q4 * 32) << 16);
    for (int x = 0; x < XS; x++) ptx::tmem_st16_zero(lane_base + (uint32_t)x * 16u);
    ptx::tmem_st_wait();
    ptx::tc_fence_before();
    ptx::mbar_arrive(zeroed);
    const int row = q4 * 32 + lane;
    const int m = (P + 1) + 128 * j + row;
    const int qpose = m / (P * P), rem = m % (P * P);
    const int y = rem / P, z = rem % P;
    const int pose = g * p.G + qpose;
    const bool valid = qpose < p.G && pose < p.n_poses && y >= 1 && y <= D && z >= 1 && z <= D;
    const int cyz = ((y == 1 ? 0 : (y == D ? 2 : 1)) * 3 + (z == 1 ? 0 : (z == D ? 2 : 1)));
    ptx::mbar_wait(accdone, 0);
    ptx::tc_fence_after();
    for (int x = x_lo; x <= x_hi; x++) {
      uint32_t v[16];
      ptx::tmem_ld16(lane_base + (uint32_t)(x - x_lo) * 16u, v);
      ptx::tmem_ld_wait();
      if (valid) {
        const float* cr = s_corr + ((x == 0 ? 0 : (x == D - 1 ? 2 : 1)) * 9 + cyz) * 16;
        uint4 o[2];
        uint32_t* ow = reinterpret_cast<uint32_t*>(o);
#pragma unroll
        for (int c = 0; c < 8; c++) {
          const float f0 = fmaxf(__uint_as_float(v[2 * c]) + s_bias[2 * c] + cr[2 * c], 0.f);
          const float f1 = fmaxf(__uint_as_float(v[2 * c + 1]) + s_bias[2 * c + 1] + cr[2 * c + 1], 0.f);
          const __half2 h = __floats2half2_rn(f0, f1);
          ow[c] = *reinterpret_cast<const uint32_t*>(&h);
        }
        uint4* dst = p.xout + (((size_t)g * D + x) * p.C8tot + p.c8_off) * p.Lp + m;
        dst[0] = o[0];
        dst[p.Lp] = o[1];
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <int DD>
static void launch_dense_conv(const DenseLayerTc& L, const ActLayout& A, uint4* buf, int C8tot, int c8_off, int n_poses, cudaStream_t s) {
  constexpr int P = DD + 2, SL = 128 + 2 * (P + 1);
  constexpr int kStageBytes = ((2 * SL * 16) + 127) / 128 * 128;
  constexpr int kSmem = 2 * kDnWBytes + kDnStages * kStageBytes + 16 * 8 + 16 + (16 + 27 * 16) * 4 + 64;
  {
    std::lock_guard<std::mutex> init_lock(tc_init_mutex());
    static bool init[64] = {};   // per device
    int dev = 0;
    GB_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !init[dev]) {
      GB_CUDA(cudaFuncSetAttribute(dense_conv_tc_kernel<DD>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
      init[dev] = true;
    }
  }
  DenseConvParams p;
  p.xin = buf; p.xout = buf; p.wp = L.wp; p.bias = L.bias; p.corr = L.corr;
  p.KS = L.cin / 16; p.C8tot = C8tot; p.c8_off = c8_off; p.G = A.G; p.T = A.T; p.Lp = A.Lp; p.n_poses = n_poses;
  p.n_groups = (n_poses + A.G - 1) / A.G;
  const int grid = p.n_groups * A.T * (DD / DenseCfg<DD>::XS);
  dense_conv_tc_kernel<DD><<<grid, 192, kSmem, s>>>(p);
}

// ------------------------------------------------------------------------------------------------------------
// Bottleneck: 1x1x1 conv C -> C + bias + ReLU + 2x2x2 MAX pool, chunk-planar in (D, Gin) -> chunk-planar out (D/2, Gout),
// writing chunks [0, C/8) of the next block buffer.  mma.sync m16n8k16, transposed formulation (rows = output
// channels from shared-memory weights, columns = the 8 fine voxels of one pooled voxel); ReLU and max commute.
__device__ __forceinline__ void mma_16816_d(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int C>
__global__ void __launch_bounds__(256) bottleneck_maxpool_kernel(const uint4* __restrict__ xin, int D, int Gin, int Lpin, int C8in,
                                                                  const __half* __restrict__ w, const float* __restrict__ bias,
                                                                  __half* __restrict__ xout, int Gout, int Lpout, int C8out, int n_poses) {
  constexpr int MT = C / 16, KS = C / 16, WP = C + 8;  // padded weight rows: conflict-free fragment loads
  extern __shared__ __align__(16) uint8_t smem_b[];
  __half* s_w = reinterpret_cast<__half*>(smem_b);
  float* s_b = reinterpret_cast<float*>(smem_b + (size_t)C * WP * sizeof(__half));
  for (int e = threadIdx.x; e < C * C; e += blockDim.x) s_w[(e / C) * WP + (e % C)] = w[e];
  for (int e = threadIdx.x; e < C; e += blockDim.x) s_b[e] = bias[e];
  __syncthreads();
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int Pin = D + 2, Dn = D / 2, Pn = Dn + 2;
  const int di = (g >> 2) & 1, dj = (g >> 1) & 1, dk = g & 1;
  const int n_pv = n_poses * Dn * Dn * Dn;
  const int warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), n_warps = gridDim.x * (blockDim.x >> 5);
  const __half* xh = reinterpret_cast<const __half*>(xin);
  for (int pv = warp_global; pv < n_pv; pv += n_warps) {
    const int z0 = pv % Dn;
    int r = pv / Dn;
    const int y0 = r % Dn; r /= Dn;
    const int x0 = r % Dn;
    const int pose = r / Dn;
    const int gi = pose / Gin, qi = pose % Gin;
    // B fragments: column n = g is fine voxel (di,dj,dk); element (k = channel) lives at chunk k/8
    const size_t pos_in = (size_t)qi * Pin * Pin + (size_t)(2 * y0 + dj + 1) * Pin + (2 * z0 + dk + 1);
    const __half* base = xh + ((((size_t)gi * D + (2 * x0 + di)) * C8in) * Lpin + pos_in) * 8;
    uint32_t bfr[KS][2];
#pragma unroll
    for (int ks = 0; ks < KS; ks++) {
      // channels ks*16 + 2t, +1 -> chunk 2ks, offset 2t ; channels ks*16 + 8 + 2t -> chunk 2ks+1
      bfr[ks][0] = *reinterpret_cast<const uint32_t*>(base + ((size_t)(2 * ks) * Lpin) * 8 + 2 * t);
      bfr[ks][1] = *reinterpret_cast<const uint32_t*>(base + ((size_t)(2 * ks + 1) * Lpin) * 8 + 2 * t);
    }
    const int go = pose / Gout, qo = pose % Gout;
    const size_t pos_out = (size_t)qo * Pn * Pn + (size_t)(y0 + 1) * Pn + (z0 + 1);
#pragma unroll 1
    for (int mt = 0; mt < MT; mt++) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int ks = 0; ks < KS; ks++) {
        uint32_t a[4];
        const __half* w0 = s_w + (size_t)(mt * 16 + g) * WP + ks * 16 + 2 * t;
        a[0] = *reinterpret_cast<const uint32_t*>(w0);
        a[1] = *reinterpret_cast<const uint32_t*>(w0 + 8 * WP);
        a[2] = *reinterpret_cast<const uint32_t*>(w0 + 8);
        a[3] = *reinterpret_cast<const uint32_t*>(w0 + 8 * WP + 8);
        mma_16816_d(acc, a, bfr[ks][0], bfr[ks][1]);
      }
      float s0 = fmaxf(acc[0], acc[1]), s1 = fmaxf(acc[2], acc[3]);  // max over voxels, then bias + ReLU (monotone)
      s0 = fmaxf(s0, __shfl_xor_sync(0xffffffffu, s0, 1));
      s1 = fmaxf(s1, __shfl_xor_sync(0xffffffffu, s1, 1));
      s0 = fmaxf(s0, __shfl_xor_sync(0xffffffffu, s0, 2));
      s1 = fmaxf(s1, __shfl_xor_sync(0xffffffffu, s1, 2));
      if (t == 0) {
        __half* o0 = xout + ((((size_t)go * Dn + x0) * C8out + 2 * mt) * Lpout + pos_out) * 8 + g;
        o0[0] = __float2half(fmaxf(s0 + s_b[mt * 16 + g], 0.f));
        o0[(size_t)Lpout * 8] = __float2half(fmaxf(s1 + s_b[mt * 16 + g + 8], 0.f));
      }
    }
  }
}

// global max pool over the 6^3 volume of the 224-channel block buffer + FC heads
__global__ void __launch_bounds__(256) dense_heads_kernel(const uint4* __restrict__ b2, int G, int Lp, int C8, const float* __restrict__ w,
                                                          const float* __restrict__ bias, float* __restrict__ out3) {
  __shared__ float feat[224];
  __shared__ float red[3][8];
  const int pose = blockIdx.x, gi = pose / G, q = pose % G;
  const __half* xh = reinterpret_cast<const __half*>(b2);
  const int c = threadIdx.x;
  if (c < 224) {
    float v = -INFINITY;
    for (int x = 0; x < 6; x++)
      for (int y = 0; y < 6; y++)
        for (int z = 0; z < 6; z++) {
          const size_t pos = (size_t)q * 64 + (size_t)(y + 1) * 8 + (z + 1);
          v = fmaxf(v, __half2float(xh[((((size_t)gi * 6 + x) * C8 + c / 8) * Lp + pos) * 8 + (c % 8)]));
        }
    feat[c] = v;
  }
  __syncthreads();
  float a0 = 0, a1 = 0, a2 = 0;
  if (c < 224) { a0 = feat[c] * w[c]; a1 = feat[c] * w[224 + c]; a2 = feat[c] * w[448 + c]; }
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
    for (int k = 0; k < 8; k++) v += red[threadIdx.x][k];
    out3[(size_t)pose * 3 + threadIdx.x] = v;
  }
}

template <int C>
static void launch_bottleneck(const PointwiseGen& pw, const uint4* xin, const ActLayout& Ain, uint4* xout, const ActLayout& Aout,
                              int n_poses, cudaStream_t s) {
  const int smem = C * (C + 8) * (int)sizeof(__half) + C * (int)sizeof(float);
  {
    std::lock_guard<std::mutex> lk(tc_init_mutex());
    static bool init[64] = {};   // per device
    int dev = 0;
    GB_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !init[dev]) {
      GB_CUDA(cudaFuncSetAttribute(bottleneck_maxpool_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      init[dev] = true;
    }
  }
  GB_CHECK(pw.c == C, "bottleneck channels");
  bottleneck_maxpool_kernel<C><<<148 * 3, 256, smem, s>>>(xin, Ain.D, Ain.G, Ain.Lp, Ain.C8, pw.w, pw.bias,
                                                          reinterpret_cast<__half*>(xout), Aout.G, Aout.Lp, Aout.C8, n_poses);
}

int tc_forward_dense(const Model& m, const TcPoseBatch& pb, const void* x0max, TcWorkspace& ws, float* out3, cudaStream_t s,
                     Profiler* prof, cudaEvent_t x0_consumed) {
  auto tw = get_dense_weights(m);
  const int nb = pb.n_poses;
  const int nb_alloc = std::max(nb, 64);  // allocation floor, see tc_prepare_grid
  const ActLayout A0 = make_layout(24, 1, 96), A1 = make_layout(12, 2, 160), A2 = make_layout(6, 2, 224);
  ws.ensure(4, act_bytes(A0, nb_alloc));
  ws.ensure(5, act_bytes(A1, nb_alloc));
  ws.ensure(6, act_bytes(A2, nb_alloc));
  uint4* B0 = reinterpret_cast<uint4*>(ws.buf[4]);
  uint4* B1 = reinterpret_cast<uint4*>(ws.buf[5]);
  uint4* B2 = reinterpret_cast<uint4*>(ws.buf[6]);
  int launches = 0;
  {
    ProfScope ps(prof, "tcd_init_conv_28x32_d24", s);
    launch_conv_tc_32_24_planar(tw->init, reinterpret_cast<const uint4*>(x0max), B0, A0.C8, 0, A0.Lp, nb, s);
  }
  if (x0_consumed) GB_CUDA(cudaEventRecord(x0_consumed, s));
  launches++;
  for (int i = 0; i < 4; i++) {
    ProfScope ps(prof, "tcd_block0_conv_d24", s);
    launch_dense_conv<24>(tw->layer[0][i], A0, B0, A0.C8, 4 + 2 * i, nb, s);
    launches++;
  }
  {
    ProfScope ps(prof, "tcd_bottleneck0_maxpool", s);
    launch_bottleneck<96>(tw->bott[0], B0, A0, B1, A1, nb, s);
    launches++;
  }
  for (int i = 0; i < 4; i++) {
    ProfScope ps(prof, "tcd_block1_conv_d12", s);
    launch_dense_conv<12>(tw->layer[1][i], A1, B1, A1.C8, 12 + 2 * i, nb, s);
    launches++;
  }
  {
    ProfScope ps(prof, "tcd_bottleneck1_maxpool", s);
    launch_bottleneck<160>(tw->bott[1], B1, A1, B2, A2, nb, s);
    launches++;
  }
  for (int i = 0; i < 4; i++) {
    ProfScope ps(prof, "tcd_block2_conv_d6", s);
    launch_dense_conv<6>(tw->layer[2][i], A2, B2, A2.C8, 20 + 2 * i, nb, s);
    launches++;
  }
  {
    ProfScope ps(prof, "tcd_globalmax_heads", s);
    dense_heads_kernel<<<nb, 256, 0, s>>>(B2, A2.G, A2.Lp, A2.C8, tw->fcw, tw->fcb, out3);
    launches++;
  }
  tc_debug_set(0, x0max, act_bytes(make_layout(24, 1, 32), nb));
  tc_debug_set(5, B0, act_bytes(A0, nb));
  tc_debug_set(6, B1, act_bytes(A1, nb));
  tc_debug_set(7, B2, act_bytes(A2, nb));
  return launches;
}

}  // namespace gb
