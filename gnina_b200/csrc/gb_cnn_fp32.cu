// fp32 validation path of the CNN forward (N1/N2/N3 of SURVEY.md §8a): plain CUDA-core kernels on NCDHW fp32
// tensors, numerically equivalent to the TorchScript graphs the reference runs at gninasrc/lib/torch_model.cpp:185
// up to fp32 summation order.  It exists to (a) pin the fast tensor-core path, (b) provide a <=1e-4 validation mode.
#include <cstdio>
#include "gb_internal.h"

namespace gb {

void Fp32Workspace::ensure(int i, size_t n) {
  if (cap[i] >= n) return;
  if (buf[i]) cudaFree(buf[i]);
  GB_CUDA(cudaMalloc(&buf[i], n * sizeof(float)));
  cap[i] = n;
}
void Fp32Workspace::ensure_feat(size_t n) {
  if (feat_cap >= n) return;
  if (feat) cudaFree(feat);
  GB_CUDA(cudaMalloc(&feat, n * sizeof(float)));
  feat_cap = n;
}
Fp32Workspace::~Fp32Workspace() {
  for (auto p : buf)
    if (p) cudaFree(p);
  if (feat) cudaFree(feat);
}

// ---------------------------------------------------------------------------------------------------------
// Direct 3D convolution, stride 1, "same" zero padding (KS=3,pad 1 | KS=1,pad 0), optional per-input-channel
// affine (eval-mode BatchNorm folded to scale/shift, applied to in-bounds inputs only, so padded taps stay 0
// exactly like BN -> conv(pad=1) in the dense blocks), bias + optional ReLU.
// in : [B][in_ctot][D][D][D]  (first Cin channels are read)
// out: [B][out_ctot][D][D][D] (channels [out_coff, out_coff+Cout) are written)
// w  : [Cin][KS^3][Cout]
// CTA = 8x8x8 output voxels x COT output channels; input channels streamed in chunks of CIC through smem.
template <int KS, int COT>
__global__ void __launch_bounds__(512) conv3d_f32_kernel(const float* __restrict__ in, int in_ctot, int Cin,
                                                         const float* __restrict__ w, const float* __restrict__ bias,
                                                         const float* __restrict__ bn_scale,
                                                         const float* __restrict__ bn_shift, float* __restrict__ out,
                                                         int out_ctot, int out_coff, int Cout, int D, int relu,
                                                         const float* __restrict__ mask,
                                                         const float* __restrict__ out_scale = nullptr,
                                                         int accumulate = 0) {
  constexpr int CIC = 4;
  constexpr int H = KS / 2;
  constexpr int TW = 8 + 2 * H;
  constexpr int K3 = KS * KS * KS;
  __shared__ float s_in[CIC][TW][TW][TW];
  __shared__ __align__(16) float s_w[CIC][K3][COT];
  const int tiles = (D + 7) / 8;
  const int t = blockIdx.x;
  const int ti = t / (tiles * tiles), tj = (t / tiles) % tiles, tk = t % tiles;
  const int co0 = blockIdx.y * COT;
  const int b = blockIdx.z;
  const int li = threadIdx.x >> 6, lj = (threadIdx.x >> 3) & 7, lk = threadIdx.x & 7;
  const int i = ti * 8 + li, j = tj * 8 + lj, k = tk * 8 + lk;
  const size_t vol = (size_t)D * D * D;
  const float* inb = in + (size_t)b * in_ctot * vol;
  float acc[COT];
#pragma unroll
  for (int c = 0; c < COT; c++) acc[c] = 0.f;

  for (int c0 = 0; c0 < Cin; c0 += CIC) {
    // stage input tile (with halo) and weight slice
    for (int e = threadIdx.x; e < CIC * TW * TW * TW; e += 512) {
      const int ci = e / (TW * TW * TW);
      int r = e % (TW * TW * TW);
      const int a = r / (TW * TW), bb = (r / TW) % TW, cc = r % TW;
      const int gi = ti * 8 + a - H, gj = tj * 8 + bb - H, gk = tk * 8 + cc - H;
      float v = 0.f;
      if (c0 + ci < Cin && gi >= 0 && gi < D && gj >= 0 && gj < D && gk >= 0 && gk < D) {
        const size_t gidx = (size_t)(c0 + ci) * vol + ((size_t)gi * D + gj) * D + gk;
        v = inb[gidx];
        if (bn_scale) v = fmaf(v, bn_scale[c0 + ci], bn_shift[c0 + ci]);
        // ReLU backward folded into the load: mask is the forward activation of the same shape
        if (mask && !(mask[(size_t)b * in_ctot * vol + gidx] > 0.f)) v = 0.f;
      }
      (&s_in[0][0][0][0])[e] = v;
    }
    for (int e = threadIdx.x; e < CIC * K3 * COT; e += 512) {
      const int ci = e / (K3 * COT);
      const int r = e % (K3 * COT);
      const int tap = r / COT, co = r % COT;
      float v = 0.f;
      if (c0 + ci < Cin && co0 + co < Cout) v = w[((size_t)(c0 + ci) * K3 + tap) * Cout + co0 + co];
      s_w[ci][tap][co] = v;
    }
    __syncthreads();
#pragma unroll
    for (int ci = 0; ci < CIC; ci++) {
#pragma unroll
      for (int a = 0; a < KS; a++)
#pragma unroll
        for (int bb = 0; bb < KS; bb++)
#pragma unroll
          for (int cc = 0; cc < KS; cc++) {
            const float v = s_in[ci][li + a][lj + bb][lk + cc];
            const float4* wr = reinterpret_cast<const float4*>(&s_w[ci][(a * KS + bb) * KS + cc][0]);
#pragma unroll
            for (int q = 0; q < COT / 4; q++) {
              const float4 ww = wr[q];
              acc[4 * q + 0] = fmaf(v, ww.x, acc[4 * q + 0]);
              acc[4 * q + 1] = fmaf(v, ww.y, acc[4 * q + 1]);
              acc[4 * q + 2] = fmaf(v, ww.z, acc[4 * q + 2]);
              acc[4 * q + 3] = fmaf(v, ww.w, acc[4 * q + 3]);
            }
          }
    }
    __syncthreads();
  }
  if (i < D && j < D && k < D) {
    float* ob = out + ((size_t)b * out_ctot + out_coff + co0) * vol + ((size_t)i * D + j) * D + k;
#pragma unroll
    for (int c = 0; c < COT; c++) {
      if (co0 + c < Cout) {
        float v = acc[c] + bias[co0 + c];
        if (relu) v = fmaxf(v, 0.f);
        // dense-block backward: chain rule through the folded BatchNorm (x scale) and sum into the concat gradient
        if (out_scale) v *= out_scale[co0 + c];
        if (accumulate) v += ob[(size_t)c * vol];
        ob[(size_t)c * vol] = v;
      }
    }
  }
}

static Profiler* g_prof_tls();
static int launch_conv(const ConvF32& c, const float* in, int in_ctot, float* out, int out_ctot, int out_coff, int D,
                       int B, bool relu, cudaStream_t s, bool backward = false, const float* mask = nullptr,
                       bool bn_accumulate = false) {
  char nm[64];
  snprintf(nm, sizeof nm, backward ? "f32_dgrad%d_%dx%d_d%d" : "f32_conv%d_%dx%d_d%d", c.ks, c.cin, c.cout, D);
  ProfScope ps(g_prof_tls(), nm, s);
  const int tiles = (D + 7) / 8;
  // backward-data = the same convolution with the transposed/flipped weights: roles of cin and cout swap
  const int kin = backward ? c.cout : c.cin, kout = backward ? c.cin : c.cout;
  const float* kw = backward ? c.wT : c.w;
  const float* kb = backward ? c.zero_bias : c.bias;
  const int cot = (kout % 32 == 0) ? 32 : 16;
  dim3 g(tiles * tiles * tiles, (kout + cot - 1) / cot, B);
#define GB_LAUNCH(KS, COT)                                                                                     \
  conv3d_f32_kernel<KS, COT><<<g, 512, 0, s>>>(in, in_ctot, kin, kw, kb, backward ? nullptr : c.bn_scale,      \
                                               backward ? nullptr : c.bn_shift, out, out_ctot, out_coff, kout, D, \
                                               relu ? 1 : 0, mask, bn_accumulate ? c.bn_scale : nullptr,     \
                                               bn_accumulate ? 1 : 0)
  if (c.ks == 3 && cot == 32) GB_LAUNCH(3, 32);
  else if (c.ks == 3) GB_LAUNCH(3, 16);
  else if (c.ks == 1 && cot == 32) GB_LAUNCH(1, 32);
  else if (c.ks == 1) GB_LAUNCH(1, 16);
  else throw Error(GB_ERR_INTERNAL, "unsupported conv kernel size");
#undef GB_LAUNCH
  return 1;
}

// 2x2x2 stride-2 pooling (avg: default2018; max: default2017 / dense), channel-strided in/out
__global__ void pool2_f32_kernel(const float* __restrict__ in, int in_ctot, float* __restrict__ out, int out_ctot, int C,
                                 int Din, int is_max, size_t total) {
  const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int Do = Din / 2;
  size_t r = idx;
  const int k = r % Do; r /= Do;
  const int j = r % Do; r /= Do;
  const int i = r % Do; r /= Do;
  const int c = r % C;
  const int b = r / C;
  const float* p = in + (((size_t)b * in_ctot + c) * Din + 2 * i) * Din * Din + (size_t)(2 * j) * Din + 2 * k;
  float v = is_max ? -INFINITY : 0.f;
#pragma unroll
  for (int a = 0; a < 2; a++)
#pragma unroll
    for (int bb = 0; bb < 2; bb++)
#pragma unroll
      for (int cc = 0; cc < 2; cc++) {
        const float x = p[(size_t)a * Din * Din + bb * Din + cc];
        v = is_max ? fmaxf(v, x) : v + x;
      }
  if (!is_max) v *= 0.125f;
  out[(((size_t)b * out_ctot + c) * Do + i) * Do * Do + (size_t)j * Do + k] = v;
}

static int launch_pool(const float* in, int in_ctot, float* out, int out_ctot, int C, int Din, int B, bool is_max,
                       cudaStream_t s) {
  ProfScope ps(g_prof_tls(), "f32_pool", s);
  const int Do = Din / 2;
  const size_t total = (size_t)B * C * Do * Do * Do;
  pool2_f32_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(in, in_ctot, out, out_ctot, C, Din, is_max ? 1 : 0,
                                                                    total);
  return 1;
}

// global max over the D^3 volume: in [B][C][vol] -> feat [B][C]
__global__ void global_max_kernel(const float* __restrict__ in, float* __restrict__ feat, int vol) {
  const float* p = in + (size_t)blockIdx.x * vol;
  float v = -INFINITY;
  for (int e = threadIdx.x; e < vol; e += 32) v = fmaxf(v, p[e]);
  for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  if (threadIdx.x == 0) feat[blockIdx.x] = v;
}

// heads: feat [B][F] x fc_w [3][F] + fc_b -> out3 [B][3]
__global__ void __launch_bounds__(256) fc3_kernel(const float* __restrict__ feat, const float* __restrict__ w,
                                                  const float* __restrict__ bias, int F, float* __restrict__ out3) {
  __shared__ float red[3][8];
  const float* f = feat + (size_t)blockIdx.x * F;
  float a0 = 0, a1 = 0, a2 = 0;
  for (int e = threadIdx.x; e < F; e += 256) {
    const float x = f[e];
    a0 = fmaf(x, w[e], a0);
    a1 = fmaf(x, w[F + e], a1);
    a2 = fmaf(x, w[2 * (size_t)F + e], a2);
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

// torch_model.cpp:188-195.  The TorchScript module returns log_softmax(logits); softmax(log_softmax(z)) ==
// softmax(z), so pose = softmax(z)[1]; with skip_softmax the reference reads the module output [0,1] itself,
// i.e. log_softmax(z)[1].  loss = CE(module output, label 1) = -log_softmax(z)[1]; with apply_logistic_loss
// the reference takes -log(module output[0,1]) = -log(log_softmax(z)[1]).
// raw_output: out3 already holds the module's output (the overlay test model returns (0, score), not a log-softmax)
__global__ void head_post_kernel(const float* __restrict__ out3, int B, int skip_softmax, int logistic, int raw_output,
                                 float* __restrict__ pose, float* __restrict__ aff, float* __restrict__ loss) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float z0 = out3[3 * b], z1 = out3[3 * b + 1];
  const float m = fmaxf(z0, z1);
  const float lse = m + logf(expf(z0 - m) + expf(z1 - m));
  const float lp1 = raw_output ? z1 : z1 - lse;
  pose[b] = skip_softmax ? lp1 : expf(lp1);
  aff[b] = out3[3 * b + 2];
  loss[b] = logistic ? -logf(lp1) : -lp1;
}

void launch_head_post(const float* out3, int B, bool skip_softmax, bool logistic, float* pose, float* aff, float* loss,
                      cudaStream_t s, bool raw_output) {
  if (B <= 0) return;
  head_post_kernel<<<(B + 127) / 128, 128, 0, s>>>(out3, B, skip_softmax, logistic, raw_output, pose, aff, loss);
}

// CNNTorchScorer::score accumulation, cnn_torch_scorer.cpp:117-192: score accumulates in double, affinity/loss in
// float in model order; variance = population variance of the per-model affinities around the float mean.
__global__ void ensemble_kernel(const float* __restrict__ pose, const float* __restrict__ aff,
                                const float* __restrict__ loss, int M, int B, int stride, float* __restrict__ o_score,
                                float* __restrict__ o_aff, float* __restrict__ o_loss, float* __restrict__ o_var) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double sc = 0.0;
  float a = 0.f, l = 0.f;
  for (int m = 0; m < M; m++) {
    sc += (double)pose[(size_t)m * stride + b];
    a += aff[(size_t)m * stride + b];
    l += loss[(size_t)m * stride + b];
  }
  a /= (float)M;
  l /= (float)M;
  float var = 0.f;
  if (M > 1) {
    float sum = 0.f;
    for (int m = 0; m < M; m++) {
      float d = a - aff[(size_t)m * stride + b];
      sum += d * d;
    }
    var = sum / (float)M;
  }
  o_score[b] = (float)(sc / M);
  o_aff[b] = a;
  o_loss[b] = l;
  o_var[b] = var;
}

void launch_ensemble(const float* pose, const float* aff, const float* loss, int M, int B, int stride, float* o_score,
                     float* o_aff, float* o_loss, float* o_var, cudaStream_t s) {
  if (B <= 0) return;
  ensemble_kernel<<<(B + 127) / 128, 128, 0, s>>>(pose, aff, loss, M, B, stride, o_score, o_aff, o_loss, o_var);
}

static thread_local Profiler* t_prof = nullptr;
static Profiler* g_prof_tls() { return t_prof; }

// ---------------------------------------------------------------------------------------------------------
// GB_ARCH_OVERLAP: the graph of test/gnina/data/overlap.pt (read with torch.jit.load(...).code):
//   prot = x[:, 0] * x[:, 1];  ave = avg_pool3d(prot, 48).flatten(1);  ave = where(ave > 0, ave, 1e-20)
//   return (hstack([zeros, ave]), zeros)        -> metadata: skip_softmax, apply_logistic_loss: score = ave, loss = -log ave
// One block per pose; the backward kernel writes d loss / d grid = -(1 / ave) * d ave / d grid (zero where the
// `where` took the constant branch).
__global__ void __launch_bounds__(256) overlap_forward_kernel(const float* __restrict__ grid, int vol, float* __restrict__ out3) {
  __shared__ double red[8];
  const float* rec = grid + (size_t)blockIdx.x * 2 * vol;
  const float* lig = rec + vol;
  double acc = 0.0;
  for (int i = threadIdx.x; i < vol; i += 256) acc += (double)(rec[i] * lig[i]);
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; w++) t += red[w];
    const float ave = (float)(t / vol);
    out3[3 * blockIdx.x] = 0.f;
    out3[3 * blockIdx.x + 1] = ave > 0.f ? ave : 1e-20f;
    out3[3 * blockIdx.x + 2] = 0.f;
  }
}
__global__ void overlap_backward_kernel(const float* __restrict__ grid, int vol, const float* __restrict__ out3,
                                        float* __restrict__ dgrid) {
  const int p = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= vol) return;
  const float ave = out3[3 * p + 1];
  const float coef = ave > 1e-20f ? -1.f / (ave * (float)vol) : 0.f;   // d(-log ave)/d ave * d ave / d prot
  const float* rec = grid + (size_t)p * 2 * vol;
  float* d = dgrid + (size_t)p * 2 * vol;
  d[i] = coef * rec[vol + i];
  d[vol + i] = coef * rec[i];
}
static int overlap_forward(const Model& m, const float* grid, int B, float* out3, cudaStream_t s) {
  GB_CHECK(m.n_channels == 2, "overlap model: 2 channels");
  const int vol = m.npts * m.npts * m.npts;
  overlap_forward_kernel<<<B, 256, 0, s>>>(grid, vol, out3);
  return 1;
}

int forward_fp32(const Model& m, const float* grid, int B, Fp32Workspace& ws, float* out3, cudaStream_t s,
                 Profiler* prof) {
  int launches = 0;
  const int C = m.n_channels;
  t_prof = prof;
  if (m.arch == GB_ARCH_OVERLAP) return overlap_forward(m, grid, B, out3, s);
  GB_CHECK(m.npts == 48, "CNN graphs expect a 48^3 grid");
  auto conv = [&](const std::string& k) -> const ConvF32& {
    auto it = m.convs.find(k);
    if (it == m.convs.end()) throw Error(GB_ERR_INTERNAL, "missing conv " + k);
    return it->second;
  };
  const size_t v24 = 24 * 24 * 24, v12 = 12 * 12 * 12, v6 = 6 * 6 * 6;
  const float* feat = nullptr;
  if (m.arch == GB_ARCH_DEFAULT2018) {
    ws.ensure(0, (size_t)B * 32 * v24);
    ws.ensure(1, (size_t)B * 32 * v24);
    launches += launch_pool(grid, C, ws.buf[0], C, C, 48, B, false, s);                       // [B][28][24^3]
    launches += launch_conv(conv("unit1_conv"), ws.buf[0], C, ws.buf[1], 32, 0, 24, B, true, s);
    launches += launch_conv(conv("unit2_conv"), ws.buf[1], 32, ws.buf[0], 32, 0, 24, B, true, s);
    launches += launch_pool(ws.buf[0], 32, ws.buf[1], 32, 32, 24, B, false, s);               // [B][32][12^3]
    launches += launch_conv(conv("unit3_conv"), ws.buf[1], 32, ws.buf[0], 64, 0, 12, B, true, s);
    launches += launch_conv(conv("unit4_conv"), ws.buf[0], 64, ws.buf[1], 64, 0, 12, B, true, s);
    launches += launch_pool(ws.buf[1], 64, ws.buf[0], 64, 64, 12, B, false, s);               // [B][64][6^3]
    launches += launch_conv(conv("unit5_conv"), ws.buf[0], 64, ws.buf[1], 128, 0, 6, B, true, s);
    feat = ws.buf[1];  // NCDHW flatten == view(-1, 27648)
  } else if (m.arch == GB_ARCH_DEFAULT2017) {
    ws.ensure(0, (size_t)B * 35 * v24);
    ws.ensure(1, (size_t)B * 32 * v24);
    launches += launch_pool(grid, C, ws.buf[0], C, C, 48, B, true, s);
    launches += launch_conv(conv("unit1_conv1"), ws.buf[0], C, ws.buf[1], 32, 0, 24, B, true, s);
    launches += launch_pool(ws.buf[1], 32, ws.buf[0], 32, 32, 24, B, true, s);
    launches += launch_conv(conv("unit2_conv1"), ws.buf[0], 32, ws.buf[1], 64, 0, 12, B, true, s);
    launches += launch_pool(ws.buf[1], 64, ws.buf[0], 64, 64, 12, B, true, s);
    launches += launch_conv(conv("unit3_conv1"), ws.buf[0], 64, ws.buf[1], 128, 0, 6, B, true, s);
    feat = ws.buf[1];
  } else {
    // dense: block buffers hold the running concatenation; each BN->conv3->ReLU layer appends 16 channels
    ws.ensure(0, (size_t)B * 28 * v24);
    ws.ensure(1, (size_t)B * 96 * v24);
    ws.ensure(2, (size_t)B * 96 * v24);
    launches += launch_pool(grid, C, ws.buf[0], C, C, 48, B, true, s);
    launches += launch_conv(conv("data_enc_init_conv"), ws.buf[0], C, ws.buf[1], 96, 0, 24, B, true, s);
    auto block = [&](int L, float* buf, int c0, int D) {
      for (int i = 0; i < 4; i++) {
        const ConvF32& c = conv("dense_block_" + std::to_string(L) + ".data_enc_level" + std::to_string(L) + "_conv" +
                                std::to_string(i));
        launches += launch_conv(c, buf, c0 + 64, buf, c0 + 64, c0 + 16 * i, D, B, true, s);
      }
    };
    block(0, ws.buf[1], 32, 24);                                                                // [B][96][24^3]
    launches += launch_conv(conv("data_enc_level0_bottleneck"), ws.buf[1], 96, ws.buf[2], 96, 0, 24, B, true, s);
    // pool -> first 96 channels of the level-1 block buffer [B][160][12^3]
    launches += launch_pool(ws.buf[2], 96, ws.buf[1], 160, 96, 24, B, true, s);
    block(1, ws.buf[1], 96, 12);
    launches += launch_conv(conv("data_enc_level1_bottleneck"), ws.buf[1], 160, ws.buf[2], 160, 0, 12, B, true, s);
    launches += launch_pool(ws.buf[2], 160, ws.buf[1], 224, 160, 12, B, true, s);              // [B][224][6^3]
    block(2, ws.buf[1], 160, 6);
    ws.ensure_feat((size_t)B * 224);
    global_max_kernel<<<B * 224, 32, 0, s>>>(ws.buf[1], ws.feat, (int)v6);
    launches++;
    feat = ws.feat;
  }
  (void)v12;
  {
    ProfScope ps(prof, "f32_fc_heads", s);
    fc3_kernel<<<B, 256, 0, s>>>(feat, m.fc_w, m.fc_b, m.fc_features, out3);
  }
  launches++;
  t_prof = nullptr;
  return launches;
}

// ---------------------------------------------------------------------------------------------------------
// Backward pieces (N5 of SURVEY.md §8a): avg-pool backward (each fine voxel receives 1/8 of its pooled gradient),
// heads backward for loss = CE(logits, label 1): dL/dz0 = softmax(z)[0], dL/dz1 = -softmax(z)[0], affinity head
// does not enter the loss (torch_model.cpp:195).
__global__ void unpool2_f32_kernel(const float* __restrict__ dy, float* __restrict__ dx, int C, int Dout, size_t total) {
  const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int Din = Dout / 2;
  size_t r = idx;
  const int k = r % Dout; r /= Dout;
  const int j = r % Dout; r /= Dout;
  const int i = r % Dout; r /= Dout;  // r = b*C + c
  dx[idx] = 0.125f * dy[((r * Din + (i >> 1)) * Din + (j >> 1)) * Din + (k >> 1)];
}

// max-pool backward: the gradient of a pooled voxel goes to the FIRST maximum of its 2x2x2 window in (x, y, z) scan
// order (ATen max_pool3d keeps the first index whose value is strictly greater); channel-strided like pool2.
__global__ void unpool2_max_f32_kernel(const float* __restrict__ dy, int dy_ctot, const float* __restrict__ xin, int x_ctot,
                                       float* __restrict__ dx, int dx_ctot, int C, int Din, size_t total) {
  const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int Do = Din / 2;
  size_t r = idx;
  const int k = r % Do; r /= Do;
  const int j = r % Do; r /= Do;
  const int i = r % Do; r /= Do;
  const int c = r % C;
  const int b = r / C;
  const size_t off = (size_t)(2 * i) * Din * Din + (size_t)(2 * j) * Din + 2 * k;
  const float* p = xin + ((size_t)b * x_ctot + c) * Din * Din * Din + off;
  float* q = dx + ((size_t)b * dx_ctot + c) * Din * Din * Din + off;
  float best = -INFINITY;
  int arg = 0;
#pragma unroll
  for (int e = 0; e < 8; e++) {
    const float x = p[(size_t)(e >> 2) * Din * Din + ((e >> 1) & 1) * Din + (e & 1)];
    if (x > best) { best = x; arg = e; }
  }
  const float g = dy[(((size_t)b * dy_ctot + c) * Do + i) * Do * Do + (size_t)j * Do + k];
#pragma unroll
  for (int e = 0; e < 8; e++) q[(size_t)(e >> 2) * Din * Din + ((e >> 1) & 1) * Din + (e & 1)] = e == arg ? g : 0.f;
}

// global max backward + FC backward for the dense family: d feat[c] = p0 (w0[c] - w1[c]) lands on the first maximum
// of channel c's volume, zero elsewhere.  One warp per (pose, channel).
__global__ void dense_head_backward_kernel(const float* __restrict__ out3, const float* __restrict__ w, int F,
                                           const float* __restrict__ fin, float* __restrict__ dfin, int vol) {
  const int b = blockIdx.x / F, c = blockIdx.x % F;
  const float* p = fin + (size_t)blockIdx.x * vol;
  float best = -INFINITY;
  int arg = 0x7fffffff;
  for (int e = threadIdx.x; e < vol; e += 32) {
    const float x = p[e];
    if (x > best) { best = x; arg = e; }
  }
  for (int o = 16; o; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
    if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
  }
  const float z0 = out3[3 * b], z1 = out3[3 * b + 1];
  const float m = fmaxf(z0, z1);
  const float e0 = expf(z0 - m), e1 = expf(z1 - m);
  const float g = e0 / (e0 + e1) * (w[c] - w[F + c]);
  float* q = dfin + (size_t)blockIdx.x * vol;
  for (int e = threadIdx.x; e < vol; e += 32) q[e] = e == arg ? g : 0.f;
}

__global__ void fc3_backward_kernel(const float* __restrict__ out3, const float* __restrict__ w, int F,
                                    float* __restrict__ dfeat) {
  const int b = blockIdx.y;
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  const float z0 = out3[3 * b], z1 = out3[3 * b + 1];
  const float m = fmaxf(z0, z1);
  const float e0 = expf(z0 - m), e1 = expf(z1 - m);
  const float p0 = e0 / (e0 + e1);
  dfeat[(size_t)b * F + f] = p0 * w[f] - p0 * w[F + f];
}

void Fp32GradWorkspace::ensure(int i, size_t n) {
  if (cap[i] >= n) return;
  if (a[i]) cudaFree(a[i]);
  GB_CUDA(cudaMalloc(&a[i], n * sizeof(float)));
  cap[i] = n;
}
Fp32GradWorkspace::~Fp32GradWorkspace() {
  for (auto p : a)
    if (p) cudaFree(p);
}

static int forward_backward_dense_fp32(const Model& m, const float* grid, int B, Fp32GradWorkspace& ws, float* out3,
                                       float* dgrid, cudaStream_t s);

int forward_backward_fp32(const Model& m, const float* grid, int B, Fp32GradWorkspace& ws, float* out3, float* dgrid,
                          cudaStream_t s, Profiler* prof) {
  if (m.arch == GB_ARCH_OVERLAP) {
    const int vol = m.npts * m.npts * m.npts;
    overlap_forward(m, grid, B, out3, s);
    overlap_backward_kernel<<<dim3((vol + 255) / 256, B), 256, 0, s>>>(grid, vol, out3, dgrid);
    return 2;
  }
  if (m.arch != GB_ARCH_DEFAULT2018 && m.arch != GB_ARCH_DENSE)
    throw Error(GB_ERR_USAGE, "gradient path is implemented for the default2018 and dense families only (model " + m.name + ")");
  GB_CHECK(m.npts == 48, "CNN graphs expect a 48^3 grid");
  int launches = 0;
  t_prof = prof;
  if (m.arch == GB_ARCH_DENSE) {
    launches = forward_backward_dense_fp32(m, grid, B, ws, out3, dgrid, s);
    t_prof = nullptr;
    return launches;
  }
  const int C = m.n_channels;
  auto conv = [&](const std::string& k) -> const ConvF32& { return m.convs.at(k); };
  const size_t v24 = 24 * 24 * 24, v12 = 12 * 12 * 12, v6 = 6 * 6 * 6;
  const size_t n[8] = {(size_t)B * C * v24,  (size_t)B * 32 * v24, (size_t)B * 32 * v24, (size_t)B * 32 * v12,
                       (size_t)B * 64 * v12, (size_t)B * 64 * v12, (size_t)B * 64 * v6,  (size_t)B * 128 * v6};
  for (int i = 0; i < 8; i++) ws.ensure(i, n[i]);
  ws.ensure(8, (size_t)B * 32 * v24);
  ws.ensure(9, (size_t)B * 32 * v24);
  float *x0 = ws.a[0], *y1 = ws.a[1], *y2 = ws.a[2], *x2 = ws.a[3], *y3 = ws.a[4], *y4 = ws.a[5], *x4 = ws.a[6],
        *y5 = ws.a[7], *ga = ws.a[8], *gb2 = ws.a[9];
  // forward, every activation kept
  launches += launch_pool(grid, C, x0, C, C, 48, B, false, s);
  launches += launch_conv(conv("unit1_conv"), x0, C, y1, 32, 0, 24, B, true, s);
  launches += launch_conv(conv("unit2_conv"), y1, 32, y2, 32, 0, 24, B, true, s);
  launches += launch_pool(y2, 32, x2, 32, 32, 24, B, false, s);
  launches += launch_conv(conv("unit3_conv"), x2, 32, y3, 64, 0, 12, B, true, s);
  launches += launch_conv(conv("unit4_conv"), y3, 64, y4, 64, 0, 12, B, true, s);
  launches += launch_pool(y4, 64, x4, 64, 64, 12, B, false, s);
  launches += launch_conv(conv("unit5_conv"), x4, 64, y5, 128, 0, 6, B, true, s);
  fc3_kernel<<<B, 256, 0, s>>>(y5, m.fc_w, m.fc_b, m.fc_features, out3);
  // backward
  const int F = m.fc_features;
  fc3_backward_kernel<<<dim3((F + 255) / 256, B), 256, 0, s>>>(out3, m.fc_w, F, ga);              // d y5 (pre-mask)
  launches += 2;
  launches += launch_conv(conv("unit5_conv"), ga, 128, gb2, 64, 0, 6, B, false, s, true, y5);       // d x4
  {
    const size_t tot = (size_t)B * 64 * v12;
    unpool2_f32_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(gb2, ga, 64, 12, tot);         // d y4
  }
  launches += launch_conv(conv("unit4_conv"), ga, 64, gb2, 64, 0, 12, B, false, s, true, y4);       // d y3
  launches += launch_conv(conv("unit3_conv"), gb2, 64, ga, 32, 0, 12, B, false, s, true, y3);       // d x2
  {
    const size_t tot = (size_t)B * 32 * v24;
    unpool2_f32_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(ga, gb2, 32, 24, tot);         // d y2
  }
  launches += launch_conv(conv("unit2_conv"), gb2, 32, ga, 32, 0, 24, B, false, s, true, y2);       // d y1
  launches += launch_conv(conv("unit1_conv"), ga, 32, gb2, C, 0, 24, B, false, s, true, y1);        // d x0
  {
    const size_t tot = (size_t)B * C * 48 * 48 * 48;
    unpool2_f32_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(gb2, dgrid, C, 48, tot);       // d grid
  }
  launches += 3;
  t_prof = nullptr;
  return launches;
}

// Dense family (N2) forward keeping every activation, then the backward pass (N5):
//   head -> global max -> block 2 -> max-pool -> bottleneck 1 -> block 1 -> max-pool -> bottleneck 0 -> block 0 ->
//   init conv -> max-pool -> grid.
// A dense-block layer j reads channels [0, cj) of the running concatenation F through BatchNorm (folded to
// scale/shift on the load) and appends 16 channels.  Backward, layers in reverse order:
//   g_j = dF[cj : cj+16] * [F[cj : cj+16] > 0];   dF[0 : cj] += bn_scale_j * conv^T(g_j)
// (the convolution kernel's epilogue applies bn_scale and accumulates in place).
static int forward_backward_dense_fp32(const Model& m, const float* grid, int B, Fp32GradWorkspace& ws, float* out3,
                                       float* dgrid, cudaStream_t s) {
  int launches = 0;
  const int C = m.n_channels;
  auto conv = [&](const std::string& k) -> const ConvF32& { return m.convs.at(k); };
  const size_t v24 = 24 * 24 * 24, v12 = 12 * 12 * 12, v6 = 6 * 6 * 6;
  // activations: 0 P0 [C@24], 1 F0 [96@24], 2 N0 [96@24], 3 F1 [160@12], 4 N1 [160@12], 5 F2 [224@6], 6 feat [224]
  // gradients : 7 dF0 / dP0 share nothing: 7 dF2 [224@6], 8 dN1 then dF1 ping-pong, 9 dN0/dF0 ping-pong
  ws.ensure(0, (size_t)B * C * v24);
  ws.ensure(1, (size_t)B * 96 * v24);
  ws.ensure(2, (size_t)B * 96 * v24);
  ws.ensure(3, (size_t)B * 160 * v12);
  ws.ensure(4, (size_t)B * 160 * v12);
  ws.ensure(5, (size_t)B * 224 * v6);
  ws.ensure(6, (size_t)B * 224);
  ws.ensure(7, (size_t)B * 224 * v6);
  ws.ensure(8, (size_t)B * 96 * v24);   // also holds [160@12] (smaller)
  ws.ensure(9, (size_t)B * 96 * v24);
  float *P0 = ws.a[0], *F0 = ws.a[1], *N0 = ws.a[2], *F1 = ws.a[3], *N1 = ws.a[4], *F2 = ws.a[5], *feat = ws.a[6];
  float *dF2 = ws.a[7], *ga = ws.a[8], *gb2 = ws.a[9];
  auto layer = [&](int L, int i) -> const ConvF32& {
    return conv("dense_block_" + std::to_string(L) + ".data_enc_level" + std::to_string(L) + "_conv" + std::to_string(i));
  };
  // ---- forward ----
  launches += launch_pool(grid, C, P0, C, C, 48, B, true, s);
  launches += launch_conv(conv("data_enc_init_conv"), P0, C, F0, 96, 0, 24, B, true, s);
  auto block = [&](int L, float* buf, int c0, int D) {
    for (int i = 0; i < 4; i++) launches += launch_conv(layer(L, i), buf, c0 + 64, buf, c0 + 64, c0 + 16 * i, D, B, true, s);
  };
  block(0, F0, 32, 24);
  launches += launch_conv(conv("data_enc_level0_bottleneck"), F0, 96, N0, 96, 0, 24, B, true, s);
  launches += launch_pool(N0, 96, F1, 160, 96, 24, B, true, s);
  block(1, F1, 96, 12);
  launches += launch_conv(conv("data_enc_level1_bottleneck"), F1, 160, N1, 160, 0, 12, B, true, s);
  launches += launch_pool(N1, 160, F2, 224, 160, 12, B, true, s);
  block(2, F2, 160, 6);
  global_max_kernel<<<B * 224, 32, 0, s>>>(F2, feat, (int)v6);
  fc3_kernel<<<B, 256, 0, s>>>(feat, m.fc_w, m.fc_b, m.fc_features, out3);
  launches += 2;
  // ---- backward ----
  GB_CHECK(m.fc_features == 224, "dense head features");
  dense_head_backward_kernel<<<B * 224, 32, 0, s>>>(out3, m.fc_w, 224, F2, dF2, (int)v6);
  launches++;
  // block backward on a [ctot = c0 + 64] buffer pair (F activations, dF gradients); vol = D^3
  auto block_backward = [&](int L, const float* F, float* dF, int c0, int D) {
    const size_t vol = (size_t)D * D * D;
    const int ctot = c0 + 64;
    for (int i = 3; i >= 0; i--) {
      const int cj = c0 + 16 * i;
      // input = dF channels [cj, cj+16), masked by F's same channels; output accumulates into dF channels [0, cj)
      launches += launch_conv(layer(L, i), dF + (size_t)cj * vol, ctot, dF, ctot, 0, D, B, false, s, true,
                              F + (size_t)cj * vol, true);
    }
  };
  block_backward(2, F2, dF2, 160, 6);
  // max-pool 12 -> 6 fed channels [0,160) of F2 from N1
  {
    const size_t tot = (size_t)B * 160 * v6;
    unpool2_max_f32_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(dF2, 224, N1, 160, ga, 160, 160, 12, tot);  // dN1
  }
  launches += launch_conv(conv("data_enc_level1_bottleneck"), ga, 160, gb2, 160, 0, 12, B, false, s, true, N1);       // dF1
  block_backward(1, F1, gb2, 96, 12);
  {
    const size_t tot = (size_t)B * 96 * v12;
    unpool2_max_f32_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(gb2, 160, N0, 96, ga, 96, 96, 24, tot);      // dN0
  }
  launches += launch_conv(conv("data_enc_level0_bottleneck"), ga, 96, gb2, 96, 0, 24, B, false, s, true, N0);         // dF0
  block_backward(0, F0, gb2, 32, 24);
  // init conv: gradient of its 32 output channels (first channels of F0), masked by its ReLU
  launches += launch_conv(conv("data_enc_init_conv"), gb2, 96, ga, C, 0, 24, B, false, s, true, F0);                  // dP0
  {
    const size_t tot = (size_t)B * C * v24;
    unpool2_max_f32_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(ga, C, grid, C, dgrid, C, C, 48, tot);
  }
  launches += 3;
  return launches;
}

}  // namespace gb
