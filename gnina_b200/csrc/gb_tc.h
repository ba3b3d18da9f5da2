// Fast path: fused voxelise+pool (fp16, channel-chunk-planar) and tcgen05 fp16 convolutions.  See gb_cnn_tc.cu.
#pragma once
#include "gb_internal.h"
#include <cuda_fp16.h>
#include <functional>
#include <mutex>

namespace gb {

struct TcPoseBatch {
  const float4* rec_xyzr; const int* rec_ch; int n_rec;
  const float4* lig_xyzr; const int* lig_ch; const int* lig_off;  // lig_off already offset to the chunk's first pose
  const float* centers;                                            // idem
  int n_poses, max_pose_atoms, n_channels, n_rec_channels;
  float resolution, dimension;
  const float* rot = nullptr;  // [n_poses][9] rotation about the grid centre (G3), or null; offset to the chunk like centers
};
constexpr int kFusedGroup = 8;  // poses per row group of the fused conv1 kernel's input layout (gb_cnn_tc_fused.cu)

// Pooled input grids, double buffered so that the (CUDA-core) voxeliser of chunk i+1 can run on an auxiliary
// stream while the (tensor-core) network of chunk i runs on the main stream; shared by the models of a grid group.
struct TcGridWorkspace {
  // [kind][double buffer]; kind 0 = average pool, one pose per group (gradient path, gb_cnn_tc_grad.cu), 1 = max pool
  // (dense family), 2 = average pool in row groups of kFusedGroup poses (scoring path, gb_cnn_tc_fused.cu).  The layouts
  // differ in where their zero borders are, so every kind owns its buffers.
  // kind 3 = max pool with 48 (35 used) channels: default2017
  void* x0[4][2] = {};
  size_t cap[4][2] = {};
  cudaEvent_t ready[2] = {nullptr, nullptr}, consumed[2] = {nullptr, nullptr}, started[2] = {nullptr, nullptr};
  bool started_valid = false;
  bool consumed_valid[2] = {false, false};
  unsigned iter = 0;
  float4* list_xyzr = nullptr; int* list_ch = nullptr; int* list_n = nullptr; size_t list_cap = 0, listn_cap = 0;
  void receptor_changed() {}
  void batch_done() {}
  ~TcGridWorkspace();
};

struct TcWorkspace {
  void* buf[16] = {};
  size_t cap[16] = {};
  void ensure(int i, size_t bytes);
  ~TcWorkspace();
};

// one-time initialisations (lazily packed weights, function attributes, constant tables) happen under this lock:
// clones of a scorer run on different host threads and share the models
std::mutex& tc_init_mutex();
bool tc_supported(const Model& m);
// pose lists + fused voxelise/pool of one chunk into gw.x0[kind][buf] for every pool kind in kinds_mask (bit 0 = avg
// for the default2018 family, bit 1 = max for the dense family) on stream s; returns #launches
int tc_prepare_grid(const TcPoseBatch& pb, TcGridWorkspace& gw, int buf, int kinds_mask, cudaStream_t s,
                    Profiler* prof = nullptr);
int tc_pool_kind(const Model& m);  // 0 avg, 1 max
// buffer kind tc_forward expects for model m: tc_pool_kind, or 2 when the fused scoring kernel will run
int tc_grid_kind(const Model& m, bool keep_activations);
bool tc_fused_enabled();
// network forward on the pooled grid x0 -> out3 [n_poses][3]; records x0_consumed (if non-null) once x0 has been
// read for the last time; returns the number of kernel launches
int tc_forward(const Model& m, const TcPoseBatch& pb, const void* x0, TcWorkspace& ws, float* out3, cudaStream_t s,
               Profiler* prof = nullptr, cudaEvent_t x0_consumed = nullptr, bool keep_activations = false);
// default2018 family, after tc_forward(..., keep_activations = true) on the same workspace: backward of the CE loss
// through the network and the fused voxelise/pool to the ligand atoms of the chunk; atom_grad [chunk atoms][3] (see
// gb_cnn_tc_grad.cu); returns #launches
// rec_off / rec_grad (optional, single-pose chunks): the typed receptor atoms {0, n_rec} as a second atom set of pose 0
int tc_backward(const Model& m, const TcPoseBatch& pb, TcWorkspace& ws, const float* out3, float* atom_grad, cudaStream_t s,
                Profiler* prof = nullptr, const int* rec_off = nullptr, float* rec_grad = nullptr);

// chunk-planar padded grouped activation layout (see gb_cnn_tc.cu)
struct ActLayout {
  int D, P, G, T, C8, Lp;
  size_t group_u4() const { return (size_t)D * C8 * Lp; }  // uint4 per group
};
ActLayout make_layout(int D, int G, int C);
size_t act_bytes(const ActLayout& L, int n_poses);
// dense family (gb_cnn_tc_dense.cu)
struct TcDenseWeights;
int tc_forward_dense(const Model& m, const TcPoseBatch& pb, const void* x0max, TcWorkspace& ws, float* out3, cudaStream_t s,
                     Profiler* prof, cudaEvent_t x0_consumed);
// shared launcher: 3x3x3 conv 32->32 @24^3 writing chunks [c8_off, c8_off+4) of a chunk-planar block buffer
struct ConvTc {
  int cin = 0, cout = 0;     // cin padded to a multiple of 16
  uint4* wp = nullptr;       // [cout/32][9][cin/8][96] x 16 B
  float* bias = nullptr;
};
struct PointwiseTc {
  int c = 0;
  __half* w = nullptr;       // [co][ci] row-major fp16
  float* bias = nullptr;
};
struct TcWeights {
  ConvTc conv1, conv3, conv5;
  PointwiseTc pw2, pw4;   // (unused by default2017: no 1x1x1 convolutions)
  uint4* pw2_packed = nullptr;  // unit2_conv as the K-major B operand of the fused kernel's second tcgen05.mma
  uint4* pair_w = nullptr;      // CTA-pair variant of the fused kernel: per-rank B rows of unit1_conv / unit2_conv (pack_pair_weights)
  uint4* pair_w2 = nullptr;
  float* fcw = nullptr;      // [3][216*128] channels-last order
  float* fcb = nullptr;
  std::vector<void*> allocs;
  ~TcWeights();
};
std::shared_ptr<TcWeights> get_tc_weights(const Model& m);
ConvTc make_conv_tc(std::vector<void*>& allocs, int cout, int cin, const std::function<float(int, int, int, int, int)>& wfn,
                    const float* bias);
// any instantiated (cin, D): out_planar == nullptr -> channels-last `out`, else chunk-planar (same D, P, G as the input)
void launch_conv_tc_any(int cin, int D, const ConvTc& c, const ActLayout& L, const uint4* xin, __half* out, int n_poses,
                        cudaStream_t s, uint4* out_planar, int out_c8tot, int out_c8off, int out_lp, int relu);
void launch_conv_tc_32_24_planar(const ConvTc& c, const uint4* xin, uint4* xout, int out_c8tot, int out_c8off, int out_lp,
                                 int n_poses, cudaStream_t s);
void tc_debug_set(int i, const void* p, size_t bytes);
// fused scoring kernel (gb_cnn_tc_fused.cu): x0 in the row-group layout -> X2 (input of unit3_conv)
ActLayout make_fused_x0_layout();
uint4* pack_pointwise_tc(std::vector<void*>& allocs, const float* w, int c);
void pack_pair_weights(std::vector<void*>& allocs, const uint4* wp, const uint4* w2p, uint4** pair_w, uint4** pair_w2);
void launch_conv1_pw2_pool(const ConvTc& conv1, const __half* w2, const uint4* w2p, const float* bias2, const uint4* pair_w, const uint4* pair_w2,
                           const uint4* x0, const ActLayout& L0, uint4* x2, const ActLayout& L2, int n_poses, cudaStream_t s);

// test-only access to the buffers of the most recent tc_forward on this thread: 0 x0, 1 y(3), 2 x2, 3 x4, 4 y5
const void* tc_debug_buffer(int i, size_t* bytes);

}  // namespace gb
