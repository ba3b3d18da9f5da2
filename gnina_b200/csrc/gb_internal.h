// Internal declarations shared by the translation units of libgnina_b200.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <atomic>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>
#include "../../include/gnina_b200.h"

namespace gb {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
void set_last_error(const std::string& m);
// Before the first CUDA call of the process: ask for 32 hardware work queues instead of the default 8, unless the
// user set CUDA_DEVICE_MAX_CONNECTIONS.  Every handle owns a stream; with 8 queues, streams of different host threads
// share a queue and a copy queued behind another handle's pending D2H waits for THAT handle's long kernel (measured:
// set_ligand 0.1 ms -> 200 ms with 16 docking workers).  No effect if CUDA is already initialised.
void prefer_many_hw_queues();

#define GB_CUDA(expr)                                                                                  \
  do {                                                                                                 \
    cudaError_t _e = (expr);                                                                           \
    if (_e != cudaSuccess)                                                                             \
      throw gb::Error(GB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" __FILE__ ")"); \
  } while (0)
#define GB_CHECK(cond, msg)                                                             \
  do {                                                                                  \
    if (!(cond)) throw gb::Error(GB_ERR_INTERNAL, std::string("check failed: ") + msg); \
  } while (0)

constexpr int kNumSminaTypes = 28;
extern const char* const kSminaNames[kNumSminaTypes];
extern const float kSminaXsRadius[kNumSminaTypes];

// FileMappedGninaTyper equivalent: one channel per non-empty line of the map text.
struct TypeMap {
  int n_channels = 0;
  int t2c[kNumSminaTypes];
  void parse(const std::string& text);
};

struct HostTensor {
  std::vector<int> shape;
  const float* data = nullptr;  // into Model::raw
  size_t nelem = 0;
};

// Device-side description of one convolution of the fp32 validation path.
struct ConvF32 {
  int cin = 0, cout = 0, ks = 1;
  float* w = nullptr;     // [cin][ks^3][cout]
  float* bias = nullptr;  // [cout]
  float* bn_scale = nullptr;  // [cin] or null : y = x*scale + shift applied to in-bounds inputs
  float* bn_shift = nullptr;
  float* wT = nullptr;        // backward-data weights [cout][ks^3][cin]: taps flipped, channels swapped
  float* zero_bias = nullptr; // [cin] zeros (bias of the backward-data convolution)
};

// fp16 tensor-core path weights (see gb_cnn_tc.cu, gb_cnn_tc_dense.cu)
struct TcWeights;
struct TcDenseWeights;
struct TcGradWeights;

struct Model {
  std::atomic<int> refs{1};
  int device = 0;
  int arch = 0;
  float resolution = 0.5f, dimension = 23.5f, radius_scaling = 1.f;
  bool apply_logistic_loss = false, skip_softmax = false;
  std::string name, recmap, ligmap;
  TypeMap rec, lig;
  int n_channels = 0, npts = 48;
  std::vector<char> raw;
  std::map<std::string, HostTensor> tensors;
  // device (fp32 path)
  std::vector<float*> dev_allocs;
  std::map<std::string, ConvF32> convs;
  float* fc_w = nullptr;  // [3][F] rows: pose0, pose1, affinity
  float* fc_b = nullptr;  // [3]
  int fc_features = 0;
  std::shared_ptr<TcWeights> tc;  // lazily built
  std::shared_ptr<TcDenseWeights> tc_dense;
  std::shared_ptr<TcGradWeights> tc_grad;  // backward-data weights of the fast gradient path
  const HostTensor& t(const std::string& n) const;
  ~Model();
};
Model* load_model_from_memory(const void* data, size_t n, int device, const std::string& label);

// ---------------------------------------------------------------------------------------------
// Atoms on the device.  Receptor: filtered to typed atoms, stably sorted by channel.
struct DevAtoms {
  float4* xyzr = nullptr;  // x,y,z,radius*radius_scale
  int* channel = nullptr;
  int n = 0;
};

// Signature of a voxelisation setup: models with equal signatures share typed atoms and grids.
struct GridSig {
  std::string recmap, ligmap;
  float resolution, dimension, radius_scaling;
  bool operator==(const GridSig& o) const {
    return recmap == o.recmap && ligmap == o.ligmap && resolution == o.resolution && dimension == o.dimension &&
           radius_scaling == o.radius_scaling;
  }
};

// Optional per-kernel CUDA-event timing (bench.py reads it through gb_cnn_profile_read).
struct Profiler {
  bool on = false;
  struct Pending { int id; cudaEvent_t a, b; };
  std::vector<std::string> names;
  std::vector<double> total_ms;
  std::vector<long long> count;
  std::vector<Pending> pending;
  std::vector<cudaEvent_t> pool;
  int cur = -1;
  cudaEvent_t cur_a = nullptr;
  void begin(const char* name, cudaStream_t s);
  void end(cudaStream_t s);
  void resolve();  // call after the stream is synchronised
  void reset();
  ~Profiler();
};
struct ProfScope {
  Profiler* p; cudaStream_t s;
  ProfScope(Profiler* p_, const char* n, cudaStream_t s_) : p(p_), s(s_) { if (p && p->on) p->begin(n, s); }
  ~ProfScope() { if (p && p->on) p->end(s); }
};

// kernels: gb_grid.cu
void launch_build_pose_lists(const float4* rec_xyzr, const int* rec_ch, int n_rec, const float4* lig_xyzr,
                             const int* lig_ch, const int* lig_off, const float* centers, int n_poses, float half_dim,
                             int cap, float4* list_xyzr, int* list_ch, int* list_n, cudaStream_t s, const float* rot = nullptr);
void launch_voxelize_f32(const float4* list_xyzr, const int* list_ch, const int* list_n, int cap, const float* centers,
                         int n_poses, int n_channels, int npts, float resolution, float dimension, float* grid,
                         cudaStream_t s);

// kernels: gb_cnn_fp32.cu
struct Fp32Workspace {
  float* buf[4] = {nullptr, nullptr, nullptr, nullptr};
  size_t cap[4] = {0, 0, 0, 0};
  float* feat = nullptr; size_t feat_cap = 0;
  void ensure(int i, size_t nfloats);
  void ensure_feat(size_t nfloats);
  ~Fp32Workspace();
};
// grid [B][C][48^3] fp32 -> out3 [B][3] (pose logit0, logit1, affinity); returns #kernel launches
int forward_fp32(const Model& m, const float* grid, int B, Fp32Workspace& ws, float* out3, cudaStream_t s,
                 Profiler* prof = nullptr);
// Forward of the default2018 family keeping every activation, then backward of loss = CE(logits, label 1)
// (torch_model.cpp:195-199) down to dLoss/dGrid [B][C][48^3].  Returns #kernel launches.
struct Fp32GradWorkspace {
  float* a[10] = {};   // activations: x0,y1,y2,x2,y3,y4,x4,y5 ; gradient ping-pong: a[8], a[9]
  size_t cap[10] = {};
  void ensure(int i, size_t nfloats);
  ~Fp32GradWorkspace();
};
int forward_backward_fp32(const Model& m, const float* grid, int B, Fp32GradWorkspace& ws, float* out3, float* dgrid,
                          cudaStream_t s, Profiler* prof = nullptr);
// GridMaker::backward (torch_model.cpp:203): atom gradients from dLoss/dGrid [n_poses][C][N^3] for typed atoms
// stored pose after pose (pose p owns atoms [pose_off[p], pose_off[p+1])); dgrid/centers/pose_off are those of the
// chunk; out: atom_grad[atom][3] indexed like the atom arrays.
void launch_grid_backward(const float4* atoms_xyzr, const int* atoms_ch, const int* pose_off, int max_pose_atoms,
                          const float* centers, int n_poses, int n_channels, int npts, float resolution, float dimension,
                          const float* dgrid, float* atom_grad, cudaStream_t s, const float* rot = nullptr);
// dst[i] += alpha * src[i] for i in [lo, hi)
void launch_axpy_range(const float* src, float* dst, int lo, int hi, float alpha, cudaStream_t s);
// [B][3] raw -> pose/aff/loss per torch_model.cpp:188-195
void launch_head_post(const float* out3, int B, bool skip_softmax, bool logistic, float* pose, float* aff,
                      float* loss, cudaStream_t s, bool raw_output = false);
// ensemble mean/variance (cnn_torch_scorer.cpp:117-192): per-model arrays [M][B] -> 4 x [B]
void launch_ensemble(const float* pose, const float* aff, const float* loss, int M, int B, int stride, float* o_score,
                     float* o_aff, float* o_loss, float* o_var, cudaStream_t s);

}  // namespace gb
