// Fast path: fused voxelise+pool (fp16, channel-chunk-planar) and tcgen05 fp16 convolutions.  See gb_cnn_tc.cu.
#pragma once
#include "gb_internal.h"

namespace gb {

struct TcPoseBatch {
  const float4* rec_xyzr; const int* rec_ch; int n_rec;
  const float4* lig_xyzr; const int* lig_ch; const int* lig_off;  // lig_off already offset to the chunk's first pose
  const float* centers;                                            // idem
  int n_poses, max_pose_atoms, n_channels, n_rec_channels;
  float resolution, dimension;
};

// Pooled input grids, double buffered so that the (CUDA-core) voxeliser of chunk i+1 can run on an auxiliary
// stream while the (tensor-core) network of chunk i runs on the main stream; shared by the models of a grid group.
struct TcGridWorkspace {
  void* x0[2] = {nullptr, nullptr};
  size_t cap[2] = {0, 0};
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
  void* buf[4] = {nullptr, nullptr, nullptr, nullptr};
  size_t cap[4] = {0, 0, 0, 0};
  void ensure(int i, size_t bytes);
  ~TcWorkspace();
};

bool tc_supported(const Model& m);
// pose lists + fused voxelise/avg-pool of one chunk into gw.x0[buf] on stream s; returns #launches
int tc_prepare_grid(const TcPoseBatch& pb, TcGridWorkspace& gw, int buf, cudaStream_t s, Profiler* prof = nullptr);
// network forward on the pooled grid x0 -> out3 [n_poses][3]; records x0_consumed (if non-null) once x0 has been
// read for the last time; returns the number of kernel launches
int tc_forward(const Model& m, const TcPoseBatch& pb, const void* x0, TcWorkspace& ws, float* out3, cudaStream_t s,
               Profiler* prof = nullptr, cudaEvent_t x0_consumed = nullptr);

// test-only access to the buffers of the most recent tc_forward on this thread: 0 x0, 1 y(3), 2 x2, 3 x4, 4 y5
const void* tc_debug_buffer(int i, size_t* bytes);

}  // namespace gb
