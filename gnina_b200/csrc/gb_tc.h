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

// Pooled input grids of the current chunk, shared by the models of a grid group that pool the same way.
struct TcGridWorkspace {
  void* x0[2] = {nullptr, nullptr};  // [0] avg-pooled, [1] max-pooled
  size_t cap[2] = {0, 0};
  bool valid[2] = {false, false};
  float4* list_xyzr = nullptr; int* list_ch = nullptr; int* list_n = nullptr; size_t list_cap = 0, listn_cap = 0;
  bool lists_valid = false;
  void receptor_changed() { valid[0] = valid[1] = false; lists_valid = false; }
  void batch_done() { valid[0] = valid[1] = false; lists_valid = false; }
  ~TcGridWorkspace();
};

struct TcWorkspace {
  void* buf[4] = {nullptr, nullptr, nullptr, nullptr};
  size_t cap[4] = {0, 0, 0, 0};
  void ensure(int i, size_t bytes);
  ~TcWorkspace();
};

bool tc_supported(const Model& m);
// -> out3 [n_poses][3]; returns the number of kernel launches
int tc_forward(const Model& m, const TcPoseBatch& pb, TcGridWorkspace& gw, TcWorkspace& ws, float* out3, cudaStream_t s,
               Profiler* prof = nullptr);

// test-only access to the buffers of the most recent tc_forward on this thread: 0 x0, 1 y(3), 2 x2, 3 x4, 4 y5
const void* tc_debug_buffer(int i, size_t* bytes);

}  // namespace gb
