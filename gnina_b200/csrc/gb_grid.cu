// Voxeliser, fp32 validation layout.  Replaces libmolgrid::GridMaker::forward as called at
// gninasrc/lib/torch_model.cpp:181 (and the torch::zeros memset at :179): one launch voxelises a whole batch of
// poses into float[B][C][N][N][N] (z fastest), every voxel written exactly once (no memset, no atomics).
//
// Two kernels:
//  build_pose_lists : per pose, ordered compaction of the (channel-sorted) receptor atoms whose support can touch
//                     the pose's grid box, followed by the pose's ligand atoms -> a channel-sorted atom list.
//  voxelize_f32     : one CTA per (8x8x8-voxel tile, pose); ordered compaction of the pose list against the tile,
//                     then each thread sums its voxel channel by channel (lists are channel-sorted, so a running
//                     accumulator is flushed whenever the channel advances; empty channels get explicit zeros).
// Arithmetic follows oracle/gridmaker_ref.c operation by operation (no FMA contraction) so that the validation
// grid agrees with the oracle to a few ulp.
#include "gb_internal.h"

namespace gb {

__device__ __forceinline__ int block_ordered_slot(bool pred, int* s_warp_counts, int& total) {
  // returns the position of this thread's element among all threads with pred (ordered by tid), and the total
  const unsigned mask = __ballot_sync(0xffffffffu, pred);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  if (lane == 0) s_warp_counts[warp] = __popc(mask);
  __syncthreads();
  int base = 0, tot = 0;
  for (int w = 0; w < nwarps; w++) {
    int c = s_warp_counts[w];
    if (w < warp) base += c;
    tot += c;
  }
  __syncthreads();
  total = tot;
  return base + __popc(mask & ((1u << lane) - 1u));
}

__global__ void __launch_bounds__(256) build_pose_lists_kernel(const float4* __restrict__ rec_xyzr,
                                                               const int* __restrict__ rec_ch, int n_rec,
                                                               const float4* __restrict__ lig_xyzr,
                                                               const int* __restrict__ lig_ch,
                                                               const int* __restrict__ lig_off,
                                                               const float* __restrict__ centers, float half_dim,
                                                               int cap, float4* __restrict__ list_xyzr,
                                                               int* __restrict__ list_ch, int* __restrict__ list_n,
                                                               const float* __restrict__ rot) {
  __shared__ int s_counts[8];
  const int p = blockIdx.x;
  const float cx = centers[3 * p], cy = centers[3 * p + 1], cz = centers[3 * p + 2];
  // G3: Transform(center, 0, rotate) (torch_model.cpp:170-173) -- every atom of the pose (receptor and ligand) is
  // rotated about the grid centre before gridding; rot = row-major 3x3 per pose, NULL = identity
  float R[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f};
  if (rot)
    for (int k = 0; k < 9; k++) R[k] = rot[9 * p + k];
  float4* out_a = list_xyzr + (size_t)p * cap;
  int* out_c = list_ch + (size_t)p * cap;
  int n_out = 0;
  for (int pass = 0; pass < 2; pass++) {
    const float4* src = pass == 0 ? rec_xyzr : lig_xyzr;
    const int* srcc = pass == 0 ? rec_ch : lig_ch;
    const int b = pass == 0 ? 0 : lig_off[p], e = pass == 0 ? n_rec : lig_off[p + 1];
    for (int base = b; base < e; base += blockDim.x) {
      const int i = base + threadIdx.x;
      bool keep = false;
      float4 a = make_float4(0, 0, 0, 0);
      int ch = 0;
      if (i < e) {
        a = src[i];
        ch = srcc[i];
        if (rot) {
          const float dx = a.x - cx, dy = a.y - cy, dz = a.z - cz;
          a.x = cx + (R[0] * dx + R[1] * dy + R[2] * dz);
          a.y = cy + (R[3] * dx + R[4] * dy + R[5] * dz);
          a.z = cz + (R[6] * dx + R[7] * dy + R[8] * dz);
        }
        const float reach = half_dim + 1.5f * a.w;
        keep = fabsf(a.x - cx) <= reach && fabsf(a.y - cy) <= reach && fabsf(a.z - cz) <= reach;
      }
      int tot;
      const int slot = block_ordered_slot(keep, s_counts, tot);
      if (keep && n_out + slot < cap) {
        out_a[n_out + slot] = a;
        out_c[n_out + slot] = ch;
      }
      n_out += tot;
    }
  }
  if (threadIdx.x == 0) list_n[p] = n_out < cap ? n_out : cap;
}

void launch_build_pose_lists(const float4* rec_xyzr, const int* rec_ch, int n_rec, const float4* lig_xyzr,
                             const int* lig_ch, const int* lig_off, const float* centers, int n_poses, float half_dim,
                             int cap, float4* list_xyzr, int* list_ch, int* list_n, cudaStream_t s, const float* rot) {
  if (n_poses <= 0) return;
  build_pose_lists_kernel<<<n_poses, 256, 0, s>>>(rec_xyzr, rec_ch, n_rec, lig_xyzr, lig_ch, lig_off, centers, half_dim,
                                                  cap, list_xyzr, list_ch, list_n, rot);
}

__device__ __forceinline__ float density_exact(float dx, float dy, float dz, float ar) {
  // oracle/gridmaker_ref.c:density — libmolgrid GridMaker::calc_point, G = 1, final multiple 1.5
  const float rsq = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
  const float dist = sqrtf(rsq);
  if (dist >= __fmul_rn(ar, 1.5f)) return 0.f;
  if (dist <= ar) return expf(__fdiv_rn(__fmul_rn(__fmul_rn(-2.f, dist), dist), __fmul_rn(ar, ar)));
  const float e2 = 0.13533528323661270f;  // expf(-2)
  const float A = 4.f * e2, B = -12.f * e2, C = 9.f * e2;
  const float dr = __fdiv_rn(dist, ar);
  const float q = __fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(A, dr), B), dr), C);
  return q > 0.f ? q : 0.f;
}

constexpr int kTile = 8;
constexpr int kChunk = 512;

__global__ void __launch_bounds__(512) voxelize_f32_kernel(const float4* __restrict__ list_xyzr,
                                                           const int* __restrict__ list_ch,
                                                           const int* __restrict__ list_n, int cap,
                                                           const float* __restrict__ centers, int n_channels, int npts,
                                                           float resolution, float dimension,
                                                           float* __restrict__ grid) {
  __shared__ float4 s_atom[kChunk];
  __shared__ int s_ch[kChunk];
  __shared__ int s_counts[16];
  const int p = blockIdx.y;
  const int tiles = (npts + kTile - 1) / kTile;
  const int t = blockIdx.x;
  const int ti = t / (tiles * tiles), tj = (t / tiles) % tiles, tk = t % tiles;
  const int li = threadIdx.x >> 6, lj = (threadIdx.x >> 3) & 7, lk = threadIdx.x & 7;
  const int i = ti * kTile + li, j = tj * kTile + lj, k = tk * kTile + lk;
  const bool valid = i < npts && j < npts && k < npts;
  const float half = dimension / 2.f;
  const float ox = centers[3 * p] - half, oy = centers[3 * p + 1] - half, oz = centers[3 * p + 2] - half;
  const float gx = __fadd_rn(ox, __fmul_rn((float)i, resolution));
  const float gy = __fadd_rn(oy, __fmul_rn((float)j, resolution));
  const float gz = __fadd_rn(oz, __fmul_rn((float)k, resolution));
  // tile bounds (grid coordinates of first/last voxel of the tile)
  const float lox = ox + ti * kTile * resolution, hix = lox + (kTile - 1) * resolution;
  const float loy = oy + tj * kTile * resolution, hiy = loy + (kTile - 1) * resolution;
  const float loz = oz + tk * kTile * resolution, hiz = loz + (kTile - 1) * resolution;
  const size_t vol = (size_t)npts * npts * npts;
  float* out = grid + (size_t)p * n_channels * vol + ((size_t)i * npts + j) * npts + k;
  const float4* la = list_xyzr + (size_t)p * cap;
  const int* lc = list_ch + (size_t)p * cap;
  const int n = list_n[p];
  int cur = 0;
  float acc = 0.f;
  for (int base = 0; base < n; base += kChunk) {
    const int a_i = base + threadIdx.x;
    bool keep = false;
    float4 a = make_float4(0, 0, 0, 0);
    int ch = 0;
    if (a_i < n) {
      a = la[a_i];
      ch = lc[a_i];
      const float reach = 1.5f * a.w + 1e-4f;
      keep = a.x >= lox - reach && a.x <= hix + reach && a.y >= loy - reach && a.y <= hiy + reach &&
             a.z >= loz - reach && a.z <= hiz + reach;
    }
    int tot;
    const int slot = block_ordered_slot(keep, s_counts, tot);
    if (keep) {
      s_atom[slot] = a;
      s_ch[slot] = ch;
    }
    __syncthreads();
    for (int m = 0; m < tot; m++) {
      const int ch_m = s_ch[m];
      while (cur < ch_m) {  // uniform across the CTA
        if (valid) out[(size_t)cur * vol] = acc;
        acc = 0.f;
        cur++;
      }
      const float4 am = s_atom[m];
      acc += density_exact(gx - am.x, gy - am.y, gz - am.z, am.w);
    }
    __syncthreads();
  }
  while (cur < n_channels) {
    if (valid) out[(size_t)cur * vol] = acc;
    acc = 0.f;
    cur++;
  }
}

void launch_voxelize_f32(const float4* list_xyzr, const int* list_ch, const int* list_n, int cap, const float* centers,
                         int n_poses, int n_channels, int npts, float resolution, float dimension, float* grid,
                         cudaStream_t s) {
  if (n_poses <= 0) return;
  const int tiles = (npts + kTile - 1) / kTile;
  dim3 g(tiles * tiles * tiles, n_poses);
  voxelize_f32_kernel<<<g, 512, 0, s>>>(list_xyzr, list_ch, list_n, cap, centers, n_channels, npts, resolution,
                                        dimension, grid);
}

// GridMaker::backward (call site gninasrc/lib/torch_model.cpp:203): one warp per listed atom, lanes stride over the
// atom's bounding box of voxels, warp-shuffle reduction of the three gradient components.
// d rho/d x_a = rho'(d) (x_a - v)/d ; rho'(d) = -4 d/r^2 exp(-2 d^2/r^2) (d <= r), (2 A d/r + B)/r (r < d < 1.5 r).
__global__ void __launch_bounds__(256) grid_backward_kernel(const float4* __restrict__ atoms_xyzr,
                                                            const int* __restrict__ atoms_ch,
                                                            const int* __restrict__ pose_off,
                                                            const float* __restrict__ centers, int n_channels, int npts,
                                                            float resolution, float dimension,
                                                            const float* __restrict__ dgrid, float* __restrict__ atom_grad,
                                                            const float* __restrict__ rot) {
  const int p = blockIdx.y;
  const int a_i = pose_off[p] + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (a_i >= pose_off[p + 1]) return;
  float4 a = atoms_xyzr[a_i];
  // G3: the grid was built from atoms rotated about the centre (x' = c + R (x - c)); the gradient is taken at x' and
  // rotated back, d/dx = R^T d/dx' (Transform::backward, torch_model.cpp:204-206)
  float R[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f};
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
    const float* g = dgrid + ((size_t)p * n_channels + ch) * npts * npts * npts;
    const float e2 = 0.13533528323661270f, A = 4.f * e2, Bq = -12.f * e2;
    for (int e = lane; e < ni * nj * nk; e += 32) {
      const int k = lo[2] + e % nk, j = lo[1] + (e / nk) % nj, i = lo[0] + e / (nk * nj);
      const float dx = (ox + i * resolution) - a.x, dy = (oy + j * resolution) - a.y, dz = (oz + k * resolution) - a.z;
      const float dist = sqrtf(dx * dx + dy * dy + dz * dz);
      if (dist >= reach || dist == 0.f) continue;
      float dr;
      if (dist <= ar) dr = -4.f * dist / (ar * ar) * expf(-2.f * dist * dist / (ar * ar));
      else dr = (2.f * A * (dist / ar) + Bq) / ar;
      const float sc = g[((size_t)i * npts + j) * npts + k] * dr / dist;
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
    og[0] = R[0] * gx + R[3] * gy + R[6] * gz;   // R^T g (identity without rotation)
    og[1] = R[1] * gx + R[4] * gy + R[7] * gz;
    og[2] = R[2] * gx + R[5] * gy + R[8] * gz;
  }
}

void launch_grid_backward(const float4* atoms_xyzr, const int* atoms_ch, const int* pose_off, int max_pose_atoms,
                          const float* centers, int n_poses, int n_channels, int npts, float resolution, float dimension,
                          const float* dgrid, float* atom_grad, cudaStream_t s, const float* rot) {
  if (n_poses <= 0 || max_pose_atoms <= 0) return;
  dim3 g((max_pose_atoms + 7) / 8, n_poses);
  grid_backward_kernel<<<g, 256, 0, s>>>(atoms_xyzr, atoms_ch, pose_off, centers, n_channels, npts, resolution, dimension,
                                         dgrid, atom_grad, rot);
}

__global__ void axpy_range_kernel(const float* __restrict__ src, float* __restrict__ dst, int lo, int hi, float alpha) {
  const int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
  if (i < hi) dst[i] += alpha * src[i];
}
void launch_axpy_range(const float* src, float* dst, int lo, int hi, float alpha, cudaStream_t s) {
  if (hi <= lo) return;
  axpy_range_kernel<<<(hi - lo + 255) / 256, 256, 0, s>>>(src, dst, lo, hi, alpha);
}

}  // namespace gb
