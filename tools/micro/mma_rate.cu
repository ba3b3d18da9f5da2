// Microbenchmark: cycles per tcgen05.mma (kind::f16, M=128, K=16, no-swizzle K-major operands) as a function of N,
// for one or two co-resident CTAs per SM.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../gnina_b200/csrc/gb_ptx.cuh"
using namespace gb;

__global__ void __launch_bounds__(128) k(int N, int iters, int a_rows_shift, long long* out, int smem_pad) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // 1.0h
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_mbar_init(); }
  if (warp == 0) { ptx::tmem_alloc(&s_tmem, 256); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tm = s_tmem;
  long long t0 = 0, t1 = 0;
  if (warp == 0) {
    const uint64_t hi = ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
    const uint32_t a16 = ptx::smem_u32(smem) >> 4, b16 = ptx::smem_u32(smem + 16384) >> 4;
    const uint64_t ad = hi | ((uint64_t)182 << 16) | a16;   // LBO = 182 rows like the conv1 slab
    const uint64_t bd = hi | ((uint64_t)96 << 16) | b16;
    const uint32_t id = ptx::idesc_f16(128, N);
    t0 = clock64();
    if (ptx::elect_one()) {
      for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) ptx::mma_f16_ss(tm, ad + (uint64_t)((u * a_rows_shift) & 31), bd, id, 1u);
      }
      ptx::tc_commit(&bar);
    }
    __syncwarp();
    ptx::mbar_wait(&bar, 0);
    t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(tm, 256); }
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 2000;
  for (int ctas = 1; ctas <= 2; ctas++)
    for (int shift = 0; shift <= 1; shift++)
      for (int N : {32, 64, 96, 128, 192, 256}) {
        const int smem = ctas == 1 ? 150 * 1024 : 100 * 1024;  // forces 1 or 2 CTAs per SM
        k<<<148 * ctas, 128, smem>>>(N, iters, shift, d, 0);
        cudaError_t e = cudaDeviceSynchronize();
        long long cyc; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
        const double per = (double)cyc / (iters * 8);
        printf("ctas/SM=%d a_shift=%d N=%3d : %.1f cycles/MMA  (ideal %.1f)  -> %.1f%% of nominal, %s\n", ctas, shift, N, per,
               128.0 * N * 16 / 4096 * 1.0, 100.0 * (128.0 * N * 16 / 4096) / per * (ctas), cudaGetErrorString(e));
      }
  return 0;
}
