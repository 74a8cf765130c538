// Microbenchmark: tcgen05.ld (TMEM -> registers) throughput per SM with 4 (or 8) warps reading 32x32b.x32 back to back.
// Answers "why does folding three taps into N = 192 (3x the accumulator readback) not give 1.5x?" (DESIGN.md section 6).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ldtm_rate tools/ldtm_rate.cu && ./ldtm_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../superpoint-nerf-pytorch_b200/csrc/tc_ptx.cuh"
using namespace tcptx;

__global__ void __launch_bounds__(256, 1) ldtm_kernel(int iters, int nwarps, long long* out, uint32_t* sink) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  uint32_t acc = 0;
  long long t0 = clock64();
  if (warp < nwarps) {
    const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    for (int i = 0; i < iters; ++i) {
      uint32_t v[32];
      tmem_ld32(taddr + (uint32_t)((i & 15) * 32), v);
      tmem_ld_wait();
#pragma unroll
      for (int k = 0; k < 32; ++k) acc ^= v[k];
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * 256 + threadIdx.x] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
  }
}

int main() {
  long long* d; uint32_t* sink;
  cudaMalloc(&d, 148 * sizeof(long long));
  cudaMalloc(&sink, 148 * 256 * 4);
  const int iters = 20000;
  for (int nw : {1, 4, 8}) {
    ldtm_kernel<<<148, 256>>>(iters, nw, d, sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    long long h[148];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    const double bytes = (double)iters * nw * 32 * 32 * 4;
    printf("%d warps: %.1f cycles per LDTM.x32 per warp, %.1f B/cycle/SM TMEM->RF\n", nw, (double)mx / iters, bytes / mx);
  }
  return 0;
}
