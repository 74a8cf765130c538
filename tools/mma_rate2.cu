// Microbenchmark: issue rate of tcgen05.mma.cta_group::2 (M 256 = 128 rows per CTA of a 2-CTA cluster) x N x K16,
// SWIZZLE_NONE K-major SMEM operands, B split between the two CTAs (N/2 rows each).  Companion of mma_rate.cu: is the
// pair form faster than two independent cta_group::1 M128 MMAs for the N = 64 convolution tile?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate2 tools/mma_rate2.cu && ./mma_rate2
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../superpoint-nerf-pytorch_b200/csrc/tc_ptx.cuh"
using namespace tcptx;

__device__ __forceinline__ uint32_t ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void csync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(128, 1) rate2_kernel(int N, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = ctarank();
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  asm volatile("fence.proxy.async;" ::: "memory");
  csync();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  csync();   // the peer's TMEM is allocated before the leader issues
  const uint32_t tmem = tmem_base_s;
  if (warp == 0 && rank == 0) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((256u >> 4) << 24);
    const uint32_t a_hi = (128u >> 4) | (1u << 14), b_hi = a_hi;
    const uint32_t a_lo0 = (smem_u32(smem) >> 4) | ((2048u >> 4) << 16);
    const uint32_t b_lo0 = (smem_u32(smem + 64 * 1024) >> 4) | ((((uint32_t)(N / 2) * 16) >> 4) << 16);   // this CTA's N/2 rows of B
    long long t0 = clock64();
    if (elect_one()) {
      for (int i = 0; i < iters; ++i) {
        const uint32_t ao = (uint32_t)(i & 15) * (4096 >> 4);
        const uint32_t bo = (uint32_t)(i & 7) * (8192 >> 4);
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\tsetp.ne.b32 p, %6, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(tmem), "r"(a_lo0 + ao), "r"(a_hi), "r"(b_lo0 + bo), "r"(b_hi),
            "r"(idesc), "r"(1u)
            : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)),
                   "h"((uint16_t)1)
                   : "memory");
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x >> 1] = t1 - t0;
  }
  tc_fence_before();
  csync();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  cudaFuncSetAttribute(rate2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  const int iters = 4000;
  for (int grid : {2, 148})
    for (int N : {64, 128, 192, 256}) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 160 * 1024;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      cudaError_t e = cudaLaunchKernelEx(&cfg, rate2_kernel, N, iters, d);
      if (e == cudaSuccess) e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      long long h[74];
      cudaMemcpy(h, d, (grid / 2) * sizeof(long long), cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int i = 0; i < grid / 2; ++i) mx = h[i] > mx ? h[i] : mx;
      const double cyc = (double)mx / iters;
      printf("grid %3d (pairs %2d)  cta_group::2  M 256  N %3d : %.1f cycles per MMA = %.1f per 128-row tile  (cta_group::1 M128: ideal %d, "
             "measured 48 / 64 / 96 / 128 for N 64 / 128 / 192 / 256)\n", grid, grid / 2, N, cyc, cyc / 2, 128 * N * 16 / 4096);
    }
  return 0;
}
