// Microbenchmark: issue rate of tcgen05.mma M128 x N x K16 (kind::f16, SWIZZLE_NONE K-major SMEM operands) for several N.
// Answers "is the M128xN64 convolution tile bound by shared-memory operand bandwidth?" (DESIGN.md section 6).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate tools/mma_rate.cu && ./mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../superpoint-nerf-pytorch_b200/csrc/tc_ptx.cuh"
using namespace tcptx;

__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int iters, int distinct_a, long long* out, int random_data = 0) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    // random fp16 pairs in (-2, 2): sign + exponent 0x30..0x3f + random mantissa (tensor-core power depends on the data)
    reinterpret_cast<uint32_t*>(smem)[i] = random_data ? ((h & 0x83ff83ffu) | 0x38003800u | ((h >> 3) & 0x04000400u)) : 0u;
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (warp == 0) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    // A: 128 rows x 16 K: [chunk 2][128 rows][16 B]  (LBO 2048, SBO 128); B: [chunk 2][N rows][16 B] (LBO N*16, SBO 128)
    const uint32_t a_hi = (128u >> 4) | (1u << 14), b_hi = a_hi;
    const uint32_t a_lo0 = (smem_u32(smem) >> 4) | ((2048u >> 4) << 16);
    const uint32_t b_lo0 = (smem_u32(smem + 64 * 1024) >> 4) | ((((uint32_t)N * 16) >> 4) << 16);
    long long t0 = clock64();
    if (elect_one()) {
      for (int i = 0; i < iters; ++i) {
        // rotate through 16 different A / B tiles so that no operand is trivially reused
        const uint32_t ao = distinct_a ? (uint32_t)(i & 15) * (4096 >> 4) : 0u;
        const uint32_t bo = (uint32_t)(i & 7) * (8192 >> 4);
        umma_f16_2w(tmem, a_lo0 + ao, a_hi, b_lo0 + bo, b_hi, idesc, 1u);
      }
      umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  const int iters = 4000;
  for (int grid : {1, 148})
    for (int N : {64, 128, 192, 256})
      for (int da : {1, 0}) {
        rate_kernel<<<grid, 128, 160 * 1024>>>(N, iters, da, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        long long h[148];
        cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
        const double cyc = (double)mx / iters;
        printf("grid %3d  N %3d  distinct_A %d : %.1f cycles/MMA  -> %.0f MAC/cycle/SM (ideal tensor time %d cycles, smem operand bytes %d)\n",
               grid, N, da, cyc, 128.0 * N * 16 / cyc, 128 * N * 16 / 4096, 4096 + N * 32);
      }
  // Sustained runs on all SMs: SM cycles (clock64) against wall time (events) = the clock the chip holds under this load.
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int rnd : {0, 1})
  for (int N : {64, 192, 256}) {
    const int long_iters = 400000;
    rate_kernel<<<148, 128, 160 * 1024>>>(N, long_iters, 1, d, rnd);  // warm
    cudaEventRecord(e0);
    for (int rep = 0; rep < 40; ++rep) rate_kernel<<<148, 128, 160 * 1024>>>(N, long_iters, 1, d, rnd);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h[148];
    cudaMemcpy(h, d, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    const double cyc = (double)mx / long_iters, sec = ms / 1e3 / 40;
    printf("sustained grid 148 %s data N %3d: %.1f cycles/MMA, %.2f ms per launch -> SM clock %.0f MHz, %.0f TFLOP/s\n", rnd ? "random" : "zero  ", N, cyc, sec * 1e3,
           (double)mx / sec / 1e6, 148.0 * long_iters * 2.0 * 128 * N * 16 / sec / 1e12);
  }
  return 0;
}
