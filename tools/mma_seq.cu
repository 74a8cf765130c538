// Microbenchmark: what does the fused front end's MMA stream cost per tile when nothing else runs?
// One thread issues, per "tile", the same sequence front_tc_kernel issues: 36 x (M128 N64 K16) into D2[acc], a commit,
// the bias MMA, a commit, then block_1's 2 MMAs into D1 with two commits - in variants that drop the commits, keep one
// accumulator, or drop the accumulate=0 restart, to see which of them breaks the 48-cycle issue rate.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_seq tools/mma_seq.cu && ./mma_seq
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../superpoint-nerf-pytorch_b200/csrc/tc_ptx.cuh"
using namespace tcptx;

// flags: 1 = commits as in the kernel (4 per tile), 2 = alternate accumulators per tile, 4 = accumulate=0 on the first MMA
//        8 = walk the nine tap descriptors of a 10 x 18 halo slab (SBO 160, LBO 2880) instead of one fixed A tile
template <int flags>
__global__ void __launch_bounds__(128, 1) seq_kernel(int tiles, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[8], bar_end;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    reinterpret_cast<uint32_t*>(smem)[i] = (h & 0x83ff83ffu) | 0x38003800u | ((h >> 3) & 0x04000400u);
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1);
    mbar_init(&bar_end, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
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
    const uint32_t idesc = (1u << 4) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t hi128 = (128u >> 4) | (1u << 14), hi160 = (160u >> 4) | (1u << 14);
    const uint32_t a_flat = (smem_u32(smem) >> 4) | ((2048u >> 4) << 16);
    const uint32_t a_slab = (smem_u32(smem) >> 4) | ((2880u >> 4) << 16);
    const uint32_t b_lo0 = (smem_u32(smem + 96 * 1024) >> 4) | ((1024u >> 4) << 16);
    long long t0 = clock64();
    if (elect_one()) {
      for (int t = 0; t < tiles; ++t) {
        const uint32_t d2 = tmem + ((flags & 2) ? (uint32_t)(t & 1) * 64 : 0u);
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint32_t boff = ((uint32_t)tap * 8192 + (uint32_t)kk * 2048) >> 4;
            const uint32_t acc = ((flags & 4) && tap == 0 && kk == 0) ? 0u : 1u;
            if (flags & 8) {
              const uint32_t aoff = ((uint32_t)((tap / 3) * 10 + tap % 3) * 16 + (uint32_t)kk * 2 * 2880) >> 4;
              umma_f16_2w(d2, a_slab + (uint32_t)(t & 3) * (23552 >> 4) + aoff, hi160, b_lo0 + boff, hi128, idesc, acc);
            } else {
              umma_f16_2w(d2, a_flat + (uint32_t)((tap * 4 + kk) & 15) * (4096 >> 4), hi128, b_lo0 + boff, hi128, idesc, acc);
            }
          }
        }
        if (flags & 1) umma_commit(&bars[0]);
        umma_f16_2w(d2, a_flat, hi128, b_lo0, hi128, idesc, 1u);
        if (flags & 1) umma_commit(&bars[1]);
        for (int h = 0; h < 2; ++h) umma_f16_2w(tmem + 128 + (uint32_t)((t % 3) * 2 + h) * 64, a_flat + (uint32_t)h * (2048 >> 4), hi128, b_lo0, hi128, idesc, 0u);
        if (flags & 1) { umma_commit(&bars[2]); umma_commit(&bars[3]); }
      }
      umma_commit(&bar_end);
    }
    __syncwarp();
    mbar_wait(&bar_end, 0);
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
  const int tiles = 400;
  void (*kerns[])(int, long long*) = {seq_kernel<0>, seq_kernel<1>, seq_kernel<2>, seq_kernel<4>, seq_kernel<8>, seq_kernel<3>, seq_kernel<7>, seq_kernel<15>, seq_kernel<14>};
  const int fl[] = {0, 1, 2, 4, 8, 3, 7, 15, 14};
  for (int v = 0; v < 9; ++v) {
    const int flags = fl[v];
    cudaFuncSetAttribute(kerns[v], cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int rep = 0; rep < 2; ++rep) kerns[v]<<<148, 128, 200 * 1024>>>(tiles, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    long long h[148];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("flags %2d (%s%s%s%s): %.1f cycles per tile of 39 MMAs = %.2f cycles/MMA\n", flags, flags & 1 ? "commits " : "", flags & 2 ? "alt-acc " : "",
           flags & 4 ? "acc0-restart " : "", flags & 8 ? "slab-descs" : "", (double)mx / tiles, (double)mx / tiles / 39);
  }
  return 0;
}
