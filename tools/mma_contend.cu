// Microbenchmark: does traffic from other warps (tcgen05.ld TMEM reads, st.shared) slow down M128 x N64 x K16 MMAs?
// Warp 0 issues the MMAs; warps 4-7 (one per TMEM lane quarter) loop on tcgen05.ld.32x32b.x32 when mode & 1; warps 8-11
// loop on 16-byte st.shared (conflict-free, 4 wavefronts per instruction) with `gap` dependent FMAs between stores
// when mode & 2.  Prints cycles per MMA and the achieved rates of the contending traffic.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_contend tools/mma_contend.cu && ./mma_contend
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../superpoint-nerf-pytorch_b200/csrc/tc_ptx.cuh"
using namespace tcptx;

__global__ void __launch_bounds__(384, 1) contend_kernel(int N, int iters, int mode, int gap, int sbo, int lbo, int pat, int pitch, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ volatile int done;
  __shared__ unsigned long long cnt_ld, cnt_st;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); done = 0; cnt_ld = 0; cnt_st = 0; asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
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
    const uint32_t a_hi = ((uint32_t)sbo >> 4) | (1u << 14), b_hi = (128u >> 4) | (1u << 14);
    const uint32_t a_lo0 = (smem_u32(smem) >> 4) | (((uint32_t)lbo >> 4) << 16);
    const uint32_t b_lo0 = (smem_u32(smem + 64 * 1024) >> 4) | ((((uint32_t)N * 16) >> 4) << 16);
    // nine A offsets (16-byte units), fixed before the timed loop so that the issue loop is nothing but MMAs:
    //   pat 0: nine disjoint 4 KB-aligned tiles; pat 1: 3x3 tap shifts on a slab with `pitch` pixels per row
    uint32_t aoff[9];
    for (int k = 0; k < 9; ++k) aoff[k] = pat == 0 ? (uint32_t)k * (4096 >> 4) : (uint32_t)((k % 3) + (k / 3) * pitch);
    long long t0 = clock64();
    if (elect_one()) {
      for (int i = 0; i < iters; i += 9) {
        const uint32_t bo = (uint32_t)((i / 9) & 7) * (8192 >> 4);
#pragma unroll
        for (int k = 0; k < 9; ++k) umma_f16_2w(tmem, a_lo0 + aoff[k], a_hi, b_lo0 + bo, b_hi, idesc, 1u);
      }
      umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    done = 1;
    if (threadIdx.x == 0) out[0] = t1 - t0;
  } else if (warp >= 4 && warp < 8 && (mode & 1)) {
    const uint32_t taddr = tmem + 256 + ((uint32_t)((warp & 3) * 32) << 16);  // columns 256.. : not the accumulator
    unsigned long long n = 0;
    uint32_t acc = 0;
    while (!done) {
      uint32_t v[32];
      tmem_ld32(taddr, v);
      tmem_ld_wait();
      acc ^= v[lane & 31];
      ++n;
    }
    if (lane == 0) atomicAdd(&cnt_ld, n);
    if (acc == 0x12345678u) out[3] = acc;
  } else if (warp >= 8 && (mode & 2)) {
    uint8_t* dst = smem + 128 * 1024 + (warp - 8) * 4096 + lane * 16;
    unsigned long long n = 0;
    float f = (float)lane;
    while (!done) {
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(dst + (n & 7) * 512)), "r"(__float_as_uint(f)), "r"(1u), "r"(2u), "r"(3u) : "memory");
      for (int g = 0; g < gap; ++g) f = fmaf(f, 1.0001f, 0.5f);
      ++n;
    }
    if (lane == 0) atomicAdd(&cnt_st, n);
    if (f == 123.f) out[3] = 1;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) { out[1] = (long long)cnt_ld; out[2] = (long long)cnt_st; }
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, 8 * sizeof(long long));
  cudaFuncSetAttribute(contend_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  const int iters = 19998;  // multiple of 9
  struct { int N, mode, gap, sbo, lbo, pat, pitch; } cases[] = {
      {64, 0, 0, 128, 2048, 0, 0},   {64, 1, 0, 128, 2048, 0, 0},                                   // reference points
      {64, 0, 0, 160, 2880, 1, 10},  {64, 1, 0, 160, 2880, 1, 10},                                  // front_tc today: 8-px rows in a 10-px halo slab
      {64, 0, 0, 128, 2592, 1, 16},  {64, 1, 0, 128, 2592, 1, 16},                                  // linear 16-px rows (conv_fold geometry), 9 taps
      {64, 0, 0, 128, 2560, 1, 16},  {64, 0, 0, 256, 4096, 0, 0},  {64, 0, 0, 160, 2880, 0, 0},
      {192, 0, 0, 128, 2048, 0, 0},  {192, 0, 0, 128, 2592, 1, 16}, {192, 1, 0, 128, 2592, 1, 16}};
  for (auto c : cases) {
    cudaMemset(d, 0, 8 * sizeof(long long));
    contend_kernel<<<1, 384, 160 * 1024>>>(c.N, iters, c.mode, c.gap, c.sbo, c.lbo, c.pat, c.pitch, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    long long h[4];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    const double cyc = (double)h[0];
    printf("N %3d mode %d gap %2d SBO %3d LBO %4d pat %d : %.1f cycles/MMA | tcgen05.ld.x32 %.1f B/cycle | st.shared %.1f B/cycle (%.3f wavefronts/cycle)\n", c.N, c.mode,
           c.gap, c.sbo, c.lbo, c.pat, cyc / iters, h[1] * 4096.0 / cyc, h[2] * 512.0 / cyc, h[2] * 4.0 / cyc);
  }
  return 0;
}
