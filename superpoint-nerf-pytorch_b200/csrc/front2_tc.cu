// CTA-PAIR variant of the fused front end (front_tc.cu): homography warp -> block_1 -> block_2 with every tensor-core
// instruction issued as tcgen05.mma.cta_group::2 (M = 256 = two 128-pixel tiles, one per CTA of a 2-CTA cluster).
//
// Why: the M128 x N64 x K16 MMAs of block_2 read 4 KB (A) + 2 KB (B) from shared memory per 32 tensor cycles, i.e. they
// are bound by the 128 B/cycle shared-memory pipe at 2/3 of the tensor rate (tools/mma_rate.cu: 48 cycles per MMA).  As
// a pair, each CTA supplies its own 128 rows of A and only HALF of B (32 of the 64 output channels; the peer's half
// arrives over the pair link): 5 KB per CTA per MMA -> 40 cycles, and each CTA keeps half of block_2's weights
// resident (37 KB instead of 74 KB), which pays for a fifth slab stage.
//
// What changes relative to front_tc.cu (roles, tile shape, data flow and numerics are identical):
//   * launch: clusters of 2 CTAs (cudaLaunchAttributeClusterDimension), CTA rank r of pair u owns tile 2u + r;
//   * TMEM: tcgen05.alloc / dealloc .cta_group::2 by warp 1 of BOTH CTAs (same columns in both);
//   * only the leader CTA (rank 0) issues MMAs; its two issuing warps wait on barriers that collect the arrivals of BOTH
//     CTAs' producer roles (the peer arrives through mapa + mbarrier.arrive.shared::cluster), and commit with
//     tcgen05.commit.cta_group::2 ... multicast::cluster so that each CTA's own consumer barriers are signalled;
//   * weights: each CTA bulk-loads the 512-byte halves [chunk][32 couts][8] of the existing operand-B images.
//
// Reference semantics: K.warp_perspective(image, H, bilinear) (engine_solvers/export.py:51) + VGG_Block 1 and 2
// (models/model_utils/VGG_Backbone.py:23-36, 60-63).
#include "spn_common.cuh"
#include "spn_geom.cuh"
#include "tc_ptx.cuh"

namespace {

using namespace tcptx;
using namespace spngeom;

constexpr int kTW = 8, kTH = 16;
constexpr int kPW = kTW + 2, kPH = kTH + 2;          // block_1 halo tile: 10 x 18 = 180 pixels
constexpr int kHalo = kPW * kPH;
constexpr int kQW = kTW + 4, kQH = kTH + 4;          // warped-image patch: 12 x 20 = 240 pixels
constexpr uint32_t kChStride = (uint32_t)kHalo * 16;  // slab bytes between 8-channel groups (2880)
constexpr int kSlabBytes = 8 * kHalo * 16;            // 23040
constexpr int kStageBytes = (kSlabBytes + 1023) & ~1023;
constexpr int kHalfBlk = 2 * 32 * 16;                 // one (tap, k-step) operand-B block of this CTA: [chunk 2][cout 32][8] = 1 KB
constexpr int kW2Bytes = 37 * kHalfBlk;               // 36 weight blocks + the bias block
constexpr int kOnesBytes = 4096;                      // constant operand A of the bias MMA
constexpr int kW1Bytes = kHalfBlk;                    // block_1 weights (taps + bias rows), this CTA's 32 couts
constexpr int kA1Rows = 256;
constexpr int kA1Bytes = 2 * kA1Rows * 16;            // [chunk 2][row 256][8 halfs]
constexpr int kStages = 5;
constexpr int kNA1 = 3;                               // A1 / D1 buffers: MMA1 runs two tiles ahead of MMA2
constexpr int kThreads = 576;                         // 18 warps
constexpr int kPThreads = 256;                        // P role: warps 10-17

struct Front2Params {
  const float* images;   // [n_src][H][W] fp32
  const float* hinv;     // [n_src][n_h][9] pixel-space H^-1, or null (plain forward: slot == source image)
  int n_h;               // homographies per source image (slot = src * (n_h + 1) + j, j == 0 is the identity)
  int slot_begin, n_slots;
  int H, W, tiles_x, tiles_y;
  unsigned long long magic_tpi, magic_tx;   // fast_div magics for tiles_per_img and tiles_x
  int is_bf16;
  const void* w1img;     // operand-B image of block_1: [chunk 2][cout 64][8]
  const void* w2img;     // operand-B image of block_2: 36 x [chunk 2][cout 64][8], then the bias block
  void* out;             // C8 [n_slots][8][H/2][W/2][8]
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster (release at cluster scope)
__device__ __forceinline__ void mbar_arrive_cta(uint64_t* bar, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
// wait with acquire at cluster scope (the arrivals may come from the peer CTA)
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void umma2_commit(uint64_t* bar) {  // arrives on `bar` in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t"
      "}" ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ uint16_t to16(float v, int bf) {
  if (bf) {
    __nv_bfloat16 h = __float2bfloat16_rn(v);
    return *reinterpret_cast<uint16_t*>(&h);
  }
  __half h = __float2half_rn(v);
  return *reinterpret_cast<uint16_t*>(&h);
}

__global__ void __launch_bounds__(kThreads, 1) front2_tc_kernel(const Front2Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // Barriers the leader's MMA issuers wait on collect arrivals from both CTAs (they are used in CTA 0 only); barriers
  // signalled by tcgen05.commit exist, and are used, in both CTAs.
  __shared__ __align__(8) uint64_t bar_w, bar_w_peer, bar_a1_full[kNA1], bar_a1_empty[kNA1], bar_d1_full[kNA1], bar_d1_empty[kNA1],
      bar_slab_full[kStages], bar_slab_empty[kStages], bar_d2_full[2], bar_d2_empty[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) uint16_t patch_s[kNA1][kQH * kQW];

  uint8_t* w2s = smem;                                  // 37 KB
  uint8_t* w1s = smem + kW2Bytes;                       // 1 KB
  uint8_t* a1s = w1s + kW1Bytes;                        // kNA1 x 8 KB
  uint8_t* ones = a1s + kNA1 * kA1Bytes;                // 4 KB
  uint8_t* slab0 = smem + (((kW2Bytes + kW1Bytes + kNA1 * kA1Bytes + kOnesBytes) + 1023) & ~1023);   // kStages x 23.5 KB

  griddep_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int n_tiles = p.n_slots * tiles_per_img;
  const int n_units = (n_tiles + 1) >> 1;               // a unit = two consecutive tiles, one per CTA of the pair

  if (threadIdx.x == 0) {
    mbar_init(&bar_w, 1);
    mbar_init(&bar_w_peer, 1);
    for (int i = 0; i < kNA1; ++i) {
      mbar_init(&bar_a1_full[i], 2 * kPThreads); mbar_init(&bar_a1_empty[i], 1);
      mbar_init(&bar_d1_full[i], 1);             mbar_init(&bar_d1_empty[i], 2 * 128);
    }
    for (int i = 0; i < 2; ++i) { mbar_init(&bar_d2_full[i], 1); mbar_init(&bar_d2_empty[i], 2 * 128); }
    for (int i = 0; i < kStages; ++i) { mbar_init(&bar_slab_full[i], 2 * 128); mbar_init(&bar_slab_empty[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // A1 rows >= 180 are never written again: zero the buffers once
  for (int i = threadIdx.x; i < kNA1 * kA1Bytes / 16; i += kThreads) reinterpret_cast<uint4*>(a1s)[i] = make_uint4(0, 0, 0, 0);
  {
    const uint32_t one2 = p.is_bf16 ? 0x3F803F80u : 0x3C003C00u;
    for (int i = threadIdx.x; i < kOnesBytes / 16; i += kThreads)
      reinterpret_cast<uint4*>(ones)[i] = i < 128 ? make_uint4(one2, 0u, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
  }
  asm volatile("fence.proxy.async;" ::: "memory");  // the fills above are read by the tensor core (of the pair)
  cluster_sync_all();                                // both CTAs' barriers are initialised before anyone arrives remotely
  if (warp == 1) {  // TMEM: D2 2 x 64 columns + D1 3 buffers x 2 halves x 64 columns = 512 columns, in both CTAs
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t fmt = p.is_bf16 ? 1u : 0u;
  const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((64u >> 3) << 17) | ((256u >> 4) << 24);   // M 256, N 64

  if (warp <= 1) {
    // ===================== weight loader (warp 0 of each CTA) + MMA issuers (warps 0, 1 of the leader) ==============
    if (warp == 0) {
      if (elect_one()) {
        mbar_expect_tx(&bar_w, (uint32_t)(kW2Bytes + kW1Bytes));
        // this CTA's 32 output channels of every [chunk 2][cout 64][8] block: two 512-byte pieces per block
        for (int blk = 0; blk < 37; ++blk)
          for (int j = 0; j < 2; ++j)
            bulk_load(w2s + blk * kHalfBlk + j * 512, (const uint8_t*)p.w2img + (size_t)blk * 2048 + j * 1024 + rank * 512, 512, &bar_w);
        for (int j = 0; j < 2; ++j) bulk_load(w1s + j * 512, (const uint8_t*)p.w1img + j * 1024 + rank * 512, 512, &bar_w);
      }
      __syncwarp();
      if (rank == 1) {  // tell the leader that the peer's weight halves have landed
        mbar_wait(&bar_w, 0);
        if (lane == 0) mbar_arrive_cta(&bar_w_peer, 0);
        __syncwarp();
      }
    }
    if (rank == 0) {
      const int par = warp;
      mbar_wait(&bar_w, 0);
      mbar_wait_cluster(&bar_w_peer, 0);
      const uint32_t hi_a1 = (128u >> 4) | (1u << 14);                  // SBO 128 B (8 rows x 16 B), version 1
      const uint32_t lo_a1_c = ((uint32_t)(kA1Rows * 16) >> 4) << 16;   // LBO 4096 B between the two K chunks
      const uint32_t hi_b = (128u >> 4) | (1u << 14);
      const uint32_t lo_b_c = (512u >> 4) << 16;                        // LBO 512 B: 32 couts x 16 B per K chunk
      const uint32_t hi_a2 = ((uint32_t)(kPW * 16) >> 4) | (1u << 14);  // SBO = one halo row (160 B)
      const uint32_t lo_a2_c = (kChStride >> 4) << 16;
      const uint32_t w1_lo = (smem_u32(w1s) >> 4) | lo_b_c, w2_lo = (smem_u32(w2s) >> 4) | lo_b_c;
      const uint32_t a1_addr = smem_u32(a1s), slab_addr = smem_u32(slab0);
      const uint32_t d1_col = tmem_base + 128;

      auto issue_mma1 = [&](int i) {  // block_1 for local unit i: D1[b][h] = A1[b] rows h*128.. times W1, in both CTAs
        const int b = i % kNA1;
        const uint32_t ph = (uint32_t)(i / kNA1) & 1u;
        mbar_wait_cluster(&bar_a1_full[b], ph);
        mbar_wait_cluster(&bar_d1_empty[b], ph ^ 1u);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint32_t a_lo = ((a1_addr + (uint32_t)b * kA1Bytes + (uint32_t)h * 128 * 16) >> 4) | lo_a1_c;
            umma2_f16(d1_col + (uint32_t)(b * 2 + h) * 64, a_lo, hi_a1, w1_lo, hi_b, idesc, 0u);
          }
          umma2_commit(&bar_a1_empty[b]);
          umma2_commit(&bar_d1_full[b]);
        }
        __syncwarp();
      };

      if (pair + par * n_pairs < n_units) issue_mma1(par);
      for (int i = par, u = pair + par * n_pairs; u < n_units; u += 2 * n_pairs, i += 2) {
        const int stage = i % kStages, acc = i & 1;
        const uint32_t sph = (uint32_t)(i / kStages) & 1u, aph = (uint32_t)(i >> 1) & 1u;
        mbar_wait_cluster(&bar_slab_full[stage], sph);
        mbar_wait_cluster(&bar_d2_empty[acc], aph ^ 1u);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_lo = ((slab_addr + (uint32_t)stage * kStageBytes) >> 4) | lo_a2_c;
          const uint32_t d2 = tmem_base + (uint32_t)acc * 64;
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const int ky = tap / 3, kx = tap % 3;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint32_t aoff = ((uint32_t)(ky * kPW + kx) * 16 + (uint32_t)kk * 2 * kChStride) >> 4;
              const uint32_t boff = ((uint32_t)(tap * 4 + kk) * kHalfBlk) >> 4;
              umma2_f16(d2, a_lo + aoff, hi_a2, w2_lo + boff, hi_b, idesc, (tap | kk) ? 1u : 0u);
            }
          }
          umma2_commit(&bar_slab_empty[stage]);
          // + bias2: D2 += ones[128 x 16] . biasB[64 x 16]
          umma2_f16(d2, (smem_u32(ones) >> 4) | ((2048u >> 4) << 16), (128u >> 4) | (1u << 14), w2_lo + ((36u * kHalfBlk) >> 4), hi_b,
                    idesc, 1u);
          umma2_commit(&bar_d2_full[acc]);
        }
        __syncwarp();
        if (u + 2 * n_pairs < n_units) issue_mma1(i + 2);
      }
    }
  } else if (warp >= 10) {
    // ===================== P: warped patch + im2col operand of block_1 =====================
    const int pt = threadIdx.x - 320;  // 0..255
    griddep_wait();  // the output buffer may still be read by the previous chunk's kernels; every store follows P's data
    float hm[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f};
    int hm_slot = -1;
    int i = 0;
    for (int u = pair; u < n_units; u += n_pairs, ++i) {
      const int t = 2 * u + (int)rank;
      const bool live = t < n_tiles;   // an odd tile count leaves the peer of the last unit without a tile: zero operand
      const int b = i % kNA1;
      const uint32_t ph = (uint32_t)(i / kNA1) & 1u;
      const int tt = live ? t : 0;
      const int ls = fast_div(tt, p.magic_tpi), rr = tt - ls * tiles_per_img;
      const int ty = fast_div(rr, p.magic_tx), tx = rr - ty * p.tiles_x;
      const int slot = p.slot_begin + ls;
      const int src = p.hinv ? slot / (p.n_h + 1) : slot;
      const int j = p.hinv ? slot - src * (p.n_h + 1) : 0;
      const float* img = p.images + (size_t)src * p.H * p.W;
      mbar_wait(&bar_a1_empty[b], ph ^ 1u);  // MMA1 of unit i-3 has consumed A1[b] (and patch_s[b] long before)
      if (j > 0 && slot != hm_slot) {  // warp-uniform; a CTA's consecutive tiles mostly belong to the same slot
        const float* hp = p.hinv + ((size_t)src * p.n_h + (j - 1)) * 9;
#pragma unroll
        for (int k = 0; k < 9; ++k) hm[k] = __ldg(hp + k);
        hm_slot = slot;
      }
      // 1. the 12 x 20 patch of the warped image around the tile (zero outside the frame = block_1's padding)
      for (int s = pt; s < kQH * kQW; s += kPThreads) {
        const int py = s / kQW, px = s - py * kQW;
        const int y = ty * kTH - 2 + py, x = tx * kTW - 2 + px;
        float v = 0.f;
        if (live && y >= 0 && y < p.H && x >= 0 && x < p.W) {
          if (j == 0) {
            v = __ldg(&img[(size_t)y * p.W + x]);
          } else {
            float sx, sy;
            apply_h(hm, (float)x, (float)y, sx, sy);
            v = bilinear_zero_nb(img, sx, sy, p.H, p.W);
          }
        }
        patch_s[b][s] = to16(v, p.is_bf16);
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      // 2. A1 row r = halo pixel (hy, hx): its 3x3 neighbourhood (K 0..8), then two constant-one columns that
      //    multiply the (hi, lo) bias rows of W1.  Halo pixels outside the image get an all-zero row, so block_1's
      //    output there is exactly 0 = block_2's zero padding.
      const uint32_t one16 = p.is_bf16 ? 0x3F80u : 0x3C00u;
      for (int r = pt; r < kHalo; r += kPThreads) {
        const int hy = r / kPW, hx = r - hy * kPW;
        const int gy = ty * kTH - 1 + hy, gx = tx * kTW - 1 + hx;
        uint4 c0 = make_uint4(0u, 0u, 0u, 0u), c1 = make_uint4(0u, 0u, 0u, 0u);
        if (live && gy >= 0 && gy < p.H && gx >= 0 && gx < p.W) {
          const uint16_t* q = &patch_s[b][hy * kQW + hx];
          c0.x = q[0] | ((uint32_t)q[1] << 16);
          c0.y = q[2] | ((uint32_t)q[kQW] << 16);
          c0.z = q[kQW + 1] | ((uint32_t)q[kQW + 2] << 16);
          c0.w = q[2 * kQW] | ((uint32_t)q[2 * kQW + 1] << 16);
          c1.x = q[2 * kQW + 2] | (one16 << 16);
          c1.y = one16;
        }
        uint8_t* dst = a1s + (size_t)b * kA1Bytes + (size_t)r * 16;
        *reinterpret_cast<uint4*>(dst) = c0;                    // taps 0..7
        *reinterpret_cast<uint4*>(dst + kA1Rows * 16) = c1;     // tap 8, one, one, zero padding
      }
      asm volatile("fence.proxy.async;" ::: "memory");
      mbar_arrive_cta(&bar_a1_full[b], 0);
    }
  } else if (warp >= 6) {
    // ===================== E1: block_1 epilogue -> block_2's input slab =====================
    const int q = warp & 3;
    int i = 0;
    for (int u = pair; u < n_units; u += n_pairs, ++i) {
      const int b = i % kNA1, stage = i % kStages;
      const uint32_t ph = (uint32_t)(i / kNA1) & 1u, sph = (uint32_t)(i / kStages) & 1u;
      mbar_wait(&bar_d1_full[b], ph);
      mbar_wait(&bar_slab_empty[stage], sph ^ 1u);
      tc_fence_after();
      uint8_t* slab = slab0 + (size_t)stage * kStageBytes;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (h * 128 + q * 32 >= kHalo) continue;  // warp-uniform: this lane quarter holds no halo pixel
        const int r = h * 128 + q * 32 + lane;
        uint32_t v[64];
        const uint32_t taddr = tmem_base + 128 + (uint32_t)(b * 2 + h) * 64 + ((uint32_t)(q * 32) << 16);
        tmem_ld32(taddr, v);
        tmem_ld32(taddr + 32, v + 32);
        tmem_ld_wait();
        if (r < kHalo) {  // bias is already inside D1 (ones columns of A1): ReLU + 16-bit pack + 8 x 16-byte stores
#pragma unroll
          for (int c8 = 0; c8 < 8; ++c8) {
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int c = c8 * 8 + 2 * e;
              w[e] = pack2_relu(__uint_as_float(v[c]), __uint_as_float(v[c + 1]), p.is_bf16);
            }
            *reinterpret_cast<uint4*>(slab + (size_t)c8 * kChStride + (size_t)r * 16) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
      }
      tc_fence_before();
      mbar_arrive_cta(&bar_d1_empty[b], 0);
      asm volatile("fence.proxy.async;" ::: "memory");
      mbar_arrive_cta(&bar_slab_full[stage], 0);
    }
  } else {
    // ===================== E2: block_2 epilogue (bias, ReLU, 2x2 max-pool, C8 store) =====================
    const int q = warp & 3;
    const int g = q * 4 + (lane >> 3), r = lane & 7;
    const int Ho = p.H >> 1, Wo = p.W >> 1;
    int i = 0;
    for (int u = pair; u < n_units; u += n_pairs, ++i) {
      const int t = 2 * u + (int)rank;
      const bool live = t < n_tiles;
      const int acc = i & 1;
      const uint32_t aph = (uint32_t)(i >> 1) & 1u;
      const int tt = live ? t : 0;
      const int ls = fast_div(tt, p.magic_tpi), rr = tt - ls * tiles_per_img;
      const int ty = fast_div(rr, p.magic_tx), tx = rr - ty * p.tiles_x;
      const int y = ty * kTH + g, x = tx * kTW + r;
      mbar_wait(&bar_d2_full[acc], aph);
      tc_fence_after();
      uint32_t v[64];
      const uint32_t taddr = tmem_base + (uint32_t)acc * 64 + ((uint32_t)(q * 32) << 16);
      tmem_ld32(taddr, v);
      tmem_ld32(taddr + 32, v + 32);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive_cta(&bar_d2_empty[acc], 0);
      uint32_t h2[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) h2[c] = pack2_relu(__uint_as_float(v[2 * c]), __uint_as_float(v[2 * c + 1]), p.is_bf16);
#pragma unroll
      for (int c = 0; c < 32; ++c) {  // bias is already in D2; ReLU was applied by the conversion (it commutes with the pool)
        h2[c] = max2(h2[c], __shfl_xor_sync(0xffffffffu, h2[c], 1), p.is_bf16);
        h2[c] = max2(h2[c], __shfl_xor_sync(0xffffffffu, h2[c], 8), p.is_bf16);
      }
      if (live && y < p.H && x < p.W && (g & 1) == 0 && (r & 1) == 0) {
        uint4* o = reinterpret_cast<uint4*>(p.out);
        const int oy = y >> 1, ox = x >> 1;
#pragma unroll
        for (int jj = 0; jj < 8; ++jj)
          o[(((size_t)ls * 8 + jj) * Ho + oy) * Wo + ox] = make_uint4(h2[4 * jj], h2[4 * jj + 1], h2[4 * jj + 2], h2[4 * jj + 3]);
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();  // both CTAs are done with the pair's TMEM and with each other's shared memory
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

}  // namespace

// d_out: C8 [n_slots][8][H/2][W/2][8]; weights of block_1 / block_2 from the context.
int spn_front2_tc_launch(spn_ctx* ctx, const float* d_images, const float* d_hinv, int n_h, int slot_begin, int n_slots, int H,
                         int W, int mode, const void* w1img, void* d_out, cudaStream_t s) {
  const int bf = mode == SPN_MODE_BF16 ? 1 : 0;
  const SpnLayer& L2 = ctx->layers[SPN_L_BLOCK2];
  SPN_REQUIRE(L2.cin == 64 && L2.cout == 64 && L2.ks == 3 && L2.w16[bf], "front end needs block_2 = 3x3 conv 64->64 with packed weights");
  Front2Params p;
  memset(&p, 0, sizeof(p));
  p.images = d_images; p.hinv = d_hinv; p.n_h = n_h; p.slot_begin = slot_begin; p.n_slots = n_slots;
  p.H = H; p.W = W; p.tiles_x = spn_cdiv(W, kTW); p.tiles_y = spn_cdiv(H, kTH); p.is_bf16 = bf;
  p.magic_tpi = fast_div_magic(p.tiles_x * p.tiles_y); p.magic_tx = fast_div_magic(p.tiles_x);
  SPN_REQUIRE((long long)n_slots * p.tiles_x * p.tiles_y * (p.tiles_x * p.tiles_y) < (1ll << 40), "too many tiles for one launch");
  p.w1img = w1img; p.w2img = L2.w16[bf];
  p.out = d_out;
  const size_t dyn = (size_t)(((kW2Bytes + kW1Bytes + kNA1 * kA1Bytes + kOnesBytes) + 1023) & ~1023) + (size_t)kStages * kStageBytes + 1024;
  const long long units = ((long long)n_slots * p.tiles_x * p.tiles_y + 1) / 2;
  const int max_pairs = ctx->sm_count / 2;
  const int pairs = (int)(units < max_pairs ? units : max_pairs);
  SPN_CUDA(cudaFuncSetAttribute(front2_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = dyn;
  cfg.stream = s;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = ctx->opt_pdl ? 2 : 1;
  SpnProfScope prof(ctx, SPN_L_BLOCK2, s);
  SPN_CUDA(cudaLaunchKernelEx(&cfg, front2_tc_kernel, p));
  SPN_CHECK_LAUNCH(ctx);
  return SPN_OK;
}
