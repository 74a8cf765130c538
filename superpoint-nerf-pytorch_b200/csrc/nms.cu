// box_nms + threshold + keypoint compaction, bit-exact with the reference's greedy formulation.
//
// Reference: models/model_utils/sp_utils.py:4-28 (box_nms = torchvision.ops.nms over size x size boxes centred
// on every pixel >= min_prob, optional top-k, scatter), heads.py:41 / export.py:123-125 (threshold + nonzero).
//
// torchvision's nms is the sequential greedy algorithm over a stable descending sort.  Because all boxes are
// the same size and centred on integer pixels, "box j suppresses box i" depends only on the pixel offset, so
// the greedy result equals the unique fixed point of:
//     a candidate is KEPT       iff every footprint neighbour that outranks it is SUPPRESSED,
//     a candidate is SUPPRESSED iff some footprint neighbour is KEPT,
// with rank = (score desc, row-major index asc).  The kernel iterates that rule to the fixed point with one
// __syncthreads per round; each round only touches still-undecided pixels.
// Status bytes live in a caller-invisible scratch buffer and stay L1/L2 resident.
#include <math.h>

#include <cooperative_groups.h>

#include "spn_common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int kMaxFoot = 288;  // 17x17 - 1 : box size up to 8
constexpr int kThreads = 1024;

struct NmsFoot {
  int n;
  int r;  // max |offset|
  int8_t dy[kMaxFoot];
  int8_t dx[kMaxFoot];
};

// IoU test exactly as torchvision computes it in fp32 for boxes [c - s/2, c + s/2] (no +1 on widths).
bool foot_suppresses(float size, float iou, int dy, int dx) {
  const float h = size / 2.0f;
  const float ay1 = 0.0f - h, ax1 = 0.0f - h, ay2 = 0.0f + h, ax2 = 0.0f + h;
  const float by1 = (float)dy - h, bx1 = (float)dx - h, by2 = (float)dy + h, bx2 = (float)dx + h;
  const float areaa = (ay2 - ay1) * (ax2 - ax1), areab = (by2 - by1) * (bx2 - bx1);
  const float yy1 = fmaxf(ay1, by1), xx1 = fmaxf(ax1, bx1), yy2 = fminf(ay2, by2), xx2 = fminf(ax2, bx2);
  const float w = fmaxf(0.0f, yy2 - yy1), hh = fmaxf(0.0f, xx2 - xx1);
  const float inter = w * hh;
  const float ovr = inter / (areaa + areab - inter);
  return ovr > iou;
}

__device__ __forceinline__ uint32_t ordered_key(float f) {  // monotone float -> uint
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// Phase 1: greedy NMS as a parallel fixed point, all SMs cooperating on the whole batch.
// Work unit = 32x32 pixel tile (one thread per pixel) with a halo of `foot.r` pixels staged in shared memory
// (scores + status bytes, coalesced loads).  Inside a tile the rule is iterated to a local fixed point with the halo
// frozen - every deduction made that way is valid for the global fixed point, pixels that depend on a still
// undecided halo neighbour simply stay pending - then the tile's statuses are written back and the grid
// synchronises; rounds repeat until no pixel of any image is pending.
__global__ void __launch_bounds__(kThreads)
nms_rounds_kernel(const float* __restrict__ prob_all, uint8_t* __restrict__ status_all, NmsFoot foot, int B, int H, int W,
                  float min_prob, int tiles_x, int tiles_y, unsigned* __restrict__ pend) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) uint8_t nms_smem[];
  const int r = foot.r, HT = 32 + 2 * r, HP = HT + 1;
  float* sc = reinterpret_cast<float*>(nms_smem);            // [HT][HP]
  uint8_t* st = nms_smem + (size_t)HT * HP * sizeof(float);  // [HT][HP]
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  const size_t P = (size_t)H * W;
  const int n_tiles = B * tiles_x * tiles_y;

  for (size_t i = (size_t)blockIdx.x * kThreads + tid; i < (size_t)B * P; i += (size_t)gridDim.x * kThreads)
    status_all[i] = (__ldg(&prob_all[i]) >= min_prob) ? 1 : 0;
  if (blockIdx.x == 0 && tid < 8) pend[tid] = 0;
  grid.sync();

  for (int round = 0;; ++round) {
    bool pending = false;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int b = tile / (tiles_x * tiles_y), rem = tile - b * (tiles_x * tiles_y);
      const int y0 = (rem / tiles_x) * 32 - r, x0 = (rem % tiles_x) * 32 - r;
      const float* prob = prob_all + (size_t)b * P;
      uint8_t* status = status_all + (size_t)b * P;
      bool any_undecided = false;
      for (int i = tid; i < HT * HT; i += kThreads) {
        const int hy = i / HT, hx = i - hy * HT;
        const int y = y0 + hy, x = x0 + hx;
        uint8_t v = 0;
        float s = 0.f;
        if (y >= 0 && y < H && x >= 0 && x < W) {
          v = __ldcg(&status[(size_t)y * W + x]);
          if (v == 1) s = __ldg(&prob[(size_t)y * W + x]);
          any_undecided |= (v == 1) && hy >= r && hy < HT - r && hx >= r && hx < HT - r;
        }
        st[hy * HP + hx] = v;
        sc[hy * HP + hx] = s;
      }
      if (!__syncthreads_or(any_undecided)) continue;  // tile already decided (also orders the smem fill)
      const int cy = ty + r, cx = tx + r;
      const int gy = y0 + cy, gx = x0 + cx;
      const int me = cy * HP + cx;
      // Only higher-ranked candidate neighbours ("blockers") can decide this pixel: it is suppressed iff one of
      // them is kept and kept iff all of them are suppressed.  Collect them once per visit as a bit mask.
      uint64_t blockers = 0;
      bool mine = st[me] == 1;
      if (mine) {
        const float sp = sc[me];
        if (foot.n <= 64) {
          for (int k = 0; k < foot.n; ++k) {
            const int dy = foot.dy[k], dx = foot.dx[k];
            const int q = me + dy * HP + dx;
            const uint8_t v = st[q];
            if (v == 2) { blockers = 0; st[me] = 0; mine = false; break; }
            if (v == 1) {
              const float sq = sc[q];
              if (sq > sp || (sq == sp && (dy < 0 || (dy == 0 && dx < 0)))) blockers |= 1ull << k;
            }
          }
          if (mine && blockers == 0) { st[me] = 2; mine = false; }
        }
      }
      __syncthreads();
      for (int it = 0; it < 256; ++it) {
        bool changed = false;
        if (mine) {
          if (foot.n <= 64) {
            uint64_t rem = blockers;
            int res = 1;
            while (rem) {
              const int k = __ffsll((long long)rem) - 1;
              rem &= rem - 1;
              const uint8_t v = st[me + foot.dy[k] * HP + foot.dx[k]];
              if (v == 2) { res = 0; break; }
              if (v == 0) blockers &= ~(1ull << k);
            }
            if (res == 1 && blockers == 0) res = 2;
            if (res != 1) { st[me] = (uint8_t)res; mine = false; changed = true; }
          } else {  // large footprints (box size > 4): full scan every iteration
            const float sp = sc[me];
            int res = 2;
            for (int k = 0; k < foot.n; ++k) {
              const int dy = foot.dy[k], dx = foot.dx[k];
              const int q = me + dy * HP + dx;
              const uint8_t v = st[q];
              if (v == 2) { res = 0; break; }
              if (v == 1) {
                const float sq = sc[q];
                if (sq > sp || (sq == sp && (dy < 0 || (dy == 0 && dx < 0)))) res = 1;
              }
            }
            if (res != 1) { st[me] = (uint8_t)res; mine = false; changed = true; }
          }
        }
        if (!__syncthreads_or(changed)) break;
        if (tid == 0) atomicAdd(&pend[4], 1u);  // statistics: local iterations
      }
      if (tid == 0) atomicAdd(&pend[5], 1u);    // statistics: tile visits with undecided pixels
      if (gy < H && gx < W) {
        const uint8_t v = st[me];
        status[(size_t)gy * W + gx] = v;
        pending |= (v == 1);
      }
      __syncthreads();  // smem is reused by the next tile
    }
    const bool blk_pending = __syncthreads_or(pending);
    if (tid == 0) {
      if (blk_pending) atomicAdd(&pend[round % 3], 1u);
      // slot (round+1)%3 was last read after the grid.sync of round-2; every CTA has since passed the sync of
      // round-1, so it is safe to clear it here for its next use in round+1
      if (blockIdx.x == 0) pend[(round + 1) % 3] = 0;
    }
    grid.sync();
    if (blockIdx.x == 0 && tid == 0) pend[3] = (unsigned)round + 1;  // statistics: global rounds
    if (__ldcg(&pend[round % 3]) == 0) break;
  }
}

// Same fixed point for footprints of radius RAD <= 3 (box sizes up to 4, the reference's default), with the tile's
// state held as two bit planes in shared memory - KEPT and GONE (suppressed or never a candidate), one 64-bit word
// per halo row - so that every step is a fixed, branch-free sequence per warp:
//   * warp w owns tile row w; a pixel's (2 RAD + 1)^2 neighbourhood is one funnel shift per row word, packed at
//     8 bits per row (bit (dy + RAD) * 8 + (dx + RAD)), the layout of its `blockers` mask and of the footprint mask;
//   * blockers = footprint & undecided-candidate neighbourhood & "outranks me" (one compare per tap, offsets known
//     at compile time), collected once per visit;
//   * suppressed iff (kept_nbhd & blockers) != 0, else blockers &= ~gone_nbhd and kept iff none remain;
//   * newly decided pixels are published with two ballots per warp (a single writer per row word).
// Tiles that still hold undecided pixels are appended to a compact list for the next round, so every round's work is
// spread evenly over the grid; the iteration ends when that list is empty.
template <int RAD>
__global__ void __launch_bounds__(kThreads, 2)
nms_rounds_bits_kernel(const float* __restrict__ prob_all, uint8_t* __restrict__ status_all, uint32_t foot_lo,
                       uint32_t foot_hi, int B, int H, int W, float min_prob, int tiles_x, int tiles_y,
                       unsigned* __restrict__ pend, int* __restrict__ lists) {
  cg::grid_group grid = cg::this_grid();
  constexpr int R = 2 * RAD + 1, HT = 32 + 2 * RAD, HP = HT + 1;
  constexpr uint32_t rowbits = (1u << R) - 1u;
  __shared__ uint64_t kept[HT], gone[HT];
  __shared__ float sc[HT * HP];
  __shared__ uint8_t st[HT * HP];
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  const size_t P = (size_t)H * W;
  const int n_tiles = B * tiles_x * tiles_y;

  for (size_t i = (size_t)blockIdx.x * kThreads + tid; i < (size_t)B * P; i += (size_t)gridDim.x * kThreads)
    status_all[i] = (__ldg(&prob_all[i]) >= min_prob) ? 1 : 0;
  for (int i = blockIdx.x * kThreads + tid; i < n_tiles; i += gridDim.x * kThreads) lists[i] = i;
  if (blockIdx.x == 0 && tid < 8) pend[tid] = tid == 0 ? (unsigned)n_tiles : 0u;
  grid.sync();

  // pend[k % 3] = length of the tile list consumed in round k; lists[(k & 1) * n_tiles ...] holds it
  for (int round = 0;; ++round) {
    const int n_cur = (int)__ldcg(&pend[round % 3]);
    const int* cur = lists + (round & 1) * n_tiles;
    int* nxt = lists + ((round + 1) & 1) * n_tiles;
    unsigned* n_nxt = &pend[(round + 1) % 3];
    if (blockIdx.x == 0 && tid == 0) pend[(round + 2) % 3] = 0;  // consumed in round - 1, appended to in round + 1
    for (int idx = blockIdx.x; idx < n_cur; idx += gridDim.x) {
      const int tile = __ldcg(&cur[idx]);
      const int b = tile / (tiles_x * tiles_y), rem = tile - b * (tiles_x * tiles_y);
      const int y0 = (rem / tiles_x) * 32 - RAD, x0 = (rem % tiles_x) * 32 - RAD;
      const float* prob = prob_all + (size_t)b * P;
      uint8_t* status = status_all + (size_t)b * P;
      bool any_undecided = false;
      for (int i = tid; i < HT * HT; i += kThreads) {
        const int hy = i / HT, hx = i - hy * HT;
        const int y = y0 + hy, x = x0 + hx;
        uint8_t v = 0;
        float s = 0.f;
        if (y >= 0 && y < H && x >= 0 && x < W) {
          v = __ldcg(&status[(size_t)y * W + x]);
          if (v == 1) s = __ldg(&prob[(size_t)y * W + x]);
          any_undecided |= (v == 1) && hy >= RAD && hy < HT - RAD && hx >= RAD && hx < HT - RAD;
        }
        st[hy * HP + hx] = v;
        sc[hy * HP + hx] = s;
      }
      if (!__syncthreads_or(any_undecided)) continue;  // nothing left to decide here (also orders the smem fill)
      // bit planes of the tile + halo
      for (int hy = ty; hy < HT; hy += 32) {
        const bool has1 = tx + 32 < HT;
        const uint8_t v0 = st[hy * HP + tx], v1 = has1 ? st[hy * HP + tx + 32] : (uint8_t)0;
        const unsigned k0 = __ballot_sync(0xffffffffu, v0 == 2), k1 = __ballot_sync(0xffffffffu, has1 && v1 == 2);
        const unsigned g0 = __ballot_sync(0xffffffffu, v0 == 0), g1 = __ballot_sync(0xffffffffu, !has1 || v1 == 0);
        if (tx == 0) {
          kept[hy] = (uint64_t)k0 | ((uint64_t)k1 << 32);
          gone[hy] = (uint64_t)g0 | ((uint64_t)g1 << 32);
        }
      }
      const int cy = ty + RAD, cx = tx + RAD;
      const int gy = y0 + cy, gx = x0 + cx;
      const int me = cy * HP + cx;
      int stv = st[me];
      bool mine = stv == 1;
      __syncthreads();  // planes complete
      uint32_t bl = 0, bh = 0;  // undecided neighbours that outrank this pixel and overlap it ("blockers")
      bool first = true;
      for (int it = 0;; ++it) {
        int res_new = 1;  // 0 / 2: decided in this step
        if (__any_sync(0xffffffffu, mine)) {
          uint32_t knl = 0, knh = 0, gnl = 0, gnh = 0;
#pragma unroll
          for (int i = 0; i < R; ++i) {
            const uint2 kw = *reinterpret_cast<const uint2*>(&kept[ty + i]);
            const uint2 gw = *reinterpret_cast<const uint2*>(&gone[ty + i]);
            const uint32_t kb = __funnelshift_r(kw.x, kw.y, tx) & rowbits;
            const uint32_t gb = __funnelshift_r(gw.x, gw.y, tx) & rowbits;
            if (i < 4) {
              knl |= kb << (8 * i);
              gnl |= gb << (8 * i);
            } else {
              knh |= kb << (8 * (i - 4));
              gnh |= gb << (8 * (i - 4));
            }
          }
          if (first) {  // warp-uniform: collect the blockers
            const float sp = sc[me];
            uint32_t ol = 0, oh = 0;
#pragma unroll
            for (int i = 0; i < R; ++i) {
#pragma unroll
              for (int j = 0; j < R; ++j) {
                if (i == RAD && j == RAD) continue;
                const float sq = sc[me + (i - RAD) * HP + (j - RAD)];
                const bool earlier = (i < RAD) || (i == RAD && j < RAD);  // row-major index tie break
                const bool o = earlier ? (sq >= sp) : (sq > sp);
                if (i < 4) ol |= (uint32_t)o << (8 * i + j); else oh |= (uint32_t)o << (8 * (i - 4) + j);
              }
            }
            bl = ol & ~(knl | gnl) & foot_lo;
            bh = oh & ~(knh | gnh) & foot_hi;
          }
          if (mine) {
            if (((knl & foot_lo) | (knh & foot_hi)) != 0) {
              res_new = 0;
            } else {
              bl &= ~gnl;
              bh &= ~gnh;
              if ((bl | bh) == 0) res_new = 2;
            }
            if (res_new != 1) { mine = false; stv = res_new; }
          }
        }
        first = false;
        const unsigned bk = __ballot_sync(0xffffffffu, res_new == 2), bg = __ballot_sync(0xffffffffu, res_new == 0);
        if (tx == 0 && (bk | bg)) {
          kept[cy] |= (uint64_t)bk << RAD;
          gone[cy] |= (uint64_t)bg << RAD;
        }
        if (!__syncthreads_or((bk | bg) != 0)) break;
        if (tid == 0) atomicAdd(&pend[4], 1u);  // statistics: local iterations
      }
      if (tid == 0) atomicAdd(&pend[5], 1u);    // statistics: tile visits with undecided pixels
      bool still = false;
      if (gy < H && gx < W) {
        status[(size_t)gy * W + gx] = (uint8_t)stv;
        still = stv == 1;
      }
      if (__syncthreads_or(still) && tid == 0) nxt[atomicAdd(n_nxt, 1u)] = tile;  // barrier: smem is reused next
    }
    grid.sync();
    if (blockIdx.x == 0 && tid == 0) pend[3] = (unsigned)round + 1;  // statistics: global rounds
    if (__ldcg(n_nxt) == 0) break;
  }
}

// Phase 2: ordered compaction over S CTAs per image + (optional) top-k selection on the compacted survivor list.
//
// nms_emit_kernel<MODE>: every CTA owns a contiguous pixel range of one image (row-major), counts the pixels that pass
// the predicate, publishes the count and looks back over the earlier segments of the same image (decoupled look-back:
// a CTA's segment index is an atomic ticket, so everything it waits for has already started and publishes before it
// waits itself), then writes its elements at their global row-major rank.
//   MODE 0: predicate = survived the NMS; output = (ordered key, pixel index) list for the top-k selection
//   MODE 1: predicate = survived and >= det_thresh; output = keypoint list (row, col), optional dense maps, count
// One CTA per image (the round-1 layout) made this phase 63 us for ONE 240x320 image - two latency-bound passes of 75
// iterations per warp on a single SM; with S = 2 x SMs / B segments a single image takes a few microseconds per pass.
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(kThreads)
nms_emit_kernel(const float* __restrict__ prob_all, const uint8_t* __restrict__ status_all, int H, int W, int S,
                float det_thresh, float* __restrict__ nms_all, int32_t* __restrict__ pred_all,
                int32_t* __restrict__ kp_all, int32_t* __restrict__ kp_count, int max_kp,
                uint32_t* __restrict__ list_all, int* __restrict__ list_count, unsigned* __restrict__ flags,
                unsigned* __restrict__ ticket) {
  __shared__ int s_warp[kThreads / 32];
  __shared__ unsigned s_ticket;
  __shared__ int s_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_ticket = atomicAdd(ticket, 1u);
  __syncthreads();
  const int b = (int)(s_ticket / (unsigned)S), sg = (int)(s_ticket % (unsigned)S);
  const int P = H * W;
  const float* prob = prob_all + (size_t)b * P;
  const uint8_t* status = status_all + (size_t)b * P;
  const int seg_px = (((P + S - 1) / S) + 31) & ~31;                  // CTA range, multiple of 32
  const int c_begin = min(P, sg * seg_px), c_end = min(P, c_begin + seg_px);
  const int wseg = ((c_end - c_begin + kThreads - 1) / kThreads) * 32;  // warp range, multiple of 32
  const int p_begin = min(c_end, c_begin + warp * wseg), p_end = min(c_end, p_begin + wseg);

  auto pass = [&](int p, float& v) -> bool {   // predicate + the value the outputs need
    v = (status[p] == 2) ? __ldg(&prob[p]) : 0.0f;
    return MODE == 0 ? status[p] == 2 : v >= det_thresh;
  };
  int cnt = 0;
  for (int p = p_begin + lane; p < p_end; p += 32) {
    float v;
    cnt += pass(p, v) ? 1 : 0;
  }
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) s_warp[warp] = cnt;
  __syncthreads();
  int base = 0, total = 0;
  for (int w = 0; w < kThreads / 32; ++w) {
    const int c = s_warp[w];
    if (w < warp) base += c;
    total += c;
  }
  if (tid == 0) {
    st_release_u32(&flags[b * S + sg], ((unsigned)total << 1) | 1u);
    int before = 0;
    for (int k = sg - 1; k >= 0; --k) {
      unsigned f;
      do { f = ld_acquire_u32(&flags[b * S + k]); } while (!(f & 1u));
      before += (int)(f >> 1);
    }
    s_base = before;
    if (sg == S - 1) {
      if (MODE == 0) list_count[b] = before + total;
      else if (kp_count) kp_count[b] = before + total;
    }
  }
  __syncthreads();
  base += s_base;

  float* nms = (MODE == 1 && nms_all) ? nms_all + (size_t)b * P : nullptr;
  int32_t* pred = (MODE == 1 && pred_all) ? pred_all + (size_t)b * P : nullptr;
  int32_t* kp = (MODE == 1 && kp_all) ? kp_all + (size_t)b * max_kp * 2 : nullptr;
  uint32_t* lkey = MODE == 0 ? list_all + (size_t)b * 2 * P : nullptr;   // [kept] ordered keys
  int32_t* lidx = MODE == 0 ? reinterpret_cast<int32_t*>(lkey + P) : nullptr;  // [kept] pixel indices, ascending
  for (int p0 = p_begin; p0 < p_end; p0 += 32) {
    const int p = p0 + lane;
    bool hit = false;
    float v = 0.f;
    if (p < p_end) {
      hit = pass(p, v);
      if (nms) nms[p] = v;
      if (pred) pred[p] = hit ? 1 : 0;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, hit);
    const int rank = base + __popc(bal & ((1u << lane) - 1u));
    if (hit) {
      if (MODE == 0) {
        lkey[rank] = ordered_key(v);
        lidx[rank] = p;
      } else if (kp && rank < max_kp) {
        const int y = p / W;
        kp[2 * rank] = y;
        kp[2 * rank + 1] = p - y * W;
      }
    }
    base += __popc(bal);
  }
}

// top-k over the survivors: (score desc, index asc).  One CTA per image walks the compacted (key, index) list written
// by nms_emit_kernel<0> (a few percent of the pixels): 4-pass radix select of the top_k-th key, then the ties are ranked
// in index order; the losers' status bytes are cleared so that nms_emit_kernel<1> drops them.
__global__ void __launch_bounds__(kThreads)
nms_topk_select_kernel(uint8_t* __restrict__ status_all, int P, int top_k, const uint32_t* __restrict__ list_all,
                       const int* __restrict__ list_count) {
  __shared__ int s_warp[kThreads / 32];
  __shared__ unsigned s_hist[256];
  __shared__ unsigned s_sel[2];
  const int b = blockIdx.x;
  const int kept = list_count[b];
  if (kept <= top_k) return;
  uint8_t* status = status_all + (size_t)b * P;
  const uint32_t* lkey = list_all + (size_t)b * 2 * P;
  const int32_t* lidx = reinterpret_cast<const int32_t*>(lkey + P);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t prefix = 0, maskbits = 0;
  unsigned want = (unsigned)top_k;  // rank (1-based, descending) of the threshold element
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    if (tid < 256) s_hist[tid] = 0;
    __syncthreads();
    for (int i = tid; i < kept; i += kThreads) {
      const uint32_t key = lkey[i];
      if ((key & maskbits) == prefix) atomicAdd(&s_hist[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (tid == 0) {
      unsigned cum = 0;
      int bin = 255;
      for (; bin > 0; --bin) {
        if (cum + s_hist[bin] >= want) break;
        cum += s_hist[bin];
      }
      s_sel[0] = (unsigned)bin;
      s_sel[1] = want - cum;
    }
    __syncthreads();
    prefix |= s_sel[0] << shift;
    maskbits |= 0xffu << shift;
    want = s_sel[1];
    __syncthreads();
  }
  // prefix = key of the top_k-th survivor; keep keys > prefix and the first `want` equal keys in index order.
  // Warp w owns list positions [w*lseg, (w+1)*lseg): count its ties, one block scan, then rank them in order.
  const int lseg = ((kept + kThreads - 1) / kThreads) * 32;
  const int lb = min(kept, warp * lseg), le = min(kept, lb + lseg);
  int tcnt = 0;
  for (int i = lb + lane; i < le; i += 32) {
    const uint32_t key = lkey[i];
    if (key < prefix) status[lidx[i]] = 0;
    tcnt += (key == prefix);
  }
  for (int o = 16; o > 0; o >>= 1) tcnt += __shfl_xor_sync(0xffffffffu, tcnt, o);
  if (lane == 0) s_warp[warp] = tcnt;
  __syncthreads();
  int tie_base = 0;
  for (int w = 0; w < warp; ++w) tie_base += s_warp[w];
  if (tcnt > 0) {  // warp-uniform
    for (int i0 = lb; i0 < le; i0 += 32) {
      const int i = i0 + lane;
      const bool tie = i < le && lkey[i] == prefix;
      const unsigned bal = __ballot_sync(0xffffffffu, tie);
      if (tie && (unsigned)(tie_base + __popc(bal & ((1u << lane) - 1u))) >= want) status[lidx[i]] = 0;
      tie_base += __popc(bal);
    }
  }
}

}  // namespace

extern "C" int spn_nms_stats(spn_ctx* ctx, int B, int H, int W, int64_t* h_out) {
  SPN_REQUIRE(ctx && h_out && ctx->aux, "spn_nms_stats: nothing to read");
  SpnDeviceGuard guard(ctx->device);
  const size_t status_bytes = ((size_t)B * H * W + 255) & ~(size_t)255;
  unsigned v[8];
  SPN_CUDA(cudaDeviceSynchronize());
  SPN_CUDA(cudaMemcpy(v, ctx->aux + status_bytes, sizeof(v), cudaMemcpyDeviceToHost));
  h_out[0] = v[3]; h_out[1] = v[4]; h_out[2] = v[5];
  return SPN_OK;
}

extern "C" int spn_box_nms_topk(spn_ctx* ctx, const float* d_prob, int B, int H, int W, float size, float iou,
                                float min_prob, int top_k, float det_thresh, float* d_nms, int32_t* d_pred,
                                int32_t* d_kp, int32_t* d_kp_count, int max_kp, spn_stream stream) {
  SPN_REQUIRE(ctx && d_prob, "spn_box_nms_topk: null pointer");
  SpnDeviceGuard guard(ctx->device);
  SPN_REQUIRE(B > 0 && H > 0 && W > 0 && (long long)H * W < (1ll << 30), "spn_box_nms_topk: bad shape");
  SPN_REQUIRE(size > 0.f && size <= 8.f, "spn_box_nms_topk: box size must be in (0, 8]");
  SPN_REQUIRE(!d_kp || max_kp > 0, "spn_box_nms_topk: max_kp must be > 0 when d_kp is given");
  NmsFoot foot;
  memset(&foot, 0, sizeof(foot));
  const int r = (int)ceilf(size);
  for (int dy = -r; dy <= r; ++dy)
    for (int dx = -r; dx <= r; ++dx)
      if ((dy || dx) && foot_suppresses(size, iou, dy, dx)) {
        if (foot.n >= kMaxFoot) { spn_set_error("spn_box_nms_topk: footprint too large"); return SPN_E_INVALID; }
        foot.dy[foot.n] = (int8_t)dy;
        foot.dx[foot.n] = (int8_t)dx;
        foot.n++;
      }
  foot.r = 0;
  for (int k = 0; k < foot.n; ++k) foot.r = max(foot.r, max(abs((int)foot.dy[k]), abs((int)foot.dx[k])));
  cudaStream_t s = (cudaStream_t)stream;
  const size_t status_bytes = ((size_t)B * H * W + 255) & ~(size_t)255;
  int tiles_x = spn_cdiv(W, 32), tiles_y = spn_cdiv(H, 32);
  const long long n_tiles = (long long)B * tiles_x * tiles_y;
  SPN_REQUIRE(n_tiles < (1ll << 28), "spn_box_nms_topk: batch too large");
  const size_t lists_bytes = (2 * (size_t)n_tiles * sizeof(int) + 255) & ~(size_t)255;
  const size_t topk_bytes = top_k > 0 ? (size_t)B * H * W * 8 : 0;   // (key, index) list of the survivors, per image
  // phase 2 geometry: S segments (CTAs) per image, about two CTAs per SM in total, at least 2048 pixels each
  int S = (2 * ctx->sm_count + B - 1) / B;
  S = max(1, min(S, min(32, H * W / 2048)));
  const size_t sync_bytes = ((size_t)(2 * B * S + B + 8) * sizeof(unsigned) + 255) & ~(size_t)255;  // flags x2, list counts, tickets
  int rc = spn_ensure_aux(ctx, status_bytes + 256 + lists_bytes + topk_bytes + sync_bytes, s);
  if (rc) return rc;
  uint8_t* status = (uint8_t*)ctx->aux;
  unsigned* pend = (unsigned*)(ctx->aux + status_bytes);
  int* lists = (int*)(ctx->aux + status_bytes + 256);
  uint32_t* topk_list = top_k > 0 ? (uint32_t*)(ctx->aux + status_bytes + 256 + lists_bytes) : nullptr;
  const bool bits_path = foot.r >= 1 && foot.r <= 3;  // box sizes up to 4: bit-plane kernel
  uint32_t foot_lo = 0, foot_hi = 0;
  if (bits_path)
    for (int k = 0; k < foot.n; ++k) {
      const int bit = (foot.dy[k] + foot.r) * 8 + foot.dx[k] + foot.r;
      if (bit < 32) foot_lo |= 1u << bit; else foot_hi |= 1u << (bit - 32);
    }
  const int HT = 32 + 2 * foot.r;
  const size_t smem = bits_path ? 0 : (size_t)HT * (HT + 1) * 5 + 16;
  const void* kernel = foot.r == 1 ? (const void*)nms_rounds_bits_kernel<1>
                     : foot.r == 2 ? (const void*)nms_rounds_bits_kernel<2>
                     : foot.r == 3 ? (const void*)nms_rounds_bits_kernel<3> : (const void*)nms_rounds_kernel;
  int per_sm = 0;
  SPN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, smem));
  SPN_REQUIRE(per_sm >= 1, "spn_box_nms_topk: cooperative kernel does not fit on an SM");
  int grid = per_sm * ctx->sm_count;
  if (n_tiles < grid) grid = (int)n_tiles;
  SpnProfScope prof(ctx, SPN_PROF_NMS, s);
  if (bits_path) {
    void* args[] = {(void*)&d_prob, (void*)&status, (void*)&foot_lo, (void*)&foot_hi, (void*)&B, (void*)&H, (void*)&W,
                    (void*)&min_prob, (void*)&tiles_x, (void*)&tiles_y, (void*)&pend, (void*)&lists};
    SPN_CUDA(cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(kThreads), args, smem, s));
  } else {
    void* args[] = {(void*)&d_prob, (void*)&status, (void*)&foot, (void*)&B, (void*)&H, (void*)&W, (void*)&min_prob,
                    (void*)&tiles_x, (void*)&tiles_y, (void*)&pend};
    SPN_CUDA(cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(kThreads), args, smem, s));
  }
  ctx->launches++;
  unsigned* sync = (unsigned*)(ctx->aux + status_bytes + 256 + lists_bytes + topk_bytes);
  unsigned* flags0 = sync, *flags1 = sync + (size_t)B * S, *tickets = sync + 2 * (size_t)B * S;
  int* list_count = (int*)(tickets + 8);
  SPN_CUDA(cudaMemsetAsync(sync, 0, sync_bytes, s));
  if (top_k > 0) {
    nms_emit_kernel<0><<<B * S, kThreads, 0, s>>>(d_prob, status, H, W, S, det_thresh, nullptr, nullptr, nullptr, nullptr, 0,
                                                  topk_list, list_count, flags0, tickets);
    SPN_CHECK_LAUNCH(ctx);
    nms_topk_select_kernel<<<B, kThreads, 0, s>>>(status, H * W, top_k, topk_list, list_count);
    SPN_CHECK_LAUNCH(ctx);
  }
  nms_emit_kernel<1><<<B * S, kThreads, 0, s>>>(d_prob, status, H, W, S, det_thresh, d_nms, d_pred, d_kp, d_kp_count, max_kp,
                                                nullptr, nullptr, flags1, tickets + 1);
  SPN_CHECK_LAUNCH(ctx);
  return SPN_OK;
}
