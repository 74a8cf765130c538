// On-GPU evaluation of the exports: the keypoint selection, repeatability counting and
// mutual-nearest-neighbour descriptor matching of the reference's evaluation scripts
//   evaluations/detector_evaluation.py:145-238   compute_repeatability
//   evaluations/descriptor_evaluation.py:17-90   keep_shared_points, compute_homography (matching part)
// The RANSAC homography fit that follows the matching stays with cv2 on the host, as in the reference.
//
//   select_keypoints_kernel   np.where(prob > 0) + warp_keypoints / keep_true_keypoints / filter_keypoints (fp64, as
//                             numpy) + select_k_best (ascending by probability, last k): one CTA per map, candidates
//                             compacted into shared memory and bitonic-sorted as 64-bit (prob bits, pixel index) keys
//   repeat_count_kernel       N1 x N2 Euclidean distances in fp64, nearest-neighbour counts within the threshold
//   nn_match_tc_kernel        cv2.BFMatcher(NORM_L2): the N1 x N2 x 256 distance matrix as a tcgen05 GEMM.  The fp32
//                             descriptors are split into fp16 (hi, lo) halves on the way into shared memory and the dot
//                             product is accumulated as hi*hi + hi*lo + lo*hi in fp32 TMEM (error ~2^-22, i.e. fp32
//                             quality on L2-normalised descriptors); the epilogue forms |a|^2 + |b|^2 - 2 a.b per
//                             accumulator row and publishes the row-wise arg-min with one 64-bit atomicMin
//                             (distance bits, index): ties go to the lower index like cv2.
//   cross_check_kernel        crossCheck=True: keep (i, j) iff j is i's nearest and i is j's nearest
#include <math.h>

#include "spn_common.cuh"
#include "tc_ptx.cuh"

namespace {

using namespace tcptx;

// ------------------------------------------------------------------------------------------------ keypoint selection
constexpr int kSelCap = 16384;   // candidates per map held in shared memory (128 KB of 64-bit keys)
constexpr int kSelThreads = 1024;

struct SelParams {
  const float* prob;     // [B][H][W]
  const double* warp;    // [B][9] or null: homography applied to (x = col, y = row, 1)
  int H, W;
  int bound_h, bound_w;  // the warped point must satisfy 0 <= row' < bound_h, 0 <= col' < bound_w
  int emit_warped;       // 0: output the original integer (row, col); 1: output the warped (row', col')
  int keep_k;
  double* pts;           // [B][keep_k][2]
  float* score;          // [B][keep_k]
  int* count;            // [B] points written; [B + b] = candidates before select_k_best (diagnostic / overflow check)
};

__device__ __forceinline__ void warp_point(const double* h, double x, double y, double& wx, double& wy) {
  const double z = h[6] * x + h[7] * y + h[8];
  wx = (h[0] * x + h[1] * y + h[2]) / z;
  wy = (h[3] * x + h[4] * y + h[5]) / z;
}

__global__ void __launch_bounds__(kSelThreads, 1) select_keypoints_kernel(const SelParams p) {
  extern __shared__ unsigned long long keys[];  // kSelCap
  __shared__ int n_s, warp_tot[kSelThreads / 32];
  __shared__ double hm[9];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const float* pr = p.prob + (size_t)b * p.H * p.W;
  if (tid == 0) n_s = 0;
  if (tid < 9 && p.warp) hm[tid] = p.warp[(size_t)b * 9 + tid];
  __syncthreads();
  const int npx = p.H * p.W;
  // ordered compaction (row-major, like np.where): chunks of kSelThreads pixels, block-wide exclusive scan of the flags
  for (int base = 0; base < npx; base += kSelThreads) {
    const int i = base + tid;
    bool ok = false;
    float v = 0.f;
    if (i < npx) {
      v = pr[i];
      ok = v > 0.f;
      if (ok && p.warp) {
        const int y = i / p.W, x = i - y * p.W;
        double wx, wy;
        warp_point(hm, (double)x, (double)y, wx, wy);
        ok = wy >= 0.0 && wy < (double)p.bound_h && wx >= 0.0 && wx < (double)p.bound_w;
      }
    }
    const unsigned m = __ballot_sync(0xffffffffu, ok);
    if (lane == 0) warp_tot[wid] = __popc(m);
    __syncthreads();
    int off = n_s;
    for (int w = 0; w < wid; ++w) off += warp_tot[w];
    const int pos = off + __popc(m & ((1u << lane) - 1u));
    if (ok && pos < kSelCap) keys[pos] = ((unsigned long long)__float_as_uint(v) << 32) | (unsigned)i;
    __syncthreads();
    if (tid == 0) {
      int t = 0;
      for (int w = 0; w < kSelThreads / 32; ++w) t += warp_tot[w];
      n_s += t;
    }
    __syncthreads();
  }
  const int n_all = n_s;
  const int n = min(n_all, kSelCap);
  int n2 = 1;
  while (n2 < n) n2 <<= 1;
  for (int i = n + tid; i < n2; i += kSelThreads) keys[i] = ~0ull;
  __syncthreads();
  // bitonic sort, ascending (positive floats order like their bit patterns; ties by pixel index)
  for (int k = 2; k <= n2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < n2; i += kSelThreads) {
        const int l = i ^ j;
        if (l > i) {
          const unsigned long long a = keys[i], c = keys[l];
          const bool up = (i & k) == 0;
          if ((a > c) == up) { keys[i] = c; keys[l] = a; }
        }
      }
      __syncthreads();
    }
  const int kk = min(p.keep_k, n);
  for (int i = tid; i < kk; i += kSelThreads) {  // select_k_best: ascending by probability, the last kk
    const unsigned long long key = keys[n - kk + i];
    const int px = (int)(key & 0xffffffffu);
    const int y = px / p.W, x = px - y * p.W;
    double oy = (double)y, ox = (double)x;
    if (p.warp && p.emit_warped) warp_point(hm, (double)x, (double)y, ox, oy);
    p.pts[((size_t)b * p.keep_k + i) * 2] = oy;
    p.pts[((size_t)b * p.keep_k + i) * 2 + 1] = ox;
    p.score[(size_t)b * p.keep_k + i] = __uint_as_float((unsigned)(key >> 32));
  }
  if (tid == 0) {
    p.count[b] = kk;
    p.count[gridDim.x + b] = n_all;
  }
}

// ------------------------------------------------------------------------------------------------ repeatability counts
// out[b] = {N1, N2, count1, count2}: count1 = #points of set 1 whose nearest point of set 2 is within thresh, and
// vice versa (detector_evaluation.py:214-233).
__global__ void __launch_bounds__(256) repeat_count_kernel(const double* __restrict__ p1, const int* __restrict__ n1,
                                                           const double* __restrict__ p2, const int* __restrict__ n2, int cap,
                                                           double thresh, int* __restrict__ out) {
  __shared__ int c1, c2;
  const int b = blockIdx.x;
  const int N1 = n1[b], N2 = n2[b];
  if (threadIdx.x == 0) { c1 = 0; c2 = 0; }
  __syncthreads();
  const double* a = p1 + (size_t)b * cap * 2;
  const double* c = p2 + (size_t)b * cap * 2;
  int l1 = 0, l2 = 0;
  if (N2 > 0)
    for (int i = threadIdx.x; i < N1; i += blockDim.x) {
      double best = 1e300;
      for (int j = 0; j < N2; ++j) {
        const double dy = a[2 * i] - c[2 * j], dx = a[2 * i + 1] - c[2 * j + 1];
        best = fmin(best, sqrt(dy * dy + dx * dx));
      }
      l1 += best <= thresh;
    }
  if (N1 > 0)
    for (int j = threadIdx.x; j < N2; j += blockDim.x) {
      double best = 1e300;
      for (int i = 0; i < N1; ++i) {
        const double dy = a[2 * i] - c[2 * j], dx = a[2 * i + 1] - c[2 * j + 1];
        best = fmin(best, sqrt(dy * dy + dx * dx));
      }
      l2 += best <= thresh;
    }
  atomicAdd(&c1, l1);
  atomicAdd(&c2, l2);
  __syncthreads();
  if (threadIdx.x == 0) {
    out[4 * b] = N1; out[4 * b + 1] = N2; out[4 * b + 2] = c1; out[4 * b + 3] = c2;
  }
}

// ------------------------------------------------------------------------------------------------ NN matching on tcgen05
constexpr int kMT = 128;            // query rows per CTA (TMEM lanes)
constexpr int kNT = 128;            // train rows per CTA (accumulator columns)
constexpr int kKC = 64;             // K elements staged per step (4 MMA K-steps of 16)
constexpr int kPlane = kMT * 16;    // bytes of one 8-element K chunk for 128 rows: [row][8 halfs]
constexpr int kOpBytes = (kKC / 8) * kPlane;   // 16 KB: one operand (hi or lo) of one K step

struct MatchParams {
  const float* q;        // [B][qcap][C] query descriptors
  const float* t;        // [B][tcap][C] train descriptors
  const int* nq;         // [B]
  const int* nt;         // [B]
  int qcap, tcap, C;
  unsigned long long* best;   // [B][qcap]: (distance^2 bits << 32) | train index, initialised to ~0
};

// rows [row0, row0 + 128) x K [k0, k0 + 64) of src (row-major fp32, `n` valid rows) -> hi / lo fp16 operand images
// [chunk 8][row 128][8]; returns this thread's partial |row|^2 (thread = row).
__device__ __forceinline__ float stage_operand(const float* __restrict__ src, int n, int C, int row0, int k0, uint8_t* hi, uint8_t* lo) {
  const int r = threadIdx.x;
  const bool live = row0 + r < n;
  const float4* g = reinterpret_cast<const float4*>(src + (size_t)(row0 + r) * C + k0);
  float ss = 0.f;
#pragma unroll
  for (int c = 0; c < kKC / 8; ++c) {
    float v[8];
    if (live) {
      const float4 a = __ldg(g + 2 * c), b = __ldg(g + 2 * c + 1);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = 0.f;
    }
    uint32_t h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const __half h0 = __float2half_rn(v[2 * e]), h1 = __float2half_rn(v[2 * e + 1]);
      const __half l0 = __float2half_rn(v[2 * e] - __half2float(h0)), l1 = __float2half_rn(v[2 * e + 1] - __half2float(h1));
      h[e] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
      l[e] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
      ss = fmaf(v[2 * e], v[2 * e], ss);
      ss = fmaf(v[2 * e + 1], v[2 * e + 1], ss);
    }
    *reinterpret_cast<uint4*>(hi + (size_t)c * kPlane + (size_t)r * 16) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(lo + (size_t)c * kPlane + (size_t)r * 16) = make_uint4(l[0], l[1], l[2], l[3]);
  }
  return ss;
}

__global__ void __launch_bounds__(128, 1) nn_match_tc_kernel(const MatchParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];   // A hi, A lo, B hi, B lo: 4 x 16 KB
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ float nb_s[kNT];
  const int b = blockIdx.z;
  const int q0 = blockIdx.y * kMT, t0 = blockIdx.x * kNT;
  const int nq = p.nq[b], nt = p.nt[b];
  if (q0 >= nq || t0 >= nt) return;   // uniform per CTA
  const int warp = threadIdx.x >> 5;
  uint8_t* a_hi = smem;
  uint8_t* a_lo = smem + kOpBytes;
  uint8_t* b_hi = smem + 2 * kOpBytes;
  uint8_t* b_lo = smem + 3 * kOpBytes;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const float* qb = p.q + (size_t)b * p.qcap * p.C;
  const float* tb = p.t + (size_t)b * p.tcap * p.C;
  // instruction descriptor: fp16 x fp16 -> fp32, M = 128, N = 128, both operands K-major
  const uint32_t idesc = (1u << 4) | ((uint32_t)(kNT >> 3) << 17) | ((uint32_t)(kMT >> 4) << 24);
  const uint32_t d_hi = (128u >> 4) | (1u << 14);                    // SBO: 8 rows x 16 B
  const uint32_t lbo = ((uint32_t)kPlane >> 4) << 16;               // LBO: next 8-element K chunk
  float na = 0.f, nb = 0.f;
  uint32_t phase = 0;
  for (int k0 = 0; k0 < p.C; k0 += kKC) {
    na += stage_operand(qb, nq, p.C, q0, k0, a_hi, a_lo);
    nb += stage_operand(tb, nt, p.C, t0, k0, b_hi, b_lo);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // generic-proxy stores -> tensor-core reads
    __syncthreads();
    if (threadIdx.x == 0) {
      tc_fence_after();
#pragma unroll
      for (int kk = 0; kk < kKC / 16; ++kk) {
        const uint32_t off = ((uint32_t)kk * 2 * kPlane) >> 4;
        const uint32_t ah = ((smem_u32(a_hi) >> 4) + off) | lbo, al = ((smem_u32(a_lo) >> 4) + off) | lbo;
        const uint32_t bh = ((smem_u32(b_hi) >> 4) + off) | lbo, bl = ((smem_u32(b_lo) >> 4) + off) | lbo;
        umma_f16_2w(tmem_base, ah, d_hi, bh, d_hi, idesc, (k0 | kk) ? 1u : 0u);   // hi . hi
        umma_f16_2w(tmem_base, ah, d_hi, bl, d_hi, idesc, 1u);                    // hi . lo
        umma_f16_2w(tmem_base, al, d_hi, bh, d_hi, idesc, 1u);                    // lo . hi
      }
      umma_commit(&bar);
    }
    mbar_wait(&bar, phase);   // the MMAs have read this step's operands: the buffers may be refilled
    phase ^= 1;
  }
  tc_fence_after();
  nb_s[threadIdx.x] = nb;
  __syncthreads();
  // epilogue: thread = query row = TMEM lane; arg-min of |a|^2 + |b|^2 - 2 a.b over this CTA's 128 train rows
  float best = INFINITY;
  int best_j = 0;
#pragma unroll 1
  for (int c0 = 0; c0 < kNT; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      const int j = t0 + c0 + c;
      const float d2 = fmaxf(fmaf(-2.0f, __uint_as_float(v[c]), na + nb_s[c0 + c]), 0.f);
      if (j < nt && d2 < best) { best = d2; best_j = j; }
    }
  }
  if (q0 + (int)threadIdx.x < nq)
    atomicMin(&p.best[(size_t)b * p.qcap + q0 + threadIdx.x], ((unsigned long long)__float_as_uint(best) << 32) | (unsigned)best_j);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128));
  }
}

__global__ void fill_u64_kernel(unsigned long long* p, size_t n, unsigned long long v) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

__global__ void cross_check_kernel(const unsigned long long* __restrict__ bq, const unsigned long long* __restrict__ bt,
                                   const int* __restrict__ nq, const int* __restrict__ nt, int qcap, int tcap,
                                   int* __restrict__ match, float* __restrict__ dist) {
  const int b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= qcap) return;
  int m = -1;
  float d = 0.f;
  if (i < nq[b] && nt[b] > 0) {
    const unsigned long long k = bq[(size_t)b * qcap + i];
    const int j = (int)(k & 0xffffffffu);
    if ((int)(bt[(size_t)b * tcap + j] & 0xffffffffu) == i) {
      m = j;
      d = sqrtf(__uint_as_float((unsigned)(k >> 32)));
    }
  }
  match[(size_t)b * qcap + i] = m;
  dist[(size_t)b * qcap + i] = d;
}

}  // namespace

extern "C" int spn_select_keypoints(spn_ctx* ctx, const float* d_prob, int B, int H, int W, const double* h_warp, int bound_h,
                                    int bound_w, int emit_warped, int keep_k, double* d_pts, float* d_score, int32_t* d_count,
                                    spn_stream stream) {
  SPN_REQUIRE(ctx && d_prob && d_pts && d_score && d_count, "spn_select_keypoints: null pointer");
  SPN_REQUIRE(B > 0 && B <= 65535 && H > 0 && W > 0 && keep_k > 0 && keep_k <= kSelCap, "spn_select_keypoints: bad shape / keep_k (<= %d)", kSelCap);
  SpnDeviceGuard guard(ctx->device);
  cudaStream_t s = (cudaStream_t)stream;
  SelParams p;
  memset(&p, 0, sizeof(p));
  p.prob = d_prob; p.H = H; p.W = W; p.bound_h = bound_h; p.bound_w = bound_w; p.emit_warped = emit_warped; p.keep_k = keep_k;
  p.pts = d_pts; p.score = d_score; p.count = d_count;
  if (h_warp) {
    int rc = spn_ensure_aux(ctx, (size_t)B * 9 * sizeof(double), s);
    if (rc) return rc;
    SPN_CUDA(cudaMemcpyAsync(ctx->aux, h_warp, (size_t)B * 9 * sizeof(double), cudaMemcpyHostToDevice, s));
    p.warp = (const double*)ctx->aux;
  }
  const size_t dyn = (size_t)kSelCap * sizeof(unsigned long long);
  SPN_CUDA(cudaFuncSetAttribute(select_keypoints_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
  select_keypoints_kernel<<<B, kSelThreads, dyn, s>>>(p);
  SPN_CHECK_LAUNCH(ctx);
  return SPN_OK;
}

extern "C" int spn_repeatability_counts(spn_ctx* ctx, const double* d_pts1, const int32_t* d_n1, const double* d_pts2,
                                        const int32_t* d_n2, int B, int cap, double thresh, int32_t* d_out, spn_stream stream) {
  SPN_REQUIRE(ctx && d_pts1 && d_n1 && d_pts2 && d_n2 && d_out, "spn_repeatability_counts: null pointer");
  SPN_REQUIRE(B > 0 && B <= 65535 && cap > 0, "spn_repeatability_counts: bad shape");
  SpnDeviceGuard guard(ctx->device);
  repeat_count_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(d_pts1, d_n1, d_pts2, d_n2, cap, thresh, d_out);
  SPN_CHECK_LAUNCH(ctx);
  return SPN_OK;
}

extern "C" int spn_mutual_nn_match(spn_ctx* ctx, const float* d_desc1, const int32_t* d_n1, const float* d_desc2,
                                   const int32_t* d_n2, int B, int cap1, int cap2, int C, int32_t* d_match, float* d_dist,
                                   spn_stream stream) {
  SPN_REQUIRE(ctx && d_desc1 && d_n1 && d_desc2 && d_n2 && d_match && d_dist, "spn_mutual_nn_match: null pointer");
  SPN_REQUIRE(B > 0 && B <= 65535 && cap1 > 0 && cap2 > 0 && C > 0 && C % kKC == 0, "spn_mutual_nn_match: bad shape (C must be a multiple of %d)", kKC);
  SpnDeviceGuard guard(ctx->device);
  cudaStream_t s = (cudaStream_t)stream;
  const size_t n1 = (size_t)B * cap1, n2 = (size_t)B * cap2;
  int rc = spn_ensure_aux(ctx, (n1 + n2) * sizeof(unsigned long long), s);
  if (rc) return rc;
  unsigned long long* bq = (unsigned long long*)ctx->aux;
  unsigned long long* bt = bq + n1;
  fill_u64_kernel<<<(unsigned)((n1 + n2 + 255) / 256), 256, 0, s>>>(bq, n1 + n2, ~0ull);
  SPN_CHECK_LAUNCH(ctx);
  const size_t dyn = 4 * (size_t)kOpBytes + 1024;
  SPN_CUDA(cudaFuncSetAttribute(nn_match_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
  MatchParams p;
  p.q = d_desc1; p.t = d_desc2; p.nq = d_n1; p.nt = d_n2; p.qcap = cap1; p.tcap = cap2; p.C = C; p.best = bq;
  nn_match_tc_kernel<<<dim3(spn_cdiv(cap2, kNT), spn_cdiv(cap1, kMT), B), 128, dyn, s>>>(p);   // nearest train row of every query row
  SPN_CHECK_LAUNCH(ctx);
  p.q = d_desc2; p.t = d_desc1; p.nq = d_n2; p.nt = d_n1; p.qcap = cap2; p.tcap = cap1; p.best = bt;
  nn_match_tc_kernel<<<dim3(spn_cdiv(cap1, kNT), spn_cdiv(cap2, kMT), B), 128, dyn, s>>>(p);   // and the other direction
  SPN_CHECK_LAUNCH(ctx);
  cross_check_kernel<<<dim3(spn_cdiv(cap1, 128), B), 128, 0, s>>>(bq, bt, d_n1, d_n2, cap1, cap2, d_match, d_dist);
  SPN_CHECK_LAUNCH(ctx);
  return SPN_OK;
}
