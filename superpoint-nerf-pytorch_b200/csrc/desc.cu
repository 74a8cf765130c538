// Descriptor upsampling: bicubic x grid (align_corners=False, A=-0.75, clamped taps) + L2 normalisation.
//
// Reference: models/model_utils/heads.py:65-66
//   desc = F.normalize(F.interpolate(desc_raw, scale_factor=grid, mode='bicubic', align_corners=False), p=2, dim=1)
// Dense version (API completeness; writes B*C*H*W fp32 = 315 MB at 480x640 and is HBM-write bound) and the
// sparse version that evaluates the identical expression only at keypoints, which is what the reference's
// consumers read (evaluations/descriptor_evaluation.py:67-69).
#include <math.h>

#include "spn_common.cuh"

namespace {

__device__ __forceinline__ void cubic_coeffs(float t, float* w) {
  const float A = -0.75f;
  const float x0 = t + 1.0f, x1 = t, x2 = 1.0f - t, x3 = 2.0f - t;
  w[0] = ((A * x0 - 5.0f * A) * x0 + 8.0f * A) * x0 - 4.0f * A;
  w[1] = ((A + 2.0f) * x1 - (A + 3.0f)) * x1 * x1 + 1.0f;
  w[2] = ((A + 2.0f) * x2 - (A + 3.0f)) * x2 * x2 + 1.0f;
  w[3] = ((A * x3 - 5.0f * A) * x3 + 8.0f * A) * x3 - 4.0f * A;
}

struct Taps {
  int iy[4], ix[4];
  float wy[4], wx[4];
};

__device__ __forceinline__ Taps make_taps(int y, int x, int Hc, int Wc, float inv_grid, bool bicubic) {
  Taps t;
  const float sy = ((float)y + 0.5f) * inv_grid - 0.5f, sx = ((float)x + 0.5f) * inv_grid - 0.5f;
  if (bicubic) {
    const float fy = floorf(sy), fx = floorf(sx);
    cubic_coeffs(sy - fy, t.wy);
    cubic_coeffs(sx - fx, t.wx);
    for (int k = 0; k < 4; ++k) {
      t.iy[k] = min(max((int)fy - 1 + k, 0), Hc - 1);
      t.ix[k] = min(max((int)fx - 1 + k, 0), Wc - 1);
    }
  } else {  // bilinear, align_corners=False convention (source index clamped at 0)
    const float cy = fmaxf(sy, 0.f), cx = fmaxf(sx, 0.f);
    const float fy = floorf(cy), fx = floorf(cx);
    t.wy[0] = 0.f; t.wy[1] = 1.f - (cy - fy); t.wy[2] = cy - fy; t.wy[3] = 0.f;
    t.wx[0] = 0.f; t.wx[1] = 1.f - (cx - fx); t.wx[2] = cx - fx; t.wx[3] = 0.f;
    for (int k = 0; k < 4; ++k) {
      t.iy[k] = min(max((int)fy - 1 + k, 0), Hc - 1);
      t.ix[k] = min(max((int)fx - 1 + k, 0), Wc - 1);
    }
  }
  return t;
}

__device__ __forceinline__ float interp_channel(const float* __restrict__ ch, const Taps& t, int Wc) {
  float acc = 0.f;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const float* row = ch + (size_t)t.iy[a] * Wc;
    float r = 0.f;
#pragma unroll
    for (int b = 0; b < 4; ++b) r = fmaf(__ldg(&row[t.ix[b]]), t.wx[b], r);
    acc = fmaf(r, t.wy[a], acc);
  }
  return acc;
}

// one CTA = 32 consecutive output pixels of one row x all C channels (staged in smem so the norm is computed
// once and the NCHW store is coalesced along x).
__global__ void __launch_bounds__(256)
dense_desc_kernel(const float* __restrict__ raw, int C, int Hc, int Wc, int grid, float* __restrict__ out) {
  extern __shared__ float val[];  // C x 33
  __shared__ float ssq[8][32];
  const int H = Hc * grid, W = Wc * grid;
  const int px = threadIdx.x & 31, sub = threadIdx.x >> 5;
  const int x = blockIdx.x * 32 + px, y = blockIdx.y, b = blockIdx.z;
  const float* rb = raw + (size_t)b * C * Hc * Wc;
  const bool ok = x < W;
  float s = 0.f;
  if (ok) {
    const Taps t = make_taps(y, x, Hc, Wc, 1.0f / (float)grid, true);
    for (int c = sub; c < C; c += 8) {
      const float v = interp_channel(rb + (size_t)c * Hc * Wc, t, Wc);
      val[c * 33 + px] = v;
      s = fmaf(v, v, s);
    }
  }
  ssq[sub][px] = s;
  __syncthreads();
  if (!ok) return;
  float tot = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) tot += ssq[k][px];
  const float inv = 1.0f / fmaxf(sqrtf(tot), 1e-12f);
  for (int c = sub; c < C; c += 8) out[(((size_t)b * C + c) * H + y) * W + x] = val[c * 33 + px] * inv;
}

// Separable version for grid >= 2 (the reference uses 8): the 32 output pixels of a warp need at most 4 + 31/grid + 1
// <= 20 source columns, so lanes 0..ncols-1 first interpolate their column vertically (4 loads + 4 FMAs), and every
// lane then combines four of those values horizontally through warp shuffles: 4 load instructions per channel
// instead of 16, and a quarter of the FMAs.  Same block shape / staging / normalisation as dense_desc_kernel.
__global__ void __launch_bounds__(256)
dense_desc_sep_kernel(const float* __restrict__ raw, int C, int Hc, int Wc, int grid, float* __restrict__ out) {
  extern __shared__ float val[];  // C x 33
  __shared__ float ssq[8][32];
  const int H = Hc * grid, W = Wc * grid;
  const int px = threadIdx.x & 31, sub = threadIdx.x >> 5;
  const int x0 = blockIdx.x * 32, x = x0 + px, y = blockIdx.y, b = blockIdx.z;
  const float* rb = raw + (size_t)b * C * Hc * Wc;
  const bool ok = x < W;
  const float inv_grid = 1.0f / (float)grid;
  // vertical taps: the same for the whole block
  const float sy = ((float)y + 0.5f) * inv_grid - 0.5f, fy = floorf(sy);
  float wy[4], wx[4];
  cubic_coeffs(sy - fy, wy);
  int iy[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) iy[k] = min(max((int)fy - 1 + k, 0), Hc - 1) * Wc;
  // horizontal taps of this lane, and the first source column any lane of the warp touches
  const float sx = ((float)x + 0.5f) * inv_grid - 0.5f, fx = floorf(sx);
  cubic_coeffs(sx - fx, wx);
  const int col0 = (int)floorf(((float)x0 + 0.5f) * inv_grid - 0.5f) - 1;
  const int src0 = (int)fx - 1 - col0;                  // lane that holds this pixel's first tap (0 <= src0, src0 + 3 < 32)
  const int mycol = min(max(col0 + px, 0), Wc - 1);     // column this lane interpolates vertically (clamped taps)
  float s = 0.f;
  for (int c = sub; c < C; c += 8) {
    const float* ch = rb + (size_t)c * Hc * Wc + mycol;
    float v = __ldg(ch + iy[0]) * wy[0];
    v = fmaf(__ldg(ch + iy[1]), wy[1], v);
    v = fmaf(__ldg(ch + iy[2]), wy[2], v);
    v = fmaf(__ldg(ch + iy[3]), wy[3], v);
    float r = __shfl_sync(0xffffffffu, v, src0) * wx[0];
    r = fmaf(__shfl_sync(0xffffffffu, v, src0 + 1), wx[1], r);
    r = fmaf(__shfl_sync(0xffffffffu, v, src0 + 2), wx[2], r);
    r = fmaf(__shfl_sync(0xffffffffu, v, src0 + 3), wx[3], r);
    val[c * 33 + px] = r;
    s = fmaf(r, r, s);
  }
  ssq[sub][px] = s;
  __syncthreads();
  if (!ok) return;
  float tot = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) tot += ssq[k][px];
  const float inv = 1.0f / fmaxf(sqrtf(tot), 1e-12f);
  for (int c = sub; c < C; c += 8) out[(((size_t)b * C + c) * H + y) * W + x] = val[c * 33 + px] * inv;
}

// one warp = one keypoint; lanes stride over channels.
__global__ void __launch_bounds__(256)
sparse_desc_kernel(const float* __restrict__ raw, int C, int Hc, int Wc, int grid, const int32_t* __restrict__ kp,
                   const int32_t* __restrict__ kp_count, int max_kp, int bicubic, float* __restrict__ out) {
  const int b = blockIdx.y;
  const int k = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int n = min(__ldg(&kp_count[b]), max_kp);
  if (k >= max_kp) return;
  if (k >= n) {  // slots beyond the keypoint count are zero
    for (int c = lane; c < C; c += 32) out[((size_t)b * max_kp + k) * C + c] = 0.f;
    return;
  }
  const int y = __ldg(&kp[((size_t)b * max_kp + k) * 2]), x = __ldg(&kp[((size_t)b * max_kp + k) * 2 + 1]);
  const Taps t = make_taps(y, x, Hc, Wc, 1.0f / (float)grid, bicubic != 0);
  const float* rb = raw + (size_t)b * C * Hc * Wc;
  float* o = out + ((size_t)b * max_kp + k) * C;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float v = interp_channel(rb + (size_t)c * Hc * Wc, t, Wc);
    o[c] = v;
    s = fmaf(v, v, s);
  }
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  const float inv = 1.0f / fmaxf(sqrtf(s), 1e-12f);
  for (int c = lane; c < C; c += 32) o[c] *= inv;  // same lane wrote o[c]
}

}  // namespace

extern "C" int spn_dense_descriptors(spn_ctx* ctx, const float* d_desc_raw, int B, int C, int Hc, int Wc, int grid,
                                     float* d_desc, spn_stream stream) {
  SPN_REQUIRE(ctx && d_desc_raw && d_desc, "spn_dense_descriptors: null pointer");
  SpnDeviceGuard guard(ctx->device);
  SPN_REQUIRE(B > 0 && B <= 65535 && C > 0 && Hc > 0 && Wc > 0 && grid > 0, "spn_dense_descriptors: bad shape");
  const int H = Hc * grid, W = Wc * grid;
  SPN_REQUIRE(H <= 65535, "spn_dense_descriptors: image too tall");
  const size_t smem = (size_t)C * 33 * sizeof(float);
  SPN_REQUIRE(smem <= 200 * 1024, "spn_dense_descriptors: too many channels");
  // grid >= 2: a warp's 32 pixels touch at most 4 + 31/grid + 1 <= 21 source columns -> separable kernel
  auto kern = grid >= 2 ? dense_desc_sep_kernel : dense_desc_kernel;
  if (smem > 48 * 1024) SPN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 g(spn_cdiv(W, 32), H, B);
  SpnProfScope prof(ctx, SPN_PROF_DESC, (cudaStream_t)stream);
  kern<<<g, 256, smem, (cudaStream_t)stream>>>(d_desc_raw, C, Hc, Wc, grid, d_desc);
  SPN_CHECK_LAUNCH(ctx);
  return SPN_OK;
}

extern "C" int spn_sample_descriptors(spn_ctx* ctx, const float* d_desc_raw, int B, int C, int Hc, int Wc, int grid,
                                      const int32_t* d_kp, const int32_t* d_kp_count, int max_kp, int interp,
                                      float* d_out, spn_stream stream) {
  SPN_REQUIRE(ctx && d_desc_raw && d_kp && d_kp_count && d_out, "spn_sample_descriptors: null pointer");
  SpnDeviceGuard guard(ctx->device);
  SPN_REQUIRE(B > 0 && B <= 65535 && C > 0 && Hc > 0 && Wc > 0 && grid > 0 && max_kp > 0, "spn_sample_descriptors: bad shape");
  SPN_REQUIRE(interp == 0 || interp == 1, "spn_sample_descriptors: interp must be 0 (bicubic) or 1 (bilinear)");
  dim3 g(spn_cdiv(max_kp, 8), B);
  sparse_desc_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(d_desc_raw, C, Hc, Wc, grid, d_kp, d_kp_count, max_kp,
                                                          interp == 0, d_out);
  SPN_CHECK_LAUNCH(ctx);
  return SPN_OK;
}
