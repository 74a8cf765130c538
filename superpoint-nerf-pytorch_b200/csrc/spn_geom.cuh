// Homography / sampling helpers shared by geometry.cu and front_tc.cu (kornia 0.7.0 warp_perspective semantics,
// see oracle/kornia_shim.py for the restated algorithm).
#pragma once
#include <cuda_runtime.h>

namespace spngeom {

// Correctly rounded 1/z for z well inside the normal range (2^-100 < |z| < 2^100): the MUFU.RCP + one Newton step the
// compiler emits for 1.0f / z, without its denormal/overflow guard and the branch that comes with it.
__device__ __forceinline__ float rcp_normal(float z) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(z));
  const float e = fmaf(-z, r, 1.0f);
  return fmaf(r, e, r);
}

// src = M p  (p = (x, y, 1)) with kornia's homogeneous divide: scale = |z| > 1e-8 ? 1/(z + 1e-8) : 1
__device__ __forceinline__ void apply_h(const float* __restrict__ m, float x, float y, float& sx, float& sy) {
  const float nx = fmaf(m[0], x, fmaf(m[1], y, m[2]));
  const float ny = fmaf(m[3], x, fmaf(m[4], y, m[5]));
  const float z = fmaf(m[6], x, fmaf(m[7], y, m[8]));
  // |z| > 1e-8 puts |z + 1e-8| in [~1e-15, |z| + 1e-8]: normal range for any finite homography
  const float sc = fabsf(z) > 1e-8f ? rcp_normal(z + 1e-8f) : 1.0f;
  sx = nx * sc;
  sy = ny * sc;
}

// grid_sample(mode='nearest', padding 'zeros') of an all-ones image: 1 iff the rounded coordinate is inside.
__device__ __forceinline__ int inside_nearest(float sx, float sy, int H, int W) {
  const float rx = rintf(sx), ry = rintf(sy);
  return (rx >= 0.f && rx <= (float)(W - 1) && ry >= 0.f && ry <= (float)(H - 1)) ? 1 : 0;
}

// grid_sample(mode='bilinear', padding 'zeros', align_corners=True) of one channel.
__device__ __forceinline__ float bilinear_zero(const float* __restrict__ img, float sx, float sy, int H, int W) {
  if (!(sx > -1.f && sx < (float)W && sy > -1.f && sy < (float)H)) return 0.f;  // also rejects NaN/inf
  const float fx = floorf(sx), fy = floorf(sy);
  const int x0 = (int)fx, y0 = (int)fy;
  const float wx1 = sx - fx, wy1 = sy - fy;
  const float wx0 = (fx + 1.f) - sx, wy0 = (fy + 1.f) - sy;
  const bool xin0 = x0 >= 0 && x0 < W, xin1 = x0 + 1 >= 0 && x0 + 1 < W;
  const bool yin0 = y0 >= 0 && y0 < H, yin1 = y0 + 1 >= 0 && y0 + 1 < H;
  float v = 0.f;
  if (yin0 && xin0) v = fmaf(__ldg(&img[(size_t)y0 * W + x0]), wx0 * wy0, v);
  if (yin0 && xin1) v = fmaf(__ldg(&img[(size_t)y0 * W + x0 + 1]), wx1 * wy0, v);
  if (yin1 && xin0) v = fmaf(__ldg(&img[(size_t)(y0 + 1) * W + x0]), wx0 * wy1, v);
  if (yin1 && xin1) v = fmaf(__ldg(&img[(size_t)(y0 + 1) * W + x0 + 1]), wx1 * wy1, v);
  return v;
}

// Bilinear sample whose four taps are known to be inside the image (0 <= floor(sx), floor(sx) + 1 <= W - 1, same in y):
// no bounds tests, interpolation in lerp form.  `plane` is a 32-bit element offset into `img`.
__device__ __forceinline__ float bilinear_inside(const float* __restrict__ img, int plane, float sx, float sy, int W) {
  const float fx = floorf(sx), fy = floorf(sy);
  const float ax = sx - fx, ay = sy - fy;
  const float* p = img + (plane + (int)fy * W + (int)fx);
  const float v00 = __ldg(p), v01 = __ldg(p + 1), v10 = __ldg(p + W), v11 = __ldg(p + W + 1);
  const float top = fmaf(ax, v01 - v00, v00), bot = fmaf(ax, v11 - v10, v10);
  return fmaf(ay, bot - top, top);
}

// Branch-free variant for latency-critical producers: the four taps are always loaded from clamped addresses and a
// tap outside the image gets weight 0 (same value as bilinear_zero, no divergent control flow).
__device__ __forceinline__ float bilinear_zero_nb(const float* __restrict__ img, float sx, float sy, int H, int W) {
  const bool ok = sx > -1.f && sx < (float)W && sy > -1.f && sy < (float)H;  // also rejects NaN/inf
  const float cx = ok ? sx : 0.f, cy = ok ? sy : 0.f;
  const float fx = floorf(cx), fy = floorf(cy);
  const int x0 = (int)fx, y0 = (int)fy;
  const float wx1 = cx - fx, wy1 = cy - fy;
  const float wx0 = (fx + 1.f) - cx, wy0 = (fy + 1.f) - cy;
  const float mx0 = (x0 >= 0 && ok) ? wx0 : 0.f, mx1 = (x0 + 1 < W && ok) ? wx1 : 0.f;
  const float my0 = (y0 >= 0) ? wy0 : 0.f, my1 = (y0 + 1 < H) ? wy1 : 0.f;
  const int xa = max(x0, 0), xb = min(x0 + 1, W - 1), ya = max(y0, 0), yb = min(y0 + 1, H - 1);
  // 32-bit element offsets (H * W < 2^31 is checked by every caller's shape limits): one 64-bit base + four offsets
  const unsigned ra = (unsigned)(ya * W), rb = (unsigned)(yb * W);
  const float v00 = __ldg(img + (ra + (unsigned)xa)), v01 = __ldg(img + (ra + (unsigned)xb));
  const float v10 = __ldg(img + (rb + (unsigned)xa)), v11 = __ldg(img + (rb + (unsigned)xb));
  float v = v00 * (mx0 * my0);
  v = fmaf(v01, mx1 * my0, v);
  v = fmaf(v10, mx0 * my1, v);
  v = fmaf(v11, mx1 * my1, v);
  return v;
}

// ---- kornia 0.7.0 warp_perspective coordinate chain, bit for bit (oracle/kornia_shim.py; the arithmetic below was
// fitted against torch-CPU fp32 on the same matrices: tools/kornia_chain_fit.py) -------------------------------------
//   xn = (x / (W-1) - 0.5) * 2, yn likewise                      (create_meshgrid; tabulated on the host, KGrid)
//   q  = [xn, yn, 1] @ Ainv^T  as  (a0*xn  ->  fma(a1, yn, .)  ->  + a2)   (torch.bmm, K = 3)
//   g  = q.xy * (|q.z| > 1e-8 ? 1 / (q.z + 1e-8) : 1)              (transform_points)
//   s  = (g + 1) * (size-1)/2                                      (grid_sample unnormalize, align_corners=True)
// `a` = torch.inverse(normalize_homography(M)) in normalised coordinates.  Every operation is a single IEEE fp32
// operation in the reference's order (intrinsics keep the compiler from contracting or re-associating them).
struct KGrid {
  const float* xs;  // [W] normalised x of every column
  const float* ys;  // [H]
  float hw, hh;     // (W-1)/2, (H-1)/2
};

__device__ __forceinline__ void kornia_src(const float* __restrict__ a, float xn, float yn, float hw, float hh, float& sx,
                                           float& sy) {
  const float qx = __fadd_rn(fmaf(a[1], yn, __fmul_rn(a[0], xn)), a[2]);
  const float qy = __fadd_rn(fmaf(a[4], yn, __fmul_rn(a[3], xn)), a[5]);
  const float qz = __fadd_rn(fmaf(a[7], yn, __fmul_rn(a[6], xn)), a[8]);
  const float sc = fabsf(qz) > 1e-8f ? __frcp_rn(__fadd_rn(qz, 1e-8f)) : 1.0f;
  sx = __fmul_rn(__fadd_rn(__fmul_rn(sc, qx), 1.0f), hw);
  sy = __fmul_rn(__fadd_rn(__fmul_rn(sc, qy), 1.0f), hh);
}

// F.grid_sample(mode='bilinear', padding_mode='zeros', align_corners=True) of one channel with ATen's CPU arithmetic:
// w = x - floor(x), e = 1 - w, n = y - floor(y), s = 1 - n; weights s*e, s*w, n*e, n*w; taps outside the image read 0;
// value = fma(se_v, se, fma(sw_v, sw, fma(ne_v, ne, nw_v * nw))).
__device__ __forceinline__ float bilinear_combine(float nwv, float nev, float swv, float sev, float sx, float sy, float fx,
                                                  float fy) {
  const float w = __fsub_rn(sx, fx), e = __fsub_rn(1.0f, w), n = __fsub_rn(sy, fy), s = __fsub_rn(1.0f, n);
  const float nw = __fmul_rn(s, e), ne = __fmul_rn(s, w), sw = __fmul_rn(n, e), se = __fmul_rn(n, w);
  return fmaf(sev, se, fmaf(swv, sw, fmaf(nev, ne, __fmul_rn(nwv, nw))));
}

__device__ __forceinline__ float bilinear_zero_exact(const float* __restrict__ img, float sx, float sy, int H, int W) {
  if (!(sx > -1.f && sx < (float)W && sy > -1.f && sy < (float)H)) return 0.f;  // every tap outside (also NaN/inf)
  const float fx = floorf(sx), fy = floorf(sy);
  const int x0 = (int)fx, y0 = (int)fy;
  const bool xin0 = x0 >= 0, xin1 = x0 + 1 < W, yin0 = y0 >= 0, yin1 = y0 + 1 < H;
  const float nwv = (yin0 && xin0) ? __ldg(&img[(size_t)y0 * W + x0]) : 0.f;
  const float nev = (yin0 && xin1) ? __ldg(&img[(size_t)y0 * W + x0 + 1]) : 0.f;
  const float swv = (yin1 && xin0) ? __ldg(&img[(size_t)(y0 + 1) * W + x0]) : 0.f;
  const float sev = (yin1 && xin1) ? __ldg(&img[(size_t)(y0 + 1) * W + x0 + 1]) : 0.f;
  return bilinear_combine(nwv, nev, swv, sev, sx, sy, fx, fy);
}

// Exact t / d for 0 <= t, t * d < 2^40, with m = ceil(2^40 / d) computed on the host (fast_div_magic).
__device__ __forceinline__ int fast_div(int t, unsigned long long m) {
  return (int)(((unsigned long long)(unsigned)t * m) >> 40);
}
inline unsigned long long fast_div_magic(int d) { return ((1ull << 40) + (unsigned long long)d - 1) / (unsigned long long)d; }

}  // namespace spngeom
