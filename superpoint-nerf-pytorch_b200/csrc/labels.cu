// Train-time label building that shares the hot path's data: the detector-loss labels of
// utils/losses.py:13-27 of the reference, one thread per 8x8 cell.
//   labels = argmax_c( cat([2 * pixel_unshuffle(kpts_heatmap, 8), 1]) + noise ),  noise ~ U(0, 0.1)   (random tie break)
//   valid  = prod_c( pixel_unshuffle(valid_mask, 8) )
// With a caller-supplied noise tensor the labels are bit-identical to torch.argmax on the same noise (first maximum
// wins); without one a counter-based generator keyed by (seed, element index) draws the tie-break noise.
#include "spn_common.cuh"

namespace {

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

__global__ void __launch_bounds__(128) detector_labels_kernel(const int32_t* __restrict__ kmap, const int32_t* __restrict__ valid,
                                                              const float* __restrict__ noise, uint64_t seed, int B, int Hc, int Wc,
                                                              int64_t* __restrict__ labels, float* __restrict__ valid_cells) {
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  const int n_cells = B * Hc * Wc;
  if (cell >= n_cells) return;
  const int b = cell / (Hc * Wc), r = cell - b * Hc * Wc;
  const int cy = r / Wc, cx = r - cy * Wc;
  const int W = Wc * 8;
  const size_t base = ((size_t)b * Hc * 8 + cy * 8) * W + cx * 8;
  float best = -1.f;
  int best_c = 0;
  int ok = 1;
  for (int c = 0; c < 65; ++c) {
    float v = 1.0f;   // dustbin
    if (c < 64) {
      const size_t o = base + (size_t)(c >> 3) * W + (c & 7);
      v = 2.0f * (float)kmap[o];
      if (valid) ok &= valid[o] != 0;
    }
    float nz;
    if (noise) {
      nz = noise[((size_t)b * 65 + c) * Hc * Wc + r];
    } else {
      const uint64_t h = mix64(mix64(seed) ^ (((uint64_t)cell * 65 + c) + 1));
      nz = 0.1f * (float)((h >> 40) * (1.0 / 16777216.0));
    }
    v += nz;
    if (v > best) { best = v; best_c = c; }
  }
  labels[cell] = best_c;
  valid_cells[cell] = (float)ok;
}

}  // namespace

extern "C" int spn_detector_labels(spn_ctx* ctx, const int32_t* d_kpts_heatmap, const int32_t* d_valid_mask, const float* d_noise,
                                   uint64_t seed, int B, int H, int W, int64_t* d_labels, float* d_valid_cells, spn_stream stream) {
  SPN_REQUIRE(ctx && d_kpts_heatmap && d_labels && d_valid_cells, "spn_detector_labels: null pointer");
  SPN_REQUIRE(B > 0 && H > 0 && W > 0 && H % 8 == 0 && W % 8 == 0, "spn_detector_labels: H, W must be positive multiples of 8");
  SpnDeviceGuard guard(ctx->device);
  const int cells = B * (H / 8) * (W / 8);
  detector_labels_kernel<<<spn_cdiv(cells, 128), 128, 0, (cudaStream_t)stream>>>(d_kpts_heatmap, d_valid_mask, d_noise, seed, B, H / 8,
                                                                                 W / 8, d_labels, d_valid_cells);
  SPN_CHECK_LAUNCH(ctx);
  return SPN_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// ExportNeRFDetections.step splat (engine_solvers/export.py:271-283 of the reference): for every (destination point u_i,
// source point w_i) pair IN ORDER, copy the 3x3 patch of the source heatmap around w_i onto the 3x3 patch around u_i
// (a single pixel when either point is within one pixel of the border); later pairs overwrite earlier ones.  The
// sequential "last writer wins" rule is reproduced in parallel with an owner map: pass 1 records, per destination
// pixel, the highest pair index that covers it (atomicMax), pass 2 gathers the value that pair would have written.
namespace {

__device__ __forceinline__ bool nerf_border(int u0, int u1, int w0, int w1, int H, int W) {
  return u0 <= 1 || u1 <= 1 || u0 >= H - 1 || u1 >= W - 1 || w0 <= 1 || w1 <= 1 || w0 >= H - 1 || w1 >= W - 1;
}

__global__ void nerf_owner_kernel(const float* __restrict__ dst, const int32_t* __restrict__ src, int n, int H, int W,
                                  unsigned* __restrict__ owner) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int u0 = (int)dst[2 * i], u1 = (int)dst[2 * i + 1];   // python int(): truncation toward zero
  const int w0 = src[2 * i], w1 = src[2 * i + 1];
  if (u0 < 0 || u0 >= H || u1 < 0 || u1 >= W) return;
  if (nerf_border(u0, u1, w0, w1, H, W)) {
    atomicMax(&owner[u0 * W + u1], (unsigned)i + 1u);
  } else {
    for (int dy = -1; dy <= 1; ++dy)
      for (int dx = -1; dx <= 1; ++dx) atomicMax(&owner[(u0 + dy) * W + u1 + dx], (unsigned)i + 1u);
  }
}

__global__ void nerf_gather_kernel(const float* __restrict__ prob, const float* __restrict__ dst, const int32_t* __restrict__ src, int H,
                                   int W, const unsigned* __restrict__ owner, float* __restrict__ out) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= H * W) return;
  const unsigned o = owner[p];
  float v = 0.f;
  if (o) {
    const int i = (int)o - 1;
    const int u0 = (int)dst[2 * i], u1 = (int)dst[2 * i + 1];
    const int w0 = src[2 * i], w1 = src[2 * i + 1];
    const int y = p / W, x = p - y * W;
    v = prob[(w0 + (y - u0)) * W + w1 + (x - u1)];   // (y - u0, x - u1) is (0, 0) in the single-pixel case
  }
  out[p] = v;
}

}  // namespace

extern "C" int spn_nerf_splat(spn_ctx* ctx, const float* d_prob_src, const float* d_dst_pts, const int32_t* d_src_pts, int n_pairs, int H,
                              int W, float* d_out, spn_stream stream) {
  SPN_REQUIRE(ctx && d_prob_src && d_out && (n_pairs == 0 || (d_dst_pts && d_src_pts)), "spn_nerf_splat: null pointer");
  SPN_REQUIRE(n_pairs >= 0 && H > 2 && W > 2, "spn_nerf_splat: bad shape");
  SpnDeviceGuard guard(ctx->device);
  cudaStream_t s = (cudaStream_t)stream;
  int rc = spn_ensure_aux(ctx, (size_t)H * W * sizeof(unsigned), s);
  if (rc) return rc;
  unsigned* owner = (unsigned*)ctx->aux;
  SPN_CUDA(cudaMemsetAsync(owner, 0, (size_t)H * W * sizeof(unsigned), s));
  if (n_pairs) {
    nerf_owner_kernel<<<spn_cdiv(n_pairs, 128), 128, 0, s>>>(d_dst_pts, d_src_pts, n_pairs, H, W, owner);
    SPN_CHECK_LAUNCH(ctx);
  }
  nerf_gather_kernel<<<spn_cdiv(H * W, 256), 256, 0, s>>>(d_prob_src, d_dst_pts, d_src_pts, H, W, owner, d_out);
  SPN_CHECK_LAUNCH(ctx);
  return SPN_OK;
}
