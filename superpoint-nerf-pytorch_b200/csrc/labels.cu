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
