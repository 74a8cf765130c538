"""How the arithmetic of the kornia coordinate chain was pinned (build container, CPU only; documentation + re-check).

For every step of kornia.warp_perspective (oracle/kornia_shim.py) this script emulates candidate fp32 operation orders
in numpy and counts bit-exact agreement with torch-CPU on the same inputs.  The winners are what csrc/spn_geom.cuh
(kornia_src, bilinear_combine) and csrc/geometry.cu (kornia_grid, mm3_ref) implement:

  create_meshgrid     xn = (i / (W-1) - 0.5) * 2                                 (three separately rounded ops)
  3x3 @ 3x3           c = (a0*b0 + a1*b1) + a2*b2                                 (no FMA)
  bmm, K = 3          q = fma(a1, yn, a0*xn) + a2                                 (FMA chain in k order)
  transform_points    g = q.xy * (|q.z| > 1e-8 ? 1 / (q.z + 1e-8) : 1)
  grid_sample         s = (g + 1) * (size-1)/2 ; nearest = rint ; bilinear weights s*e, s*w, n*e, n*w and
                      value = fma(se_v, se, fma(sw_v, sw, fma(ne_v, ne, nw_v * nw)))
  torch.inverse       MKL getrf/getrs: NOT reproducible operation-by-operation (approximate reciprocals inside trsm) -
                      hence the sampling matrices are computed with torch on the host (utils/kornia_geometry.py).

    python tools/kornia_chain_fit.py
"""
import sys
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import kornia_shim as K  # noqa: E402
from oracle import spn_oracle as O  # noqa: E402

f = np.float32
PARAMS = dict(translation=True, rotation=True, scaling=True, perspective=True, scaling_amplitude=0.2, perspective_amplitude_x=0.2,
              perspective_amplitude_y=0.2, allow_artifacts=True, patch_ratio=0.85, max_angle=1.57)


def fma(a, b, c):   # exact product in fp64, one rounding to fp32 at the end (double rounding is negligible here)
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(f)


def chain(a, h, w):
    xs = ((np.arange(w, dtype=f) / f(w - 1)) - f(0.5)) * f(2)
    ys = ((np.arange(h, dtype=f) / f(h - 1)) - f(0.5)) * f(2)
    X, Y = np.meshgrid(xs, ys)
    q = [(fma(a[r, 1], Y, f(a[r, 0] * X)) + a[r, 2]).astype(f) for r in range(3)]
    sc = np.where(np.abs(q[2]) > f(1e-8), f(1) / (q[2] + f(1e-8)), f(1)).astype(f)
    return (sc * q[0] + f(1)) * f((w - 1) / 2), (sc * q[1] + f(1)) * f((h - 1) / 2)


def main():
    np.random.seed(3)
    rng = np.random.RandomState(0)
    for (h, w) in [(240, 320), (120, 160), (64, 96), (40, 72)]:
        bad_mask = bad_bil = 0
        img = rng.rand(h, w).astype(f)
        for _ in range(10):
            M = O.sample_homography((h, w), **PARAMS)
            Ai = K._inverse_cast(K.normalize_homography(M, (h, w), (h, w)))
            sx, sy = chain(Ai[0].numpy(), h, w)
            ref = K.warp_perspective(torch.ones(1, 1, h, w), M, (h, w), mode="nearest")[0, 0].numpy()
            m = (np.rint(sx) >= 0) & (np.rint(sx) <= w - 1) & (np.rint(sy) >= 0) & (np.rint(sy) <= h - 1)
            bad_mask += int((m.astype(f) != ref).sum())
            refb = K.warp_perspective(torch.from_numpy(img)[None, None], M, (h, w), mode="bilinear")[0, 0].numpy()
            x0, y0 = np.floor(sx), np.floor(sy)
            wx, ny = sx - x0, sy - y0
            ex, s_ = f(1) - wx, f(1) - ny

            def tap(yy, xx):
                ok = (yy >= 0) & (yy < h) & (xx >= 0) & (xx < w)
                return np.where(ok, img[np.clip(yy, 0, h - 1).astype(int), np.clip(xx, 0, w - 1).astype(int)], f(0))
            v = fma(tap(y0 + 1, x0 + 1), ny * wx, fma(tap(y0 + 1, x0), ny * ex, fma(tap(y0, x0 + 1), s_ * wx, tap(y0, x0) * (s_ * ex))))
            bad_bil += int((v != refb).sum())
        print(f"{h}x{w}: nearest-mask mismatches {bad_mask}, bilinear value mismatches {bad_bil} (10 homographies)")


if __name__ == "__main__":
    main()
