import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"

MP_MODEL = {"script": "SuperPoint", "class_name": "SuperPoint", "model_name": "magicpoint",
            "vgg_cn": [64, 64, 64, 64, 128, 128, 128, 128],
            "detector_head": {"detector_dim": [128, 256], "grid_size": 8, "nms": 4, "det_thresh": 0.015, "top_k": 0}}
SP_MODEL = {"script": "SuperPoint", "class_name": "SuperPoint", "model_name": "superpoint",
            "vgg_cn": [64, 64, 64, 64, 128, 128, 128, 128],
            "detector_head": {"detector_dim": [128, 256], "grid_size": 8, "nms": 4, "det_thresh": 0.001, "top_k": 50},
            "descriptor_head": {"descriptor_dim": [128, 256], "grid_size": 8}}
HA_CFG = {"num": 8, "aggregation": "sum", "filter_counts": 0, "valid_border_margin": 3,
          "params": {"translation": True, "rotation": True, "scaling": True, "perspective": True,
                     "scaling_amplitude": 0.2, "perspective_amplitude_x": 0.2, "perspective_amplitude_y": 0.2,
                     "allow_artifacts": True, "patch_ratio": 0.85, "max_angle": 1.57}}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def golden():
    return lambda name: np.load(GOLDEN / name, allow_pickle=False)


def smooth_image(h, w, seed):
    """Same generator as tests/golden/make_golden.py."""
    rng = np.random.RandomState(seed)
    img = rng.rand(h, w).astype(np.float32)
    k = np.array([1, 4, 6, 4, 1], np.float32) / 16
    for _ in range(2):
        img = np.apply_along_axis(lambda r: np.convolve(r, k, mode="same"), 1, img)
        img = np.apply_along_axis(lambda r: np.convolve(r, k, mode="same"), 0, img)
    img = (img - img.min()) / (img.max() - img.min())
    return img.astype(np.float32)


def rel_err(a, b):
    """max|a-b| / max|b|  (the parity metric of SURVEY.md section 8d)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def keypoint_agreement(a, b, tol=1.0):
    """fraction of points of a with a point of b within tol px (Chebyshev), and vice versa."""
    a = np.asarray(a, np.float64).reshape(-1, 2)
    b = np.asarray(b, np.float64).reshape(-1, 2)
    if len(a) == 0 and len(b) == 0:
        return 1.0, 1.0
    if len(a) == 0 or len(b) == 0:
        return 0.0, 0.0
    d = np.abs(a[:, None, :] - b[None, :, :]).max(-1)
    return float((d.min(1) <= tol).mean()), float((d.min(0) <= tol).mean())


def make_eval_pair(seed, h=96, w=128, n_pts=220):
    """Synthetic HPatches-style evaluation pair (shared by tests/golden/make_golden.py and the GPU tests): an NMS-like
    sparse ``prob`` map, a ground-truth homography, a ``warped_prob`` whose detections are the mapped points (some
    jittered, some dropped, some spurious) and dense (H,W,256) descriptor maps whose vectors correspond across the pair
    up to noise.  numpy RandomState only, so it regenerates bit-identically anywhere."""
    rng = np.random.RandomState(seed)
    ang, sc = rng.uniform(-0.25, 0.25), rng.uniform(0.9, 1.1)
    H = np.array([[sc * np.cos(ang), -sc * np.sin(ang), rng.uniform(-6, 6)],
                  [sc * np.sin(ang), sc * np.cos(ang), rng.uniform(-5, 5)],
                  [rng.uniform(-2e-4, 2e-4), rng.uniform(-2e-4, 2e-4), 1.0]]).astype(np.float32)
    prob = np.zeros((h, w), np.float32)
    warped_prob = np.zeros((h, w), np.float32)
    desc = rng.randn(h, w, 256).astype(np.float32)
    warped_desc = rng.randn(h, w, 256).astype(np.float32)
    ys, xs = rng.randint(0, h, n_pts), rng.randint(0, w, n_pts)
    for y, x in zip(ys, xs):
        if prob[y, x] > 0:
            continue
        prob[y, x] = rng.uniform(0.05, 1.0)
        q = H.astype(np.float64) @ np.array([x, y, 1.0])
        wx, wy = q[0] / q[2], q[1] / q[2]
        r = rng.rand()
        if r < 0.25:
            continue                                    # not re-detected
        jy, jx = (rng.randint(-2, 3), rng.randint(-2, 3)) if r < 0.5 else (0, 0)
        iy, ix = int(round(wy)) + jy, int(round(wx)) + jx
        if 0 <= iy < h and 0 <= ix < w and warped_prob[iy, ix] == 0:
            warped_prob[iy, ix] = rng.uniform(0.05, 1.0)
            warped_desc[iy, ix] = desc[y, x] + 0.35 * rng.randn(256).astype(np.float32)
    for _ in range(n_pts // 5):                          # spurious detections in the warped image
        iy, ix = rng.randint(0, h), rng.randint(0, w)
        if warped_prob[iy, ix] == 0:
            warped_prob[iy, ix] = rng.uniform(0.05, 1.0)
    desc /= np.linalg.norm(desc, axis=-1, keepdims=True)
    warped_desc /= np.linalg.norm(warped_desc, axis=-1, keepdims=True)
    return {"prob": prob, "warped_prob": warped_prob, "homography": H, "desc": desc, "warped_desc": warped_desc}


def make_nerf_batch(seed=0, n_views=4, h=64, w=96):
    """Synthetic multi-view batch in the layout ExportNeRFDetections consumes (export.py:304-341): images, per-view depth
    maps with a depth edge, shared intrinsics, small relative rotations / translations.  numpy only (bit-reproducible)."""
    import torch
    rng = np.random.RandomState(seed)
    imgs = np.stack([smooth_image(h, w, 700 + seed * 10 + i) for i in range(n_views)])[:, None]
    yy, xx = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    depth = np.stack([2.5 + 0.3 * np.sin(xx / 17.0 + i) + 0.2 * np.cos(yy / 11.0) + 0.5 * (xx > w * (0.45 + 0.05 * i)) for i in range(n_views)])
    K = np.array([[85.0, 0, w / 2.0], [0, 85.0, h / 2.0], [0, 0, 1.0]], np.float32)

    def rot(ax, ay, az):
        cx, sx, cy, sy, cz, sz = np.cos(ax), np.sin(ax), np.cos(ay), np.sin(ay), np.cos(az), np.sin(az)
        return (np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]]) @ np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
                @ np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]]))
    R = np.stack([rot(*(rng.uniform(-0.03, 0.03, 3))) for _ in range(n_views)]).astype(np.float32)
    t = rng.uniform(-0.08, 0.08, (n_views, 3, 1)).astype(np.float32)
    return {"raw": {"image": torch.from_numpy(imgs.astype(np.float32)), "input_depth": torch.from_numpy(depth.astype(np.float32)),
                    "input_rotation": torch.from_numpy(R), "input_translation": torch.from_numpy(t)},
            "camera_intrinsic_matrix": torch.from_numpy(np.stack([K] * n_views)), "name": [f"view{seed}_{i}" for i in range(n_views)]}
