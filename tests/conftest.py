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
