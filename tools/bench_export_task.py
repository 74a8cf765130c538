#!/usr/bin/env python
"""Throughput of the user-facing task class ExportDetections (file writes included) on synthetic 240x320 images."""
import argparse
import copy
import json
import sys
import tempfile
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from superpoint_nerf_pytorch_b200 import settings  # noqa: E402
from superpoint_nerf_pytorch_b200.engine import SyntheticLoader  # noqa: E402
from superpoint_nerf_pytorch_b200.engine_solvers.export import ExportDetections  # noqa: E402
from superpoint_nerf_pytorch_b200.utils.get_model import get_model  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--images", type=int, default=512)
ap.add_argument("--per-launch", type=int, default=16)
ap.add_argument("--streams", type=int, default=1)
a = ap.parse_args()
mcfg = dict(copy.deepcopy(bench.MODEL_CFG), precision="f16")
model = get_model(mcfg, "cuda").eval()
model.load_state_dict(bench.random_init_state_dict())
ha = dict(copy.deepcopy(bench.HA_CFG), sampler="device", images_per_launch=a.per_launch, max_forwards=100, streams=a.streams)
cfg = {"data": {"experiment_name": "task"}, "homography_adaptation": ha, "model": mcfg}
with tempfile.TemporaryDirectory() as tmp:
    settings.EXPER_PATH = tmp
    ExportDetections(cfg, model, SyntheticLoader(2 * a.per_launch, (240, 320), "export_pseudo_labels", seed=10**6), "warm", True, "cuda")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ExportDetections(cfg, model, SyntheticLoader(a.images, (240, 320), "export_pseudo_labels"), "training", True, "cuda")
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    n = len(list(Path(tmp, "outputs", "task", "training").glob("*.npy")))
print(json.dumps({"task": "ExportDetections (240x320, 100 H, .npy files written)", "images": n, "seconds": dt, "img_per_s": n / dt,
                  "images_per_launch": a.per_launch, "streams": a.streams}))
