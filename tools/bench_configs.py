#!/usr/bin/env python
"""Timings of the secondary BASELINE.json configs (1, 3, 4) on one B200 - not the headline metric (bench.py).

    python tools/bench_configs.py [--cpu]      # --cpu also times the oracle port on the host cores (slow: NMS is O(N^2))

config 1: MagicPoint forward + box-NMS (radius 4) on 32x1x120x160, random-init weights
config 3: SuperPoint forward at 480x640, top-1000 keypoints, bicubic descriptor sampling + L2 norm (sparse and dense)
config 4: HPatches-shaped repeatability export (pairs at 480x640 and 240x320) through Export_Hpatches_Repeatability
"""
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

from superpoint_nerf_pytorch_b200 import settings  # noqa: E402
from superpoint_nerf_pytorch_b200.engine_solvers.export import Export_Hpatches_Repeatability  # noqa: E402
from superpoint_nerf_pytorch_b200.utils.get_model import get_model  # noqa: E402

MP = {"script": "SuperPoint", "class_name": "SuperPoint", "model_name": "magicpoint", "vgg_cn": [64, 64, 64, 64, 128, 128, 128, 128],
      "detector_head": {"detector_dim": [128, 256], "grid_size": 8, "nms": 4, "det_thresh": 0.001, "top_k": 0}}
SP = {"script": "SuperPoint", "class_name": "SuperPoint", "model_name": "superpoint", "vgg_cn": [64, 64, 64, 64, 128, 128, 128, 128],
      "detector_head": {"detector_dim": [128, 256], "grid_size": 8, "nms": 4, "det_thresh": 0.001, "top_k": 1000},
      "descriptor_head": {"descriptor_dim": [128, 256], "grid_size": 8}, "dense_desc": False}


def timed(fn, warmup=3, iters=10):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cpu", action="store_true")
    args = ap.parse_args()
    out = {}
    torch.manual_seed(0)
    g = torch.Generator().manual_seed(0)

    # ---- config 1 ---------------------------------------------------------------------------------------------
    x1 = torch.rand((32, 1, 120, 160), generator=g).cuda()
    for prec in ("fp32", "f16"):
        m = get_model(dict(copy.deepcopy(MP), precision=prec), "cuda").eval()
        ms = timed(lambda: m(x1))
        out[f"config1_{prec}"] = {"ms_per_batch32": ms, "img_per_s": 32e3 / ms}
    if args.cpu:
        from oracle import spn_oracle as O
        sd = {k: v.cpu() for k, v in m.state_dict().items()}
        t0 = time.perf_counter()
        O.model_forward(sd, x1[:4].cpu(), MP)
        out["config1_cpu_oracle"] = {"img_per_s": 4 / (time.perf_counter() - t0), "sample": "4 of 32 images", "cores": torch.get_num_threads()}

    # ---- config 3 ---------------------------------------------------------------------------------------------
    x3 = torch.rand((1, 1, 480, 640), generator=g).cuda()
    for prec in ("fp32", "f16"):
        m = get_model(dict(copy.deepcopy(SP), precision=prec), "cuda").eval()
        ctx = m.native()

        def fwd_sparse():   # one C-ABI call: forward + NMS/top-k once + descriptors at the 1000 keypoints
            return m(x3, keypoints=True)["descriptor_output"]["desc_sparse"]

        ms = timed(fwd_sparse, iters=50)
        d = fwd_sparse()
        out[f"config3_{prec}_sparse"] = {"ms_per_image": ms, "img_per_s": 1e3 / ms, "descriptors": list(d.shape)}
        md = get_model(dict(copy.deepcopy(SP), precision=prec, dense_desc=True), "cuda").eval()
        ms = timed(lambda: md(x3), iters=5)
        out[f"config3_{prec}_dense_desc"] = {"ms_per_image": ms, "img_per_s": 1e3 / ms, "desc_MB": 256 * 480 * 640 * 4 / 1e6}

    # ---- config 4 ---------------------------------------------------------------------------------------------
    for (H, W) in ((480, 640), (240, 320)):
        m = get_model(dict(copy.deepcopy(MP), precision="f16", detector_head=dict(MP["detector_head"], top_k=1000)), "cuda").eval()
        n = 20
        pairs = []
        for i in range(n):
            img = torch.rand((1, 1, H, W), generator=g)
            h = torch.eye(3).unsqueeze(0)
            pairs.append({"image": img, "warped_image": torch.roll(img, (3, -5), (2, 3)), "homography": h, "name": [f"p{i}"]})
        with tempfile.TemporaryDirectory() as tmp:
            settings.EXPER_PATH = tmp
            cfg = {"data": {"experiment_name": "cfg4"}, "model": MP}
            Export_Hpatches_Repeatability(cfg, m, pairs[:2], "cuda")
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            Export_Hpatches_Repeatability(cfg, m, pairs, "cuda")
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        out[f"config4_{H}x{W}"] = {"pairs_per_s": n / dt, "note": "includes D2H and np.savez_compressed of 5 arrays per pair"}
    import subprocess
    try:   # clock record for these builder-run numbers
        q = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.sw_power_cap,"
                            "clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown", "--format=csv,noheader"],
                           capture_output=True, text=True, timeout=10).stdout.strip()
        out["clocks_after_run"] = q
    except Exception as e:
        out["clocks_after_run"] = f"unavailable: {e}"
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
