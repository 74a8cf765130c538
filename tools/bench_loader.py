#!/usr/bin/env python
"""Does the loader keep the export fed?  

Writes N synthetic 640x480 JPEGs, then measures (a) the prefetching COCO loader alone (decode on host threads + resize
kernel) and (b) the ExportDetections task (f16, 100 homographies, 240x320) driven by that loader, .npy writes included.

    python tools/bench_loader.py [--n 1024] [--workers W]
"""
import argparse
import copy
import json
import shutil
import sys
import tempfile
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from superpoint_nerf_pytorch_b200 import settings  # noqa: E402
from superpoint_nerf_pytorch_b200.engine_solvers.export import ExportDetections  # noqa: E402
from superpoint_nerf_pytorch_b200.utils.data_loaders import get_loader  # noqa: E402
from superpoint_nerf_pytorch_b200.utils.get_model import get_model  # noqa: E402


def main():
    import cv2
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1024)
    ap.add_argument("--workers", type=int, default=0)
    ap.add_argument("--images-per-launch", type=int, default=64)
    ap.add_argument("--max-forwards", type=int, default=400)
    args = ap.parse_args()
    tmp = Path(tempfile.mkdtemp(prefix="spn_loader_"))
    try:
        d = tmp / "data" / "COCO" / "images" / "training"
        d.mkdir(parents=True)
        rng = np.random.RandomState(0)
        base = cv2.GaussianBlur(rng.randint(0, 256, (480, 640, 3)).astype(np.uint8), (0, 0), 2.0)
        for k in range(args.n):
            cv2.imwrite(str(d / f"im{k:06d}.jpg"), np.roll(base, k * 7, axis=1), [cv2.IMWRITE_JPEG_QUALITY, 90])
        settings.DATA_PATH, settings.EXPER_PATH = str(tmp / "data"), str(tmp / "exper")
        data = {"name": "COCO", "class_name": "COCO", "experiment_name": "loader_bench", "preprocessing": {"resize": [240, 320]},
                "has_labels": False, "warped_pair": False, "batch_size": 1, "truncate": False,
                "augmentation": {"photometric": {"enable": False}, "homographic": {"enable": False}}}
        if args.workers:
            data["loader_workers"] = args.workers
        mcfg = dict(copy.deepcopy(bench.MODEL_CFG), precision="f16")
        cfg = {"data": data, "model": mcfg,
               "homography_adaptation": dict(copy.deepcopy(bench.HA_CFG), sampler="device", seed=1, images_per_launch=args.images_per_launch,
                                            max_forwards=args.max_forwards)}
        loader = get_loader(cfg, "export_pseudo_labels", device="cuda", export_split="training")
        for _ in zip(range(32), loader):
            pass
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = sum(b["raw"]["image"].shape[0] for b in loader)
        torch.cuda.synchronize()
        t_load = time.perf_counter() - t0
        model = get_model(mcfg, "cuda").eval()
        model.load_state_dict(bench.random_init_state_dict())
        warm = copy.deepcopy(cfg)
        warm["data"]["experiment_name"] = "warm"
        n_warm = max(64, 2 * args.images_per_launch)   # two full groups: workspace, pinned staging and allocator pools at their final size
        ExportDetections(warm, model, [b for _, b in zip(range(n_warm), loader)], "training", True, "cuda")
        torch.cuda.synchronize()
        # where the export loop's host thread spends its time: waiting for the GPU (good), launching, saving, or in the loader
        acc = {"wait_gpu": 0.0, "launch": 0.0, "finish_total": 0.0}
        from superpoint_nerf_pytorch_b200.engine_solvers import export as E
        o_wait, o_launch, o_finish = E.HomographyAdaptation.keypoints_wait, E.ExportDetections._launch, E.ExportDetections._finish

        def timed(fn, key):
            def w(*a, **k):
                t = time.perf_counter()
                try:
                    return fn(*a, **k)
                finally:
                    acc[key] += time.perf_counter() - t
            return w

        def wait_split(self, handle):
            t = time.perf_counter()
            handle["event"].synchronize()
            acc["wait_gpu"] += time.perf_counter() - t
            return o_wait(self, handle)

        E.HomographyAdaptation.keypoints_wait = wait_split
        E.ExportDetections._launch = timed(o_launch, "launch")
        E.ExportDetections._finish = timed(o_finish, "finish_total")
        t0 = time.perf_counter()
        ExportDetections(cfg, model, loader, "training", True, "cuda")
        torch.cuda.synchronize()
        t_exp = time.perf_counter() - t0
        E.HomographyAdaptation.keypoints_wait, E.ExportDetections._launch, E.ExportDetections._finish = o_wait, o_launch, o_finish
        acc["loader_and_loop"] = t_exp - acc["launch"] - acc["finish_total"]
        acc["save_and_convert"] = acc["finish_total"] - acc["wait_gpu"]
        files = len(list(Path(tmp, "exper", "outputs", "loader_bench", "training").glob("*.npy")))
        print(json.dumps({"images": n, "loader_only_img_per_s": n / t_load, "export_with_loader_img_per_s": files / t_exp,
                          "files_written": files, "host_thread_seconds": {k: round(v, 3) for k, v in acc.items()}, "seconds": round(t_exp, 3),
                          "workers": loader.workers, "images_per_launch": args.images_per_launch,
                          "max_forwards": args.max_forwards, "jpeg": "640x480 q90 -> 240x320",
                          "note": "export = decode (host threads) + resize kernel + HA x100 (f16) + NMS + .npy per image"}))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
