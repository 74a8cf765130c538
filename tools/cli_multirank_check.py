"""Multi-rank check of the CLI on real GPUs (one process per GPU, NCCL): run

    python tools/cli_multirank_check.py [--ranks 2] [--images 12]

on a box with >= `ranks` GPUs.  It runs the export task twice through `python -m torch.distributed.run ... -m
superpoint_nerf_pytorch_b200.engine --task export_pseudo_labels --synthetic N --random_init True` (packed export on):
once with `ranks` processes and once with one, into separate EXPER_PATHs, and checks that
  * every image was written exactly once, by the rank that owns its index (index = rank mod world),
  * the packed shards' offsets / global offsets are the exclusive prefix sums of the all-gathered per-rank counts,
  * packed shards and per-image .npy files hold the same keypoints,
  * the multi-rank keypoints are IDENTICAL to the single-process run (the device sampler is keyed by the global image
    index, so the homographies of an image do not depend on rank, batching or world size).
Prints one JSON line and exits non-zero on any mismatch."""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np
import yaml

ROOT = Path(__file__).resolve().parents[1]


def run_export(n_ranks: int, n_images: int, exper: Path, cfg_path: Path, port: int) -> None:
    env = dict(os.environ, SPN_EXPER_PATH=str(exper), PYTHONPATH=str(ROOT))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n_ranks}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), "-m", "superpoint_nerf_pytorch_b200.engine", "--config_path", str(cfg_path), "--task",
           "export_pseudo_labels", "--synthetic", str(n_images), "--random_init", "True"]
    subprocess.run(cmd, check=True, cwd=ROOT, env=env, timeout=300, stdout=subprocess.DEVNULL)


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--ranks", type=int, default=2)
    ap.add_argument("--images", type=int, default=12)
    args = ap.parse_args()
    sys.path.insert(0, str(ROOT))
    from superpoint_nerf_pytorch_b200.engine_solvers.export import load_packed_labels

    cfg = yaml.safe_load(open(ROOT / "superpoint-nerf-pytorch_b200" / "configs" / "magicpoint_coco_export.yaml"))
    cfg["data"].update(packed=True, experiment_name="multirank")
    cfg["homography_adaptation"].update(images_per_launch=4, seed=7)
    with tempfile.TemporaryDirectory() as tmp:
        tmp = Path(tmp)
        cfg_path = tmp / "cfg.yaml"
        cfg_path.write_text(yaml.safe_dump(cfg))
        run_export(args.ranks, args.images, tmp / "multi", cfg_path, 29631)
        run_export(1, args.images, tmp / "single", cfg_path, 29632)
        out_m = tmp / "multi" / "outputs" / "multirank" / "training"
        out_s = tmp / "single" / "outputs" / "multirank" / "training"
        names = [f"synthetic_{i:08d}" for i in range(args.images)]
        problems = []
        shards = sorted(out_m.glob("packed_rank*.npz"))
        if len(shards) != args.ranks:
            problems.append(f"{len(shards)} packed shards for {args.ranks} ranks")
        total, base = 0, 0
        for r, f in enumerate(shards):
            z = np.load(f)
            want_names = names[r::args.ranks]
            if [str(n) for n in z["names"]] != want_names:
                problems.append(f"rank {r} wrote {list(z['names'])}, owns {want_names}")
            counts = z["counts_all_ranks"]
            if int(z["global_offset"]) != int(counts[:r, 1].sum()) or int(z["global_offset"]) != base:
                problems.append(f"rank {r}: global_offset {int(z['global_offset'])}, prefix sum {int(counts[:r, 1].sum())}, running {base}")
            if int(counts[r, 0]) != len(want_names) or int(counts[r, 1]) != len(z["keypoints"]) or int(z["offsets"][-1]) != len(z["keypoints"]):
                problems.append(f"rank {r}: counts {counts[r].tolist()} vs {len(want_names)} images / {len(z['keypoints'])} keypoints")
            base += len(z["keypoints"])
        packed_m, packed_s = load_packed_labels(out_m), load_packed_labels(out_s)
        for n in names:
            a, b = np.load(out_m / f"{n}.npy"), np.load(out_s / f"{n}.npy")
            total += len(a)
            if not np.array_equal(a, packed_m.get(n)) or not np.array_equal(b, packed_s.get(n)):
                problems.append(f"{n}: packed shard and .npy differ")
            if not np.array_equal(a, b):
                problems.append(f"{n}: {args.ranks}-rank run has {len(a)} keypoints, single-process run {len(b)} (or different ones)")
        print(json.dumps({"ranks": args.ranks, "images": args.images, "keypoints": total, "problems": problems}))
        return 1 if problems or total == 0 else 0


if __name__ == "__main__":
    sys.exit(main())
