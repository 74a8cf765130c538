"""Dump device-sampled homographies (spn_sample_homographies) for offline comparison with a host build of the same code."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from conftest import HA_CFG  # noqa: E402
from superpoint_nerf_pytorch_b200 import _native  # noqa: E402

ctx = _native.Context(0)
cases = {"export": HA_CFG["params"], "noart": dict(HA_CFG["params"], allow_artifacts=False, patch_ratio=0.5),
         "wide": dict(HA_CFG["params"], allow_artifacts=False, patch_ratio=0.7, scaling_amplitude=0.3, max_angle=0.8, n_angles=9, n_scales=3)}
out = {}
for tag, p in cases.items():
    h, hinv = ctx.sample_homographies(p, seed=17, first_index=0, count=2000, H=240, W=320)
    out[tag] = h.cpu().numpy()
Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
np.savez(ROOT / "gpurun_out" / "sampler_dump.npz", **out)
print("ok")
