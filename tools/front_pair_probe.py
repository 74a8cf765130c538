"""A/B of the fused front end: single-CTA kernel (front_tc.cu) vs the 2-CTA cluster kernel (front2_tc.cu, option
front_pair): identical results required (same operands, same accumulation order per output), then timing of 100 slots."""
import copy
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import bench  # noqa: E402
from conftest import smooth_image  # noqa: E402
from superpoint_nerf_pytorch_b200.utils.get_model import get_model  # noqa: E402

m = get_model(dict(copy.deepcopy(bench.MODEL_CFG), precision="f16"), "cuda").eval()
m.load_state_dict(bench.random_init_state_dict())
ctx = m.native()
ok = True
for (NI, H, W, n_h) in [(1, 64, 96, 3), (2, 240, 320, 5), (1, 40, 72, 2), (3, 120, 160, 4), (1, 8, 8, 0), (1, 16, 8, 2)]:
    imgs = torch.from_numpy(np.stack([smooth_image(H, W, 60 + i) for i in range(NI)])).cuda()
    hinv = None
    if n_h:
        h, hinv = ctx.sample_homographies(bench.HA_CFG["params"], seed=3, first_index=0, count=NI * n_h, H=H, W=W)
        hinv = hinv.view(NI, n_h, 3, 3)
    B = NI * (n_h + 1)
    outs = []
    for pair in (0, 1):
        ctx.set_option("front_pair", pair)
        outs.append(m.prob_heatmap_ha(imgs, hinv, 0, B).clone())
        if B > 3:
            part = m.prob_heatmap_ha(imgs, hinv, 1, B - 2).clone()
            assert torch.equal(part, outs[-1][1:B - 1]), "slot sub-range differs"
    torch.cuda.synchronize()
    same = torch.equal(outs[0], outs[1])
    err = float((outs[0] - outs[1]).abs().max())
    print(f"{NI}x{H}x{W} n_h={n_h}: pair == single: {same} (max abs diff {err:.3e})", flush=True)
    ok &= same
NI, H, W, n_h = 4, 240, 320, 99
imgs = torch.rand((NI, H, W), device="cuda")
h, hinv = ctx.sample_homographies(bench.HA_CFG["params"], seed=1, first_index=0, count=NI * n_h, H=H, W=W)
hinv = hinv.view(NI, n_h, 3, 3)
for pair in (0, 1, 0, 1):
    ctx.set_option("front_pair", pair)
    for _ in range(2):
        for i in range(NI):
            ctx.encoder_forward_ha(imgs, hinv, i * 100, 100, 1)
    torch.cuda.synchronize()
    ctx.profile_enable(True)
    ctx.profile_read()
    for _ in range(5):
        for i in range(NI):
            ctx.encoder_forward_ha(imgs, hinv, i * 100, 100, 1)
    prof = ctx.profile_read()
    ctx.profile_enable(False)
    t, n = prof["backbone.block_2"]
    print(f"front_pair={pair}: front kernel {t / n:.4f} ms per 100 forwards ({n} launches)", flush=True)
print("ALL EQUAL" if ok else "MISMATCH")
