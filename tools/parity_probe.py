"""Fast-mode parity probe on the headline workload (tools only; the gates live in tests/test_gpu_tc.py).

For {peaky, random-init} weights x {fp32, f16[, f16x3]} at 240x320 with N numpy-sampled homographies: rel_err of the
aggregated heatmap over ALL pixels (max|a-b| / max|b|), keypoint agreement within 1 px both ways, logits error of a
plain forward.  Writes gpurun_out/parity_probe.json.
"""
import copy
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from conftest import HA_CFG, MP_MODEL, keypoint_agreement, rel_err, smooth_image  # noqa: E402
from oracle import spn_oracle as O  # noqa: E402


def random_init_sd():
    from superpoint_nerf_pytorch_b200.utils.get_model import get_model
    torch.manual_seed(0)
    return {k: v.detach().cpu() for k, v in get_model(copy.deepcopy(MP_MODEL), "cpu").state_dict().items()}


def main():
    from superpoint_nerf_pytorch_b200.engine_solvers.export import HomographyAdaptation
    from superpoint_nerf_pytorch_b200.utils.get_model import get_model
    n_h = int(sys.argv[1]) if len(sys.argv) > 1 else 25
    modes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["fp32", "f16"]
    out = []
    H, W = 240, 320
    for tag in ("peaky", "random"):
        sd = O.make_state_dict("magicpoint", seed=5, logit_gain=10.0) if tag == "peaky" else random_init_sd()
        ha = copy.deepcopy(HA_CFG)
        ha["num"] = n_h + 1
        cfg = {"homography_adaptation": ha, "model": copy.deepcopy(MP_MODEL)}
        if tag == "peaky":
            img = torch.from_numpy(smooth_image(H, W, 8))[None, None]
        else:
            img = torch.rand((1, 1, H, W), generator=torch.Generator().manual_seed(0))
        np.random.seed(2)
        t0 = time.time()
        want = O.homography_adaptation(sd, img, cfg, nms_fn=O.box_nms_c)
        t_or = time.time() - t0
        ref = want["mean_prob"].numpy()
        for prec in modes:
            c = copy.deepcopy(MP_MODEL)
            c["precision"] = prec
            m = get_model(c, "cuda").eval()
            m.load_state_dict(sd)
            eng = HomographyAdaptation({"homography_adaptation": ha, "model": c}, m, "cuda")
            heat, _ = eng.heatmaps(img.cuda(), homographies=want["homographies"].view(1, n_h, 3, 3))
            agg = heat[0].cpu().numpy()
            kp = eng.keypoints(heat)[0]
            a, b = keypoint_agreement(kp, want["keypoints"])
            err = np.abs(agg - ref) / np.abs(ref).max()
            fw = m(img.cuda())["detector_output"]
            wl = O.model_forward(sd, img, MP_MODEL, nms_fn=O.box_nms_c)["detector_output"]
            rec = {"weights": tag, "precision": prec, "n_h": n_h, "agg_rel_err": float(err.max()),
                   "agg_rel_err_p9999": float(np.quantile(err, 0.9999)), "frac_gt_1e-4": float((err > 1e-4).mean()),
                   "kp_ours": int(len(kp)), "kp_ref": int(len(want["keypoints"])), "kp_within_1px": [a, b],
                   "heat_min_max": [float(ref.min()), float(ref.max())],
                   "logits_rel_err": rel_err(fw["logits"].cpu().numpy(), wl["logits"].numpy()),
                   "prob_rel_err": rel_err(fw["prob_heatmap"].cpu().numpy(), wl["prob_heatmap"].numpy()),
                   "oracle_seconds": t_or}
            print(json.dumps(rec), flush=True)
            out.append(rec)
    Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "parity_probe.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
