"""GPU tests of the train-time callers that reuse the hot path's kernels against golden outputs of
the unmodified reference: Homographic_aug.__call__ / compute_valid_mask (homographic_augmentation.py:109-151) and the
detector-loss label building (utils/losses.py:13-27)."""
import numpy as np
import pytest
import torch

from conftest import HA_CFG

pytestmark = pytest.mark.gpu


def _aug(margin=2):
    from superpoint_nerf_pytorch_b200.data.data_utils.homographic_augmentation import Homographic_aug
    return Homographic_aug({"params": dict(HA_CFG["params"], allow_artifacts=False, patch_ratio=0.7), "valid_border_margin": margin}, "cuda")


def test_homographic_aug_call_vs_reference_golden(golden):
    g = golden("train_reuse.npz")
    aug = _aug()
    np.random.seed(int(g["aug_np_seed"]))
    out = aug(torch.from_numpy(g["aug_image"]).cuda(), torch.from_numpy(g["aug_points"]).cuda())
    assert torch.equal(out["homography"].cpu(), torch.from_numpy(g["aug_homography"]))       # same numpy RNG order
    assert np.array_equal(out["warp"]["valid_mask"].cpu().numpy(), g["aug_valid_mask"])       # bit-exact eroded mask
    assert out["warp"]["valid_mask"].dtype == torch.int32
    assert np.array_equal(out["warp"]["image"].cpu().numpy(), g["aug_warped"])                # bit-exact bilinear warp
    assert out["warp"]["kpts"].shape == g["aug_kpts"].shape
    assert np.abs(out["warp"]["kpts"].cpu().numpy() - g["aug_kpts"]).max() < 1e-3
    assert np.array_equal(out["warp"]["kpts_heatmap"].cpu().numpy(), g["aug_heatmap"])
    assert 0 < g["aug_valid_mask"].mean() < 1


def test_compute_valid_mask_vs_reference_golden(golden):
    g = golden("train_reuse.npz")
    aug = _aug()
    Hb = torch.from_numpy(g["vm_H"]).cuda()
    for er in (0, 2, 3):                                   # erosion 0 = no erosion (the 1x1 structuring element)
        m = aug.compute_valid_mask((96, 128), Hb, erosion=er)
        assert m.dtype == torch.int32 and tuple(m.shape) == (3, 1, 96, 128)
        assert np.array_equal(m.cpu().numpy().astype(np.uint8), g[f"vm_e{er}"]), er
    one = aug.compute_valid_mask((96, 128), Hb[1], erosion=2)                                # (3,3) input
    assert np.array_equal(one.cpu().numpy().astype(np.uint8), g["vm_e2"][1:2])


def test_detector_labels_vs_reference(golden):
    from superpoint_nerf_pytorch_b200.utils.losses import detector_labels
    g = golden("train_reuse.npz")
    kmap, valid, noise = (torch.from_numpy(g[k]).cuda() for k in ("lab_kmap", "lab_valid", "lab_noise"))
    labels, cells = detector_labels(kmap, valid, include_mask=True, noise=noise)
    assert labels.dtype == torch.int64 and np.array_equal(labels.cpu().numpy(), g["lab_labels"])
    assert np.array_equal(cells.cpu().numpy(), g["lab_cells"])
    # and the loss computed from these labels is the reference's number
    logits = torch.from_numpy(g["lab_logits"]).cuda()
    det = torch.nn.functional.cross_entropy(logits, labels, reduction="none")
    loss = torch.mean(torch.sum(det * cells, dim=(1, 2)) / (torch.sum(cells, dim=(1, 2)) + 1e-10))
    assert abs(float(loss) - float(g["loss_ref_masked"])) < 1e-5
    # device-drawn tie-break noise: labels differ only inside cells that hold several keypoints; dustbin iff empty
    lab2, cells2 = detector_labels(kmap, valid, include_mask=False, seed=5)
    k = torch.pixel_unshuffle(kmap.unsqueeze(1).float(), 8)
    n_kp = k.sum(1)
    assert bool(((lab2 == 64) == (n_kp == 0)).all()) and bool((cells2 == 1).all())
    picked = torch.gather(torch.cat([k, torch.zeros_like(k[:, :1])], 1), 1, lab2.unsqueeze(1)).squeeze(1)
    assert bool((picked[n_kp > 0] == 1).all())                   # a non-empty cell's label is one of its keypoints
    assert bool((lab2[n_kp <= 1] == labels[n_kp <= 1]).all())


def test_nerf_splat_matches_sequential_loop():
    """spn_nerf_splat vs the reference's sequential Python loop (export.py:271-283) as restated by the oracle, with overlapping
    patches (order matters: later pairs overwrite) and border points (single-pixel case)."""
    import superpoint_nerf_pytorch_b200 as P
    ctx = P.get_context("cuda:0")
    rng = np.random.RandomState(4)
    H, W, n = 48, 64, 400
    prob = rng.rand(H, W).astype(np.float32)
    src = np.stack([rng.randint(0, H, n), rng.randint(0, W, n)], 1).astype(np.int32)
    dst = np.stack([rng.uniform(0, H - 1.001, n), rng.uniform(0, W - 1.001, n)], 1).astype(np.float32)
    dst[:6] = [[0.2, 5.7], [1.9, 9.0], [H - 1.5, 3.3], [10.5, W - 1.2], [20.0, 20.0], [20.9, 21.2]]
    from oracle import eval_oracle as EO
    want = EO.nerf_splat(prob, dst, src)            # pinned to the reference's step by tests/test_oracle_golden.py
    got = ctx.nerf_splat(torch.from_numpy(prob).cuda(), torch.from_numpy(dst).cuda(), torch.from_numpy(src).cuda()).cpu().numpy()
    assert np.array_equal(got, want)
    empty = ctx.nerf_splat(torch.from_numpy(prob).cuda(), torch.zeros((0, 2), device="cuda"), torch.zeros((0, 2), dtype=torch.int32, device="cuda"))
    assert float(empty.abs().max()) == 0.0


def test_export_nerf_detections_vs_reference_golden(golden, tmp_path, monkeypatch):
    """ExportNeRFDetections on the synthetic multi-view batches of conftest.make_nerf_batch, same python random seed as the
    reference run that produced the golden keypoint files: >= 99 % of the keypoints within 1 px, both directions."""
    import copy
    import random
    from conftest import MP_MODEL, keypoint_agreement, make_nerf_batch
    from oracle import spn_oracle as O
    from superpoint_nerf_pytorch_b200 import settings
    from superpoint_nerf_pytorch_b200.engine_solvers.export import ExportNeRFDetections
    from superpoint_nerf_pytorch_b200.utils.get_model import get_model
    g = golden("nerf_export.npz")
    mcfg = copy.deepcopy(MP_MODEL)
    mcfg["detector_head"]["top_k"] = 300
    mcfg["precision"] = "fp32"
    m = get_model(mcfg, "cuda").eval()
    m.load_state_dict(O.make_state_dict("magicpoint", seed=int(g["seed"]), logit_gain=float(g["gain"])))
    monkeypatch.setattr(settings, "EXPER_PATH", str(tmp_path))
    batches = [make_nerf_batch(0), make_nerf_batch(1, n_views=5)]
    random.seed(int(g["py_seed"]))
    ExportNeRFDetections({"data": {"experiment_name": "nerf"}, "model": mcfg}, m, batches, "training", "cuda")
    worst = 1.0
    for bt in batches:
        for nm in bt["name"]:
            kp = np.load(tmp_path / "outputs" / "nerf" / "training" / f"{nm}.npy")
            assert kp.dtype == np.int64 and kp.shape[1] == 2
            a, b = keypoint_agreement(kp, g[nm])
            worst = min(worst, a, b)
            assert abs(len(kp) - len(g[nm])) <= max(2, len(g[nm]) // 50), (nm, len(kp), len(g[nm]))
    print(f"NeRF export: worst keypoint agreement {worst:.4f}")
    assert worst >= 0.99


def test_nerf_reproject_vs_reference_step_golden(golden):
    """ExportNeRFDetections.reproject (box_nms keypoint list -> depth-aware re-projection -> spn_nerf_splat) against the
    map ExportNeRFDetections.step of the unmodified reference produced from the same heatmap (tests/golden/nerf_step.npz).
    The re-projection runs in CUDA fp32 (3x3 inverses and 3xN products), so a coordinate that lands within an ulp of an
    integer may truncate differently: at most 0.5 % of the pixels may differ (measured: none)."""
    import copy
    from conftest import MP_MODEL, make_nerf_batch
    from superpoint_nerf_pytorch_b200.engine_solvers.export import ExportNeRFDetections
    from superpoint_nerf_pytorch_b200.utils.get_model import get_model
    g = golden("nerf_step.npz")
    m = get_model(dict(copy.deepcopy(MP_MODEL), precision="fp32"), "cuda").eval()
    ex = object.__new__(ExportNeRFDetections)
    ex.model, ex.device = m, "cuda"
    ctx = m.native()
    det = float(g["det_thresh"])
    worst = 0.0
    for c in range(int(g["n"])):
        bseed, nv, j, k = (int(v) for v in g[f"case{c}"])
        bt = make_nerf_batch(bseed, n_views=nv)
        raw = {kk: v.cuda() for kk, v in bt["raw"].items()}
        heat = torch.from_numpy(g[f"heat{c}"]).cuda()
        r = ctx.box_nms(heat[None], float(g["nms"]), 0.1, det, int(g["top_k"]), det_thresh=det, want_map=False, max_kp=4096)
        n = int(r["kp_count"][0])
        got = ex.reproject(heat, r["kp"][0, :n], raw["input_depth"][k], bt["camera_intrinsic_matrix"][j].cuda(),
                           raw["input_rotation"][k], raw["input_translation"][k], raw["input_rotation"][j],
                           raw["input_translation"][j]).cpu().numpy()
        bad = float((got != g[f"splat{c}"]).mean())
        worst = max(worst, bad)
        assert bad <= 5e-3, (c, bad)
    print(f"NeRF reproject: worst fraction of differing pixels {worst:.2e}")
