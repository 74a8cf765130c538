"""Generate the golden vectors under tests/golden/ by RUNNING THE UNMODIFIED REFERENCE.

Build-container only: imports /root/reference/superpoint (read-only) with oracle/kornia_shim.py standing in
for the missing kornia package.  The GPU box has no /root/reference; tests read the
committed .npz files only.

    python tests/golden/make_golden.py

Everything is seeded; weights come from oracle.spn_oracle.make_state_dict (numpy RNG) so the tests can rebuild
the identical state dict without torch's init RNG.
"""
from __future__ import annotations

import copy
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
REF = Path("/root/reference/superpoint")
OUT = Path(__file__).resolve().parent

from oracle import kornia_shim  # noqa: E402
from oracle import spn_oracle as O  # noqa: E402

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


def smooth_image(h, w, seed):
    """Blurred-noise 'COCO-shaped' test image in [0,1] (numpy only, deterministic)."""
    rng = np.random.RandomState(seed)
    img = rng.rand(h, w).astype(np.float32)
    k = np.array([1, 4, 6, 4, 1], np.float32) / 16
    for _ in range(2):
        img = np.apply_along_axis(lambda r: np.convolve(r, k, mode="same"), 1, img)
        img = np.apply_along_axis(lambda r: np.convolve(r, k, mode="same"), 0, img)
    img = (img - img.min()) / (img.max() - img.min())
    return img.astype(np.float32)


def main():
    exper = tempfile.mkdtemp(prefix="spn_golden_")
    kornia_shim.install(exper_path=exper)
    sys.path.insert(0, str(REF))
    from superpoint.models.SuperPoint import SuperPoint
    from superpoint.models.model_utils.sp_utils import box_nms
    from superpoint.data.data_utils.homographic_augmentation import Homographic_aug
    import superpoint.engine_solvers.export as refexport

    torch.set_grad_enabled(False)

    # ---- 1. box_nms known answers -------------------------------------------------------------
    rng = np.random.RandomState(1)
    cases = {}
    ci = 0
    for (h, w, size, min_prob, topk, kind) in [
        (48, 64, 4, 0.015, 0, "dense"), (48, 64, 4, 0.001, 20, "dense"), (40, 56, 3, 0.3, 0, "uniform"),
        (40, 56, 8, 0.0, 0, "uniform"), (32, 48, 4, 0.1, 0, "ties"), (32, 48, 4, 0.1, 7, "ramp"),
        (24, 24, 4, 0.5, 0, "tie3"), (16, 16, 4, 2.0, 0, "empty"), (64, 80, 2, 0.2, 0, "uniform"),
        (30, 40, 5, 0.2, 11, "uniform"),
    ]:
        if kind == "dense":
            p = (1 / 65 + 0.001 * rng.randn(h, w)).astype(np.float32)
        elif kind == "uniform":
            p = rng.rand(h, w).astype(np.float32)
        elif kind == "ties":
            p = (np.round(rng.rand(h, w) * 8) / 8).astype(np.float32)
        elif kind == "ramp":
            p = (np.arange(h * w, dtype=np.float32).reshape(h, w)[::-1, ::-1] / (h * w)).copy()
        elif kind == "tie3":
            p = np.zeros((h, w), np.float32)
            p[5, 5] = p[5, 6] = p[6, 5] = 0.5
            p[20, 3] = 0.9
        else:
            p = rng.rand(h, w).astype(np.float32)
        out = box_nms(torch.from_numpy(p), size, min_prob=min_prob, keep_top_k=topk).numpy()
        cases[f"c{ci}_prob"] = p
        cases[f"c{ci}_out"] = out
        cases[f"c{ci}_par"] = np.array([size, 0.1, min_prob, topk], np.float64)
        ci += 1
    cases["n"] = np.array(ci)
    np.savez_compressed(OUT / "nms_cases.npz", **cases)

    # ---- 2. model forward ---------------------------------------------------------------------
    for tag, mcfg, seed, gain, shape in [("magicpoint", MP_MODEL, 3, 12.0, (2, 1, 48, 64)),
                                         ("superpoint", SP_MODEL, 4, 12.0, (1, 1, 40, 48))]:
        sd = O.make_state_dict(mcfg["model_name"], seed=seed, logit_gain=gain)
        model = SuperPoint(copy.deepcopy(mcfg)).eval()
        model.load_state_dict(sd)
        x = torch.from_numpy(np.stack([smooth_image(shape[2], shape[3], 10 + i) for i in range(shape[0])])[:, None])
        out = model(x)
        d = {"x": x.numpy(), "seed": np.array(seed), "gain": np.array(gain),
             "logits": out["detector_output"]["logits"].numpy(),
             "prob_heatmap": out["detector_output"]["prob_heatmap"].numpy(),
             "prob_heatmap_nms": out["detector_output"]["prob_heatmap_nms"].numpy(),
             "pred_pts": out["detector_output"]["pred_pts"].numpy()}
        if "descriptor_output" in out:
            d["desc_raw"] = out["descriptor_output"]["desc_raw"].numpy()
            desc = out["descriptor_output"]["desc"].numpy()
            pr = np.random.RandomState(5)
            ys = pr.randint(0, shape[2], 64)
            xs = pr.randint(0, shape[3], 64)
            ys[:4] = [0, 0, shape[2] - 1, shape[2] - 1]
            xs[:4] = [0, shape[3] - 1, 0, shape[3] - 1]
            d["desc_pts"] = np.stack([ys, xs], 1).astype(np.int64)
            d["desc_at_pts"] = desc[0][:, ys, xs].T.copy()  # (64,256)
            d["desc_shape"] = np.array(desc.shape)
        np.savez_compressed(OUT / f"forward_{tag}.npz", **d)

    # ---- 3. homography sampler ----------------------------------------------------------------
    aug = Homographic_aug({"params": HA_CFG["params"], "valid_border_margin": 3}, "cpu")
    hs = {}
    for seed in range(6):
        np.random.seed(seed)
        hs[f"s{seed}"] = torch.cat([aug.sample_homography((240, 320), **HA_CFG["params"]) for _ in range(3)]).numpy()
    np.random.seed(100)
    p2 = dict(HA_CFG["params"], allow_artifacts=False, patch_ratio=0.5, n_scales=5, n_angles=25, translation_overflow=0.0)
    hs["noartifact"] = torch.cat([aug.sample_homography((120, 160), **p2) for _ in range(3)]).numpy()
    np.savez_compressed(OUT / "homographies.npz", **hs)

    # ---- 4. one ExportDetections.step: warped image / mask / count / projection -----------------
    img = torch.from_numpy(smooth_image(120, 160, 21))[None, None]
    np.random.seed(7)
    Hs = torch.cat([aug.sample_homography((120, 160), **HA_CFG["params"]) for _ in range(3)])
    steps = {"image": img.numpy(), "H": Hs.numpy()}

    class FakeModel:
        def __init__(self):
            self.seen = []

        def eval(self):
            return self

        def __call__(self, x):
            self.seen.append(x.clone())
            # a smooth, position-dependent "probability" so the projection is informative
            return {"detector_output": {"prob_heatmap": (0.25 + 0.5 * x[:, 0]).clone()}}

    exp = object.__new__(refexport.ExportDetections)
    exp.config = {"homography_adaptation": HA_CFG}
    exp.device = "cpu"
    exp.model = FakeModel()
    for i in range(3):
        class One:
            def sample_homography(self, shape, **kw):
                return Hs[i:i + 1]
        exp.one_homography = One()
        probs0 = torch.zeros((1, 1, 120, 160))
        counts0 = torch.ones((1, 1, 120, 160))
        probs, counts = exp.step(img, probs0, counts0)
        steps[f"warped{i}"] = exp.model.seen[-1][0, 0].numpy()
        steps[f"proj{i}"] = probs[0, 1].numpy()
        steps[f"count{i}"] = counts[0, 1].numpy().astype(np.int32)
        # mask = erosion(nearest warp of ones by H): recover it through a second call whose "model" returns ones
        m = refexport.kornia.morphology.erosion(
            refexport.K.warp_perspective(torch.ones_like(img), Hs[i:i + 1], dsize=(120, 160), mode="nearest", align_corners=True),
            O.erosion_kernel(3))
        steps[f"mask{i}"] = m[0, 0].numpy().astype(np.int32)
    # the normalised sampling matrices kornia derives inside warp_perspective for H and H^-1 (what the device kernels
    # take as input: LAPACK's inverse is not reproducible on the device, the coordinate chain after it is)
    steps["ainv"] = torch.cat([kornia_shim._inverse_cast(kornia_shim.normalize_homography(Hs[i:i + 1], (120, 160), (120, 160)))
                               for i in range(3)]).numpy()
    steps["ainv_back"] = torch.cat([kornia_shim._inverse_cast(kornia_shim.normalize_homography(torch.inverse(Hs[i:i + 1]), (120, 160), (120, 160)))
                                    for i in range(3)]).numpy()
    np.savez_compressed(OUT / "ha_step.npz", **steps)

    # ---- 4b. many validity masks / counts (bit-exact gate for the geometry kernels), several sizes and margins ----
    mm = {}
    cases = [(240, 320, 3, 12, 31), (120, 160, 3, 12, 32), (64, 96, 2, 8, 33), (40, 72, 1, 8, 34), (96, 128, 4, 6, 35)]
    for ci, (h, w, margin, n, seed) in enumerate(cases):
        np.random.seed(seed)
        Hc = torch.cat([aug.sample_homography((h, w), **HA_CFG["params"]) for _ in range(n)])
        ones = torch.ones((1, 1, h, w))
        ker = O.erosion_kernel(margin)
        masks, counts, fw, bw = [], [], [], []
        for i in range(n):
            Hi = Hc[i:i + 1]
            Hinv = torch.inverse(Hi)
            mk = refexport.kornia.morphology.erosion(refexport.K.warp_perspective(ones, Hi, dsize=(h, w), mode="nearest", align_corners=True), ker)
            ct = refexport.kornia.morphology.erosion(refexport.K.warp_perspective(ones, Hinv, dsize=(h, w), mode="nearest", align_corners=True), ker)
            masks.append(mk[0, 0].numpy().astype(np.uint8))
            counts.append(ct[0, 0].numpy().astype(np.uint8))
            fw.append(kornia_shim._inverse_cast(kornia_shim.normalize_homography(Hi, (h, w), (h, w))))
            bw.append(kornia_shim._inverse_cast(kornia_shim.normalize_homography(Hinv, (h, w), (h, w))))
        mm[f"c{ci}_par"] = np.array([h, w, margin, n])
        mm[f"c{ci}_H"] = Hc.numpy()
        mm[f"c{ci}_ainv"] = torch.cat(fw).numpy()
        mm[f"c{ci}_ainv_back"] = torch.cat(bw).numpy()
        mm[f"c{ci}_mask"] = np.packbits(np.stack(masks), axis=-1)
        mm[f"c{ci}_count"] = np.packbits(np.stack(counts), axis=-1)
    mm["n"] = np.array(len(cases))
    np.savez_compressed(OUT / "ha_masks.npz", **mm)

    # ---- 5. end-to-end ExportDetections (HA, 8 homographies, 120x160) ---------------------------
    sd = O.make_state_dict("magicpoint", seed=11, logit_gain=12.0)
    model = SuperPoint(copy.deepcopy(MP_MODEL)).eval()
    model.load_state_dict(sd)
    config = {"data": {"experiment_name": "golden"}, "homography_adaptation": HA_CFG, "model": copy.deepcopy(MP_MODEL)}
    img = torch.from_numpy(smooth_image(120, 160, 33))[None, None]
    captured = {}
    orig_nms = refexport.box_nms

    def spy_nms(prob, **kw):
        captured["agg"] = prob.clone()
        r = orig_nms(prob=prob, **kw)
        captured["nms"] = r.clone()
        return r

    used_h = []
    orig_sample = Homographic_aug.sample_homography

    def spy_sample(self, *a, **kw):
        h = orig_sample(self, *a, **kw)
        used_h.append(h.clone())
        return h

    refexport.box_nms = spy_nms
    Homographic_aug.sample_homography = spy_sample
    np.random.seed(123)
    refexport.ExportDetections(config, model, [{"raw": {"image": img}, "name": ["img0"]}], "training", True, "cpu")
    refexport.box_nms = orig_nms
    Homographic_aug.sample_homography = orig_sample
    kp = np.load(Path(exper, "outputs", "golden", "training", "img0.npy"))
    Hu = torch.cat(used_h)
    np.savez_compressed(OUT / "ha_export.npz", image=img.numpy(), H=Hu.numpy(), agg=captured["agg"].numpy(),
                        nms=captured["nms"].numpy(), keypoints=kp, seed=np.array(11), gain=np.array(12.0), np_seed=np.array(123),
                        ainv=torch.cat([kornia_shim._inverse_cast(kornia_shim.normalize_homography(Hu[i:i + 1], (120, 160), (120, 160)))
                                        for i in range(len(Hu))]).numpy(),
                        ainv_back=torch.cat([kornia_shim._inverse_cast(kornia_shim.normalize_homography(torch.inverse(Hu[i:i + 1]), (120, 160), (120, 160)))
                                             for i in range(len(Hu))]).numpy())
    print("ha_export keypoints", kp.shape, kp.dtype)

    # ---- 6. HPatches-style exports: file layout --------------------------------------------------
    sdp = O.make_state_dict("superpoint", seed=4, logit_gain=12.0)
    model = SuperPoint(copy.deepcopy(SP_MODEL)).eval()
    model.load_state_dict(sdp)
    im1 = torch.from_numpy(smooth_image(40, 48, 50))[None, None]
    np.random.seed(9)
    Hgt = aug.sample_homography((40, 48), **HA_CFG["params"])
    im2 = kornia_shim.warp_perspective(im1, Hgt, (40, 48))
    loader = [{"image": im1, "warped_image": im2, "homography": Hgt, "name": ["pair0"]}]
    cfg = {"data": {"experiment_name": "golden_hp"}, "model": copy.deepcopy(SP_MODEL)}
    refexport.Export_Hpatches_Repeatability(cfg, model, loader, "cpu")
    refexport.Export_Hpatches_Descriptors(cfg, model, loader, "cpu")
    rep = np.load(Path(exper, "repeatability", "golden_hp", "pair0.npz"))
    des = np.load(Path(exper, "descriptors", "golden_hp", "pair0.npz"))
    layout = {"image": im1.numpy(), "warped_image": im2.numpy(), "homography": Hgt.numpy()}
    for k in rep.files:
        layout[f"rep__{k}__shape"] = np.array(rep[k].shape)
        layout[f"rep__{k}__dtype"] = np.array(str(rep[k].dtype))
    for k in des.files:
        layout[f"des__{k}__shape"] = np.array(des[k].shape)
        layout[f"des__{k}__dtype"] = np.array(str(des[k].dtype))
    layout["rep_prob"] = rep["prob"]
    layout["rep_warped_prob"] = rep["warped_prob"]
    pr = np.random.RandomState(6)
    ys, xs = pr.randint(0, 40, 32), pr.randint(0, 48, 32)
    layout["des_pts"] = np.stack([ys, xs], 1)
    layout["des_desc_at_pts"] = des["desc"][ys, xs]          # (32,256) from the (H,W,256) layout
    layout["des_warped_desc_at_pts"] = des["warped_desc"][ys, xs]
    np.savez_compressed(OUT / "hpatches_export.npz", **layout)
    # ---- 7. loader pre-processing: COCO.ratio_preserving_resize, HPatches.adapt_homography_to_resize
    from superpoint.data.COCO import COCO
    from superpoint.data.HPatches import HPatches
    pre = {}
    rngp = np.random.RandomState(12)
    cases = [((96, 128), (48, 64)), ((107, 160), (48, 64)), ((94, 125), (48, 64)), ((160, 120), (48, 64)), ((83, 125), (60, 80)),
             ((61, 80), (60, 80)), ((50, 70), (48, 64))]
    for k, (src, tgt) in enumerate(cases):
        ds = object.__new__(COCO)
        ds.config = {"preprocessing": {"resize": list(tgt)}}
        u8 = rngp.randint(0, 256, size=src).astype(np.uint8)
        out = ds.ratio_preserving_resize(torch.from_numpy(u8).to(torch.float32)) / 255.0
        pre[f"img{k}"] = u8
        pre[f"tgt{k}"] = np.array(tgt)
        pre[f"out{k}"] = out.numpy()
    pre["n"] = np.array(len(cases))
    hp = object.__new__(HPatches)
    hp.config = {"preprocessing": {"resize": [240, 320]}}
    hp.device = "cpu"
    Hs_in = rngp.randn(3, 3).astype(np.float32) * 0.01 + np.eye(3, dtype=np.float32)
    hom = {"homography": torch.from_numpy(Hs_in), "image_shape": torch.tensor([480.0, 640.0]),
           "warped_image_shape": torch.tensor([427.0, 600.0])}
    pre["h_in"] = Hs_in
    pre["h_out"] = hp.adapt_homography_to_resize(hom).numpy()
    np.savez_compressed(OUT / "preprocess.npz", **pre)

    # ---- 8. export evaluations: detector_evaluation.compute_repeatability and
    #         descriptor_evaluation.compute_homography of the unmodified reference on synthetic pairs ---------------
    sys.path.insert(0, str(ROOT / "tests"))
    from conftest import make_eval_pair
    from superpoint.evaluations import descriptor_evaluation as ref_desc
    from superpoint.evaluations import detector_evaluation as ref_det
    import os
    ev = {}
    seeds = [102, 103, 105, 107]        # chosen so that every nearest neighbour is separated by > 5e-5 (see the assert below)
    ev["seeds"] = np.array(seeds)
    rep_dir = Path(exper, "repeatability", "golden_eval")
    os.makedirs(rep_dir, exist_ok=True)
    ref_det.EXPER_PATH = exper
    for k, seed in enumerate(seeds):
        d = make_eval_pair(seed)
        np.savez(rep_dir / f"pair{k}.npz", prob=d["prob"], warped_prob=d["warped_prob"], homography=d["homography"])
        for kk, thr in ((300, 3), (50, 1)):                      # per-pair values through a one-file experiment
            one = Path(exper, "repeatability", f"golden_eval_one{k}")
            os.makedirs(one, exist_ok=True)
            np.savez(one / "p.npz", prob=d["prob"], warped_prob=d["warped_prob"], homography=d["homography"])
            ev[f"rep{k}_k{kk}_t{thr}"] = np.array(ref_det.compute_repeatability(f"golden_eval_one{k}", keep_k_points=kk, distance_thresh=thr))
        for kk in (1000, 60):
            r = ref_desc.compute_homography(d, keep_k_points=kk)
            ev[f"hom{k}_k{kk}_kp1"] = r["keypoints1"]
            ev[f"hom{k}_k{kk}_kp2"] = r["keypoints2"]
            ev[f"hom{k}_k{kk}_matches"] = np.array([[m.queryIdx, m.trainIdx] for m in r["matches"]], np.int64).reshape(-1, 2)
            ev[f"hom{k}_k{kk}_dist"] = np.array([m.distance for m in r["matches"]], np.float32)
            ev[f"hom{k}_k{kk}_score"] = np.array(r.get("matching_score", 0.0))
            ev[f"hom{k}_k{kk}_correct"] = np.array(r["correctness"])
            # robustness of the golden: the nearest neighbours must be separated far beyond fp32 rounding
            a = d["desc"][r["keypoints1"][:, 0], r["keypoints1"][:, 1]].astype(np.float64)
            b = d["warped_desc"][r["keypoints2"][:, 0], r["keypoints2"][:, 1]].astype(np.float64)
            dm = np.sqrt(np.maximum(((a[:, None, :] - b[None, :, :]) ** 2).sum(-1), 0))
            for mat in (dm, dm.T):
                srt = np.sort(mat, axis=1)
                assert (srt[:, 1] - srt[:, 0]).min() > 5e-5, "near-tie in the golden matching case: pick another seed"
    ev["rep_all_k300_t3"] = np.array(ref_det.compute_repeatability("golden_eval", keep_k_points=300, distance_thresh=3))
    np.savez_compressed(OUT / "eval_cases.npz", **ev)
    print({k: (v.tolist() if v.size == 1 else v.shape) for k, v in ev.items() if k.startswith("rep") or k.endswith("score")})

    # ---- 9. train-time reuse: Homographic_aug.__call__ / compute_valid_mask, detector_loss labels --
    from superpoint.utils.losses import detector_loss as ref_detector_loss
    tr = {}
    aug_t = Homographic_aug({"params": dict(HA_CFG["params"], allow_artifacts=False, patch_ratio=0.7), "valid_border_margin": 2}, "cpu")
    img255 = torch.from_numpy(smooth_image(96, 128, 71) * 255.0)[None, None]
    prs = np.random.RandomState(8)
    pts = torch.from_numpy(np.stack([prs.uniform(0, 95, 60), prs.uniform(0, 127, 60)], 1).astype(np.float32))
    np.random.seed(31)
    outa = aug_t(img255, pts)
    tr["aug_image"] = img255.numpy()
    tr["aug_points"] = pts.numpy()
    tr["aug_np_seed"] = np.array(31)
    tr["aug_warped"] = outa["warp"]["image"].numpy()
    tr["aug_kpts"] = outa["warp"]["kpts"].numpy()
    tr["aug_heatmap"] = outa["warp"]["kpts_heatmap"].numpy()
    tr["aug_valid_mask"] = outa["warp"]["valid_mask"].numpy()
    tr["aug_homography"] = outa["homography"].numpy()
    np.random.seed(32)
    Hb = torch.cat([aug_t.sample_homography((96, 128), **HA_CFG["params"]) for _ in range(3)])
    tr["vm_H"] = Hb.numpy()
    for er in (0, 2, 3):
        tr[f"vm_e{er}"] = aug_t.compute_valid_mask((96, 128), Hb, erosion=er).numpy().astype(np.uint8)
    g9 = torch.Generator().manual_seed(9)
    kmap = (torch.rand((2, 48, 64), generator=g9) < 0.03).to(torch.int32)
    vmask = (torch.rand((2, 48, 64), generator=g9) < 0.97).to(torch.int32)
    logits9 = torch.randn((2, 65, 6, 8), generator=g9)
    torch.manual_seed(77)
    tr["loss_ref_masked"] = ref_detector_loss(logits9, kmap, vmask, include_mask=True).numpy()
    torch.manual_seed(77)
    labels9, cells9, noise9 = O.detector_labels(kmap, vmask, include_mask=True)
    torch.manual_seed(77)
    assert torch.equal(O.detector_loss(logits9, kmap, vmask, include_mask=True), torch.as_tensor(tr["loss_ref_masked"]))
    tr.update(lab_kmap=kmap.numpy(), lab_valid=vmask.numpy(), lab_logits=logits9.numpy(), lab_noise=noise9.numpy(),
              lab_labels=labels9.numpy(), lab_cells=cells9.numpy())
    np.savez_compressed(OUT / "train_reuse.npz", **tr)

    # ---- 10. ExportNeRFDetections (export.py:225-366) on synthetic multi-view batches ---------------------------------
    import random
    from conftest import make_nerf_batch
    nf = {}
    sdn = O.make_state_dict("magicpoint", seed=13, logit_gain=12.0)
    mcfg_n = copy.deepcopy(MP_MODEL)
    mcfg_n["detector_head"]["top_k"] = 300
    modeln = SuperPoint(copy.deepcopy(mcfg_n)).eval()
    modeln.load_state_dict(sdn)
    cfgn = {"data": {"experiment_name": "golden_nerf"}, "model": mcfg_n}
    batches = [make_nerf_batch(0), make_nerf_batch(1, n_views=5)]
    random.seed(5)
    refexport.ExportNeRFDetections(cfgn, modeln, batches, "training", "cpu")
    for bt in batches:
        for nm in bt["name"]:
            nf[nm] = np.load(Path(exper, "outputs", "golden_nerf", "training", f"{nm}.npy"))
    nf["seed"] = np.array(13)
    nf["gain"] = np.array(12.0)
    nf["py_seed"] = np.array(5)
    np.savez_compressed(OUT / "nerf_export.npz", **nf)
    print("nerf export keypoints", {k: v.shape for k, v in nf.items() if k.startswith("view")})

    # ---- 11. ExportNeRFDetections.step (export.py:246-300) in isolation: a stub model returns a fixed heatmap, so the
    #          golden holds exactly what NMS -> nonzero -> warp_points_NeRF -> filter_points -> the splat loop produce ----
    ns = {}
    stepper = object.__new__(refexport.ExportNeRFDetections)
    stepper.config = {"model": {"detector_head": {"nms": 4, "det_thresh": 0.05, "top_k": 0}}}
    stepper.device = "cpu"
    ns["nms"], ns["det_thresh"], ns["top_k"] = np.array(4), np.array(0.05), np.array(0)
    for case, (bseed, nv, j, k) in enumerate([(0, 4, 0, 1), (0, 4, 2, 3), (1, 5, 4, 0)]):
        bt = make_nerf_batch(bseed, n_views=nv)
        h, w = bt["raw"]["image"].shape[-2:]
        rs = np.random.RandomState(900 + case)
        heat = np.zeros((h, w), np.float32)
        n_det = 150
        ry, rx = rs.randint(0, h, n_det), rs.randint(0, w, n_det)          # includes border detections (single-pixel copies)
        heat[ry, rx] = rs.uniform(0.06, 1.0, n_det).astype(np.float32)
        heat += rs.uniform(0.0, 0.01, (h, w)).astype(np.float32)           # background below det_thresh (patch contents)
        heat_t = torch.from_numpy(heat)
        stepper.model = lambda x, _h=heat_t: {"detector_output": {"prob_heatmap": _h.unsqueeze(0)}}
        probs0 = torch.zeros((1, 1, h, w))
        p_out, _ = stepper.step(bt["raw"]["image"][k:k + 1], probs0, torch.ones((1, 1, h, w)),
                                bt["raw"]["input_rotation"][j], bt["raw"]["input_translation"][j],
                                bt["raw"]["input_rotation"][k], bt["raw"]["input_translation"][k],
                                bt["raw"]["input_depth"][k], bt["camera_intrinsic_matrix"][j])
        ns[f"case{case}"] = np.array([bseed, nv, j, k])
        ns[f"heat{case}"] = heat
        ns[f"splat{case}"] = p_out[0, 1].numpy()
    ns["n"] = np.array(3)
    np.savez_compressed(OUT / "nerf_step.npz", **ns)
    print("nerf step splats", [int((ns[f"splat{c}"] > 0).sum()) for c in range(3)])

    # ---- 12. ExportDetections variants the GPU tests check against the oracle: aggregation 'max' (export.py:107-110)
    #          and enable_HA False (export.py:93-95: one plain forward, then the same NMS / threshold / nonzero) -----------
    hv = {}
    sdv = O.make_state_dict("magicpoint", seed=11, logit_gain=12.0)
    modelv = SuperPoint(copy.deepcopy(MP_MODEL)).eval()
    modelv.load_state_dict(sdv)
    imgv = torch.from_numpy(smooth_image(120, 160, 34))[None, None]
    orig_nms_v = refexport.box_nms
    orig_sample_v = Homographic_aug.sample_homography
    for tag, agg_mode, enable in (("max", "max", True), ("noha", "sum", False)):
        cap, used = {}, []

        def spy_nms_v(prob, _c=cap, **kw):
            _c["agg"] = prob.clone()
            return orig_nms_v(prob=prob, **kw)

        def spy_sample_v(self, *a, _u=used, **kw):
            h = orig_sample_v(self, *a, **kw)
            _u.append(h.clone())
            return h

        refexport.box_nms = spy_nms_v
        Homographic_aug.sample_homography = spy_sample_v
        cfgv = {"data": {"experiment_name": f"golden_{tag}"}, "homography_adaptation": dict(HA_CFG, aggregation=agg_mode),
                "model": copy.deepcopy(MP_MODEL)}
        np.random.seed(321)
        refexport.ExportDetections(cfgv, modelv, [{"raw": {"image": imgv}, "name": ["img0"]}], "training", enable, "cpu")
        refexport.box_nms = orig_nms_v
        Homographic_aug.sample_homography = orig_sample_v
        hv[f"{tag}_agg"] = cap["agg"].numpy()
        hv[f"{tag}_keypoints"] = np.load(Path(exper, "outputs", f"golden_{tag}", "training", "img0.npy"))
        hv[f"{tag}_H"] = torch.cat(used).numpy() if used else np.zeros((0, 3, 3), np.float32)
    hv.update(image=imgv.numpy(), seed=np.array(11), gain=np.array(12.0), np_seed=np.array(321))
    np.savez_compressed(OUT / "ha_export_variants.npz", **hv)
    print("ha_export variants", {k: v.shape for k, v in hv.items() if k.endswith("keypoints") or k.endswith("_H")})

    for f in sorted(OUT.glob("*.npz")):
        print(f.name, f.stat().st_size)


if __name__ == "__main__":
    main()
