"""GPU tests of the callers and data formats either side of the hot path: dataset
classes + prefetching loader against the oracle's restatement of the reference loaders, packed keypoint export and the
sparse descriptor layout against the reference layouts."""
import copy
from pathlib import Path

import numpy as np
import pytest
import torch

from conftest import HA_CFG, MP_MODEL, SP_MODEL, smooth_image
from oracle import spn_oracle as O

pytestmark = pytest.mark.gpu

DATA_CFG = {"name": "COCO", "class_name": "COCO", "experiment_name": "io", "preprocessing": {"resize": [120, 160]}, "has_labels": False,
            "warped_pair": False, "batch_size": 1, "truncate": False,
            "augmentation": {"photometric": {"enable": False}, "homographic": {"enable": False}}}


def _model(cfg, sd, precision="fp32"):
    from superpoint_nerf_pytorch_b200.utils.get_model import get_model
    m = get_model(dict(copy.deepcopy(cfg), precision=precision), "cuda").eval()
    m.load_state_dict(sd)
    return m


def test_coco_loader_feeds_export_like_the_reference(tmp_path, monkeypatch):
    """JPEG files -> get_loader(export_pseudo_labels) -> the reference's batch dicts; pixels == the oracle's restatement of
    COCO.read_image + ratio_preserving_resize + /255 on the same decoded image (2e-6); then the export task end to end
    with per-image .npy AND the packed shard, which must hold the same arrays."""
    import cv2
    import torchvision
    from superpoint_nerf_pytorch_b200 import settings
    from superpoint_nerf_pytorch_b200.engine_solvers.export import ExportDetections, load_packed_labels
    from superpoint_nerf_pytorch_b200.utils.data_loaders import get_loader
    d = tmp_path / "data" / "COCO" / "images" / "training"
    d.mkdir(parents=True)
    sizes = [(240, 320), (213, 320), (320, 240), (187, 250), (120, 160), (480, 640)]
    for k, (h, w) in enumerate(sizes):
        cv2.imwrite(str(d / f"img{k:02d}.jpg"), (smooth_image(h, w, 300 + k) * 255).astype(np.uint8), [cv2.IMWRITE_JPEG_QUALITY, 92])
    monkeypatch.setattr(settings, "DATA_PATH", str(tmp_path / "data"))
    monkeypatch.setattr(settings, "EXPER_PATH", str(tmp_path / "exper"))
    cfg = {"data": dict(copy.deepcopy(DATA_CFG), packed=True), "homography_adaptation": dict(copy.deepcopy(HA_CFG), num=4, sampler="device", seed=3,
                                                                                             images_per_launch=4),
           "model": copy.deepcopy(MP_MODEL)}
    loader = get_loader(cfg, "export_pseudo_labels", device="cuda", export_split="training")
    assert len(loader) == len(sizes)
    seen = {}
    for batch in loader:
        img, name = batch["raw"]["image"], batch["name"]
        assert img.is_cuda and tuple(img.shape) == (1, 1, 120, 160) and img.dtype == torch.float32 and len(name) == 1
        dec = torchvision.io.decode_image(torchvision.io.read_file(str(d / f"{name[0]}.jpg")), torchvision.io.ImageReadMode.GRAY)
        want = O.ratio_preserving_resize(dec.squeeze(0).to(torch.float32), (120, 160)).numpy()
        assert np.abs(img[0, 0].cpu().numpy() - want).max() < 2e-6, name
        seen[name[0]] = True
    assert len(seen) == len(sizes)
    sd = O.make_state_dict("magicpoint", seed=5, logit_gain=10.0)
    ExportDetections(cfg, _model(MP_MODEL, sd, "f16"), loader, "training", True, "cuda")
    out = Path(tmp_path, "exper", "outputs", "io", "training")
    packed = load_packed_labels(out)
    assert sorted(packed) == sorted(seen)
    for name in seen:
        kp = np.load(out / f"{name}.npy")
        assert kp.dtype == np.int64 and np.array_equal(kp, packed[name]) and len(kp) > 0
    z = np.load(out / "packed_rank000.npz")
    assert int(z["global_offset"]) == 0 and z["offsets"][-1] == len(z["keypoints"]) == sum(len(v) for v in packed.values())
    assert z["counts_all_ranks"].tolist() == [[len(sizes), len(z["keypoints"])]]


def test_hpatches_loader_and_sparse_descriptor_layout(tmp_path, monkeypatch):
    """PPM pairs + H_1_k files -> HPatches items like the reference (images vs the oracle resize at 2e-6, homography vs
    the oracle's adapt_homography_to_resize bit-equal); Export_Hpatches_Descriptors in the reference's dense layout and in
    the sparse layout give the same matches through evaluations.descriptor_evaluation.compute_homography."""
    import cv2
    from superpoint_nerf_pytorch_b200 import settings
    from superpoint_nerf_pytorch_b200.engine_solvers.export import Export_Hpatches_Descriptors
    from superpoint_nerf_pytorch_b200.evaluations import descriptor_evaluation as E
    from superpoint_nerf_pytorch_b200.utils.data_loaders import get_loader
    root = tmp_path / "data" / "HPatches"
    rng = np.random.RandomState(0)
    for seq, (h, w) in (("v_a", (150, 200)), ("v_b", (135, 180)), ("i_c", (150, 200))):
        (root / seq).mkdir(parents=True)
        base = (smooth_image(h, w, 500 + len(seq)) * 255).astype(np.uint8)
        base = np.stack([base, np.roll(base, 3, 0), np.roll(base, 5, 1)], -1)            # .ppm holds BGR
        assert cv2.imwrite(str(root / seq / "1.ppm"), base)
        for i in range(2, 7):
            Hm = np.array([[1 + 0.02 * i, 0.01 * i, 2.0 * i], [-0.01 * i, 1 - 0.01 * i, -1.5 * i], [1e-5 * i, -2e-5, 1.0]])
            assert cv2.imwrite(str(root / seq / f"{i}.ppm"), cv2.warpPerspective(base, Hm, (w, h)))
            np.savetxt(str(root / seq / f"H_1_{i}"), Hm)
    monkeypatch.setattr(settings, "DATA_PATH", str(tmp_path / "data"))
    monkeypatch.setattr(settings, "EXPER_PATH", str(tmp_path / "exper"))
    dcfg = {"name": "HPatches", "class_name": "HPatches", "experiment_name": "hp_dense", "alteration": "v", "batch_size": 1,
            "preprocessing": {"resize": [120, 160]}}
    mcfg = copy.deepcopy(SP_MODEL)
    mcfg["detector_head"]["top_k"] = 200
    cfg = {"data": dcfg, "model": mcfg}
    loader = get_loader(cfg, "export_HPatches_Descriptors", device="cuda")
    items = list(loader)
    assert len(items) == 10 and all(it["name"][0].startswith("v_") for it in items)          # alteration filter: 2 sequences x 5
    it = items[0]
    seq, _, k = it["name"][0].rsplit("_", 2)
    raw1 = cv2.imread(str(root / seq / "1.ppm"), cv2.IMREAD_GRAYSCALE)
    rawk = cv2.imread(str(root / seq / f"{k}.ppm"), cv2.IMREAD_GRAYSCALE)
    assert tuple(it["image"].shape) == (1, 1, 120, 160) and tuple(it["homography"].shape) == (1, 3, 3)
    assert np.abs(it["image"][0, 0].cpu().numpy() - O.ratio_preserving_resize(torch.from_numpy(raw1).float(), (120, 160)).numpy()).max() < 2e-6
    assert np.abs(it["warped_image"][0, 0].cpu().numpy() - O.ratio_preserving_resize(torch.from_numpy(rawk).float(), (120, 160)).numpy()).max() < 2e-6
    want_h = O.adapt_homography_to_resize(np.loadtxt(str(root / seq / f"H_1_{k}")), torch.tensor(raw1.shape, dtype=torch.float32),
                                          torch.tensor(rawk.shape, dtype=torch.float32), (120, 160))
    assert torch.equal(it["homography"][0].cpu(), want_h)
    sd = O.make_state_dict("superpoint", seed=4, logit_gain=12.0)
    m = _model(mcfg, sd)
    Export_Hpatches_Descriptors(cfg, m, items[:3], "cuda")
    cfg_s = {"data": dict(dcfg, experiment_name="hp_sparse", sparse=True), "model": mcfg}
    Export_Hpatches_Descriptors(cfg_s, m, items[:3], "cuda")
    for it in items[:3]:
        name = it["name"][0]
        dense = np.load(Path(tmp_path, "exper", "descriptors", "hp_dense", f"{name}.npz"))
        sparse = np.load(Path(tmp_path, "exper", "descriptors", "hp_sparse", f"{name}.npz"))
        assert "desc" in dense.files and dense["desc"].shape == (120, 160, 256)
        assert "desc" not in sparse.files and sparse["desc_sparse"].shape[1] == 256 and len(sparse["keypoints"]) <= 200
        assert np.array_equal(sparse["keypoints"], np.argwhere(sparse["prob"] > 0))
        assert np.array_equal(sparse["prob"], dense["prob"]) and np.array_equal(sparse["warped_prob"], dense["warped_prob"])
        kp = sparse["keypoints"]
        assert np.abs(sparse["desc_sparse"] - dense["desc"][kp[:, 0], kp[:, 1]]).max() < 3e-5
        rd = E.compute_homography(dense, keep_k_points=150)
        rs = E.compute_homography(sparse, keep_k_points=150)
        md = [(a.queryIdx, a.trainIdx) for a in rd["matches"]]
        ms = [(a.queryIdx, a.trainIdx) for a in rs["matches"]]
        assert md == ms and np.array_equal(rd["keypoints1"], rs["keypoints1"]) and len(md) > 10
    import os
    sz_d = os.path.getsize(Path(tmp_path, "exper", "descriptors", "hp_dense", f"{items[0]['name'][0]}.npz"))
    sz_s = os.path.getsize(Path(tmp_path, "exper", "descriptors", "hp_sparse", f"{items[0]['name'][0]}.npz"))
    assert sz_s * 10 < sz_d
