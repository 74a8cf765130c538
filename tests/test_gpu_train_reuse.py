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
