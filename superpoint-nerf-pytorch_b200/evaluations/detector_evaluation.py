"""Detector repeatability on the GPU, with the reference's interface (evaluations/detector_evaluation.py:10-18,145-238).

``compute_repeatability(exper_name, keep_k_points, distance_thresh)`` reads the ``.npz`` files written by
``Export_Hpatches_Repeatability`` (keys ``prob``, ``warped_prob``, ``homography``) and returns the mean repeatability.
Per pair, np.where / warp_keypoints / keep_true_keypoints / filter_keypoints / select_k_best run in
``spn_select_keypoints`` (fp64 warps like numpy) and the N1 x N2 nearest-neighbour counting in
``spn_repeatability_counts``; pairs of equal shape are batched into one launch each.
"""
from glob import glob
from os import path as osp

import numpy as np
import torch

from .. import settings
from .._native import get_context


def get_paths(exper_name, repeatability=False, MP_det_eval=False):
    """detector_evaluation.py:10-18."""
    if repeatability:
        return glob(osp.join(settings.EXPER_PATH, "repeatability/{}/*.npz".format(exper_name)))
    if MP_det_eval:
        return glob(osp.join(settings.EXPER_PATH, "MP_det_eval/{}/*.npz".format(exper_name)))
    return glob(osp.join(settings.EXPER_PATH, "outputs/{}/*.npz".format(exper_name)))


def repeatability_of_pairs(probs, warped_probs, homographies, keep_k_points=300, distance_thresh=3, device="cuda"):
    """probs / warped_probs: (B,H,W) / (B,H2,W2) NMS'd heatmaps (numpy or tensors), homographies (B,3,3) ->
    (repeatability per pair (B,) float64 with NaN where N1 + N2 == 0, counts (B,4) int {N1, N2, count1, count2})."""
    ctx = get_context(device)
    dev = torch.device("cuda", ctx.device)
    p1 = torch.as_tensor(np.asarray(probs), dtype=torch.float32).to(dev)
    p2 = torch.as_tensor(np.asarray(warped_probs), dtype=torch.float32).to(dev)
    H = np.asarray(homographies)
    Hinv = np.stack([np.linalg.inv(h) for h in H])          # detector_evaluation.py:195 (same dtype as the file's H)
    B = p1.shape[0]
    # warped_keypoints: detections of the warped image whose pre-image lies inside the first image (:192-196)
    w_pts, _, w_cnt = ctx.select_keypoints(p2, warp=Hinv.astype(np.float64), bounds=p1.shape[1:], emit_warped=False, keep_k=keep_k_points)
    # true_warped_keypoints: detections of the first image mapped by H, kept if inside the warped image (:198-203)
    t_pts, _, t_cnt = ctx.select_keypoints(p1, warp=H.astype(np.float64), bounds=p2.shape[1:], emit_warped=True, keep_k=keep_k_points)
    cand = torch.cat([w_cnt[B:], t_cnt[B:]])
    if int(cand.max()) > ctx.SELECT_CAP:
        raise RuntimeError(f"more than {ctx.SELECT_CAP} detections in a map: apply NMS / top_k before evaluating")
    counts = ctx.repeatability_counts(t_pts, t_cnt[:B].contiguous(), w_pts, w_cnt[:B].contiguous(), distance_thresh).cpu().numpy()
    n = counts[:, 0] + counts[:, 1]
    with np.errstate(invalid="ignore", divide="ignore"):
        rep = np.where(n > 0, (counts[:, 2] + counts[:, 3]) / np.maximum(n, 1), np.nan)
    return rep, counts


def compute_repeatability(exper_name, keep_k_points=300, distance_thresh=3, verbose=False, device="cuda"):
    """detector_evaluation.py:145-238."""
    paths = get_paths(exper_name, repeatability=True)
    groups = {}
    for path in paths:                                   # batch the pairs by shape
        data = np.load(path)
        key = (data["prob"].shape, data["warped_prob"].shape)
        groups.setdefault(key, []).append((data["prob"], data["warped_prob"], data["homography"]))
    repeatability, N1s, N2s = [], [], []
    for items in groups.values():
        rep, counts = repeatability_of_pairs(np.stack([i[0] for i in items]), np.stack([i[1] for i in items]),
                                             np.stack([i[2] for i in items]), keep_k_points, distance_thresh, device)
        N1s += counts[:, 0].tolist()
        N2s += counts[:, 1].tolist()
        repeatability += [r for r in rep if not np.isnan(r)]
    if verbose:
        print("Average number of points in the first image: " + str(np.mean(N1s)))
        print("Average number of points in the second image: " + str(np.mean(N2s)))
    return np.mean(repeatability)
