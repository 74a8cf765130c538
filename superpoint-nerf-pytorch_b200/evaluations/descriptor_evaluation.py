"""Descriptor matching / homography estimation with the reference's interface
(evaluations/descriptor_evaluation.py:10-137).

``keep_shared_points`` (np.where + warp filter + select_k_best) runs in ``spn_select_keypoints``; the
``cv2.BFMatcher(cv2.NORM_L2, crossCheck=True)`` match of the two 1000 x 256 descriptor sets runs in
``spn_mutual_nn_match`` (tcgen05 GEMM + arg-min + cross check).  The RANSAC fit (``cv2.findHomography``) and the
corner-error correctness stay on the host, as in the reference.  Besides the reference's dense ``desc`` /
``warped_desc`` (H,W,256) arrays, ``compute_homography`` also accepts the sparse layout written by
``Export_Hpatches_Descriptors`` with ``sparse: true`` (keys ``keypoints``/``desc_sparse`` and the ``warped_`` twins).
"""
from glob import glob
from os import path as osp

import numpy as np
import torch

from .. import settings
from .._native import get_context


def get_paths(exper_name):
    """descriptor_evaluation.py:10-14."""
    return glob(osp.join(settings.EXPER_PATH, "descriptors/{}/*.npz".format(exper_name)))


def keep_shared_points(keypoint_map, H, keep_k_points=1000, device="cuda"):
    """descriptor_evaluation.py:17-52: (N,2) int (row, col), ascending probability, at most keep_k_points."""
    ctx = get_context(device)
    m = torch.as_tensor(np.asarray(keypoint_map), dtype=torch.float32).to(torch.device("cuda", ctx.device)).unsqueeze(0)
    pts, _, cnt = ctx.select_keypoints(m, warp=np.asarray(H, dtype=np.float64)[None], bounds=m.shape[1:], emit_warped=False,
                                       keep_k=keep_k_points)
    cnt = cnt.cpu().numpy()
    if cnt[1] > ctx.SELECT_CAP:
        raise RuntimeError(f"more than {ctx.SELECT_CAP} detections in the map: apply NMS / top_k before evaluating")
    return pts[0, :cnt[0]].cpu().numpy().astype(int)


def _lookup(data, key_dense, key_sparse, key_kp, keypoints):
    if key_dense in data:
        return np.ascontiguousarray(data[key_dense][keypoints[:, 0], keypoints[:, 1]], dtype=np.float32)
    kp = np.asarray(data[key_kp]).astype(np.int64).reshape(-1, 2)     # sparse layout: descriptors stored at their keypoints
    if len(keypoints) == 0:
        return np.zeros((0, np.asarray(data[key_sparse]).shape[-1]), np.float32)
    w = int(kp[:, 1].max()) + 1
    index = {int(r) * w + int(c): i for i, (r, c) in enumerate(kp)}
    rows = [index[int(r) * w + int(c)] for r, c in keypoints]
    return np.ascontiguousarray(np.asarray(data[key_sparse])[rows], dtype=np.float32)


def mutual_nn_matches(desc, warped_desc, device="cuda"):
    """cv2.BFMatcher(cv2.NORM_L2, crossCheck=True).match(desc, warped_desc) -> list of cv2.DMatch in query order."""
    import cv2
    ctx = get_context(device)
    dev = torch.device("cuda", ctx.device)
    if len(desc) == 0 or len(warped_desc) == 0:
        return []
    d1 = torch.as_tensor(desc, dtype=torch.float32).to(dev).unsqueeze(0)
    d2 = torch.as_tensor(warped_desc, dtype=torch.float32).to(dev).unsqueeze(0)
    n1 = torch.tensor([d1.shape[1]], dtype=torch.int32, device=dev)
    n2 = torch.tensor([d2.shape[1]], dtype=torch.int32, device=dev)
    match, dist = ctx.mutual_nn_match(d1, n1, d2, n2)
    match, dist = match[0].cpu().numpy(), dist[0].cpu().numpy()
    return [cv2.DMatch(int(i), int(j), float(dist[i])) for i, j in enumerate(match) if j >= 0]


def compute_homography(data, keep_k_points=1000, correctness_thresh=3, orb=False, device="cuda"):
    """descriptor_evaluation.py:55-137."""
    import cv2
    if orb:
        raise NotImplementedError("ORB (Hamming) matching is not part of the SuperPoint hot path")
    shape = data["prob"].shape
    real_H = data["homography"]
    keypoints = keep_shared_points(data["prob"], real_H, keep_k_points, device)
    warped_keypoints = keep_shared_points(data["warped_prob"], np.linalg.inv(real_H), keep_k_points, device)
    desc = _lookup(data, "desc", "desc_sparse", "keypoints", keypoints)
    warped_desc = _lookup(data, "warped_desc", "warped_desc_sparse", "warped_keypoints", warped_keypoints)
    matches = mutual_nn_matches(desc, warped_desc, device)
    matches = sorted(matches, key=lambda x: (x.distance < 0.25))
    matches_idx = np.array([m.queryIdx for m in matches])
    if len(matches_idx) == 0:
        return {"correctness": 0., "keypoints1": keypoints, "keypoints2": warped_keypoints, "matches": [], "inliers": [],
                "homography": None}
    m_keypoints = keypoints[matches_idx, :]
    matches_idx = np.array([m.trainIdx for m in matches])
    m_warped_keypoints = warped_keypoints[matches_idx, :]
    H, inliers = cv2.findHomography(m_keypoints[:, [1, 0]], m_warped_keypoints[:, [1, 0]], cv2.RANSAC, maxIters=3000)
    if H is None:
        return {"correctness": 0., "keypoints1": keypoints, "keypoints2": warped_keypoints, "matches": matches,
                "inliers": inliers, "homography": H}
    inliers = inliers.flatten()
    corners = np.array([[0, 0, 1], [shape[1] - 1, 0, 1], [0, shape[0] - 1, 1], [shape[1] - 1, shape[0] - 1, 1]])
    real_warped_corners = np.dot(corners, np.transpose(real_H))
    real_warped_corners = real_warped_corners[:, :2] / real_warped_corners[:, 2:]
    warped_corners = np.dot(corners, np.transpose(H))
    warped_corners = warped_corners[:, :2] / warped_corners[:, 2:]
    mean_dist = np.mean(np.linalg.norm(real_warped_corners - warped_corners, axis=1))
    correctness = float(mean_dist <= correctness_thresh)
    matching_score = len(m_keypoints) / len(keypoints)
    return {"correctness": correctness, "keypoints1": keypoints, "keypoints2": warped_keypoints, "matches": matches,
            "matching_score": matching_score, "mean_dist": mean_dist, "inliers": inliers, "homography": H}


def homography_estimation(exper_name, keep_k_points=1000, correctness_thresh=3, orb=False, device="cuda"):
    """descriptor_evaluation.py:140-157: mean correctness over the experiment's files."""
    correctness = []
    for path in get_paths(exper_name):
        data = np.load(path)
        correctness.append(compute_homography(data, keep_k_points, correctness_thresh, orb, device)["correctness"])
    return np.mean(correctness)
