"""TEST INFRASTRUCTURE - CPU oracle for the rows SURVEY.md section 8 marks "next" (f-3 export evaluation, f-4 NeRF
splat).  numpy restatement of the reference's algorithms; only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg may import this package (the product never does).

Pinned: tests/test_oracle_golden.py checks every function below against tests/golden/eval_cases.npz and
nerf_step.npz, which tests/golden/make_golden.py produced by running the UNMODIFIED reference.

Reference code restated (file:line under /root/reference/superpoint/superpoint):
  select_points          evaluations/detector_evaluation.py:152-184 + descriptor_evaluation.py:23-50 (np.where ->
                         warp -> inside filter -> select_k_best)
  repeatability_pair     evaluations/detector_evaluation.py:186-233
  keep_shared_points     evaluations/descriptor_evaluation.py:17-52
  mutual_nn              evaluations/descriptor_evaluation.py:71-75 (cv2.BFMatcher(NORM_L2, crossCheck=True).match;
                         OpenCV is a third-party dependency: batchDistance + arg-min per query, cross check against the
                         arg-min per train descriptor, first index wins ties)
  order_matches          evaluations/descriptor_evaluation.py:76 (stable sort by ``distance < 0.25``)
  nerf_splat             engine_solvers/export.py:271-283 (sequential 3x3 patch / single pixel copies, later pairs
                         overwrite)
  nerf_point_depths,
  nerf_reproject_points  data/data_utils/kp_utils.py:68-127 (warp_points_NeRF), :3-20 (filter_points)
  nerf_step              engine_solvers/export.py:246-300 (ExportNeRFDetections.step minus the model call)
"""
from __future__ import annotations

import numpy as np


def _project(points_xy: np.ndarray, H: np.ndarray) -> np.ndarray:
    """(n,2) (x, y) -> (n,2) (x, y) under the 3x3 homography, numpy's promotion rules (float64 ones column)."""
    hom = np.concatenate([points_xy, np.ones((points_xy.shape[0], 1))], axis=1)
    q = hom @ np.transpose(H)
    return q[:, :2] / q[:, 2:]


def _inside(rc: np.ndarray, shape) -> np.ndarray:
    return (rc[:, 0] >= 0) & (rc[:, 0] < shape[0]) & (rc[:, 1] >= 0) & (rc[:, 1] < shape[1])


def _k_best(rows: np.ndarray, k: int) -> np.ndarray:
    """rows (n,3) = (a, b, probability): the min(k, n) most probable rows in ASCENDING probability, probability
    stripped.  np.argsort's default (unstable introsort) is what the reference calls; ties are not exercised by the
    goldens (continuous probabilities)."""
    ranked = rows[rows[:, 2].argsort(), :2]
    return ranked[ranked.shape[0] - min(k, ranked.shape[0]):, :]


def select_points(keypoint_map: np.ndarray, H: np.ndarray, bounds, k: int, emit_warped: bool) -> np.ndarray:
    """Detections (map > 0) in row-major order, mapped by H; kept if the mapped point lies inside ``bounds``;
    the k most probable returned as (row, col) float64 - the ORIGINAL coordinates (emit_warped False: the reference's
    keep_true_keypoints) or the MAPPED ones (emit_warped True: warp_keypoints + filter_keypoints)."""
    r, c = np.where(keypoint_map > 0)
    p = keypoint_map[r, c]
    mapped = _project(np.stack([c, r], axis=-1), H)[:, ::-1]           # back to (row, col)
    keep = _inside(mapped, bounds)
    coords = mapped if emit_warped else np.stack([r, c], axis=-1).astype(np.float64)
    return _k_best(np.concatenate([coords[keep], p[keep, None].astype(np.float64)], axis=1), k)


def repeatability_pair(prob, warped_prob, H, keep_k_points=300, distance_thresh=3):
    """-> (repeatability or nan, N1, N2, count1, count2) of one exported pair."""
    H = np.asarray(H)
    warped_kp = select_points(warped_prob, np.linalg.inv(H), prob.shape, keep_k_points, emit_warped=False)
    true_kp = select_points(prob, H, warped_prob.shape, keep_k_points, emit_warped=True)
    n1, n2 = true_kp.shape[0], warped_kp.shape[0]
    dist = np.linalg.norm(true_kp[:, None, :] - warped_kp[None, :, :], axis=2)
    c1 = int(np.sum(dist.min(axis=1) <= distance_thresh)) if n2 and n1 else 0
    c2 = int(np.sum(dist.min(axis=0) <= distance_thresh)) if n1 and n2 else 0
    rep = (c1 + c2) / (n1 + n2) if n1 + n2 > 0 else float("nan")
    return rep, n1, n2, c1, c2


def keep_shared_points(keypoint_map, H, keep_k_points=1000) -> np.ndarray:
    return select_points(keypoint_map, np.asarray(H), keypoint_map.shape, keep_k_points, emit_warped=False).astype(int)


def mutual_nn(desc1: np.ndarray, desc2: np.ndarray):
    """Cross-checked nearest neighbours under the L2 norm -> (pairs (m,2) int64 (query, train) in query order,
    distances (m,) float32)."""
    if len(desc1) == 0 or len(desc2) == 0:
        return np.zeros((0, 2), np.int64), np.zeros((0,), np.float32)
    a, b = np.asarray(desc1, np.float32), np.asarray(desc2, np.float32)
    # squared distances accumulated in fp32 like OpenCV's normL2Sqr, in blocks to bound memory
    d2 = np.empty((len(a), len(b)), np.float32)
    for s in range(0, len(a), 64):
        diff = a[s:s + 64, None, :] - b[None, :, :]
        d2[s:s + 64] = np.einsum("ijk,ijk->ij", diff, diff)
    fwd = d2.argmin(axis=1)                  # first minimum wins, like OpenCV's strict '<' scan
    bwd = d2.argmin(axis=0)
    q = np.arange(len(a))
    ok = bwd[fwd] == q
    return np.stack([q[ok], fwd[ok]], axis=1).astype(np.int64), np.sqrt(d2[q[ok], fwd[ok]]).astype(np.float32)


def order_matches(pairs: np.ndarray, dist: np.ndarray):
    """Python's sorted(matches, key=lambda m: m.distance < 0.25): stable, False (far) before True (near)."""
    order = np.argsort((dist < 0.25).astype(np.int8), kind="stable")
    return pairs[order], dist[order]


def nerf_splat(prob_src: np.ndarray, dst_pts: np.ndarray, src_pts: np.ndarray) -> np.ndarray:
    """out = zeros; for every (destination (row, col) float, source (row, col) int) pair IN ORDER: if either truncated
    point is within one pixel of the border copy the single source value, otherwise copy the source's 3x3 patch of
    ``prob_src`` onto the patch around the destination.  Later pairs overwrite earlier ones (order is part of the result)."""
    H, W = prob_src.shape
    out = np.zeros_like(prob_src)
    for d, s in zip(np.asarray(dst_pts), np.asarray(src_pts)):
        dy, dx, sy, sx = int(d[0]), int(d[1]), int(s[0]), int(s[1])
        near_border = min(dy, dx, sy, sx) <= 1 or max(dy, sy) >= H - 1 or max(dx, sx) >= W - 1
        if near_border:
            out[dy, dx] = prob_src[sy, sx]
        else:
            out[dy - 1:dy + 2, dx - 1:dx + 2] = prob_src[sy - 1:sy + 2, sx - 1:sx + 2]
    return out


def nerf_point_depths(points_rc, depth):
    """data/data_utils/kp_utils.py:83-106: depth used to un-project a detection = the depth map at the point, unless a
    full 5x5 window fits around it and spans >= 0.03 (a depth edge), in which case the window's minimum."""
    import torch
    H, W = depth.shape
    vals = []
    for r, c in points_rc.tolist():
        r, c = int(r), int(c)
        own = depth[r, c]
        if r <= 2 or c <= 2 or r >= H - 2 or c >= W - 2:
            vals.append(own)
            continue
        win = depth[r - 2:r + 3, c - 2:c + 3]
        lo, hi = win.min(), win.max()
        vals.append(lo if (hi - lo) >= 0.03 else own)
    return torch.tensor(vals) if vals else torch.zeros((0,))


def nerf_reproject_points(points_rc, depth_k, K, R_k, t_k, R_j, t_j):
    """data/data_utils/kp_utils.py:68-127 for one view pair: (N,2) int (row, col) detections of view k -> (N,2) fp32
    (row, col) in view j: pixel ray through K^-1, normalised, scaled by the point depth, k's pose -> world -> j's pose ->
    K.  The torch calls and their shapes follow the reference so the fp32 results (and their truncation) are its own."""
    import torch
    z = nerf_point_depths(points_rc, depth_k).unsqueeze(0).unsqueeze(1)                       # (1,1,N)
    xy1 = torch.cat((torch.fliplr(points_rc), torch.ones((points_rc.shape[0], 1))), dim=1)   # (N,3) (x, y, 1), fp32
    Kb, Rk, tk, Rj, tj = (m.unsqueeze(0) for m in (K, R_k, t_k, R_j, t_j))
    rays = torch.tensordot(torch.linalg.inv(Kb), xy1, dims=([2], [1]))                       # (1,3,N)
    rays /= torch.linalg.norm(rays, dim=1, keepdim=True)
    rays *= z
    world = Rk @ rays + tk
    cam_j = torch.linalg.inv(Rj) @ world - (torch.linalg.inv(Rj) @ tj)
    pix = (Kb @ cam_j).transpose(2, 1)
    pix = pix[:, :, :2] / pix[:, :, 2:]
    return torch.flip(pix, dims=(2,)).squeeze(0)


def nerf_step(prob_k, depth_k, K, R_k, t_k, R_j, t_j, nms, det_thresh, top_k, nms_fn):
    """engine_solvers/export.py:246-300 without the model call: view k's heatmap (H,W) torch fp32 -> the map splatted into
    view j's frame.  ``nms_fn`` = the box_nms oracle.  Keeps the reference's pairing quirk: the border filter (rows <
    H-1, cols < W-1, both >= 0) shortens the re-projected list, which is then zipped with the UNFILTERED detections."""
    import torch
    H, W = prob_k.shape
    kept = nms_fn(prob_k, nms, min_prob=det_thresh, keep_top_k=top_k)
    pts_k = torch.nonzero(torch.ge(kept, det_thresh).to(torch.int32), as_tuple=False)
    if len(pts_k) == 0:
        return np.zeros((H, W), np.float32), pts_k.numpy()
    moved = nerf_reproject_points(pts_k, depth_k, K, R_k, t_k, R_j, t_j)
    ok = (moved[:, 0] >= 0) & (moved[:, 0] < H - 1) & (moved[:, 1] >= 0) & (moved[:, 1] < W - 1)
    moved = moved[ok]
    return nerf_splat(prob_k.numpy(), moved.numpy(), pts_k.numpy()[:len(moved)]), pts_k.numpy()
