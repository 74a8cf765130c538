"""GPU parity of the on-GPU export evaluations (csrc/eval.cu) against golden values produced by the UNMODIFIED reference
(evaluations/detector_evaluation.py: compute_repeatability, evaluations/descriptor_evaluation.py: compute_homography)
on the synthetic pairs of conftest.make_eval_pair, plus size-independent properties at HPatches size."""
import os

import numpy as np
import pytest
import torch

from conftest import make_eval_pair

pytestmark = pytest.mark.gpu


def test_repeatability_vs_reference_golden(golden, tmp_path, monkeypatch):
    from superpoint_nerf_pytorch_b200 import settings
    from superpoint_nerf_pytorch_b200.evaluations import detector_evaluation as E
    g = golden("eval_cases.npz")
    monkeypatch.setattr(settings, "EXPER_PATH", str(tmp_path))
    d_all = tmp_path / "repeatability" / "all"
    os.makedirs(d_all)
    for k, seed in enumerate(g["seeds"]):
        d = make_eval_pair(int(seed))
        np.savez(d_all / f"pair{k}.npz", prob=d["prob"], warped_prob=d["warped_prob"], homography=d["homography"])
        for kk, thr in ((300, 3), (50, 1)):
            rep, counts = E.repeatability_of_pairs(d["prob"][None], d["warped_prob"][None], d["homography"][None], kk, thr)
            assert rep[0] == float(g[f"rep{k}_k{kk}_t{thr}"]), (k, kk, thr, rep, counts)      # an exact rational: bit-equal
    assert abs(E.compute_repeatability("all", keep_k_points=300, distance_thresh=3) - float(g["rep_all_k300_t3"])) < 1e-12


def test_keep_shared_points_and_matching_vs_reference_golden(golden):
    """keep_shared_points: identical point lists (same order); cv2.BFMatcher(crossCheck) matches: identical pairs,
    matching_score identical; distances within 1e-5 (the tensor core truncates when it aligns products to the fp32
    accumulator: ~1e-6 systematic error on a 256-long dot product of unit vectors, i.e. ~20 bits, vs cv2's SIMD fp32)."""
    from superpoint_nerf_pytorch_b200.evaluations import descriptor_evaluation as E
    g = golden("eval_cases.npz")
    for k, seed in enumerate(g["seeds"]):
        d = make_eval_pair(int(seed))
        for kk in (1000, 60):
            r = E.compute_homography(d, keep_k_points=kk)
            assert np.array_equal(r["keypoints1"], g[f"hom{k}_k{kk}_kp1"])
            assert np.array_equal(r["keypoints2"], g[f"hom{k}_k{kk}_kp2"])
            got = np.array([[m.queryIdx, m.trainIdx] for m in r["matches"]], np.int64).reshape(-1, 2)
            want = g[f"hom{k}_k{kk}_matches"]
            # the reference sorts by (distance < 0.25) with a stable sort, so order is part of the contract
            assert np.array_equal(got, want), (k, kk)
            dist = np.array([m.distance for m in r["matches"]], np.float32)
            assert np.abs(dist - g[f"hom{k}_k{kk}_dist"]).max() < 1e-5
            assert r["matching_score"] == float(g[f"hom{k}_k{kk}_score"])
            assert r["correctness"] == float(g[f"hom{k}_k{kk}_correct"])
    # sparse file layout (keypoints + descriptors at the keypoints) gives the same matches as the dense maps
    d = make_eval_pair(int(g["seeds"][0]))
    kp1, kp2 = np.argwhere(d["prob"] > 0), np.argwhere(d["warped_prob"] > 0)
    sparse = {"prob": d["prob"], "warped_prob": d["warped_prob"], "homography": d["homography"], "keypoints": kp1,
              "desc_sparse": d["desc"][kp1[:, 0], kp1[:, 1]], "warped_keypoints": kp2,
              "warped_desc_sparse": d["warped_desc"][kp2[:, 0], kp2[:, 1]]}
    r = E.compute_homography(sparse, keep_k_points=1000)
    assert np.array_equal(np.array([[m.queryIdx, m.trainIdx] for m in r["matches"]]).reshape(-1, 2), g["hom0_k1000_matches"])


def test_mutual_nn_match_vs_cv2_and_properties():
    """1000 x 1000 x 256 (BASELINE config 3 size): against cv2.BFMatcher on the host where its nearest neighbours are
    well separated, plus properties that hold at any size: a matched pair is mutual, a set matched with itself is the
    identity with distance ~0, and padding rows beyond the counts are ignored."""
    import cv2
    import superpoint_nerf_pytorch_b200 as P
    ctx = P.get_context("cuda:0")
    rng = np.random.RandomState(0)
    a = rng.randn(1000, 256).astype(np.float32)
    b = np.concatenate([a[:700] + 0.3 * rng.randn(700, 256).astype(np.float32), rng.randn(300, 256).astype(np.float32)])[rng.permutation(1000)]
    a /= np.linalg.norm(a, axis=1, keepdims=True)
    b /= np.linalg.norm(b, axis=1, keepdims=True)
    da, db = torch.from_numpy(a).cuda()[None], torch.from_numpy(b).cuda()[None]
    pad = torch.full((1, 1100, 256), 7.0, device="cuda")             # garbage rows beyond the counts
    pad[:, :1000] = db
    n = torch.tensor([1000], dtype=torch.int32, device="cuda")
    match, dist = ctx.mutual_nn_match(da, n, pad, n)
    match, dist = match[0].cpu().numpy(), dist[0].cpu().numpy()
    ref = cv2.BFMatcher(cv2.NORM_L2, crossCheck=True).match(a, b)
    dm = np.sqrt(np.maximum(2 - 2 * (a.astype(np.float64) @ b.astype(np.float64).T), 0))
    gap_r = np.sort(dm, 1)[:, 1] - np.sort(dm, 1)[:, 0]
    gap_c = np.sort(dm.T, 1)[:, 1] - np.sort(dm.T, 1)[:, 0]
    want = {m.queryIdx: (m.trainIdx, m.distance) for m in ref}
    checked = 0
    for i in range(1000):
        j = want.get(i, (-1, 0))[0]
        jj = int(dm[i].argmin())
        if gap_r[i] > 3e-5 and gap_c[jj] > 3e-5 and (j < 0 or gap_c[j] > 3e-5):   # skip near-ties (fp32 rounding decides them)
            assert match[i] == j, (i, match[i], j)
            if j >= 0:
                assert abs(dist[i] - want[i][1]) < 1e-5
            checked += 1
    assert checked > 900 and (match >= 0).sum() > 500
    # mutual: a's match j has a as its nearest, in exact arithmetic up to the gap
    for i in np.where(match >= 0)[0][:200]:
        assert dm[i, match[i]] <= dm[i].min() + 3e-5 and dm[i, match[i]] <= dm[:, match[i]].min() + 3e-5
    # identity
    m2, d2 = ctx.mutual_nn_match(da, n, da, n)
    assert np.array_equal(m2[0].cpu().numpy(), np.arange(1000)) and float(d2.max()) < 4e-3   # sqrt of a ~4e-6 residual
    # ragged batch: counts smaller than the capacity, an empty set
    n1 = torch.tensor([37, 0], dtype=torch.int32, device="cuda")
    n2 = torch.tensor([129, 50], dtype=torch.int32, device="cuda")
    m3, _ = ctx.mutual_nn_match(torch.cat([da, da])[:, :200].contiguous(), n1, torch.cat([pad, pad])[:, :200].contiguous(), n2)
    m3 = m3.cpu().numpy()
    assert np.all(m3[0, 37:] == -1) and np.all(m3[1] == -1) and np.all(m3[0, :37] < 129)
    ref3 = {m.queryIdx: m.trainIdx for m in cv2.BFMatcher(cv2.NORM_L2, crossCheck=True).match(a[:37], b[:129])}
    agree = sum(int(m3[0, i] == ref3.get(i, -1)) for i in range(37))
    assert agree >= 36


def test_select_keypoints_full_size_properties():
    """480x640 NMS'd map with ~9000 detections: selection is the top-k by probability of the points whose warp stays
    inside, in ascending order; an oversize candidate set is reported."""
    import superpoint_nerf_pytorch_b200 as P
    ctx = P.get_context("cuda:0")
    rng = np.random.RandomState(3)
    H, W = 480, 640
    prob = np.zeros((H, W), np.float32)
    idx = rng.choice(H * W, 9000, replace=False)
    prob.flat[idx] = rng.uniform(0.01, 1.0, 9000).astype(np.float32)
    Hm = np.array([[0.95, 0.08, 12.0], [-0.06, 1.02, -9.0], [1e-4, -5e-5, 1.0]])
    pts, score, cnt = ctx.select_keypoints(torch.from_numpy(prob).cuda()[None], warp=Hm[None], emit_warped=False, keep_k=1000)
    cnt = cnt.cpu().numpy()
    ys, xs = np.where(prob > 0)
    q = np.stack([xs, ys, np.ones_like(xs)], 1) @ Hm.T
    wx, wy = q[:, 0] / q[:, 2], q[:, 1] / q[:, 2]
    inside = (wy >= 0) & (wy < H) & (wx >= 0) & (wx < W)
    assert cnt[1] == inside.sum() and cnt[0] == 1000
    pr = prob[ys[inside], xs[inside]]
    order = np.argsort(pr, kind="stable")[-1000:]
    want = np.stack([ys[inside][order], xs[inside][order]], 1)
    assert np.array_equal(pts[0].cpu().numpy().astype(np.int64), want)
    assert np.array_equal(score[0].cpu().numpy(), pr[order])
    dense = torch.rand((1, 200, 200), device="cuda") + 0.1           # 40000 candidates > 16384: must be flagged
    _, _, c2 = ctx.select_keypoints(dense, keep_k=10)
    assert int(c2[1]) == 40000 > ctx.SELECT_CAP


def test_repeatability_vs_oracle_fresh_pairs_and_degenerate_maps():
    """Batched spn_select_keypoints + spn_repeatability_counts against the CPU oracle (oracle/eval_oracle.py, pinned to the
    reference by tests/test_oracle_golden.py) on pairs the goldens do not hold: 240x320 (the repeatability config's size),
    several k / thresholds, one pair without detections in the first image and one without any."""
    from oracle import eval_oracle as O
    from superpoint_nerf_pytorch_b200.evaluations import detector_evaluation as E
    pairs = [make_eval_pair(s, h=240, w=320, n_pts=900) for s in (11, 12, 13)]
    empty_first = dict(pairs[0], prob=np.zeros_like(pairs[0]["prob"]))
    empty_both = dict(empty_first, warped_prob=np.zeros_like(pairs[0]["prob"]))
    pairs += [empty_first, empty_both]
    probs, warped, hs = (np.stack([p[k] for p in pairs]) for k in ("prob", "warped_prob", "homography"))
    for kk, thr in ((300, 3), (1000, 1), (25, 5)):
        rep, counts = E.repeatability_of_pairs(probs, warped, hs, kk, thr)
        for i, p in enumerate(pairs):
            want = O.repeatability_pair(p["prob"], p["warped_prob"], p["homography"], kk, thr)
            assert tuple(int(v) for v in counts[i]) == want[1:], (i, kk, thr, counts[i], want)
            assert (np.isnan(rep[i]) and np.isnan(want[0])) or rep[i] == want[0], (i, kk, thr)


def test_keep_shared_points_and_matches_vs_oracle_fresh_pairs():
    """keep_shared_points + mutual_nn_matches against the CPU oracle at 240x320 with ~700 detections per image.  The seeds
    are chosen so that every nearest neighbour is separated from the runner-up by > 8e-5 (fp32 rounding cannot decide a
    match; the tensor-core distances are within ~1e-6)."""
    from oracle import eval_oracle as O
    from superpoint_nerf_pytorch_b200.evaluations import descriptor_evaluation as E
    for seed in (23, 25):
        d = make_eval_pair(seed, h=240, w=320, n_pts=900)
        kp1 = E.keep_shared_points(d["prob"], d["homography"], 1000)
        kp2 = E.keep_shared_points(d["warped_prob"], np.linalg.inv(d["homography"]), 1000)
        assert np.array_equal(kp1, O.keep_shared_points(d["prob"], d["homography"], 1000))
        assert np.array_equal(kp2, O.keep_shared_points(d["warped_prob"], np.linalg.inv(d["homography"]), 1000))
        a, b = d["desc"][kp1[:, 0], kp1[:, 1]], d["warped_desc"][kp2[:, 0], kp2[:, 1]]
        want_pairs, want_dist = O.mutual_nn(a, b)
        got = E.mutual_nn_matches(a, b)
        got_pairs = np.array([[m.queryIdx, m.trainIdx] for m in got], np.int64).reshape(-1, 2)
        assert np.array_equal(got_pairs, want_pairs), seed
        assert np.abs(np.array([m.distance for m in got], np.float32) - want_dist).max() < 1e-5
        assert len(got) > 300
