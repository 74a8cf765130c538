"""CPU restatement of the reference hot path (TEST INFRASTRUCTURE - see oracle/__init__.py).

Every function cites the reference file:line it follows (relative to
/root/reference/superpoint/superpoint/).  Arithmetic is torch-CPU fp32 (the reference's own CPU
arithmetic) plus numpy / scipy / cv2 / torchvision calls exactly where the reference makes them.

Nothing here is imported by the product package.
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

from . import kornia_shim as K

_HERE = Path(__file__).resolve().parent

# --------------------------------------------------------------------------------------------
# Layer table (models/model_utils/VGG_Backbone.py:44-58, models/model_utils/heads.py:11-13,54-56)
# name, cin, cout, ksize, relu, pool
# --------------------------------------------------------------------------------------------

def layer_table(vgg_cn=(64, 64, 64, 64, 128, 128, 128, 128), det_dim=(128, 256), grid=8, desc_dim=(128, 256),
                superpoint=False):
    c = list(vgg_cn)
    t = [
        ("backbone.block_1", 1, c[0], 3, True, False),
        ("backbone.block_2", c[0], c[1], 3, True, True),
        ("backbone.block_3", c[1], c[2], 3, True, False),
        ("backbone.block_4", c[2], c[3], 3, True, True),
        ("backbone.block_5", c[3], c[4], 3, True, False),
        ("backbone.block_6", c[4], c[5], 3, True, True),
        ("backbone.block_7", c[5], c[6], 3, True, False),
        ("backbone.block_8", c[6], c[7], 3, True, False),
        ("detector_head.convPa", det_dim[0], det_dim[1], 3, True, False),
        ("detector_head.convPb", det_dim[1], grid * grid + 1, 1, False, False),
    ]
    if superpoint:
        t += [
            ("descriptor_head.convDa", desc_dim[0], desc_dim[1], 3, True, False),
            ("descriptor_head.convDb", desc_dim[1], desc_dim[1], 1, False, False),
        ]
    return t


def make_state_dict(model_name="magicpoint", seed=0, logit_gain=1.0, randomize_bn=True):
    """Deterministic weights with the reference's state-dict keys/shapes (engine.py:108-117 loads by key).

    Generated with numpy so they do not depend on torch's init RNG.  ``randomize_bn`` gives non-trivial
    running stats / affine so that BN folding is exercised; ``logit_gain`` scales the last detector BN
    gamma so heatmaps become peaky (sparse NMS candidates) instead of ~1/65 flat.
    """
    rng = np.random.RandomState(seed)
    sd = {}
    for name, cin, cout, k, _relu, _pool in layer_table(superpoint=(model_name.lower() == "superpoint")):
        fan_in = cin * k * k
        bound = 1.0 / math.sqrt(fan_in)
        # kaiming-uniform-like scale so activations stay O(1) through the stack
        w = rng.uniform(-bound * math.sqrt(3.0), bound * math.sqrt(3.0), size=(cout, cin, k, k)).astype(np.float32)
        b = rng.uniform(-bound, bound, size=(cout,)).astype(np.float32)
        sd[f"{name}.conv2d.weight"] = torch.from_numpy(w)
        sd[f"{name}.conv2d.bias"] = torch.from_numpy(b)
        if randomize_bn:
            gamma = rng.uniform(0.8, 1.6, size=(cout,)).astype(np.float32)
            beta = rng.uniform(-0.2, 0.2, size=(cout,)).astype(np.float32)
            mean = rng.uniform(-0.2, 0.2, size=(cout,)).astype(np.float32)
            var = rng.uniform(0.5, 1.5, size=(cout,)).astype(np.float32)
        else:
            gamma = np.ones(cout, np.float32)
            beta = np.zeros(cout, np.float32)
            mean = np.zeros(cout, np.float32)
            var = np.ones(cout, np.float32)
        if name.endswith("convPb"):
            gamma = gamma * np.float32(logit_gain)
        sd[f"{name}.norm.weight"] = torch.from_numpy(gamma)
        sd[f"{name}.norm.bias"] = torch.from_numpy(beta)
        sd[f"{name}.norm.running_mean"] = torch.from_numpy(mean)
        sd[f"{name}.norm.running_var"] = torch.from_numpy(var)
        sd[f"{name}.norm.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
    return sd


# --------------------------------------------------------------------------------------------
# Model forward (models/SuperPoint.py:17-30)
# --------------------------------------------------------------------------------------------

def vgg_block(sd, name, x, k, relu, pool):
    """conv -> BN(eval) -> ReLU -> maxpool   (models/model_utils/VGG_Backbone.py:23-36)."""
    x = F.conv2d(x, sd[f"{name}.conv2d.weight"], sd[f"{name}.conv2d.bias"], stride=1, padding=(k - 1) // 2)
    x = F.batch_norm(x, sd[f"{name}.norm.running_mean"], sd[f"{name}.norm.running_var"],
                     sd[f"{name}.norm.weight"], sd[f"{name}.norm.bias"], training=False, eps=1e-5)
    if relu:
        x = F.relu(x)
    if pool:
        x = F.max_pool2d(x, kernel_size=2, stride=2)
    return x


def backbone_forward(sd, x):
    """VGG_BACKBONE.forward (models/model_utils/VGG_Backbone.py:60-71)."""
    for name, _ci, _co, k, relu, pool in layer_table()[:8]:
        x = vgg_block(sd, name, x, k, relu, pool)
    return x


def box_nms(prob, size, iou=0.1, min_prob=0.01, keep_top_k=0):
    """box_nms (models/model_utils/sp_utils.py:4-28): candidates >= min_prob, boxes pts -/+ size/2,
    torchvision.ops.nms(iou), optional top-k, scatter into a zero map."""
    import torchvision

    pts = torch.nonzero(prob >= min_prob, as_tuple=False)
    ptsf = pts.to(torch.float32)
    half = torch.tensor(size / 2.0)
    boxes = torch.cat((ptsf - half, ptsf + half), dim=1).to(torch.float32)
    scores = prob[pts[:, 0], pts[:, 1]]
    keep = torchvision.ops.nms(boxes=boxes, scores=scores, iou_threshold=iou)
    pts = pts[keep]
    scores = scores[keep]
    if keep_top_k:
        k = min(scores.shape[0], keep_top_k)
        scores, idx = torch.topk(scores, k)
        pts = pts[idx]
    out = torch.zeros_like(prob)
    out[pts[:, 0], pts[:, 1]] = scores
    return out


def detector_head_forward(sd, feat, grid=8, nms=0, det_thresh=0.015, top_k=0, nms_fn=None):
    """Detector_head.forward (models/model_utils/heads.py:17-44)."""
    nms_fn = nms_fn or box_nms
    out = {}
    x = vgg_block(sd, "detector_head.convPa", feat, 3, True, False)
    logits = vgg_block(sd, "detector_head.convPb", x, 1, False, False)
    out["logits"] = logits
    p = torch.softmax(logits, dim=1)[:, :-1]
    p = F.pixel_shuffle(p, grid).squeeze(1)
    out["prob_heatmap"] = p
    if nms:
        p = torch.stack([nms_fn(pb, nms, min_prob=det_thresh, keep_top_k=top_k) for pb in p])
        out["prob_heatmap_nms"] = p
    out["pred_pts"] = (p >= det_thresh).to(torch.int32)
    return out


def descriptor_head_forward(sd, feat, grid=8, dense=True):
    """Descriptor_head.forward (models/model_utils/heads.py:57-69)."""
    out = {}
    x = vgg_block(sd, "descriptor_head.convDa", feat, 3, True, False)
    raw = vgg_block(sd, "descriptor_head.convDb", x, 1, False, False)
    out["desc_raw"] = raw
    if dense:
        d = F.interpolate(raw, scale_factor=grid, mode="bicubic", align_corners=False)
        out["desc"] = F.normalize(d, p=2, dim=1)
    return out


@torch.no_grad()
def model_forward(sd, x, model_cfg, dense_desc=True, nms_fn=None):
    """SuperPoint.forward (models/SuperPoint.py:17-30).  ``model_cfg`` is the YAML ``model`` section."""
    dh = model_cfg["detector_head"]
    feat = backbone_forward(sd, x)
    out = {"detector_output": detector_head_forward(sd, feat, dh["grid_size"], dh["nms"], dh["det_thresh"],
                                                    dh["top_k"], nms_fn=nms_fn)}
    if model_cfg["model_name"].lower() == "superpoint":
        out["descriptor_output"] = descriptor_head_forward(sd, feat, model_cfg["descriptor_head"]["grid_size"], dense_desc)
    return out


# --------------------------------------------------------------------------------------------
# Greedy NMS restated in plain C (oracle/nms_ref.c), independent of torchvision
# --------------------------------------------------------------------------------------------

_nms_lib = None


def build_c(force=False):
    """Compile oracle/nms_ref.c -> oracle/libspn_oracle.so (gcc)."""
    so = _HERE / "libspn_oracle.so"
    src = _HERE / "nms_ref.c"
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", str(so), str(src), "-lm"])
    return so


def _load_c():
    global _nms_lib
    if _nms_lib is None:
        lib = ctypes.CDLL(str(build_c()))
        lib.spn_oracle_box_nms.restype = ctypes.c_int
        lib.spn_oracle_box_nms.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                           ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.c_void_p]
        _nms_lib = lib
    return _nms_lib


def box_nms_c(prob, size, iou=0.1, min_prob=0.01, keep_top_k=0):
    """Same contract as box_nms (sp_utils.py:4-28), greedy sorted NMS in C: stable sort by score desc,
    keep a candidate iff no earlier *kept* candidate has IoU > iou (torchvision nms semantics), top-k by
    (score desc, index asc)."""
    lib = _load_c()
    p = np.ascontiguousarray(prob.detach().cpu().numpy() if torch.is_tensor(prob) else prob, dtype=np.float32)
    out = np.zeros_like(p)
    n = lib.spn_oracle_box_nms(p.ctypes.data, p.shape[0], p.shape[1], float(size), float(iou), float(min_prob),
                               int(keep_top_k), out.ctypes.data)
    if n < 0:
        raise RuntimeError("spn_oracle_box_nms failed")
    return torch.from_numpy(out)


# --------------------------------------------------------------------------------------------
# Homography sampler (data/data_utils/homographic_augmentation.py:21-106)
# --------------------------------------------------------------------------------------------

def sample_homography_corners(translation=True, rotation=True, scaling=True, perspective=True, scaling_amplitude=0.1,
                              n_scales=5, n_angles=25, perspective_amplitude_x=0.1, perspective_amplitude_y=0.1,
                              patch_ratio=0.5, max_angle=1.57, allow_artifacts=False, translation_overflow=0.0):
    """Unit-square corner pairs (pts1, pts2) drawn from numpy's *global* RNG in the reference's call order
    (homographic_augmentation.py:28-95): truncnorm.rvs(1) x3, truncnorm.rvs(n_scales), randint, uniform x2, randint."""
    from scipy.stats import truncnorm

    margin = (1 - patch_ratio) / 2
    pts1 = margin + np.array([[0, 0], [0, patch_ratio], [patch_ratio, patch_ratio], [patch_ratio, 0]], dtype=np.float64)
    pts2 = pts1.copy()
    if perspective:
        ax, ay = perspective_amplitude_x, perspective_amplitude_y
        if not allow_artifacts:
            ax, ay = min(ax, margin), min(ay, margin)
        dy = truncnorm(-2, 2, loc=0.0, scale=ay / 2).rvs(1)
        dl = truncnorm(-2, 2, loc=0.0, scale=ax / 2).rvs(1)
        dr = truncnorm(-2, 2, loc=0.0, scale=ax / 2).rvs(1)
        pts2 += np.array([[dl, dy], [dl, -dy], [dr, dy], [dr, -dy]]).squeeze()
    if scaling:
        s = truncnorm(-2, 2, loc=1, scale=scaling_amplitude / 2).rvs(n_scales)
        s = np.concatenate((np.array([1]), s), axis=0)
        c = np.mean(pts2, axis=0, keepdims=True)
        scaled = (pts2 - c)[None] * s[:, None, None] + c
        if allow_artifacts:
            valid = np.arange(1, n_scales + 1)
        else:
            valid = np.where(((scaled >= 0.0) * (scaled <= 1.0)).prod(axis=1).prod(axis=1))[0]
        idx = valid[np.random.randint(valid.shape[0], size=1)].squeeze().astype(int)
        pts2 = scaled[idx]
    if translation:
        tmin, tmax = np.min(pts2, axis=0), np.min(1 - pts2, axis=0)
        if allow_artifacts:
            tmin = tmin + translation_overflow
            tmax = tmax + translation_overflow
        pts2 = pts2 + np.array([np.random.uniform(-tmin[0], tmax[0], 1), np.random.uniform(-tmin[1], tmax[1], 1)]).T
    if rotation:
        ang = np.linspace(-max_angle, max_angle, num=n_angles)
        ang = np.concatenate((np.array([0.0]), ang), axis=0)
        c = np.mean(pts2, axis=0, keepdims=True)
        rot = np.reshape(np.stack([np.cos(ang), -np.sin(ang), np.sin(ang), np.cos(ang)], axis=1), [-1, 2, 2])
        rotated = np.matmul((pts2 - c)[None], rot) + c
        if allow_artifacts:
            valid = np.arange(1, n_angles + 1)
        else:
            valid = np.where(((rotated >= 0.0) * (rotated <= 1.0)).prod(axis=1).prod(axis=1))[0]
        idx = valid[np.random.randint(valid.shape[0], size=1)].squeeze().astype(int)
        pts2 = rotated[idx]
    return pts1, pts2


def sample_homography(shape, **params):
    """Homographic_aug.sample_homography (homographic_augmentation.py:21-106) -> (1,3,3) fp32.
    Scale corners to pixels by (W,H), cv2.getPerspectiveTransform(f32,f32), fp32 torch.inverse."""
    import cv2

    pts1, pts2 = sample_homography_corners(**params)
    wh = np.array(tuple(shape)[::-1], dtype=np.float64)
    M = cv2.getPerspectiveTransform(np.float32(pts1 * wh[None]), np.float32(pts2 * wh[None]))
    return torch.inverse(torch.as_tensor(M, dtype=torch.float32).unsqueeze(0))


# --------------------------------------------------------------------------------------------
# Homography adaptation (engine_solvers/export.py:42-129)
# --------------------------------------------------------------------------------------------

def erosion_kernel(margin):
    """cv2.getStructuringElement(MORPH_ELLIPSE, (2*margin,)*2) as fp32 (export.py:59-60)."""
    import cv2

    return torch.as_tensor(cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (margin * 2,) * 2), dtype=torch.float32)


def ha_masks(H, shape, margin):
    """mask / count of one step (export.py:49-66): nearest warps of ones by H and H^-1, eroded, int32 (1,H,W)."""
    H_inv = torch.inverse(H)
    ones = torch.ones((1, 1, *shape), dtype=torch.float32, device=H.device)
    mask = K.warp_perspective(ones, H, dsize=shape, mode="nearest", align_corners=True)
    count = K.warp_perspective(ones, H_inv, dsize=shape, mode="nearest", align_corners=True)
    ker = erosion_kernel(margin).to(H.device)
    mask = K.erosion(mask, ker).to(torch.int32).squeeze(1)
    count = K.erosion(count, ker).to(torch.int32).squeeze(1)
    return mask, count, H_inv


@torch.no_grad()
def ha_step(prob_fn, image, H, margin):
    """One ExportDetections.step (export.py:42-79) for a given H: returns (prob_proj (1,H,W), count (1,H,W) int32,
    warped_image, mask)."""
    shape = tuple(image.shape[2:])
    mask, count, H_inv = ha_masks(H, shape, margin)
    warped = K.warp_perspective(image, H, dsize=shape, align_corners=True)
    prob = prob_fn(warped)  # (1,H,W)
    prob = prob * mask
    proj = K.warp_perspective(prob.unsqueeze(0), H_inv, dsize=shape, mode="bilinear", align_corners=True).squeeze(1)
    proj = proj * count
    return proj, count, warped, mask


@torch.no_grad()
def homography_adaptation(sd, image, config, homographies=None, full_forward=False, nms_fn=None, device=None):
    """ExportDetections.homography_adaptation body for one image (export.py:93-125).

    image (1,1,H,W) fp32.  ``homographies``: optional (num-1,3,3) fp32; when None they are drawn with
    ``sample_homography`` from numpy's global RNG, one per step, as the reference does (export.py:47).
    ``full_forward=True`` also runs the in-model box_nms whose output export.py:69 discards (used for the
    CPU baseline so that it pays what the reference pays).
    ``device``: run the same code on that torch device (the reference with ``device="cuda"``: cuDNN convolutions with
    torch's TF32 default, torchvision's CUDA nms) - used for bench.py's reference-on-GPU figure only.
    Returns dict(mean_prob (H,W), nms_prob (H,W), keypoints (N,2) int64, homographies (num-1,3,3))."""
    if device is not None:
        sd = {k: v.to(device) for k, v in sd.items()}
        image = image.to(device)
    ha = config["homography_adaptation"]
    mcfg = config["model"]
    dh = mcfg["detector_head"]
    shape = tuple(image.shape[2:])

    def prob_fn(x):
        if full_forward:
            return model_forward(sd, x, mcfg, dense_desc=True, nms_fn=nms_fn)["detector_output"]["prob_heatmap"]
        feat = backbone_forward(sd, x)
        return detector_head_forward(sd, feat, dh["grid_size"], nms=0)["prob_heatmap"]

    probs = [prob_fn(image)]
    counts = [torch.ones_like(probs[0])]
    used = []
    for i in range(ha["num"] - 1):
        H = (homographies[i:i + 1] if homographies is not None else sample_homography(shape, **ha["params"])).to(image.device)
        used.append(H)
        proj, count, _, _ = ha_step(prob_fn, image, H, ha["valid_border_margin"])
        probs.append(proj)
        counts.append(count.to(torch.float32))
    probs = torch.stack(probs, dim=1)  # 1,num,H,W
    counts = torch.stack(counts, dim=1)
    csum = torch.sum(counts, dim=1)
    if ha["aggregation"] == "max":
        agg = torch.max(probs, dim=1)[0]
    else:
        agg = torch.sum(probs, dim=1) / csum
    nms_fn = nms_fn or box_nms
    nmsp = nms_fn(agg[0], dh["nms"], min_prob=dh["det_thresh"], keep_top_k=dh["top_k"])
    kp = torch.nonzero((nmsp >= dh["det_thresh"]).to(torch.int32), as_tuple=False)
    return {"mean_prob": agg[0], "nms_prob": nmsp, "keypoints": kp.cpu().numpy(),
            "homographies": torch.cat(used, 0).cpu() if used else torch.zeros((0, 3, 3))}


# --------------------------------------------------------------------------------------------
# Pixel-space restatement of the kornia warp (SURVEY.md Appendix A "net effect") - used to cross-check the shim
# --------------------------------------------------------------------------------------------

def warp_pixelspace(src, M, mode="bilinear"):
    """out(p) = interp(src, M^-1 p) in pixel coordinates, zeros outside, fp64 coordinates.  src (1,1,H,W)."""
    _, _, H, W = src.shape
    Minv = torch.inverse(M[0].double())
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float64), torch.arange(W, dtype=torch.float64), indexing="ij")
    z = Minv[2, 0] * xs + Minv[2, 1] * ys + Minv[2, 2]
    sx = (Minv[0, 0] * xs + Minv[0, 1] * ys + Minv[0, 2]) / z
    sy = (Minv[1, 0] * xs + Minv[1, 1] * ys + Minv[1, 2]) / z
    img = src[0, 0].double()

    def at(yy, xx):
        ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
        v = img[yy.clamp(0, H - 1).long(), xx.clamp(0, W - 1).long()]
        return torch.where(ok, v, torch.zeros_like(v))

    if mode == "nearest":
        return at(torch.round(sy), torch.round(sx)).float()[None, None]
    x0, y0 = torch.floor(sx), torch.floor(sy)
    fx, fy = sx - x0, sy - y0
    out = (at(y0, x0) * (1 - fx) * (1 - fy) + at(y0, x0 + 1) * fx * (1 - fy)
           + at(y0 + 1, x0) * (1 - fx) * fy + at(y0 + 1, x0 + 1) * fx * fy)
    return out.float()[None, None]


# --------------------------------------------------------------------------------------------
# Sparse descriptors: dense desc[:, y, x] evaluated only at keypoints (heads.py:65-66 restated per point)
# --------------------------------------------------------------------------------------------

def _cubic_coeffs(t, A=-0.75):
    def c1(x):  # |x| <= 1
        return ((A + 2) * x - (A + 3)) * x * x + 1

    def c2(x):  # 1 < |x| < 2
        return ((A * x - 5 * A) * x + 8 * A) * x - 4 * A

    return [c2(t + 1.0), c1(t), c1(1.0 - t), c2(2.0 - t)]


def sparse_descriptors(desc_raw, pts, grid=8):
    """desc_raw (C,Hc,Wc) fp32, pts (N,2) int (row,col) at full resolution -> (N,C) L2-normalised.
    16-tap bicubic (A=-0.75, src=(dst+0.5)/grid-0.5, clamped indices) == F.interpolate(bicubic, align_corners=False)
    followed by F.normalize (eps 1e-12)."""
    C, Hc, Wc = desc_raw.shape
    out = torch.zeros((len(pts), C), dtype=torch.float32)
    for n, (r, c) in enumerate(np.asarray(pts).tolist()):
        sy = (r + 0.5) / grid - 0.5
        sx = (c + 0.5) / grid - 0.5
        iy, ix = math.floor(sy), math.floor(sx)
        wy = _cubic_coeffs(np.float32(sy - iy))
        wx = _cubic_coeffs(np.float32(sx - ix))
        acc = torch.zeros(C, dtype=torch.float32)
        for a in range(4):
            yy = min(max(iy - 1 + a, 0), Hc - 1)
            row = torch.zeros(C, dtype=torch.float32)
            for b in range(4):
                xx = min(max(ix - 1 + b, 0), Wc - 1)
                row = row + desc_raw[:, yy, xx] * float(wx[b])
            acc = acc + row * float(wy[a])
        out[n] = acc / acc.norm().clamp_min(1e-12)
    return out


# --------------------------------------------------------------------------------------------
# Loader pre-processing: data/COCO.py:66-76 + :135, data/HPatches.py:74-100
# --------------------------------------------------------------------------------------------

def ratio_preserving_resize(image, target, normalize=True):
    """COCO.ratio_preserving_resize (data/COCO.py:66-76) on a decoded grayscale image (H0,W0) fp32 [0,255], then the
    "/= 255." of COCO.py:135."""
    import torchvision.transforms.functional as TF

    target_size = torch.as_tensor(list(target), dtype=torch.int32)
    scales = torch.divide(target_size, torch.as_tensor(image.shape, dtype=torch.float32))
    new_size = (torch.as_tensor(image.shape[:2], dtype=torch.float32) * torch.max(scales)).to(torch.int32)
    out = K.resize(image, size=[new_size[0], new_size[1]], interpolation="bilinear", align_corners=False)
    out = TF.center_crop(out, output_size=[int(target_size[0]), int(target_size[1])])
    return out / 255.0 if normalize else out


def adapt_homography_to_resize(homography, image_shape, warped_image_shape, target):
    """HPatches.adapt_homography_to_resize (data/HPatches.py:74-100)."""
    source_size = torch.as_tensor(image_shape, dtype=torch.float32)
    source_warped_size = torch.as_tensor(warped_image_shape, dtype=torch.float32)
    target_size = torch.as_tensor(list(target), dtype=torch.float32)
    s = torch.max(torch.divide(target_size, source_size))
    up = torch.diag(torch.stack([1.0 / s, 1.0 / s, torch.tensor(1.0)]))
    ws = torch.max(torch.divide(target_size, source_warped_size))
    down = torch.diag(torch.stack([ws, ws, torch.tensor(1.0)]))
    t = torch.eye(3)
    t[0, -1] = ((source_size[1] * s - target_size[1]) / torch.tensor(2.0)).to(torch.int32)
    t[1, -1] = ((source_size[0] * s - target_size[0]) / torch.tensor(2.0)).to(torch.int32)
    wt = torch.eye(3)
    wt[0, -1] = -((source_warped_size[1] * ws - target_size[1]) / torch.tensor(2.0)).to(torch.int32)
    wt[1, -1] = -((source_warped_size[0] * ws - target_size[0]) / torch.tensor(2.0)).to(torch.int32)
    return wt @ down @ torch.as_tensor(homography, dtype=torch.float32) @ up @ t


# --------------------------------------------------------------------------------------------
# Train-time reuse: detector-loss label building, utils/losses.py:6-38
# --------------------------------------------------------------------------------------------

def detector_labels(kpts_heatmap, valid_mask=None, grid_size=8, include_mask=False, noise=None):
    """losses.py:13-27: per-cell class labels (argmax with a random U(0, 0.1) tie break) and the valid-cell mask.
    ``noise`` None draws it exactly like the reference (torch.distributions under torch's global RNG)."""
    labels = kpts_heatmap.unsqueeze(1).to(torch.float32)
    labels = torch.pixel_unshuffle(labels, grid_size)
    B, _, Hc, Wc = labels.shape
    labels = torch.cat([2 * labels, torch.ones(size=[B, 1, Hc, Wc])], dim=1)
    if noise is None:
        noise = torch.distributions.uniform.Uniform(0, 0.1).sample(labels.shape)
    labels = torch.argmax(labels + noise, dim=1)
    vm = torch.ones_like(kpts_heatmap) if (include_mask is False or valid_mask is None) else valid_mask
    vm = torch.prod(torch.pixel_unshuffle(vm.unsqueeze(1).to(torch.float32), grid_size), dim=1)
    return labels, vm, noise


def detector_loss(logits, kpts_heatmap, valid_mask, grid_size=8, include_mask=False, noise=None):
    """losses.py:6-38 (the loss itself, used only to pin ``detector_labels`` against the reference)."""
    labels, vm, _ = detector_labels(kpts_heatmap, valid_mask, grid_size, include_mask, noise)
    det = F.cross_entropy(logits, labels, reduction="none")
    return torch.mean(torch.divide(torch.sum(det * vm, dim=(1, 2)), torch.sum(vm, dim=(1, 2)) + 1e-10))
