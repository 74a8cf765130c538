"""Homography sampler with the reference's interface (data/data_utils/homographic_augmentation.py:14-106).

``sample_homography`` keeps the reference's host-side semantics *including its numpy-global-RNG draw order*, so
``np.random.seed(s)`` reproduces the reference's matrices bit for bit; ``sample_homographies_device`` is the
batched device sampler (spn_sample_homographies) used by the throughput path.  ``compute_valid_mask`` and ``__call__``
(the train-time augmentation of homographic_augmentation.py:109-151) run on the same spn_warp_batch kernel as the export:
warped image and eroded validity mask in one pass, bit-identical to the reference's CPU result.
"""
import numpy as np
import torch

from ..._native import get_context
from ...utils.kornia_geometry import sampling_matrices
from .kp_utils import compute_keypoint_map, filter_points, warp_points

_TN_LO, _TN_HI = -2.0, 2.0  # std_trunc = 2 (homographic_augmentation.py:25)


def _truncnorm(loc, scale, n):
    # same draws and values as scipy.stats.truncnorm(-2, 2, loc, scale).rvs(n): n uniforms -> ppf
    from scipy.stats import truncnorm

    return truncnorm.ppf(np.random.uniform(size=n), _TN_LO, _TN_HI) * scale + loc


def _pick(cands, valid_idx):
    return cands[valid_idx[np.random.randint(valid_idx.shape[0], size=1)].squeeze().astype(int)]


def _inside_unit(c):
    return np.where(((c >= 0.0) * (c <= 1.0)).prod(axis=1).prod(axis=1))[0]


def sample_corners(translation=True, rotation=True, scaling=True, perspective=True, scaling_amplitude=0.1, n_scales=5,
                   n_angles=25, perspective_amplitude_x=0.1, perspective_amplitude_y=0.1, patch_ratio=0.5, max_angle=1.57,
                   allow_artifacts=False, translation_overflow=0.0):
    """Unit-square source / destination corners (homographic_augmentation.py:28-95)."""
    m = (1 - patch_ratio) / 2
    src = m + np.array([[0, 0], [0, patch_ratio], [patch_ratio, patch_ratio], [patch_ratio, 0]], dtype=np.float64)
    dst = src.copy()
    if perspective:
        ax, ay = perspective_amplitude_x, perspective_amplitude_y
        if not allow_artifacts:
            ax, ay = min(ax, m), min(ay, m)
        py = _truncnorm(0.0, ay / 2, 1)
        hl = _truncnorm(0.0, ax / 2, 1)
        hr = _truncnorm(0.0, ax / 2, 1)
        dst += np.array([[hl, py], [hl, -py], [hr, py], [hr, -py]]).squeeze()
    if scaling:
        s = np.concatenate((np.array([1]), _truncnorm(1, scaling_amplitude / 2, n_scales)), axis=0)
        c = np.mean(dst, axis=0, keepdims=True)
        cands = (dst - c)[np.newaxis] * s[:, np.newaxis, np.newaxis] + c
        dst = _pick(cands, np.arange(1, n_scales + 1) if allow_artifacts else _inside_unit(cands))
    if translation:
        lo, hi = np.min(dst, axis=0), np.min(1 - dst, axis=0)
        if allow_artifacts:
            lo, hi = lo + translation_overflow, hi + translation_overflow
        dst = dst + np.array([np.random.uniform(-lo[0], hi[0], 1), np.random.uniform(-lo[1], hi[1], 1)]).T
    if rotation:
        a = np.concatenate((np.array([0.0]), np.linspace(-max_angle, max_angle, num=n_angles)), axis=0)
        c = np.mean(dst, axis=0, keepdims=True)
        R = np.reshape(np.stack([np.cos(a), -np.sin(a), np.sin(a), np.cos(a)], axis=1), [-1, 2, 2])
        cands = np.matmul((dst - c)[np.newaxis], R) + c
        dst = _pick(cands, np.arange(1, n_angles + 1) if allow_artifacts else _inside_unit(cands))
    return src, dst


def perspective_from_corners(src_px, dst_px):
    """cv2.getPerspectiveTransform(src, dst) restated: 8x8 linear solve in fp64 on fp32-rounded points."""
    s = np.asarray(src_px, np.float32).astype(np.float64)
    d = np.asarray(dst_px, np.float32).astype(np.float64)
    A = np.zeros((8, 8))
    b = np.zeros(8)
    for i in range(4):
        A[i, 0:3] = (s[i, 0], s[i, 1], 1.0)
        A[i, 6:8] = (-s[i, 0] * d[i, 0], -s[i, 1] * d[i, 0])
        A[i + 4, 3:6] = (s[i, 0], s[i, 1], 1.0)
        A[i + 4, 6:8] = (-s[i, 0] * d[i, 1], -s[i, 1] * d[i, 1])
        b[i], b[i + 4] = d[i, 0], d[i, 1]
    x = np.linalg.solve(A, b)
    return np.append(x, 1.0).reshape(3, 3)


class Homographic_aug:
    def __init__(self, config, device="cuda"):
        self.config = config["params"]
        self.erosion = config["valid_border_margin"]
        self.device = device

    def sample_homography(self, shape, **params):
        """-> (1,3,3) fp32 on ``self.device`` (homographic_augmentation.py:21-106)."""
        return self.sample_homography_host(shape, **params).to(self.device)

    def sample_homography_host(self, shape, **params):
        """The same matrix, left on the host (the export loop batches the upload of all homographies of a group)."""
        src, dst = sample_corners(**params)
        wh = np.array(tuple(shape)[::-1], dtype=np.float64)[np.newaxis]
        try:
            import cv2
            M = cv2.getPerspectiveTransform(np.float32(src * wh), np.float32(dst * wh))
        except ImportError:  # same linear system, numpy solver
            M = perspective_from_corners(src * wh, dst * wh)
        return torch.inverse(torch.as_tensor(M, dtype=torch.float32).unsqueeze(0))  # 3x3 on the host: plumbing

    def sample_homographies_device(self, shape, count, seed=0, first_index=0, **params):
        """Batched device sampler: -> (H (count,3,3), H_inv (count,3,3)) fp32 CUDA tensors."""
        p = dict(self.config)
        p.update(params)
        return get_context(self.device).sample_homographies(p, seed, first_index, count, int(shape[0]), int(shape[1]))

    def _warp(self, images, homography, erosion, want_warped):
        """images (B,H,W) CUDA fp32, homography (B,3,3) -> (warped (B,H,W) or None, mask (B,H,W) u8): one homography per
        image through spn_warp_batch (slot layout: image b owns slots 2b (identity) and 2b+1 (its warp))."""
        B, H, W = images.shape
        fwd, _ = sampling_matrices(homography, (H, W))                      # the reference's torch calls, on the host
        ctx = get_context(images.device)
        warped, mask = ctx.warp_batch(images, fwd.to(images.device).view(B, 1, 3, 3), int(erosion), want_warped=want_warped)
        return (warped.view(B, 2, H, W)[:, 1] if want_warped else None), mask.view(B, 2, H, W)[:, 1]

    def compute_valid_mask(self, shape, homography, erosion=2):
        """homographic_augmentation.py:109-127 -> (B,1,H,W) int32."""
        if homography.dim() == 2:
            homography = homography.unsqueeze(0)
        dev = torch.device(self.device)
        if dev.type != "cuda":
            raise RuntimeError("Homographic_aug runs its warps on CUDA only (no CPU fallback)")
        ones = torch.empty((homography.shape[0], int(shape[0]), int(shape[1])), dtype=torch.float32, device=dev)
        _, mask = self._warp(ones, homography, erosion, want_warped=False)
        return mask.unsqueeze(1).to(torch.int32)

    def __call__(self, image, points):
        """homographic_augmentation.py:130-151: image (1,1,H,W) on the device, points (N,2) (row, col)."""
        H, W = (int(v) for v in image.shape[2:])
        homography = self.sample_homography_host((H, W), **self.config)          # (1,3,3), numpy RNG order of the reference
        img = image.detach().to(self.device, torch.float32).contiguous().view(1, H, W)
        warped_image, mask = self._warp(img, homography, self.erosion, want_warped=True)
        homography = homography.to(img.device)
        warped_points = warp_points(points.to(img.device), homography, device=img.device)
        warped_points = filter_points(warped_points, (H, W), device=img.device)
        heatmap = compute_keypoint_map(warped_points, (H, W), device=img.device)
        return {"warp": {"image": warped_image.squeeze(), "kpts": warped_points, "kpts_heatmap": heatmap,
                         "valid_mask": mask.squeeze().to(torch.int32)},
                "homography": homography.squeeze()}
