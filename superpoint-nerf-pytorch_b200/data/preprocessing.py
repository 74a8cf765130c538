"""Loader pre-processing that feeds the hot path, mirroring the reference's dataset methods:

    ratio_preserving_resize      <- data/COCO.py:66-76, data/HPatches.py:64-72 (+ the "/= 255." of COCO.py:135)
    adapt_homography_to_resize   <- data/HPatches.py:74-100

Image decoding stays with torchvision on the host (as in the reference); everything after the decode is one kernel
(spn_resize_crop) that takes the decoded uint8 image, so the host->device copy is 1 byte/pixel.
"""
import torch

from .._native import get_context


def resize_geometry(src_shape, target):
    """(new_h, new_w, crop_top, crop_left) exactly as the reference computes them: fp32 scale arithmetic truncated to
    int32 (COCO.py:70-73), torchvision centre-crop offsets int(round((size - crop) / 2))."""
    target_t = torch.as_tensor(list(target), dtype=torch.int32)
    shape_f = torch.as_tensor(list(src_shape), dtype=torch.float32)
    scales = torch.divide(target_t, shape_f)
    new = (shape_f * torch.max(scales)).to(torch.int32)
    nh, nw = int(new[0]), int(new[1])
    H, W = int(target[0]), int(target[1])
    # torchvision.transforms.functional.center_crop pads a too-small image symmetrically ((c - s) // 2 first) then crops
    pad_top = (H - nh) // 2 if nh < H else 0
    pad_left = (W - nw) // 2 if nw < W else 0
    ph, pw = max(nh, H), max(nw, W)
    top = int(round((ph - H) / 2.0)) - pad_top
    left = int(round((pw - W) / 2.0)) - pad_left
    return nh, nw, top, left


def ratio_preserving_resize(image, target, normalize=True):
    """image (H0,W0) uint8 or fp32 (CUDA, or CPU - it is uploaded) -> (H,W) fp32 CUDA, /255 when ``normalize``."""
    if image.dim() == 3 and image.shape[0] == 1:
        image = image[0]
    if not image.is_cuda:
        image = image.cuda(non_blocking=True)
    nh, nw, top, left = resize_geometry(image.shape, target)
    return get_context(image.device).resize_crop(image, nh, nw, top, left, int(target[0]), int(target[1]), 255.0 if normalize else 1.0)


def adapt_homography_to_resize(homography, image_shape, warped_image_shape, target):
    """HPatches ground-truth homography expressed between the resized + centre-cropped images (HPatches.py:74-100).
    3x3 host arithmetic in fp32, in the reference's operation order."""
    source_size = torch.as_tensor(image_shape, dtype=torch.float32)
    source_warped_size = torch.as_tensor(warped_image_shape, dtype=torch.float32)
    target_size = torch.as_tensor(list(target), dtype=torch.float32)
    s = torch.max(torch.divide(target_size, source_size))
    up_scale = torch.diag(torch.stack([1.0 / s, 1.0 / s, torch.tensor(1.0)]))
    warped_s = torch.max(torch.divide(target_size, source_warped_size))
    down_scale = torch.diag(torch.stack([warped_s, warped_s, torch.tensor(1.0)]))

    def shift(size, scale, sign):
        t = torch.eye(3, dtype=torch.float32)
        t[0, -1] = sign * ((size[1] * scale - target_size[1]) / torch.tensor(2.0)).to(torch.int32)
        t[1, -1] = sign * ((size[0] * scale - target_size[0]) / torch.tensor(2.0)).to(torch.int32)
        return t

    return shift(source_warped_size, warped_s, -1.0) @ down_scale @ torch.as_tensor(homography, dtype=torch.float32) @ up_scale \
        @ shift(source_size, s, 1.0)
