"""Restatement of the two kornia 0.7.0 functions the reference's export path calls.

TEST INFRASTRUCTURE (see oracle/__init__.py).  kornia is pinned by the reference at
0.7.0 (``/root/reference/superpoint/requirements.txt:2``) and cannot be installed in
this image.  The reference calls

* ``kornia.geometry.transform.warp_perspective`` at ``engine_solvers/export.py:51,53,55,72``
  and ``data/data_utils/homographic_augmentation.py:116,134``;
* ``kornia.morphology.erosion`` at ``engine_solvers/export.py:62,65`` and
  ``data/data_utils/homographic_augmentation.py:123``.

This module restates the published 0.7.0 algorithm of both in plain torch, and
``install()`` registers stub modules in ``sys.modules`` so the *unmodified* reference can
be imported and run (used by ``tests/golden/make_golden.py`` only).

Published algorithm restated (kornia 0.7.0):
  warp_perspective(src, M, dsize, mode, padding_mode='zeros', align_corners=True)
    N(h,w)   = [[2/(w-1), 0, -1], [0, 2/(h-1), -1], [0, 0, 1]]             (normal_transform_pixel)
    A        = N_dst @ (M @ inverse(N_src))                                 (normalize_homography)
    Ainv     = inverse(A)                                                   (fp32 torch.inverse)
    grid     = meshgrid of xs=(arange(w)/(w-1)-0.5)*2, ys likewise          (create_meshgrid)
    g        = [x, y, 1] @ Ainv^T ; g_xy = g[:2] * where(|g_z|>1e-8, 1/(g_z+1e-8), 1)   (transform_points)
    out      = F.grid_sample(src, g_xy, mode, 'zeros', align_corners)
  morphology.erosion(t, kernel)  (border_type='geodesic', max_val=1e4, engine='unfold')
    origin   = (kh//2, kw//2); pad (left=ox, right=kw-ox-1, top=oy, bottom=kh-oy-1) with +1e4
    out      = min over window of (padded - neighbourhood), neighbourhood = 0 where kernel==1, -1e4 where kernel==0
"""
from __future__ import annotations

import sys
import types

import torch
import torch.nn.functional as F


def _normal_transform_pixel(h: int, w: int, dtype, device) -> torch.Tensor:
    t = torch.tensor([[1.0, 0.0, -1.0], [0.0, 1.0, -1.0], [0.0, 0.0, 1.0]], dtype=dtype, device=device)
    wd = 1e-14 if w == 1 else w - 1.0
    hd = 1e-14 if h == 1 else h - 1.0
    t[0, 0] = t[0, 0] * 2.0 / wd
    t[1, 1] = t[1, 1] * 2.0 / hd
    return t.unsqueeze(0)


def _inverse_cast(x: torch.Tensor) -> torch.Tensor:
    dt = x.dtype
    if dt not in (torch.float32, torch.float64):
        x = x.to(torch.float32)
    return torch.inverse(x).to(dt)


def normalize_homography(M: torch.Tensor, dsize_src, dsize_dst) -> torch.Tensor:
    sh, sw = dsize_src
    dh, dw = dsize_dst
    n_src = _normal_transform_pixel(sh, sw, M.dtype, M.device)
    n_src_inv = _inverse_cast(n_src)
    n_dst = _normal_transform_pixel(dh, dw, M.dtype, M.device)
    return n_dst @ (M @ n_src_inv)


def create_meshgrid(h: int, w: int, dtype, device) -> torch.Tensor:
    xs = torch.linspace(0, w - 1, w, device=device, dtype=dtype)
    ys = torch.linspace(0, h - 1, h, device=device, dtype=dtype)
    xs = (xs / (w - 1) - 0.5) * 2
    ys = (ys / (h - 1) - 0.5) * 2
    gy, gx = torch.meshgrid(ys, xs, indexing="ij")
    return torch.stack([gx, gy], dim=-1).unsqueeze(0)  # 1,H,W,2 (x,y)


def transform_points(trans: torch.Tensor, pts: torch.Tensor) -> torch.Tensor:
    """trans (B,3,3); pts (B,H,W,2) -> (B,H,W,2)."""
    B, H, W, _ = pts.shape
    p = pts.reshape(B, H * W, 2)
    ph = F.pad(p, (0, 1), value=1.0)
    q = torch.bmm(ph, trans.permute(0, 2, 1))
    z = q[..., -1:]
    scale = torch.where(z.abs() > 1e-8, 1.0 / (z + 1e-8), torch.ones_like(z))
    return (scale * q[..., :-1]).reshape(B, H, W, 2)


def warp_perspective(src, M, dsize, mode="bilinear", padding_mode="zeros", align_corners=True, fill_value=None):
    if src.dim() != 4:
        raise ValueError(f"Input src must be a BxCxHxW tensor. Got {tuple(src.shape)}")
    if M.dim() != 3 or M.shape[-2:] != (3, 3):
        raise ValueError(f"Input M must be a Bx3x3 tensor. Got {tuple(M.shape)}")
    B, _, H, W = src.shape
    h_out, w_out = int(dsize[0]), int(dsize[1])
    A = normalize_homography(M, (H, W), (h_out, w_out))
    Ainv = _inverse_cast(A)
    grid = create_meshgrid(h_out, w_out, src.dtype, src.device).repeat(B, 1, 1, 1)
    grid = transform_points(Ainv, grid)
    return F.grid_sample(src, grid, mode=mode, padding_mode=padding_mode, align_corners=align_corners)


def erosion(tensor, kernel, structuring_element=None, origin=None, border_type="geodesic",
            border_value=0.0, max_val=1e4, engine="unfold"):
    if tensor.dim() != 4:
        raise ValueError(f"Input size must have 4 dimensions. Got {tensor.dim()}")
    kh, kw = kernel.shape
    if origin is None:
        origin = [kh // 2, kw // 2]
    pad = [origin[1], kw - origin[1] - 1, origin[0], kh - origin[0] - 1]
    if border_type == "geodesic":
        border_value = max_val
        border_type = "constant"
    out = F.pad(tensor, pad, mode=border_type, value=border_value)
    nb = torch.zeros_like(kernel)
    nb[kernel == 0] = -max_val
    out = out.unfold(2, kh, 1).unfold(3, kw, 1)
    out, _ = torch.min(out - nb, 4)
    out, _ = torch.min(out, 4)
    return out


def resize(input, size, interpolation="bilinear", align_corners=None, side="short", antialias=False):
    """kornia 0.7.0 geometry.transform.resize for an explicit (h, w) size (call sites: data/COCO.py:74,
    data/HPatches.py:70): promote to BCHW, F.interpolate(size, mode, align_corners, antialias), restore the rank."""
    if not isinstance(size, (tuple, list)):
        raise NotImplementedError("restated for explicit (h, w) sizes only (the reference's call sites)")
    shape = input.shape
    x = input.reshape((1,) * (4 - input.dim()) + tuple(shape)) if input.dim() < 4 else input
    h, w = int(size[0]), int(size[1])
    if tuple(x.shape[-2:]) == (h, w):
        return input
    out = F.interpolate(x, size=(h, w), mode=interpolation, align_corners=align_corners, antialias=antialias)
    return out.reshape(tuple(shape[:-2]) + (h, w))


def install(exper_path: str = "/tmp/spn_exper", data_path: str = "/tmp/spn_data", ckpt_path: str = "/tmp/spn_ckpt") -> None:
    """Register stub ``kornia`` / ``matplotlib`` / ``superpoint.settings`` modules so the unmodified
    reference imports.  Build-container use only."""
    k = types.ModuleType("kornia")
    kg = types.ModuleType("kornia.geometry")
    kgt = types.ModuleType("kornia.geometry.transform")
    km = types.ModuleType("kornia.morphology")
    kgt.warp_perspective = warp_perspective
    kgt.resize = resize
    km.erosion = erosion
    k.geometry, kg.transform, k.morphology = kg, kgt, km
    for name, mod in (("kornia", k), ("kornia.geometry", kg), ("kornia.geometry.transform", kgt), ("kornia.morphology", km)):
        sys.modules.setdefault(name, mod)
    try:
        import matplotlib.pyplot  # noqa: F401
    except Exception:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules.setdefault("matplotlib", mpl)
        sys.modules.setdefault("matplotlib.pyplot", plt)
    st = types.ModuleType("superpoint.settings")
    st.DATA_PATH, st.CKPT_PATH, st.EXPER_PATH = data_path, ckpt_path, exper_path
    sys.modules["superpoint.settings"] = st
