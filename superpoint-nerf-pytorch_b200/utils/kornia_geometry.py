"""Host-side mirror of the matrix algebra inside ``kornia.geometry.transform.warp_perspective`` (kornia 0.7.0, the
version the reference pins in requirements.txt:2; call sites engine_solvers/export.py:51-55,72).

``K.warp_perspective(src, M, dsize)`` samples ``src`` at ``Ainv @ p`` in NORMALISED coordinates, where
``Ainv = torch.inverse(normalize_homography(M, (H, W), dsize))``.  The device kernels (csrc/geometry.cu) evaluate
that coordinate chain bit for bit, so given the reference's ``Ainv`` their validity masks / counts / warped pixels are
bit-identical to the reference's CPU run.  The 3x3 algebra below therefore uses exactly the torch calls kornia makes, one
(1,3,3) matrix at a time like the reference (LAPACK's inverse cannot be reproduced bit-exactly on the device); it is
plumbing on ~100 tiny matrices per image, not part of the hot path.  Homographies sampled on the device go through
``spn_kornia_matrices`` instead (same algebra, device arithmetic).
"""
import torch


def normal_transform_pixel(height, width, dtype=torch.float32):
    """kornia.geometry.conversions.normal_transform_pixel: pixel -> [-1, 1] (1,3,3)."""
    t = torch.tensor([[1.0, 0.0, -1.0], [0.0, 1.0, -1.0], [0.0, 0.0, 1.0]], dtype=dtype)
    width_denom = 1e-14 if width == 1 else width - 1.0
    height_denom = 1e-14 if height == 1 else height - 1.0
    t[0, 0] = t[0, 0] * 2.0 / width_denom
    t[1, 1] = t[1, 1] * 2.0 / height_denom
    return t.unsqueeze(0)


def normalize_homography(M, dsize_src, dsize_dst):
    """kornia normalize_homography: dst_norm_trans_dst_pixel @ (M @ src_pixel_trans_src_norm)."""
    src_norm_trans_src_pixel = normal_transform_pixel(dsize_src[0], dsize_src[1], M.dtype)
    src_pixel_trans_src_norm = torch.inverse(src_norm_trans_src_pixel)
    dst_norm_trans_dst_pixel = normal_transform_pixel(dsize_dst[0], dsize_dst[1], M.dtype)
    return dst_norm_trans_dst_pixel @ (M @ src_pixel_trans_src_norm)


def sampling_matrices(homographies, shape):
    """homographies (n,3,3) fp32 on the HOST (the H of export.py:47) -> (fwd, bwd), each (n,3,3) fp32:
    fwd[i] = inverse(normalize_homography(H_i))            the grid of warp(image / ones, H_i)     (export.py:51,53)
    bwd[i] = inverse(normalize_homography(inverse(H_i)))   the grid of warp(prob / ones, H_i^-1)   (export.py:49,55,72)"""
    h = homographies.detach().to("cpu", torch.float32).reshape(-1, 3, 3)
    shape = (int(shape[0]), int(shape[1]))
    fwd, bwd = torch.empty_like(h), torch.empty_like(h)
    for i in range(h.shape[0]):
        m = h[i:i + 1]
        fwd[i] = torch.inverse(normalize_homography(m, shape, shape))[0]
        bwd[i] = torch.inverse(normalize_homography(torch.inverse(m), shape, shape))[0]
    return fwd, bwd
