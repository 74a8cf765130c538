"""Keypoint helpers with the reference's interface (data/data_utils/kp_utils.py:3-71): O(N) tensor plumbing on the
keypoint lists that accompany the warp kernels (points are (row, col))."""
import torch


def filter_points(points, shape, device="cpu", return_mask=False):
    """Keep the points with 0 <= row < H-1 and 0 <= col < W-1 (kp_utils.py:3-20)."""
    if len(points) == 0:
        return points
    H, W = shape
    mask = (points[:, 0] >= 0) & (points[:, 0] < H - 1) & (points[:, 1] >= 0) & (points[:, 1] < W - 1)
    return (points[mask], mask) if return_mask else points[mask]


def compute_keypoint_map(points, shape, device="cpu"):
    """(N,2) points -> (H,W) int32 map with ones at the rounded in-range points (kp_utils.py:23-37)."""
    H, W = shape
    coord = torch.round(points).to(torch.int32)
    mask = (coord[:, 0] >= 0) & (coord[:, 0] < H - 1) & (coord[:, 1] >= 0) & (coord[:, 1] < W - 1)
    k_map = torch.zeros(tuple(shape), dtype=torch.int32, device=points.device if torch.is_tensor(points) else device)
    k_map[coord[mask, 0].long(), coord[mask, 1].long()] = 1
    return k_map


def warp_points(points, homography, device="cpu"):
    """(N,2) (row, col) points through (B,3,3) homographies acting on (x, y, 1) -> (B,N,2), squeezed for B = 1
    (kp_utils.py:40-71)."""
    if len(points.shape) == 0:
        return points
    pts = torch.fliplr(points)
    batch_size = homography.shape[0]
    pts = torch.cat((pts, torch.ones((pts.shape[0], 1), device=pts.device, dtype=pts.dtype)), dim=1)
    warped = torch.tensordot(homography, pts.transpose(1, 0), dims=([2], [0]))
    warped = warped.reshape([batch_size, 3, -1]).transpose(2, 1)
    warped = warped[:, :, :2] / warped[:, :, 2:]
    return torch.flip(warped, dims=(2,)).squeeze(0)
