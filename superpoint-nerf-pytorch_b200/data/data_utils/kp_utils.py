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


def warp_points_NeRF(points, depth, cam_intrinsic_matrix, input_rotation, input_translation, warp_rotation, warp_translation, device="cpu"):
    """Reproject (N,2) (row, col) points of the view with pose (input_rotation, input_translation) and depth map ``depth``
    (B,H,W) into the view with pose (warp_rotation, warp_translation) (kp_utils.py:62-125 of the reference).

    Depth per point as in the reference: the pixel's own depth, except that away from the border the MINIMUM of the 5x5
    depth patch is taken when the patch spans >= 0.03 (an object edge).  The reference evaluates that rule in a Python
    loop over points; here it is the same rule on tensors (5x5 min / max by pooling, gathered at the points)."""
    import torch.nn.functional as F
    if len(points.shape) == 0:
        return points
    dev = points.device
    pi = points.to(torch.int64)                                       # int(p[0]), int(p[1]): truncation (points are >= 0)
    B, H, W = depth.shape
    dmax = F.max_pool2d(depth.unsqueeze(1), 5, stride=1, padding=2).squeeze(1)
    dmin = -F.max_pool2d(-depth.unsqueeze(1), 5, stride=1, padding=2).squeeze(1)
    own = depth[:, pi[:, 0], pi[:, 1]]                                # (B,N)
    mn, mx = dmin[:, pi[:, 0], pi[:, 1]], dmax[:, pi[:, 0], pi[:, 1]]
    border = (pi[:, 0] <= 2) | (pi[:, 1] <= 2) | (pi[:, 0] >= H - 2) | (pi[:, 1] >= W - 2)
    depth_values = torch.where(border.unsqueeze(0) | ((mx - mn) < 0.03), own, mn).unsqueeze(1)   # (B,1,N)
    pts = torch.fliplr(points)
    pts = torch.cat((pts, torch.ones((pts.shape[0], 1), device=dev, dtype=pts.dtype)), dim=1)
    warped = torch.tensordot(torch.linalg.inv(cam_intrinsic_matrix), pts, dims=([2], [1]))
    warped = warped / torch.linalg.norm(warped, dim=(1), keepdim=True)
    warped = warped * depth_values
    warped = input_rotation @ warped + input_translation
    warped = torch.linalg.inv(warp_rotation) @ warped - (torch.linalg.inv(warp_rotation) @ warp_translation)
    warped = cam_intrinsic_matrix @ warped
    warped = warped.transpose(2, 1)
    warped = warped[:, :, :2] / warped[:, :, 2:]
    return torch.flip(warped, dims=(2,)).squeeze(0)
