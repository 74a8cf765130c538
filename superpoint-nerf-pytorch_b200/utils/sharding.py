"""Image sharding across ranks: every image with its homographies is independent, so rank r
takes image ids congruent to r modulo the world size and there is no data-path collective.  The only exchange is
one all_gather of (images written, keypoints written) per rank for the export summary / packed-export offsets."""
import torch
import torch.distributed as dist


def shard_indices(n_items: int, rank: int, world: int):
    return range(rank, n_items, world)


def gather_export_counts(n_images: int, n_keypoints: int, device="cpu"):
    """all_gather of int64[2] per rank -> (world,2) tensor on ``device``; offsets = exclusive cumsum of column 1."""
    mine = torch.tensor([n_images, n_keypoints], dtype=torch.int64, device=device)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        allc = mine.unsqueeze(0)
    else:
        parts = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
        dist.all_gather(parts, mine)
        allc = torch.stack(parts)
    offsets = torch.cumsum(allc[:, 1], 0) - allc[:, 1]
    return allc, offsets
