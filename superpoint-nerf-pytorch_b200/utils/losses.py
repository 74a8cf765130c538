"""Label building of the reference's detector loss (utils/losses.py:13-27) on the GPU kernel spn_detector_labels - the part
of the train-time path that shares data with the export (keypoint heatmaps in, per-cell class labels + valid-cell mask out).
The loss / optimiser themselves are outside this package."""
import torch

from .._native import get_context


def detector_labels(kpts_heatmap, valid_mask=None, grid_size=8, include_mask=False, noise=None, seed=0):
    """kpts_heatmap (B,H,W) int, valid_mask (B,H,W) int -> (labels (B,H/8,W/8) int64, valid cells (B,H/8,W/8) fp32).

    labels = argmax(cat([2 * pixel_unshuffle(heatmap), ones]) + U(0, 0.1) noise); ``noise`` (B,65,H/8,W/8) makes the
    random tie break reproducible (bit-identical to torch.argmax on the same noise), otherwise it is drawn on the device
    from ``seed``.  ``include_mask`` False ignores ``valid_mask`` like the reference (losses.py:23)."""
    if grid_size != 8:
        raise ValueError("only grid_size 8 is supported")
    ctx = get_context(kpts_heatmap.device)
    vm = valid_mask.to(torch.int32) if (include_mask and valid_mask is not None) else None
    return ctx.detector_labels(kpts_heatmap.to(torch.int32), vm, noise, seed)
