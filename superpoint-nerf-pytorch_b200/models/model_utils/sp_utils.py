"""box_nms with the reference's signature (models/model_utils/sp_utils.py:4-28), computed by spn_box_nms_topk."""
import torch

from ..._native import NativeError, get_context


def box_nms(prob, size, iou=0.1, min_prob=0.01, keep_top_k=0):
    """prob (H,W) fp32 CUDA tensor -> (H,W) fp32: scores of the boxes that survive greedy IoU-NMS over
    ``size`` x ``size`` boxes centred on every pixel >= ``min_prob`` (optionally only the ``keep_top_k`` best),
    zeros elsewhere.  Bit-exact with the reference; ties in top-k resolve to the lower row-major index."""
    if not (torch.is_tensor(prob) and prob.is_cuda):
        raise NativeError("box_nms: prob must be a CUDA tensor (no CPU fallback)")
    if prob.dim() != 2:
        raise ValueError(f"box_nms expects (H,W), got {tuple(prob.shape)}")
    ctx = get_context(prob.device)
    p = prob.detach().to(torch.float32).contiguous().unsqueeze(0)
    return ctx.box_nms(p, float(size), float(iou), float(min_prob), int(keep_top_k), det_thresh=float(min_prob))["nms"][0]
