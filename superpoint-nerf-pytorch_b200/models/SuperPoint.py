"""Drop-in for the reference's ``models/SuperPoint.py:5-30`` (+ VGG_Backbone.py, heads.py): same constructor
config, same state-dict keys and shapes, same output dictionary - computed by the sm_100a kernels.

The nn.Conv2d / nn.BatchNorm2d sub-modules only *hold* the parameters (so ``state_dict`` /
``load_state_dict`` / ``.to`` behave like the reference, engine.py:108-117); forward never calls them.
Inference only (the reference's export tasks call ``model.eval()``, export.py:21): BN uses running stats.

Extension keys (all optional) in the ``model`` config: ``precision`` in {'f16','fp32','bf16'} (default from
$SPN_B200_PRECISION, else 'f16': tcgen05 convolutions on fp16 operands with fp32 accumulation - the same 10-bit mantissa
as the TF32 convolutions the reference itself runs on any Ampere+ GPU, parity gate 5e-3 / >= 99 % keypoints, tested on
the export workload; 'fp32' selects the strict FFMA path with the 1e-4 gate), ``dense_desc`` (default True, as the reference), ``keypoints`` (default False): add
``detector_output['keypoints']`` (B,max_kp,2) int32 (row, col; row-major order like torch.nonzero of prob_heatmap_nms),
``detector_output['keypoint_count']`` (B,) and ``descriptor_output['desc_sparse']`` (B,max_kp,256) = desc[:, y, x] at the
keypoints - computed by ONE C-ABI call (spn_detect_describe: NMS once, no dense 315 MB descriptor map unless
``dense_desc`` is also set).  ``forward(x, keypoints=True)`` does the same per call.
"""
import itertools
import os

import torch
import torch.nn as nn

from .._native import MODES, Context, NativeError


class _Block(nn.Module):
    """Parameter holder with the reference's VGG_Block attribute names (VGG_Backbone.py:11-14)."""

    def __init__(self, cin, cout, k=3):
        super().__init__()
        self.conv2d = nn.Conv2d(cin, cout, kernel_size=k, stride=1, padding=k // 2)
        self.norm = nn.BatchNorm2d(cout)


class _Backbone(nn.Module):
    def __init__(self, cn):
        super().__init__()
        dims = [1] + list(cn)
        for i in range(8):
            setattr(self, f"block_{i + 1}", _Block(dims[i], dims[i + 1]))


class _DetectorHead(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.config = cfg
        self.convPa = _Block(cfg["detector_dim"][0], cfg["detector_dim"][1])
        self.convPb = _Block(cfg["detector_dim"][1], cfg["grid_size"] ** 2 + 1, k=1)


class _DescriptorHead(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.config = cfg
        self.convDa = _Block(cfg["descriptor_dim"][0], cfg["descriptor_dim"][1])
        self.convDb = _Block(cfg["descriptor_dim"][1], cfg["descriptor_dim"][1], k=1)


class SuperPoint(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        if list(config["vgg_cn"]) != [64, 64, 64, 64, 128, 128, 128, 128]:
            raise ValueError("only the reference's vgg_cn [64,64,64,64,128,128,128,128] is supported")
        if config["detector_head"]["grid_size"] != 8 or list(config["detector_head"]["detector_dim"]) != [128, 256]:
            raise ValueError("only grid_size 8 / detector_dim [128,256] is supported")
        self.backbone = _Backbone(config["vgg_cn"])
        self.detector_head = _DetectorHead(config["detector_head"])
        if config["model_name"].lower() == "superpoint":
            if list(config["descriptor_head"]["descriptor_dim"]) != [128, 256] or config["descriptor_head"]["grid_size"] != 8:
                raise ValueError("only descriptor_dim [128,256] / grid_size 8 is supported")
            self.descriptor_head = _DescriptorHead(config["descriptor_head"])
        prec = config.get("precision", os.environ.get("SPN_B200_PRECISION", "f16"))
        if prec not in MODES:
            raise ValueError(f"precision must be one of {sorted(MODES)}, got {prec!r}")
        self.mode = MODES[prec]
        self._ctxs = {}        # slot -> Context (slot 0 is the default; extra slots serve concurrent streams)
        self._packed_keys = {}

    # ---- native state ----------------------------------------------------------------------------
    def _weights_key(self):
        # (version counter, storage address) of every parameter and buffer: changes whenever load_state_dict / an
        # in-place update / .to() touches the weights, without building a state_dict on every call
        return tuple((v._version, v.data_ptr()) for v in itertools.chain(self.parameters(), self.buffers()))

    def native(self, slot: int = 0) -> Context:
        """The Context holding this model's packed weights (re-packed when parameters change).  Each ``slot`` is an
        independent context (own workspace), so different slots can run concurrently on different CUDA streams."""
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise NativeError("SuperPoint (b200) runs on CUDA only: move the model with .to('cuda') (no CPU fallback)")
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        ctx = self._ctxs.get(slot)
        if ctx is None or ctx.device != idx:
            with torch.cuda.device(idx):
                ctx = Context(idx)
            self._ctxs[slot] = ctx
            self._packed_keys[slot] = None
        key = self._weights_key()
        if key != self._packed_keys.get(slot):
            with torch.cuda.device(idx):
                ctx.load_state_dict(self.state_dict())
            self._packed_keys[slot] = key
        return ctx

    def train(self, mode=True):
        if mode:
            raise NativeError("the b200 SuperPoint is inference-only (training is out of scope, SURVEY.md section 2 #8)")
        return super().train(False)

    # ---- forward (models/SuperPoint.py:17-30) ----------------------------------------------------
    @torch.no_grad()
    def forward(self, x, mask=None, keypoints=None):
        if not (torch.is_tensor(x) and x.is_cuda):
            raise NativeError("input must be a CUDA tensor (no CPU fallback)")
        if x.dim() != 4 or x.shape[1] != 1:
            raise ValueError(f"expected (B,1,H,W), got {tuple(x.shape)}")
        B, _, H, W = x.shape
        ctx = self.native()
        dh = self.detector_head.config
        if keypoints is None:
            keypoints = bool(self.config.get("keypoints", False))
        if keypoints and dh["nms"] and mask is None:
            return self._forward_keypoints(ctx, x, B, H, W, dh)
        with torch.cuda.device(x.device):
            img = x.detach().to(torch.float32).contiguous().view(B, H, W)
            ctx.encoder_forward(img, self.mode)
            prob, logits = ctx.detector_head_forward(B, H, W, self.mode, mask=mask, want_logits=True)
            det = {"logits": logits, "prob_heatmap": prob}
            if dh["nms"]:
                r = ctx.box_nms(prob, float(dh["nms"]), 0.1, float(dh["det_thresh"]), int(dh["top_k"]),
                                det_thresh=float(dh["det_thresh"]), want_map=True, want_pred=True)
                det["prob_heatmap_nms"] = r["nms"]
                det["pred_pts"] = r["pred"]
            else:
                det["pred_pts"] = torch.ge(prob, dh["det_thresh"]).to(torch.int32)
            out = {"detector_output": det}
            if hasattr(self, "descriptor_head"):
                raw = ctx.descriptor_head_forward(B, H, W, self.mode)
                d = {"desc_raw": raw}
                if self.config.get("dense_desc", True):
                    d["desc"] = ctx.dense_descriptors(raw, self.descriptor_head.config["grid_size"])
                out["descriptor_output"] = d
        return out

    def _forward_keypoints(self, ctx, x, B, H, W, dh):
        """forward + post-NMS keypoints (+ descriptors at the keypoints) through one spn_detect_describe call."""
        sp = hasattr(self, "descriptor_head")
        with torch.cuda.device(x.device):
            img = x.detach().to(torch.float32).contiguous().view(B, H, W)
            r = ctx.detect_describe(img, self.mode, float(dh["nms"]), float(dh["det_thresh"]), int(dh["top_k"]), descriptors=sp,
                                    interp=self.config.get("desc_interp", "bicubic"))
            out = {"detector_output": {"logits": r["logits"], "prob_heatmap": r["prob"], "prob_heatmap_nms": r["nms"],
                                       "pred_pts": r["pred"], "keypoints": r["kp"], "keypoint_count": r["kp_count"]}}
            if sp:
                d = {"desc_raw": r["desc_raw"], "desc_sparse": r["desc_sparse"]}
                if self.config.get("dense_desc", False):
                    d["desc"] = ctx.dense_descriptors(r["desc_raw"], self.descriptor_head.config["grid_size"])
                out["descriptor_output"] = d
        return out

    @torch.no_grad()
    def prob_heatmap(self, images, mask=None, out=None, slot=0):
        """images (B,H,W) fp32 CUDA -> prob_heatmap (B,H,W), optionally multiplied by a u8 mask (export.py:69-70).
        Skips everything ExportDetections.step discards (logits copy, in-model NMS)."""
        ctx = self.native(slot)
        B, H, W = images.shape
        with torch.cuda.device(images.device):
            ctx.encoder_forward(images, self.mode)
            prob, _ = ctx.detector_head_forward(B, H, W, self.mode, mask=mask, want_logits=False, out=out)
        return prob

    @torch.no_grad()
    def prob_heatmap_ha(self, images, hinv, slot_begin, n_slots, mask=None, out=None, slot=0):
        """Heatmaps of the homography-adaptation slots [slot_begin, slot_begin + n_slots) of (images (NI,H,W),
        hinv (NI,n_h,3,3)) with the warp fused into the first convolution kernel (tensor-core modes only)."""
        ctx = self.native(slot)
        _, H, W = images.shape
        with torch.cuda.device(images.device):
            ctx.encoder_forward_ha(images, hinv, slot_begin, n_slots, self.mode)
            prob, _ = ctx.detector_head_forward(n_slots, H, W, self.mode, mask=mask, want_logits=False, out=out)
        return prob
