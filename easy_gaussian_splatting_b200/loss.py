"""§8f-4: fused L1 + SSIM photometric loss — the reference's ``LossComputer``
(/root/reference/model/gaussian.py:415-453) as two hand-written sm_100a kernels.

``LossComputer.get_loss_dict(render_img, gt_img, mask)`` blends the mask
(``mask * gt + (1 - mask) * render``, gaussian.py:428-429), takes ``F.l1_loss`` (:447-448) and
``1 - torchmetrics.StructuralSimilarityIndexMeasure(data_range=1.0)(gt, render)`` (:419, :450-453) and mixes them
``(1 - lambda_ssim) * l1 + lambda_ssim * ssim`` (:437) — about 25 torch kernels forward and as many backward per
step at full resolution.  Here the forward is one kernel (per-image sums + the SSIM partial-derivative maps) and
the backward one kernel; images stay ``[H, W, 3]`` exactly as ``rasterization()`` returns them.
There is no CPU path: CPU tensors raise.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

from . import _lib
from .stages import _f32c, _ptr, _stream

__all__ = ["fused_l1_ssim_loss", "FusedLossComputer"]

_WIN = 11


class _L1SSIM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, render, gt, mask, lambda_ssim):
        lib = _lib.load()
        C, H, W, _ = render.shape
        dev = render.device
        need_grad = render.requires_grad
        sums = torch.zeros(C, 2, dtype=torch.float64, device=dev)
        maps = torch.empty(3, C, 3, H - _WIN + 1, W - _WIN + 1, dtype=torch.float32, device=dev) if need_grad else None
        with torch.cuda.device(dev):
            rc = lib.egs_l1_ssim_fwd(C, H, W, _ptr(render), _ptr(gt), _ptr(mask), _ptr(maps), _ptr(sums), _stream(dev))
        _lib.check(rc, "egs_l1_ssim_fwd")
        l1 = (sums[:, 0] / float(3 * H * W)).float()
        ssim = (1.0 - sums[:, 1] / float(3 * (H - _WIN + 1) * (W - _WIN + 1))).float()
        total = (1.0 - lambda_ssim) * l1 + lambda_ssim * ssim
        ctx.lambda_ssim = float(lambda_ssim)
        ctx.save_for_backward(render, gt, mask, maps)
        ctx.mark_non_differentiable(l1, ssim)
        return total, l1, ssim

    @staticmethod
    def backward(ctx, v_total, _v_l1, _v_ssim):
        render, gt, mask, maps = ctx.saved_tensors
        lib = _lib.load()
        C, H, W, _ = render.shape
        dev = render.device
        v_total = v_total.to(torch.float32).contiguous()
        v_render = torch.empty_like(render)
        with torch.cuda.device(dev):
            rc = lib.egs_l1_ssim_bwd(C, H, W, _ptr(render), _ptr(gt), _ptr(mask), _ptr(maps), ctx.lambda_ssim,
                                     _ptr(v_total), _ptr(v_render), _stream(dev))
        _lib.check(rc, "egs_l1_ssim_bwd")
        return v_render, None, None, None


def fused_l1_ssim_loss(render_img: Tensor, gt_img: Tensor, mask: Optional[Tensor] = None,
                       lambda_ssim: float = 0.2) -> Tuple[Tensor, Tensor, Tensor]:
    """``render_img``, ``gt_img``: ``[H,W,3]`` or ``[C,H,W,3]`` fp32 CUDA; ``mask``: ``[H,W]`` / ``[C,H,W]`` (1 = take
    the ground truth, gaussian.py:428-429) or None.  Returns ``(total, l1, ssim_loss)`` — scalars for a single image,
    ``[C]`` vectors for a batch (one loss per image, as the reference computes one per step).  ``total`` carries the
    gradient to ``render_img``; ``l1`` and ``ssim_loss`` are the logging values of ``loss_dict`` and are detached."""
    single = render_img.dim() == 3
    if single:
        render_img, gt_img = render_img[None], gt_img[None]
        mask = None if mask is None else mask[None]
    if render_img.dim() != 4 or render_img.shape[-1] != 3 or gt_img.shape != render_img.shape:
        raise ValueError(f"render_img / gt_img must both be [H,W,3] or [C,H,W,3], got {tuple(render_img.shape)} "
                         f"and {tuple(gt_img.shape)}")
    C, H, W, _ = render_img.shape
    if H < _WIN or W < _WIN:
        raise ValueError(f"images must be at least {_WIN}x{_WIN} (the SSIM window), got {W}x{H}")
    if mask is not None and mask.shape != (C, H, W):
        raise ValueError(f"mask must be [H,W] / [C,H,W] matching the images, got {tuple(mask.shape)}")
    render_c = _f32c(render_img, "render_img")
    gt_c = _f32c(gt_img.detach(), "gt_img")
    mask_c = None if mask is None else _f32c(mask.detach(), "mask")
    if gt_c.device != render_c.device or (mask_c is not None and mask_c.device != render_c.device):
        raise RuntimeError("render_img, gt_img and mask must be on the same device")
    total, l1, ssim = _L1SSIM.apply(render_c, gt_c, mask_c, float(lambda_ssim))
    if single:
        return total[0], l1[0], ssim[0]
    return total, l1, ssim


class FusedLossComputer:
    """Drop-in for the reference's ``LossComputer`` (/root/reference/model/gaussian.py:415-445): same constructor,
    same ``get_loss_dict(render_img, gt_img, mask) -> {"l1", "ssim", ["scale_reg"], "total"}``."""

    def __init__(self, model, lambda_ssim: float, lambda_scale: float):
        self.model = model
        self.lambda_ssim = lambda_ssim
        self.lambda_scale = lambda_scale

    def get_loss_dict(self, render_img: Tensor, gt_img: Tensor, mask: Tensor) -> Dict[str, Tensor]:
        total, l1, ssim = fused_l1_ssim_loss(render_img, gt_img, mask, self.lambda_ssim)
        loss_dict = {"l1": l1, "ssim": ssim}
        regularization_dict = self.model.get_regularization_dict() if self.model is not None else {}
        if "scale_reg" in regularization_dict:
            loss_dict["scale_reg"] = regularization_dict["scale_reg"]
            total = total + self.lambda_scale * regularization_dict["scale_reg"]
        loss_dict["total"] = total
        return loss_dict
