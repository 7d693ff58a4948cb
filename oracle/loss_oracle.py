"""CPU oracle for the photometric loss of the reference (TEST INFRASTRUCTURE ONLY — same import rules as
``gsplat_oracle.py``: only ``tests/``, ``__graft_entry__.smoke()`` and bench baselines may use it).

Follows ``LossComputer`` at /root/reference/model/gaussian.py:415-453 line by line.  Its SSIM term is
``torchmetrics.image.StructuralSimilarityIndexMeasure(data_range=1.0)`` (gaussian.py:5, :419), a third-party
dependency that is NOT installed in this image and not vendored in the reference (torchmetrics is unpinned in
/root/reference/requirements.txt) — PARITY UNPINNED: ``ssim`` below restates the published algorithm of
``torchmetrics.functional.image.ssim._ssim_update`` (defaults gaussian_kernel=True, sigma=1.5, kernel_size=11,
k1=0.01, k2=0.03, reduction="elementwise_mean") from recollection: reflect-pad by 5, depthwise conv with the
outer-product Gaussian over the stack (p, t, p*p, t*t, p*t), SSIM map, crop the padded border, mean.
(Later torchmetrics releases clamp the two variances at 0 before forming the denominator; with a 11x11 Gaussian
window the variances are non-negative up to rounding, so the two variants agree to fp32 rounding.)
Pure torch, any float dtype (fp64 for gradient checks).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F
from torch import Tensor


def gaussian_window(kernel_size: int = 11, sigma: float = 1.5, dtype=torch.float32) -> Tensor:
    """torchmetrics ``_gaussian``: exp(-(d / sigma)^2 / 2), d = (1-k)/2 .. (k-1)/2, normalised to sum 1."""
    dist = torch.arange((1 - kernel_size) / 2, (1 + kernel_size) / 2, 1, dtype=dtype)
    gauss = torch.exp(-torch.pow(dist / sigma, 2) / 2)
    return gauss / gauss.sum()


def ssim(preds: Tensor, target: Tensor, data_range: float = 1.0, sigma: float = 1.5, k1: float = 0.01,
         k2: float = 0.03) -> Tensor:
    """preds, target [B,C,H,W] -> scalar mean SSIM (torchmetrics defaults, see module docstring)."""
    c1 = (k1 * data_range) ** 2
    c2 = (k2 * data_range) ** 2
    channel = preds.shape[1]
    ks = int(3.5 * sigma + 0.5) * 2 + 1  # 11 for sigma 1.5
    pad = (ks - 1) // 2
    g = gaussian_window(ks, sigma, preds.dtype).to(preds.device)
    kernel = (g[:, None] * g[None, :]).expand(channel, 1, ks, ks)
    p = F.pad(preds, (pad, pad, pad, pad), mode="reflect")
    t = F.pad(target, (pad, pad, pad, pad), mode="reflect")
    stack = torch.cat((p, t, p * p, t * t, p * t))
    out = F.conv2d(stack, kernel, groups=channel)
    mu_p, mu_t, e_pp, e_tt, e_pt = out.split(preds.shape[0])
    mu_p_sq, mu_t_sq, mu_pt = mu_p * mu_p, mu_t * mu_t, mu_p * mu_t
    s_pp, s_tt, s_pt = e_pp - mu_p_sq, e_tt - mu_t_sq, e_pt - mu_pt
    upper = 2 * s_pt + c2
    lower = s_pp + s_tt + c2
    full = ((2 * mu_pt + c1) * upper) / ((mu_p_sq + mu_t_sq + c1) * lower)
    crop = full[..., pad:-pad, pad:-pad]
    return crop.reshape(crop.shape[0], -1).mean(-1).mean()


def loss_dict(render_img: Tensor, gt_img: Tensor, mask: Optional[Tensor], lambda_ssim: float) -> Dict[str, Tensor]:
    """LossComputer.get_loss_dict (gaussian.py:422-445) without the scale regulariser (a model-side term)."""
    if mask is not None:
        m3 = mask.unsqueeze(2).repeat(1, 1, 3)                      # gaussian.py:428
        render_img = m3 * gt_img + (1.0 - m3) * render_img          # gaussian.py:429
    l1 = F.l1_loss(render_img, gt_img)                              # gaussian.py:447-448
    r = render_img.permute(2, 0, 1)[None]                           # gaussian.py:451
    g = gt_img.permute(2, 0, 1)[None]                               # gaussian.py:452
    ssim_loss = 1.0 - ssim(g, r)                                    # gaussian.py:453 (preds = gt, target = render)
    total = (1.0 - lambda_ssim) * l1 + lambda_ssim * ssim_loss      # gaussian.py:437
    return {"l1": l1, "ssim": ssim_loss, "total": total}
