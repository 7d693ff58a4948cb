"""§8f-3: fused Adam for the reference's parameter groups.

Drop-in for the ``torch.optim.Adam`` the reference builds in ``build_optimizers``
(/root/reference/model/gaussian.py:389-412: six named groups, one learning rate each, default betas/eps, no
weight decay) and steps at /root/reference/train.py:156-157.  One kernel launch updates every group; parameters
whose ``.grad`` is None are skipped exactly like torch does (the reference relies on that on densify steps,
SURVEY.md §3.1).  State lives in ``self.state[p]`` under torch's key names (``step``, ``exp_avg``,
``exp_avg_sq``), so the reference's optimizer surgery in ``densify_and_prune`` keeps working on it.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        lib = _lib.load()
        # groups that share (betas, eps, step) go out in one launch of up to 8 tensors
        batches = {}
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and p.grad.is_contiguous()):
                    raise RuntimeError("FusedAdam needs contiguous float32 CUDA parameters and gradients")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p)
                    st["exp_avg_sq"] = torch.zeros_like(p)
                st["step"] = int(st["step"]) + 1
                key = (p.device, group["betas"], group["eps"], st["step"])
                batches.setdefault(key, []).append((p, p.grad, st["exp_avg"], st["exp_avg_sq"], float(group["lr"])))
        for (dev, betas, eps, step), items in batches.items():
            for i in range(0, len(items), 8):
                chunk = items[i:i + 8]
                n = len(chunk)
                arr = lambda j: (ctypes.c_void_p * n)(*[t[j].data_ptr() for t in chunk])
                numels = (ctypes.c_int64 * n)(*[t[0].numel() for t in chunk])
                lrs = (ctypes.c_float * n)(*[t[4] for t in chunk])
                with torch.cuda.device(dev):
                    rc = lib.egs_fused_adam(n, arr(0), arr(1), arr(2), arr(3), numels, lrs, float(betas[0]), float(betas[1]),
                                            float(eps), int(step),
                                            ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
                _lib.check(rc, "egs_fused_adam")
        return loss
