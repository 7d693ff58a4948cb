"""Rendering several views of one training step (new capability; the reference renders one view per step,
/root/reference/train.py:93-104, and BASELINE config 4 asks for 64 views per step).

``ViewPipeline`` alternates consecutive views between two CUDA streams.  The Gaussian parameters do not change
inside a step, so the forward pass of view i+1 — projection, binning with its one host sync, blending — may overlap
the backward pass of view i; only the backward passes are ordered, because they accumulate into the same
``.grad`` buffers: view i+1's stream waits for view i's before its backward is enqueued.  Numerically this is the
same sum of per-view gradients as a sequential loop (up to the reduction order inside the kernels).
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import torch
from torch import Tensor

from .rendering import rasterization


class ViewPipeline:
    def __init__(self, device, enabled: bool = True):
        self.device = torch.device(device)
        self.enabled = enabled
        self.streams = [torch.cuda.Stream(device=self.device), torch.cuda.Stream(device=self.device)] if enabled else None

    def fork(self) -> None:
        """Call once before the first view of a step: the view streams pick up after the current stream."""
        if self.enabled:
            cur = torch.cuda.current_stream(self.device)
            for s in self.streams:
                s.wait_stream(cur)

    def join(self) -> None:
        """Call once after the last view: the current stream continues after both view streams."""
        if self.enabled:
            cur = torch.cuda.current_stream(self.device)
            for s in self.streams:
                cur.wait_stream(s)

    def stream_for(self, slot: int):
        return self.streams[slot % 2] if self.enabled else torch.cuda.current_stream(self.device)

    def render_backward(self, slot: int, params: Sequence[Tensor], viewmat: Tensor, K: Tensor, width: int, height: int,
                        loss_fn: Callable[[Tensor, Tensor], Tensor], sh_degree: Optional[int] = 3,
                        backgrounds: Optional[Tensor] = None, absgrad: bool = True,
                        after_backward: Optional[Callable[[dict], None]] = None,
                        render: Optional[Callable] = None) -> Tensor:
        """Forward + ``loss_fn(render_colors, render_alphas).backward()`` of one view in pipeline slot ``slot``.
        ``after_backward(meta)`` runs on the view's stream after the backward pass (e.g. the densify-stat update).
        ``render(viewmat, K)`` may replace the default call (``rasterization`` on the five activated tensors in
        ``params``), e.g. with ``rasterization_from_parameters`` on raw parameters."""

        def run():
            if render is not None:
                rc, ra, meta = render(viewmat, K)
            else:
                means, quats, scales, opacities, colors = params
                rc, ra, meta = rasterization(means, quats, scales, opacities, colors, viewmat, K, width, height,
                                             sh_degree=sh_degree, packed=False, absgrad=absgrad, backgrounds=backgrounds)
            loss = loss_fn(rc, ra)
            if self.enabled:
                me = self.streams[slot % 2]
                me.wait_stream(self.streams[(slot + 1) % 2])  # the previous view's gradient accumulation is complete
            loss.backward()
            if after_backward is not None:
                after_backward(meta)
            return loss

        if not self.enabled:
            return run()
        with torch.cuda.stream(self.streams[slot % 2]):
            return run()
