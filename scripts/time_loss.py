#!/usr/bin/env python
"""GPU: CUDA-event time of the two fused-loss kernels alone (raw C-ABI calls, no Python-side tensor work)."""
import ctypes, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from easy_gaussian_splatting_b200 import _lib

lib = _lib.load()
C, H, W = 4, 1080, 1920
dev = "cuda"
r = torch.rand(C, H, W, 3, device=dev); g = torch.rand(C, H, W, 3, device=dev); m = (torch.rand(C, H, W, device=dev) < 0.1).float()
maps = torch.empty(3, C, 3, H - 10, W - 10, device=dev); sums = torch.zeros(C, 2, dtype=torch.float64, device=dev)
vt = torch.ones(C, device=dev); vr = torch.empty_like(r)
P = lambda t: ctypes.c_void_p(t.data_ptr())
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
def fwd(): lib.egs_l1_ssim_fwd(C, H, W, P(r), P(g), P(m), P(maps), P(sums), st)
def bwd(): lib.egs_l1_ssim_bwd(C, H, W, P(r), P(g), P(m), P(maps), 0.2, P(vt), P(vr), st)
for name, fn in (("fwd", fwd), ("bwd", bwd)):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); e1.synchronize()
    ms = e0.elapsed_time(e1) / 10
    b = (64 if name == "fwd" else 76) * C * H * W
    print(f"l1_ssim_{name}: {ms:.4f} ms per launch ({C} x {W}x{H}), {b / ms / 1e6:.0f} GB/s algorithmic")
