#!/usr/bin/env python
"""GPU box helper: where does one view per call (the reference's pattern) spend its time?  Host enqueue time per view
(no synchronisation inside the loop) against device time per view.   c1_breakdown.py metric|cfg2 [views]"""
import sys, time
sys.path.insert(0, ".")
import torch
import bench
from easy_gaussian_splatting_b200 import rasterization
from easy_gaussian_splatting_b200.distributed import DensifyStats
from easy_gaussian_splatting_b200.synthetic import loss_weights, make_config_scene

wl = sys.argv[1]
nv = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = torch.device("cuda", 0)
sc = make_config_scene(wl, n_views=nv)
W, H, N = sc.width, sc.height, sc.means.shape[0]
params = [getattr(sc, k).to(dev).requires_grad_(True) for k in ("means", "quats", "scales", "opacities", "colors")]
vms, Ks, bg = sc.viewmats.to(dev), sc.Ks.to(dev), sc.background[None].to(dev)
Wc, Wa = (t.to(dev) for t in loss_weights(sc.seed, 1, H, W))
stats = DensifyStats(N, dev)


def view(v):
    for p_ in params:
        p_.grad = None
    t0 = time.perf_counter()
    rc, ra, meta = rasterization(*params, vms[v:v + 1], Ks[v:v + 1], W, H, sh_degree=3, packed=False, absgrad=True, backgrounds=bg)
    t1 = time.perf_counter()
    loss = bench.linear_functional(rc, ra, Wc, Wa)
    t2 = time.perf_counter()
    loss.backward()
    t3 = time.perf_counter()
    stats.update_local(meta["radii"], meta["means2d"].absgrad, W, H)
    return t1 - t0, t2 - t1, t3 - t2, time.perf_counter() - t3


for _ in range(3):
    for v in range(nv):
        view(v)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
acc = [0.0] * 4
t_start = time.perf_counter()
e0.record()
n = 0
for _ in range(6):
    for v in range(nv):
        for i, t in enumerate(view(v)):
            acc[i] += t
        n += 1
e1.record()
host = time.perf_counter() - t_start
torch.cuda.synchronize()
print(f"{wl}: device {e0.elapsed_time(e1) / n:.3f} ms/view, host enqueue {host / n * 1e3:.3f} ms/view "
      f"(forward {acc[0] / n * 1e3:.3f}, functional {acc[1] / n * 1e3:.3f}, backward {acc[2] / n * 1e3:.3f}, stats {acc[3] / n * 1e3:.3f})")
