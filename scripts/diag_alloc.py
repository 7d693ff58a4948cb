#!/usr/bin/env python
"""GPU diagnostic: caching-allocator growth across training steps of the cfg4 workload (8 views per call)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from easy_gaussian_splatting_b200.synthetic import make_config_scene, loss_weights, CONFIGS
from easy_gaussian_splatting_b200.training import ViewPipeline
from easy_gaussian_splatting_b200.distributed import FlatGradBucket

pipelined = "--no-pipe" not in sys.argv
dev = torch.device("cuda", 0)
cfg = CONFIGS["cfg4"]; W, H = cfg["width"], cfg["height"]
sc = make_config_scene("cfg4", n_views=64)
params = [getattr(sc, k).to(dev).requires_grad_(True) for k in ("means", "quats", "scales", "opacities", "colors")]
bucket = FlatGradBucket(params)
C = 8
bg = sc.background[None].expand(C, 3).contiguous().to(dev)
Wc, Wa = (t.to(dev) for t in loss_weights(sc.seed, C, H, W))
views = [(sc.viewmats[g * C:(g + 1) * C].to(dev), sc.Ks[g * C:(g + 1) * C].to(dev)) for g in range(8)]
pipe = ViewPipeline(dev, enabled=pipelined)
for step in range(8):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    bucket.zero_()
    pipe.fork()
    for i, (vm, K) in enumerate(views):
        pipe.render_backward(i, params, vm, K, W, H, lambda rc, ra: (rc * Wc).sum() + (ra * Wa).sum(), sh_degree=3,
                             backgrounds=bg, absgrad=True)
    pipe.join()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    s = torch.cuda.memory_stats(dev)
    print(f"step {step}: {dt * 1e3:7.1f} ms  allocated {s['allocated_bytes.all.current'] / 1e9:6.2f} GB  peak {s['allocated_bytes.all.peak'] / 1e9:6.2f}"
          f"  reserved {s['reserved_bytes.all.current'] / 1e9:6.2f} GB  cudaMallocs {s['num_device_alloc']}  inactive_split {s['inactive_split_bytes.all.current'] / 1e9:6.2f}")
