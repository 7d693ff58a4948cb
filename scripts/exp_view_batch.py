"""Experiment: C views per rasterization() call vs one call per view (same total views)."""
import sys, time, torch
sys.path.insert(0, ".")
from easy_gaussian_splatting_b200 import rasterization
from easy_gaussian_splatting_b200.synthetic import make_config_scene, loss_weights

dev = torch.device("cuda:0")
for wl in sys.argv[1:]:
    sc = make_config_scene(wl, n_views=8)
    W, H = sc.width, sc.height
    P = [getattr(sc, k).to(dev).requires_grad_(True) for k in ("means", "quats", "scales", "opacities", "colors")]
    vm, Ks = sc.viewmats.to(dev), sc.Ks.to(dev)
    bg = sc.background[None].to(dev)
    Wc, Wa = [t.to(dev) for t in loss_weights(sc.seed, 1, H, W)]
    for C in (1, 2, 4, 8):
        def step():
            for p in P:
                p.grad = None
            for v in range(0, 8, C):
                rc, ra, meta = rasterization(*P, vm[v:v + C], Ks[v:v + C], W, H, sh_degree=3, packed=False, absgrad=True,
                                             backgrounds=bg.expand(C, 3))
                ((rc * Wc).sum() + (ra * Wa).sum()).backward()
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"{wl} C={C}: {ms:.3f} ms / 8 views  -> {8 * W * H / ms / 1e3:.1f} Mpix/s", flush=True)
