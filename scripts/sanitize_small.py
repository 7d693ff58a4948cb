"""Small end-to-end forward+backward for compute-sanitizer runs (memcheck / racecheck / synccheck)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from easy_gaussian_splatting_b200 import rasterization
from easy_gaussian_splatting_b200.synthetic import make_scene, loss_weights

for kw in (dict(kind="blob", N=3000, width=131, height=77, fx=120.0, seed=7, n_views=2),
           dict(kind="outdoor", N=20000, width=320, height=200, fx=200.0, seed=2, n_views=1)):
    sc = make_scene(**kw).to("cuda")
    p = [t.clone().requires_grad_(True) for t in (sc.means, sc.quats, sc.scales, sc.opacities, sc.colors)]
    C = sc.viewmats.shape[0]
    rc, ra, meta = rasterization(*p, sc.viewmats, sc.Ks, sc.width, sc.height, sh_degree=3, packed=False, absgrad=True,
                                 backgrounds=sc.background[None].expand(C, 3).contiguous())
    Wc, Wa = loss_weights(1, C, sc.height, sc.width)
    ((rc * Wc.cuda()).sum() + (ra * Wa.cuda()).sum()).backward()
    _ = meta["isect_ids"]
    torch.cuda.synchronize()
    print("ok", kw["kind"], float(rc.mean()), float(p[0].grad.abs().mean()))
