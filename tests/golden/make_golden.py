#!/usr/bin/env python
"""Generates tests/golden/*.npz: small frozen input/output vectors of the CPU oracle.

There is nothing in /root/reference to generate vectors FROM (the reference has no rasterizer code and no
tests; gsplat is not installable here), so these fixtures freeze the oracle's own outputs:
they pin the oracle against regressions and give the CUDA path a fixed target that does not depend on
re-running the oracle.  Re-run only when the oracle's semantics are deliberately changed:
    python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from easy_gaussian_splatting_b200.synthetic import loss_weights, make_scene  # noqa: E402
from oracle import gsplat_oracle as O  # noqa: E402

CASES = {
    "blob_c1": dict(kind="blob", N=96, width=48, height=32, fx=40.0, seed=101, n_views=1, sh_degree=3),
    "blob_c2_ragged": dict(kind="blob", N=64, width=37, height=21, fx=30.0, seed=102, n_views=2, sh_degree=2),
    "object_white": dict(kind="object", N=120, width=40, height=40, fx=60.0, seed=103, n_views=1, sh_degree=1, white_background=True),
    # rasterize_mode="antialiased" (adds in_antialiased = 1 and the compensated opacities)
    "blob_antialiased": dict(kind="blob", N=80, width=40, height=24, fx=34.0, seed=104, n_views=1, sh_degree=3, antialiased=True),
}


def run(case):
    kw = dict(case)
    deg = kw.pop("sh_degree")
    mode = "antialiased" if kw.pop("antialiased", False) else "classic"
    sc = make_scene(**kw)
    names = ("means", "quats", "scales", "opacities", "colors")
    leaves = {k: getattr(sc, k).clone().requires_grad_(True) for k in names}
    C = sc.viewmats.shape[0]
    bg = sc.background[None].expand(C, 3).contiguous()
    counters = {}
    rc, ra, meta = O.rasterization(leaves["means"], leaves["quats"], leaves["scales"], leaves["opacities"], leaves["colors"],
                                   sc.viewmats, sc.Ks, sc.width, sc.height, sh_degree=deg, packed=False, absgrad=True,
                                   backgrounds=bg, counters=counters, rasterize_mode=mode)
    Wc, Wa = loss_weights(sc.seed, C, sc.height, sc.width)
    ((rc * Wc).sum() + (ra * Wa).sum()).backward()
    out = {f"in_{k}": getattr(sc, k).numpy() for k in names}
    out.update(in_viewmats=sc.viewmats.numpy(), in_Ks=sc.Ks.numpy(), in_background=bg.numpy(),
               in_width=np.int32(sc.width), in_height=np.int32(sc.height), in_sh_degree=np.int32(deg),
               in_Wc=Wc.numpy(), in_Wa=Wa.numpy(),
               render_colors=rc.detach().numpy(), render_alphas=ra.detach().numpy(),
               radii=meta["radii"].numpy(), means2d=meta["means2d"].detach().numpy(), depths=meta["depths"].detach().numpy(),
               conics=meta["conics"].detach().numpy(), tiles_per_gauss=meta["tiles_per_gauss"].numpy(),
               isect_ids=meta["isect_ids"].numpy(), flatten_ids=meta["flatten_ids"].numpy(),
               isect_offsets=meta["isect_offsets"].numpy(), last_ids=meta["last_ids"].numpy(),
               absgrad=meta["means2d"].absgrad.numpy(), borderline=counters["borderline"].numpy(),
               P_eval=np.int64(counters["P_eval"]), P_acc=np.int64(counters["P_acc"]))
    out.update({f"grad_{k}": leaves[k].grad.numpy() for k in names})
    if mode == "antialiased":
        out.update(in_antialiased=np.int32(1), opacities=meta["opacities"].detach().numpy())
    return out


if __name__ == "__main__":
    only = set(sys.argv[1:])  # optional: names of the cases to (re)generate; default = all
    for name, case in CASES.items():
        if only and name not in only:
            continue
        out = run(case)
        path = Path(__file__).parent / f"{name}.npz"
        np.savez_compressed(path, **out)
        print(name, path.stat().st_size, "bytes; n_isects", out["flatten_ids"].shape[0], "P_acc", int(out["P_acc"]))
