"""Shared helpers for the parity tests (oracle vs CUDA path)."""
import torch

from easy_gaussian_splatting_b200.synthetic import make_scene, loss_weights
from oracle import gsplat_oracle as O

PARAMS = ("means", "quats", "scales", "opacities", "colors")


def oracle_run(sc, sh_degree=3, absgrad=True, backward=True, with_bg=True, dtype=torch.float32, weights=None,
               rasterize_mode="classic"):
    """Runs the CPU oracle fwd (+bwd of the benchmark functional).  Returns a dict."""
    leaves = {k: getattr(sc, k).detach().clone().to(dtype).requires_grad_(backward) for k in PARAMS}
    C = sc.viewmats.shape[0]
    bg = sc.background[None].expand(C, 3).contiguous().to(dtype) if with_bg else None
    counters = {}
    rc, ra, meta = O.rasterization(leaves["means"], leaves["quats"], leaves["scales"], leaves["opacities"],
                                   leaves["colors"], sc.viewmats.to(dtype), sc.Ks.to(dtype), sc.width, sc.height,
                                   sh_degree=sh_degree, packed=False, absgrad=absgrad, backgrounds=bg,
                                   counters=counters, rasterize_mode=rasterize_mode)
    out = dict(colors=rc.detach(), alphas=ra.detach(), meta=meta, counters=counters)
    if backward:
        Wc, Wa = weights if weights is not None else loss_weights(sc.seed, C, sc.height, sc.width)
        ((rc * Wc.to(dtype)).sum() + (ra * Wa.to(dtype)).sum()).backward()
        out["grads"] = {k: v.grad for k, v in leaves.items()}
        if absgrad:
            out["absgrad"] = meta["means2d"].absgrad
    return out


def cuda_run(sc, sh_degree=3, absgrad=True, backward=True, with_bg=True, weights=None, device="cuda",
             rasterize_mode="classic", packed=False):
    from easy_gaussian_splatting_b200 import rasterization

    leaves = {k: getattr(sc, k).detach().clone().to(device).requires_grad_(backward) for k in PARAMS}
    C = sc.viewmats.shape[0]
    bg = sc.background[None].expand(C, 3).contiguous().to(device) if with_bg else None
    rc, ra, meta = rasterization(leaves["means"], leaves["quats"], leaves["scales"], leaves["opacities"],
                                 leaves["colors"], sc.viewmats.to(device), sc.Ks.to(device), sc.width, sc.height,
                                 sh_degree=sh_degree, packed=packed, absgrad=absgrad, backgrounds=bg,
                                 rasterize_mode=rasterize_mode)
    out = dict(colors=rc.detach().cpu(), alphas=ra.detach().cpu(), meta=meta)
    if backward:
        Wc, Wa = weights if weights is not None else loss_weights(sc.seed, C, sc.height, sc.width)
        ((rc * Wc.to(device)).sum() + (ra * Wa.to(device)).sum()).backward()
        out["grads"] = {k: v.grad.detach().cpu() for k, v in leaves.items()}
        if absgrad:
            out["absgrad"] = meta["means2d"].absgrad.detach().cpu()
    return out


def rel_err(a, b):
    """norm-wise relative error ||a-b|| / ||b||"""
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def image_report(test, ref, borderline):
    """max abs error over non-borderline pixels, number and max error of borderline pixels."""
    err = (test - ref).abs().amax(-1)
    clean = err[~borderline]
    border = err[borderline]
    return dict(max_clean=clean.max().item() if clean.numel() else 0.0,
                n_border=int(borderline.sum()),
                max_border=border.max().item() if border.numel() else 0.0,
                n_bad_clean=int((clean > 1e-4).sum()))
