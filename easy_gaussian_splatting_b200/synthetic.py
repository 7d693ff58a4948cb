"""Seeded synthetic scenes for parity tests and benchmarks.

All draws happen on CPU in fp32 from ``torch.Generator().manual_seed(seed)`` in a fixed
order (means, log_scales, quats, logit_opac, sh_dc, sh_rest, cameras) so the CPU oracle
and the CUDA path consume bit-identical inputs.  The tensors mirror what the reference
feeds ``rasterization`` (/root/reference/model/gaussian.py:353-367): activated scales
(exp), activated opacities (sigmoid), ``colors = cat(sh_0, sh_rest)`` with K = 16, and
OpenCV-convention world->camera matrices (/root/reference/scene/data_class.py:64,107-140).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import torch
from torch import Tensor


@dataclass
class Scene:
    means: Tensor  # [N,3]
    quats: Tensor  # [N,4] wxyz, unnormalised
    scales: Tensor  # [N,3]  (already exp-activated)
    opacities: Tensor  # [N]  (already sigmoid-activated)
    colors: Tensor  # [N,16,3] SH coefficients
    viewmats: Tensor  # [V,4,4]
    Ks: Tensor  # [V,3,3]
    width: int
    height: int
    background: Tensor  # [3]
    kind: str
    seed: int

    def to(self, device) -> "Scene":
        kw = {k: (v.to(device) if isinstance(v, Tensor) else v) for k, v in self.__dict__.items()}
        return Scene(**kw)


def look_at(eye: Tensor, target: Tensor, up_world=(0.0, 1.0, 0.0)) -> Tensor:
    """World->camera [4,4], OpenCV axes: x right, y down, z forward."""
    fwd = target - eye
    fwd = fwd / fwd.norm()
    upw = torch.tensor(up_world, dtype=eye.dtype)
    right = torch.linalg.cross(fwd, upw)
    right = right / right.norm()
    down = torch.linalg.cross(fwd, right)
    R = torch.stack([right, down, fwd], 0)  # rows = camera axes in world coords
    V = torch.eye(4, dtype=eye.dtype)
    V[:3, :3] = R
    V[:3, 3] = -(R @ eye)
    return V


def intrinsics(fx: float, fy: float, cx: float, cy: float) -> Tensor:
    return torch.tensor([[fx, 0.0, cx], [0.0, fy, cy], [0.0, 0.0, 1.0]], dtype=torch.float32)


def make_scene(kind: str, N: int, width: int, height: int, fx: float, seed: int, n_views: int = 1,
               white_background: bool = False, fy: Optional[float] = None) -> Scene:
    g = torch.Generator().manual_seed(seed)
    f32 = torch.float32

    def randn(*s):
        return torch.randn(*s, generator=g, dtype=f32)

    def rand(*s):
        return torch.rand(*s, generator=g, dtype=f32)

    if kind == "blob":
        means = randn(N, 3) * 0.8
        nrm = means.norm(dim=-1, keepdim=True).clamp_min(1e-9)
        means = means * torch.clamp(2.0 / nrm, max=1.0)
        log_scales = math.log(0.03) + 0.5 * randn(N, 3)
    elif kind == "object":
        nb = 12
        lo = rand(nb, 3) * 1.3 - 0.65
        hi = rand(nb, 3) * 1.3 - 0.65
        bmin, bmax = torch.minimum(lo, hi), torch.maximum(lo, hi)
        n_surf = int(0.7 * N)
        box = torch.randint(0, nb, (n_surf,), generator=g)
        u = rand(n_surf, 3)
        p = bmin[box] + u * (bmax[box] - bmin[box])
        axis = torch.randint(0, 3, (n_surf,), generator=g)
        side = torch.randint(0, 2, (n_surf,), generator=g).to(f32)
        face = bmin[box] + side[:, None] * (bmax[box] - bmin[box])
        onehot = torch.nn.functional.one_hot(axis, 3).to(torch.bool)
        p = torch.where(onehot, face, p) + 0.003 * randn(n_surf, 3)
        inner = rand(N - n_surf, 3) * 1.3 - 0.65
        means = torch.cat([p, inner], 0)
        log_scales = math.log(0.005) + 0.6 * randn(N, 3)
    elif kind == "outdoor":
        n_fg = int(0.6 * N)
        fg = (rand(n_fg, 3) * 2.0 - 1.0) * torch.tensor([2.0, 1.0, 3.0])
        n_bg = N - n_fg
        rad = torch.exp(math.log(5.0) + rand(n_bg) * (math.log(50.0) - math.log(5.0)))
        d = randn(n_bg, 3)
        d = d / d.norm(dim=-1, keepdim=True).clamp_min(1e-9)
        bgp = d * rad[:, None]
        means = torch.cat([fg, bgp], 0)
        ls_fg = math.log(0.01) + 0.6 * randn(n_fg, 3)
        ls_bg = torch.log(0.01 * rad / 3.0)[:, None] + 0.6 * randn(n_bg, 3)
        log_scales = torch.cat([ls_fg, ls_bg], 0)
    else:
        raise ValueError(f"unknown scene kind {kind!r}")

    quats = randn(N, 4)
    logit_opac = 1.5 * randn(N)
    sh_dc = 0.5 * randn(N, 1, 3)
    sh_rest = 0.1 * randn(N, 15, 3)
    scales = torch.exp(log_scales)
    opacities = torch.sigmoid(logit_opac)
    colors = torch.cat([sh_dc, sh_rest], 1).contiguous()

    views = []
    origin = torch.zeros(3, dtype=f32)
    if kind == "blob":
        for i in range(n_views):
            ang = 2.0 * math.pi * i / max(n_views, 1) + 0.3
            eye = torch.tensor([4.0 * math.sin(ang), -0.5, -4.0 * math.cos(ang)], dtype=f32)
            eye = eye / eye.norm() * 4.0
            views.append(look_at(eye, origin, up_world=(0.0, -1.0, 0.0)))
    elif kind == "object":
        az = rand(n_views) * 2.0 * math.pi
        el = rand(n_views) * (math.pi / 3.0)
        for i in range(n_views):
            r = 4.031
            eye = torch.tensor([r * math.cos(el[i]) * math.cos(az[i]), -r * math.sin(el[i]),
                                r * math.cos(el[i]) * math.sin(az[i])], dtype=f32)
            views.append(look_at(eye, origin, up_world=(0.0, -1.0, 0.0)))
    else:
        ph0 = float(rand(1)) * 2.0 * math.pi
        for i in range(n_views):
            ang = ph0 + 2.0 * math.pi * i / max(n_views, 1)
            eye = torch.tensor([5.0 * math.cos(ang), -0.5, 5.0 * math.sin(ang)], dtype=f32)
            views.append(look_at(eye, origin, up_world=(0.0, -1.0, 0.0)))
    viewmats = torch.stack(views, 0).contiguous()
    K = intrinsics(fx, fx if fy is None else fy, width / 2.0, height / 2.0)
    Ks = K[None].repeat(n_views, 1, 1).contiguous()
    bg = torch.full((3,), 1.0 if white_background else 0.0, dtype=f32)
    return Scene(means.contiguous(), quats.contiguous(), scales.contiguous(), opacities.contiguous(), colors,
                 viewmats, Ks, width, height, bg, kind, seed)


# the named configurations of BASELINE.md section 4
CONFIGS = {
    "cfg1": dict(kind="blob", N=10_000, width=256, height=256, fx=274.5, seed=0, white_background=False),
    "cfg2": dict(kind="object", N=300_000, width=800, height=800, fx=1111.11, seed=1, white_background=True),
    "cfg3": dict(kind="outdoor", N=2_500_000, width=979, height=546, fx=581.0, seed=2, white_background=False),
    "metric": dict(kind="outdoor", N=1_000_000, width=1920, height=1080, fx=1662.77, seed=3, white_background=False),
    "cfg4": dict(kind="outdoor", N=3_000_000, width=1920, height=1080, fx=1662.77, seed=4, white_background=False),
    "cfg5": dict(kind="outdoor", N=6_000_000, width=3840, height=2160, fx=3325.54, seed=5, white_background=False),
}


def make_config_scene(name: str, n_views: int = 1, N: Optional[int] = None) -> Scene:
    kw = dict(CONFIGS[name])
    if N is not None:
        kw["N"] = N
    return make_scene(n_views=n_views, **kw)


def loss_weights(scene_seed: int, C: int, height: int, width: int):
    """Upstream gradient of the benchmark's linear functional L = sum(colors*Wc) + sum(alphas*Wa)
    (SURVEY.md section 8d): U(0,1) weights, seed = scene seed + 1000."""
    g = torch.Generator().manual_seed(scene_seed + 1000)
    Wc = torch.rand(C, height, width, 3, generator=g, dtype=torch.float32)
    Wa = torch.rand(C, height, width, 1, generator=g, dtype=torch.float32)
    return Wc, Wa
