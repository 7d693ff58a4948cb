"""TEST INFRASTRUCTURE: the CALLER side of the hot path — what sits above ``gsplat.rendering.rasterization`` in
li199603/easy_gaussian_splatting — in two forms:

1. ``import_reference()``: the reference's OWN ``model/gaussian.py`` imported from ``/root/reference`` (only where
   that tree exists, i.e. in the build container; never on the GPU box), with ``sys.modules`` stand-ins for the
   dependencies this image lacks (``torchmetrics``; ``scene`` is reduced to the reference's real ``Pointcloud`` class)
   and ``gsplat`` resolved to ``shim/gsplat`` — so ``GaussianModel.forward`` / ``update_statistics`` /
   ``densify_and_prune`` / ``reset_opacities`` / ``build_optimizers`` / ``LossComputer`` run UNCHANGED against this
   repo's rasterizer.
2. ``CallerModel``: a compact restatement of exactly those members (each citing the reference line it follows), for
   machines without ``/root/reference``.  tests/test_reference_swap.py proves on CPU that both produce bit-identical
   training trajectories through the same rasterizer, which is what entitles the ``-m gpu`` tests to use the
   restatement on the real kernels.

``train_loop`` restates the loop body of /root/reference/train.py:93-157 (forward, loss, backward, then under
``no_grad``: statistics, densify / prune, opacity reset, SH degree, learning rate; ``optimizer.step()`` and
``zero_grad()``) for either model.
"""
from __future__ import annotations

import contextlib
import importlib
import importlib.util
import sys
import types
from pathlib import Path
from typing import Any, Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F
from torch import Tensor, nn

ROOT = Path(__file__).resolve().parent.parent
REFERENCE = Path("/root/reference")

# configs/nerf_synthetic.yaml of the reference, with the schedule compressed so that a few hundred iterations see
# every event (densify + prune several times, an opacity reset, SH degree steps).
TINY_CFG = dict(
    sh_degree=3, sh_degree_interval=40, means_lr_init=1e-3, means_lr_final=1e-5, means_lr_schedule_max_steps=300,
    log_scales_lr=1e-2, quats_lr=1e-3, sh_0_lr=2.5e-3, sh_rest_lr=1.25e-4, logit_opacities_lr=5e-2,
    refine_start=20, refine_stop=260, refine_every=30, reset_opacities_every=60, min_opacity=0.005,
    densify_grad_thresh=5e-4, densify_scale_thresh=0.1, num_splits=2, prune_radii_ratio_thresh=0.15,
    prune_scale_thresh=1.0, lambda_ssim=0.2, use_scale_regularization=False, max_scale_ratio=10.0, lambda_scale=0.1,
    white_background=True,
)


def reference_available() -> bool:
    return (REFERENCE / "model" / "gaussian.py").exists()


@contextlib.contextmanager
def cuda_means_cpu():
    """The reference hard-codes ``device="cuda"`` (gaussian.py:57-64, 84, 92, 135, 166, 302) and calls ``.cuda()``.
    On a machine without a GPU, map those to the CPU for the duration of a test; a no-op where CUDA exists."""
    if torch.cuda.is_available():
        yield
        return
    names = ("zeros", "ones", "randn", "tensor", "full", "full_like", "zeros_like", "empty")
    saved = {n: getattr(torch, n) for n in names}
    saved_cuda = nn.Module.cuda

    def wrap(fn):
        def inner(*a, **k):
            if isinstance(k.get("device"), str) and k["device"].startswith("cuda"):
                k["device"] = "cpu"
            return fn(*a, **k)
        return inner

    try:
        for n, fn in saved.items():
            setattr(torch, n, wrap(fn))
        nn.Module.cuda = lambda self, device=None: self
        yield
    finally:
        for n, fn in saved.items():
            setattr(torch, n, fn)
        nn.Module.cuda = saved_cuda


def _ssim_module(ssim_fn):
    class StructuralSimilarityIndexMeasure(nn.Module):  # the one torchmetrics name gaussian.py:9 imports
        def __init__(self, data_range=1.0):
            super().__init__()
            self.data_range = data_range

        def forward(self, preds, target):
            return ssim_fn(preds, target, data_range=self.data_range)

    m = types.ModuleType("torchmetrics.image")
    m.StructuralSimilarityIndexMeasure = StructuralSimilarityIndexMeasure
    return m


def import_reference(ssim_fn):
    """-> the reference's ``model.gaussian`` module (fresh import), wired to this repo's rasterizer."""
    if not reference_available():
        raise RuntimeError("/root/reference is not present on this machine")
    spec = importlib.util.spec_from_file_location("_ref_scene_data_class", REFERENCE / "scene" / "data_class.py")
    data_class = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(data_class)
    scene = types.ModuleType("scene")  # scene/__init__.py would pull in the COLMAP loader (pyquaternion: absent)
    scene.Pointcloud = data_class.Pointcloud
    tm = types.ModuleType("torchmetrics")
    tm.image = _ssim_module(ssim_fn)
    for name in [n for n in sys.modules if n == "model" or n.startswith("model.") or n == "gsplat" or n.startswith("gsplat.")]:
        del sys.modules[name]
    sys.modules.update({"scene": scene, "torchmetrics": tm, "torchmetrics.image": tm.image})
    for p in (str(ROOT / "shim"), str(REFERENCE)):
        if p not in sys.path:
            sys.path.insert(0, p)
    mod = importlib.import_module("model.gaussian")
    assert Path(mod.__file__).resolve() == (REFERENCE / "model" / "gaussian.py").resolve()
    import easy_gaussian_splatting_b200 as egs
    assert mod.rasterization is egs.rasterization, "the reference's import line must resolve to this repo's rasterizer"
    return mod, data_class.Pointcloud


def forget_reference() -> None:
    for name in [n for n in sys.modules if n in ("scene", "torchmetrics", "torchmetrics.image", "model", "gsplat")
                 or n.startswith(("model.", "gsplat."))]:
        del sys.modules[name]
    for p in (str(ROOT / "shim"), str(REFERENCE)):
        while p in sys.path:
            sys.path.remove(p)


# ------------------------------------------------------------------------------------------------------------------
# restated caller
# ------------------------------------------------------------------------------------------------------------------
NAMES = ("means", "log_scales", "quats", "sh_0", "sh_rest", "logit_opacities")  # gaussian.py:109-111


def _knn_mean_dist(xyz: np.ndarray, k: int = 3) -> np.ndarray:
    """model/utils.py:8-11 (sklearn NearestNeighbors, euclidean, self excluded), then the mean of gaussian.py:35."""
    from sklearn.neighbors import NearestNeighbors
    d, _ = NearestNeighbors(n_neighbors=k + 1, metric="euclidean").fit(xyz).kneighbors(xyz)
    return np.mean(d[:, 1:].astype(np.float32), axis=1, keepdims=True)


def _quat_to_rotmat(q: Tensor) -> Tensor:
    """model/utils.py:31-49: normalise, wxyz -> rotation matrix."""
    w, x, y, z = F.normalize(q, dim=-1).unbind(-1)
    rows = [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
            2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
            2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]
    return torch.stack(rows, -1).reshape(q.shape[:-1] + (3, 3))


class CallerModel:
    """What GaussianModel holds and does around the rasterization call (gaussian.py:12-374), no nn.Module needed."""

    def __init__(self, xyzs: np.ndarray, rgbs: np.ndarray, cfg: Dict[str, Any], device, rasterize=None):
        self.dev = torch.device(device)
        self.cfg = cfg
        n = xyzs.shape[0]
        t = lambda a: torch.tensor(a, dtype=torch.float32)
        spread = np.repeat(_knn_mean_dist(xyzs), 3, axis=1)                               # gaussian.py:34-36
        quats = torch.zeros(n, 4)
        quats[:, 0] = 1.0                                                                 # gaussian.py:40-41
        K = (cfg["sh_degree"] + 1) ** 2
        sh = torch.zeros(n, K, 3)
        sh[:, 0] = t((rgbs / 255.0 - 0.5) / 0.28209479177387814)                          # gaussian.py:46-48, utils.py:14-16
        init = dict(means=t(xyzs), log_scales=torch.log(t(spread) / 2.0), quats=quats, sh_0=sh[:, 0:1], sh_rest=sh[:, 1:],
                    logit_opacities=torch.logit(0.8 * torch.ones(n)))                     # gaussian.py:33-54
        self.p: Dict[str, Tensor] = {k: init[k].to(self.dev).requires_grad_(True) for k in NAMES}
        self._reset_stats()
        self.active_sh_degree = 0 if cfg["sh_degree_interval"] != 0 else cfg["sh_degree"]  # gaussian.py:67
        self.background = torch.full((3,), 1.0 if cfg["white_background"] else 0.0, device=self.dev)  # gaussian.py:82-90
        lr0 = dict(means=cfg["means_lr_init"], log_scales=cfg["log_scales_lr"], quats=cfg["quats_lr"], sh_0=cfg["sh_0_lr"],
                   sh_rest=cfg["sh_rest_lr"], logit_opacities=cfg["logit_opacities_lr"])
        self.optimizer = torch.optim.Adam([{"params": [self.p[k]], "lr": lr0[k], "name": k} for k in NAMES])  # :389-412
        if rasterize is None:
            from easy_gaussian_splatting_b200 import rasterization as rasterize
        self.rasterize = rasterize

    # gaussian.py:93-107
    n = property(lambda self: self.p["means"].shape[0])
    scales = property(lambda self: torch.exp(self.p["log_scales"]))
    opacities = property(lambda self: torch.sigmoid(self.p["logit_opacities"]))
    shs = property(lambda self: torch.cat([self.p["sh_0"], self.p["sh_rest"]], dim=1))

    def _reset_stats(self):                                                               # gaussian.py:56-64, 328-337
        self.grad_norm_accum, self.collecting_counts, self.max_radii = (torch.zeros(self.n, device=self.dev) for _ in range(3))

    def forward(self, data: Dict[str, Any]) -> Dict[str, Tensor]:                         # gaussian.py:351-374
        imgs, _, meta = self.rasterize(
            means=self.p["means"], quats=self.p["quats"], scales=self.scales, opacities=self.opacities, colors=self.shs,
            sh_degree=self.active_sh_degree, viewmats=data["w2c"][None], Ks=data["K"][None], width=data["width"],
            height=data["height"], backgrounds=self.background[None], absgrad=True, packed=False)
        return {"render_img": torch.clamp(imgs[0], min=0.0, max=1.0), "batch_xys": meta["means2d"], "batch_radii": meta["radii"]}

    __call__ = forward

    def update_statistics(self, data, out):                                               # gaussian.py:188-197
        max_hw = max(data["height"], data["width"])
        radii = out["batch_radii"].detach()[0] / max_hw
        absgrad = out["batch_xys"].absgrad.detach()[0]
        vis = radii > 0.0
        self.max_radii[vis] = torch.max(self.max_radii[vis], radii[vis])
        self.grad_norm_accum[vis] = self.grad_norm_accum[vis] + (torch.norm(absgrad, dim=-1) * max_hw)[vis]
        self.collecting_counts[vis] = self.collecting_counts[vis] + 1

    def _rebind(self, new: Dict[str, Tensor], state_fn):
        """Replace every parameter by a fresh leaf and carry the Adam state over (gaussian.py:199-257)."""
        for group in self.optimizer.param_groups:
            name, old = group["name"], group["params"][0]
            st = self.optimizer.state[old]
            st["exp_avg"], st["exp_avg_sq"] = state_fn(name, st["exp_avg"]), state_fn(name, st["exp_avg_sq"])
            del self.optimizer.state[old]
            self.p[name] = new[name].detach().requires_grad_(True)
            group["params"][0] = self.p[name]
            self.optimizer.state[self.p[name]] = st

    def densify_and_prune(self) -> Dict[str, Any]:                                        # gaussian.py:259-349
        c = self.cfg
        avg = self.grad_norm_accum / (self.collecting_counts + 1e-8)
        avg[avg.isnan()] = 0.0
        hot = avg >= c["densify_grad_thresh"]
        big = self.scales.amax(dim=-1) >= c["densify_scale_thresh"]
        split, clone = big & hot, ~big & hot
        extra: List[Dict[str, Tensor]] = []
        if torch.sum(split) != 0:                                                         # gaussian.py:164-186
            k = c["num_splits"]
            rep = lambda t: t[split].repeat(k, *([1] * (t.dim() - 1)))
            noise = torch.randn((int(torch.sum(split)) * k, 3), device=self.dev)
            offs = torch.bmm(_quat_to_rotmat(rep(self.p["quats"])), (rep(self.scales) * noise).unsqueeze(-1)).squeeze(-1)
            extra.append(dict(means=rep(self.p["means"]) + offs, log_scales=torch.log(rep(self.scales) / (0.8 * k)),
                              quats=rep(self.p["quats"]), sh_0=rep(self.p["sh_0"]), sh_rest=rep(self.p["sh_rest"]),
                              logit_opacities=rep(self.p["logit_opacities"])))
        if torch.sum(clone) != 0:                                                         # gaussian.py:146-162
            extra.append({k_: self.p[k_][clone] for k_ in NAMES})
        if extra:                                                                         # gaussian.py:199-234
            add = {k_: torch.cat([e[k_] for e in extra], dim=0) for k_ in NAMES}
            self._rebind({k_: torch.cat([self.p[k_], add[k_]], dim=0) for k_ in NAMES},
                         lambda name, s: torch.cat([s, torch.zeros_like(add[name])], dim=0))
        grown = self.n - self.max_radii.shape[0]
        if grown > 0:                                                                     # gaussian.py:296-313
            pad = torch.zeros(grown, device=self.dev)
            self.max_radii = torch.cat([self.max_radii, pad])
            split = torch.cat([split, pad])
        prune = self.opacities < c["min_opacity"]                                         # gaussian.py:314-326
        prune = prune | (self.max_radii > c["prune_radii_ratio_thresh"])
        prune = prune | (self.scales.amax(dim=-1) > c["prune_scale_thresh"])
        prune = torch.logical_or(prune, split)
        if torch.sum(prune) != 0:                                                         # gaussian.py:236-257
            keep = ~prune
            self._rebind({k_: self.p[k_][keep] for k_ in NAMES}, lambda name, s: s[keep])
        self._reset_stats()
        return {"n": self.n, "split": int(split.sum()), "clone": int(clone.sum())}

    def reset_opacities(self):                                                            # gaussian.py:129-144
        cap = torch.full_like(self.p["logit_opacities"], self.cfg["min_opacity"] * 2.0)
        new = torch.logit(torch.min(self.opacities * 0.5, cap))
        for group in self.optimizer.param_groups:
            if group["name"] == "logit_opacities":
                old = group["params"][0]
                st = self.optimizer.state[old]
                st["exp_avg"], st["exp_avg_sq"] = torch.zeros_like(st["exp_avg"]), torch.zeros_like(st["exp_avg_sq"])
                del self.optimizer.state[old]
                self.p["logit_opacities"] = new.detach().requires_grad_(True)
                group["params"][0] = self.p["logit_opacities"]
                self.optimizer.state[self.p["logit_opacities"]] = st

    def up_sh_degree(self):                                                               # gaussian.py:118-119
        self.active_sh_degree = min(self.active_sh_degree + 1, self.cfg["sh_degree"])

    def update_learning_rate(self, step: int):                                            # gaussian.py:121-127, utils.py:19-28
        c = self.cfg
        t = min(1.0, step / c["means_lr_schedule_max_steps"])
        lr = np.exp(np.log(c["means_lr_init"]) * (1 - t) + np.log(c["means_lr_final"]) * t)
        for group in self.optimizer.param_groups:
            if group["name"] == "means":
                group["lr"] = lr

    def parameters_dict(self) -> Dict[str, Tensor]:
        return {k: v.detach() for k, v in self.p.items()}


class ReferenceModel:
    """The reference's own GaussianModel + optimizer behind the same small surface as CallerModel."""

    def __init__(self, ref_module, pointcloud_cls, xyzs, rgbs, cfg):
        c = cfg
        with cuda_means_cpu():
            self.model = ref_module.GaussianModel(
                pointcloud_cls(xyzs, rgbs), c["sh_degree"], c["sh_degree_interval"], c["means_lr_init"], c["means_lr_final"],
                c["means_lr_schedule_max_steps"], c["densify_grad_thresh"], c["densify_scale_thresh"], c["num_splits"],
                c["prune_radii_ratio_thresh"], c["prune_scale_thresh"], c["min_opacity"], c["use_scale_regularization"],
                c["max_scale_ratio"], c["white_background"])                               # train.py:52-69
            self.optimizer = ref_module.build_optimizers(self.model, c["means_lr_init"], c["log_scales_lr"], c["quats_lr"],
                                                         c["sh_0_lr"], c["sh_rest_lr"], c["logit_opacities_lr"])  # train.py:70-78
            self.loss_computer = ref_module.LossComputer(self.model, c["lambda_ssim"], c["lambda_scale"])  # train.py:79-81

    n = property(lambda self: self.model.nbr_gaussians)

    def __call__(self, data):
        return self.model(data)

    def __getattr__(self, name):  # update_statistics, densify_and_prune, reset_opacities, up_sh_degree, ...
        return getattr(self.model, name)

    def parameters_dict(self):
        return {k: getattr(self.model, k).detach() for k in NAMES}


def photometric_loss(ssim_fn, render_img: Tensor, gt_img: Tensor, mask: Tensor, lambda_ssim: float) -> Tensor:
    """LossComputer.get_loss_dict()["total"] without the scale regulariser (gaussian.py:422-453)."""
    m3 = mask.unsqueeze(2).repeat(1, 1, 3)
    img = m3 * gt_img + (1.0 - m3) * render_img
    l1 = F.l1_loss(img, gt_img)
    s = 1.0 - ssim_fn(gt_img.permute(2, 0, 1)[None], img.permute(2, 0, 1)[None], data_range=1.0)
    return (1.0 - lambda_ssim) * l1 + lambda_ssim * s


def train_loop(model, frames: List[Dict[str, Any]], cfg: Dict[str, Any], steps: int, loss_fn, on_step=None) -> Dict[str, Any]:
    """train.py:93-157 for ``steps`` iterations over ``frames`` (round robin).  ``loss_fn(render, gt, mask)``."""
    hist = {"loss": [], "n": [], "events": []}
    ctx = cuda_means_cpu()
    with ctx:
        for step in range(1, steps + 1):
            data = frames[(step - 1) % len(frames)]
            out = model(data)                                                             # train.py:98
            loss = loss_fn(out["render_img"], data["image"], data["mask"])                # train.py:99-103
            loss.backward()                                                               # train.py:104
            hist["loss"].append(float(loss.item()))                                       # train.py:106-108
            with torch.no_grad():
                if cfg["refine_start"] < step <= cfg["refine_stop"]:                      # train.py:129-136
                    model.update_statistics(data, out)
                    if (step - cfg["refine_start"]) % cfg["refine_every"] == 0:
                        info = model.densify_and_prune()
                        hist["events"].append((step, "densify", model.n))
                    if (step - cfg["refine_start"]) % cfg["reset_opacities_every"] == 0:
                        model.reset_opacities()
                        hist["events"].append((step, "reset", model.n))
                if cfg["sh_degree_interval"] != 0 and step % cfg["sh_degree_interval"] == 0:  # train.py:138-139
                    model.up_sh_degree()
                model.update_learning_rate(step)                                          # train.py:141
            model.optimizer.step()                                                        # train.py:156
            model.optimizer.zero_grad()                                                   # train.py:157
            hist["n"].append(model.n)
            if on_step is not None:
                on_step(step, model)
    return hist


def viewer_render(model, data: Dict[str, Any]) -> np.ndarray:
    """The viewer closure (train.py:172-183, launch_viewer.py:29-37): a no_grad render that ends on the host."""
    with torch.no_grad():
        return model({k: data[k] for k in ("w2c", "K", "height", "width")})["render_img"].cpu().numpy()


# ------------------------------------------------------------------------------------------------------------------
# tiny synthetic dataset: target images are renders of a hidden "ground truth" Gaussian set
# ------------------------------------------------------------------------------------------------------------------
def make_dataset(n_gt: int, n_init: int, width: int, height: int, n_views: int, seed: int, device, render):
    """-> (xyzs[n_init,3] float64, rgbs[n_init,3] uint8, frames).  ``render(sc) -> [V,H,W,3]`` draws the targets (the
    oracle on CPU, this repo's rasterizer on the GPU)."""
    from easy_gaussian_splatting_b200.synthetic import make_scene
    fx = 0.9 * width
    sc = make_scene("blob", n_gt, width, height, fx, seed, n_views=n_views, white_background=True)
    imgs = render(sc)
    g = torch.Generator().manual_seed(seed + 77)
    pick = torch.randperm(n_gt, generator=g)[:n_init]
    xyz = (sc.means[pick] + 0.02 * torch.randn(n_init, 3, generator=g)).double().numpy()
    rgb = (torch.clamp(sc.colors[pick, 0] * 0.28209479177387814 + 0.5, 0, 1) * 255).round().to(torch.uint8).numpy()
    frames = []
    for v in range(n_views):
        frames.append({"K": sc.Ks[v].to(device), "w2c": sc.viewmats[v].to(device), "height": height, "width": width,
                       "image": imgs[v].clamp(0, 1).to(device), "mask": torch.zeros(height, width, device=device)})
    return xyz, rgb, frames
