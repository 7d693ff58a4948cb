"""CPU tests: the reference's OWN caller (``/root/reference/model/gaussian.py``, imported unmodified) trains through
this repo's ``rasterization()`` — signature, ``meta`` contract, ``.absgrad`` tagging, N changing between calls,
parameters re-created by densify / prune, the ``.grad is None`` optimizer step, a viewer-thread render — with the
CUDA stage operators emulated by the oracle (tests/oracle_stages.py; there is no GPU in the build container and no
``/root/reference`` on the GPU box, so this is the one place where both sides can meet).

It also pins tests/ref_caller.CallerModel — the restated caller that the ``-m gpu`` swap test drives on the real
kernels — to the reference: same seed, same rasterizer => bit-identical training trajectory.
(SURVEY.md §4 "end-to-end GaussianModel.forward swap test"; VERDICT r1 "What's missing" #1, #7.)
"""
import threading

import pytest
import torch

from oracle import gsplat_oracle as O
from oracle import loss_oracle
from tests import oracle_stages, ref_caller
from tests.ref_caller import TINY_CFG

needs_reference = pytest.mark.skipif(not ref_caller.reference_available(), reason="/root/reference is not on this machine")


def _oracle_targets(sc):
    with torch.no_grad():
        rc, _, _ = O.rasterization(sc.means, sc.quats, sc.scales, sc.opacities, sc.colors, sc.viewmats, sc.Ks, sc.width,
                                   sc.height, sh_degree=3, packed=False, backgrounds=sc.background[None].expand(sc.viewmats.shape[0], 3))
    return rc


@pytest.fixture()
def emulated(monkeypatch):
    oracle_stages.install(monkeypatch)
    yield
    ref_caller.forget_reference()


def test_emulated_stages_reproduce_the_oracle(emulated):
    """The emulation is only a re-plumbing of the oracle: rendering.rasterization through it must equal
    oracle.rasterization (images bit for bit, gradients to summation order)."""
    from easy_gaussian_splatting_b200 import rasterization
    from easy_gaussian_splatting_b200.synthetic import loss_weights, make_scene
    from tests.util import PARAMS, rel_err
    sc = make_scene("blob", 300, 70, 45, 60.0, 5, n_views=2)
    C = 2
    Wc, Wa = loss_weights(sc.seed, C, sc.height, sc.width)
    bg = sc.background[None].expand(C, 3).contiguous()
    outs = []
    for fn in (rasterization, O.rasterization):
        leaves = [getattr(sc, k).clone().requires_grad_(True) for k in PARAMS]
        rc, ra, meta = fn(*leaves, sc.viewmats, sc.Ks, sc.width, sc.height, sh_degree=2, packed=False, absgrad=True, backgrounds=bg)
        ((rc * Wc).sum() + (ra * Wa).sum()).backward()
        outs.append((rc.detach(), ra.detach(), meta, [t.grad for t in leaves], meta["means2d"].absgrad))
    a, b = outs
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    for k in ("radii", "isect_ids", "flatten_ids", "isect_offsets", "tiles_per_gauss"):
        assert torch.equal(a[2][k], b[2][k]), k
    for ga, gb, k in zip(a[3], b[3], PARAMS):
        assert rel_err(ga, gb) <= 1e-5, k
    assert rel_err(a[4], b[4]) <= 1e-6


def _dataset():
    return ref_caller.make_dataset(n_gt=160, n_init=60, width=48, height=40, n_views=4, seed=11, device="cpu",
                                   render=_oracle_targets)


STEPS = 140


@needs_reference
def test_reference_gaussian_model_trains_unchanged_through_this_rasterizer(emulated):
    ref, Pointcloud = ref_caller.import_reference(loss_oracle.ssim)
    xyz, rgb, frames = _dataset()
    torch.manual_seed(0)
    model = ref_caller.ReferenceModel(ref, Pointcloud, xyz, rgb, TINY_CFG)
    lc = model.loss_computer
    seen = {"absgrad_shapes": [], "grad_none_steps": 0, "thread": None}

    def on_step(step, m):
        # /root/reference/viewer/viewer_runtime.py:20,100: renders come from a client thread while training runs
        if step == 45:
            out = {}
            t = threading.Thread(target=lambda: out.update(img=ref_caller.viewer_render(m, frames[1])))
            t.start()
            t.join()
            seen["thread"] = out["img"]

    def loss_fn(render, gt, mask):
        return lc.get_loss_dict(render, gt, mask)["total"]  # the reference's own LossComputer (gaussian.py:415-445)

    n0 = model.n
    hist = ref_caller.train_loop(model, frames, TINY_CFG, STEPS, loss_fn, on_step=on_step)
    ns = hist["n"]
    assert len(set(ns)) >= 3, f"N must change between calls (densify + prune): {sorted(set(ns))}"
    assert any(kind == "reset" for _, kind, _ in hist["events"])
    assert all(torch.isfinite(torch.tensor(hist["loss"])))
    first, last = sum(hist["loss"][:8]) / 8, sum(hist["loss"][-8:]) / 8
    assert last < 0.8 * first, (first, last)
    assert seen["thread"] is not None and seen["thread"].shape == (40, 48, 3)
    assert model.model.grad_norm_accum.shape == (model.n,) and model.model.max_radii.shape == (model.n,)
    assert model.model.active_sh_degree == 3
    assert n0 == 60
    for k, v in model.parameters_dict().items():
        assert torch.isfinite(v).all(), k


@needs_reference
def test_restated_caller_is_bit_identical_to_the_reference_caller(emulated):
    """Pins tests/ref_caller.CallerModel (used on the GPU box, where /root/reference does not exist)."""
    ref, Pointcloud = ref_caller.import_reference(loss_oracle.ssim)
    xyz, rgb, frames = _dataset()
    torch.manual_seed(0)
    a = ref_caller.ReferenceModel(ref, Pointcloud, xyz, rgb, TINY_CFG)
    ha = ref_caller.train_loop(a, frames, TINY_CFG, STEPS, lambda r, g, m: a.loss_computer.get_loss_dict(r, g, m)["total"])
    torch.manual_seed(0)
    b = ref_caller.CallerModel(xyz, rgb, TINY_CFG, "cpu")
    hb = ref_caller.train_loop(b, frames, TINY_CFG, STEPS,
                               lambda r, g, m: ref_caller.photometric_loss(loss_oracle.ssim, r, g, m, TINY_CFG["lambda_ssim"]))
    assert ha["n"] == hb["n"] and ha["events"] == hb["events"]
    assert ha["loss"] == hb["loss"]
    pa, pb = a.parameters_dict(), b.parameters_dict()
    for k in ref_caller.NAMES:
        assert torch.equal(pa[k], pb[k]), k
    for name in ("grad_norm_accum", "collecting_counts", "max_radii"):
        assert torch.equal(getattr(a.model, name), getattr(b, name)), name
    # optimizer state followed the surgery identically
    for ga, gb in zip(a.optimizer.param_groups, b.optimizer.param_groups):
        sa, sb = a.optimizer.state[ga["params"][0]], b.optimizer.state[gb["params"][0]]
        assert ga["name"] == gb["name"] and ga["lr"] == gb["lr"]
        assert torch.equal(sa["exp_avg"], sb["exp_avg"]) and torch.equal(sa["exp_avg_sq"], sb["exp_avg_sq"])


def test_restated_caller_runs_without_the_reference_tree(emulated):
    """What the GPU box runs (there with the real kernels): the restated caller alone, shorter."""
    xyz, rgb, frames = _dataset()
    torch.manual_seed(0)
    m = ref_caller.CallerModel(xyz, rgb, TINY_CFG, "cpu")
    hist = ref_caller.train_loop(m, frames, TINY_CFG, 60,
                                 lambda r, g, mk: ref_caller.photometric_loss(loss_oracle.ssim, r, g, mk, 0.2))
    assert len(set(hist["n"])) >= 2 and all(l == l for l in hist["loss"])
