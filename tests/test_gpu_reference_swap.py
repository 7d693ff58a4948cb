"""-m gpu: the reference's call pattern on the real kernels (VERDICT r1 "What's missing" #1, #7).

``/root/reference`` does not exist on the GPU box, so the caller is tests/ref_caller.CallerModel — the restatement of
``GaussianModel`` (/root/reference/model/gaussian.py:12-374) and of the loop body of /root/reference/train.py:93-157 that
tests/test_reference_swap.py proves bit-identical to the reference's own code on CPU.  If the reference tree IS present
(a developer box with a GPU), the reference's own module is driven instead.

Covered: >= 300 iterations of forward -> L1 + SSIM loss -> backward -> update_statistics -> densify_and_prune (N
changes, every parameter re-created, the optimizer step that follows sees ``.grad is None``) -> reset_opacities ->
optimizer.step; a viewer render from a non-main thread in the middle of training (viewer_runtime.py:20,100); renders
from a second thread running concurrently with the training thread's forward + backward.
"""
import threading

import pytest
import torch

from oracle import loss_oracle
from tests import ref_caller

pytestmark = pytest.mark.gpu

CFG = dict(ref_caller.TINY_CFG, refine_start=30, refine_stop=330, refine_every=40, reset_opacities_every=120,
           sh_degree_interval=60, means_lr_schedule_max_steps=360, densify_scale_thresh=0.03)


def _gpu_targets(sc):
    from easy_gaussian_splatting_b200 import rasterization
    d = sc.to("cuda")
    C = d.viewmats.shape[0]
    with torch.no_grad():
        rc, _, _ = rasterization(d.means, d.quats, d.scales, d.opacities, d.colors, d.viewmats, d.Ks, d.width, d.height,
                                 sh_degree=3, packed=False, backgrounds=d.background[None].expand(C, 3).contiguous())
    return rc.cpu()


def _model(xyz, rgb):
    if ref_caller.reference_available():
        ref, Pointcloud = ref_caller.import_reference(loss_oracle.ssim)
        m = ref_caller.ReferenceModel(ref, Pointcloud, xyz, rgb, CFG)
        return m, (lambda r, g, k: m.loss_computer.get_loss_dict(r, g, k)["total"])
    m = ref_caller.CallerModel(xyz, rgb, CFG, "cuda")
    return m, (lambda r, g, k: ref_caller.photometric_loss(loss_oracle.ssim, r, g, k, CFG["lambda_ssim"]))


def test_reference_training_loop_on_the_real_kernels():
    xyz, rgb, frames = ref_caller.make_dataset(n_gt=20_000, n_init=6_000, width=208, height=160, n_views=8, seed=3,
                                               device="cuda", render=_gpu_targets)
    torch.manual_seed(0)
    model, loss_fn = _model(xyz, rgb)
    seen = {}

    def on_step(step, m):
        if step in (100, 250):  # a client thread of the viewer asks for a frame while training is under way
            out = {}

            def client():
                out["img"] = ref_caller.viewer_render(m, frames[step % len(frames)])
            t = threading.Thread(target=client)
            t.start()
            t.join()
            seen[step] = out["img"]
            # the same frame from the training thread: bit-identical (the forward pass is deterministic)
            same = ref_caller.viewer_render(m, frames[step % len(frames)])
            assert (same == out["img"]).all()

    hist = ref_caller.train_loop(model, frames, CFG, 340, loss_fn, on_step=on_step)
    torch.cuda.synchronize()
    ns = hist["n"]
    print("N over time:", sorted(set(ns)), "events:", hist["events"][:12])
    print("loss first/last 10:", sum(hist["loss"][:10]) / 10, sum(hist["loss"][-10:]) / 10)
    assert len(set(ns)) >= 4 and max(ns) > ns[0], "densify / prune must change N several times"
    assert any(n % 4 for n in ns), "some N must be odd-sized (alignment of every per-Gaussian buffer)"
    assert any(kind == "reset" for _, kind, _ in hist["events"])
    assert all(l == l and l < 1e3 for l in hist["loss"])
    assert sum(hist["loss"][-10:]) < 0.75 * sum(hist["loss"][:10]), "the loop must actually optimise the scene"
    assert set(seen) == {100, 250} and all(v.shape == (160, 208, 3) for v in seen.values())
    for k, v in model.parameters_dict().items():
        assert torch.isfinite(v).all(), k
    ref_caller.forget_reference()


def test_second_thread_renders_while_the_training_thread_runs():
    """Two host threads inside rasterization() at the same time (one under no_grad, one with autograd and a direct
    gradient bucket): each must get exactly what it would have got alone.  The render thread never calls
    ``torch.cuda.set_device`` (viewer threads do not)."""
    from easy_gaussian_splatting_b200 import rasterization
    from easy_gaussian_splatting_b200.distributed import FlatGradBucket
    from easy_gaussian_splatting_b200.synthetic import loss_weights, make_scene
    from tests.util import PARAMS, rel_err
    sc = make_scene("outdoor", 60_000, 489, 273, 290.0, 2, n_views=2).to("cuda")
    Wc, Wa = (t.cuda() for t in loss_weights(sc.seed, 1, sc.height, sc.width))
    bg = sc.background[None]
    params = [getattr(sc, k).clone().requires_grad_(True) for k in PARAMS]
    bucket = FlatGradBucket(params)

    def train_once():
        with bucket.direct():
            rc, ra, _ = rasterization(*params, sc.viewmats[:1], sc.Ks[:1], sc.width, sc.height, sh_degree=3, packed=False,
                                      absgrad=True, backgrounds=bg)
            ((rc * Wc).sum() + (ra * Wa).sum()).backward()
        return [p.grad.clone() for p in params]

    def view_once():
        with torch.no_grad():
            return rasterization(*[p.detach() for p in params], sc.viewmats[1:], sc.Ks[1:], sc.width, sc.height, sh_degree=3,
                                 packed=False, backgrounds=bg)[0]

    g_ref, img_ref = train_once(), view_once()
    torch.cuda.synchronize()
    imgs, errors = [], []

    def client():
        try:
            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                for _ in range(25):
                    imgs.append(view_once())
            s.synchronize()
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    t = threading.Thread(target=client)
    t.start()
    grads = [train_once() for _ in range(25)]
    t.join()
    torch.cuda.synchronize()
    assert not errors, errors
    assert len(imgs) == 25 and all(torch.equal(i, img_ref) for i in imgs)
    for g in grads:
        for a, b, k in zip(g, g_ref, PARAMS):
            assert rel_err(a.cpu(), b.cpu()) <= 1e-5, k
