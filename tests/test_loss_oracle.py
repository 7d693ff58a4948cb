"""CPU tests of the photometric-loss oracle (oracle/loss_oracle.py; reference LossComputer,
/root/reference/model/gaussian.py:415-453).  torchmetrics is absent, so the SSIM restatement is pinned by
independent computations: a direct (non-convolutional) evaluation of the published formula, known answers, and
fp64 finite differences."""
import numpy as np
import pytest
import torch

from oracle import loss_oracle as L


def _images(H, W, seed, dtype=torch.float64):
    g = torch.Generator().manual_seed(seed)
    gt = torch.rand(H, W, 3, generator=g, dtype=dtype)
    # smooth-ish render correlated with gt, in [0,1]
    render = (0.7 * gt + 0.3 * torch.rand(H, W, 3, generator=g, dtype=dtype)).clamp(0, 1)
    mask = (torch.rand(H, W, generator=g) < 0.2).to(dtype)
    return render, gt, mask


def test_window_is_the_published_gaussian():
    w = L.gaussian_window().double().numpy()
    d = np.arange(-5, 6, dtype=np.float64)
    ref = np.exp(-(d / 1.5) ** 2 / 2)
    ref /= ref.sum()
    assert w.shape == (11,) and abs(w.sum() - 1) < 1e-6 and np.abs(w - ref).max() < 1e-7


def test_ssim_equals_direct_valid_window_evaluation():
    """reflect-pad + conv + crop == the SSIM formula over every fully inside 11x11 window, evaluated with plain
    loops over numpy slices (no conv2d, no padding): the identity the CUDA kernel relies on."""
    render, gt, _ = _images(19, 23, 0)
    val = L.ssim(gt.permute(2, 0, 1)[None], render.permute(2, 0, 1)[None]).item()
    w = L.gaussian_window(dtype=torch.float64).numpy()
    w2 = np.outer(w, w)
    x, y = render.numpy(), gt.numpy()
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    acc = []
    for ch in range(3):
        for i in range(19 - 10):
            for j in range(23 - 10):
                a, b = x[i:i + 11, j:j + 11, ch], y[i:i + 11, j:j + 11, ch]
                mx, my = (w2 * a).sum(), (w2 * b).sum()
                sxx, syy, sxy = (w2 * a * a).sum() - mx * mx, (w2 * b * b).sum() - my * my, (w2 * a * b).sum() - mx * my
                acc.append((2 * mx * my + c1) * (2 * sxy + c2) / ((mx * mx + my * my + c1) * (sxx + syy + c2)))
    assert abs(val - float(np.mean(acc))) < 1e-12


def test_known_answers():
    render, gt, mask = _images(16, 16, 1)
    same = L.loss_dict(gt.clone(), gt, None, 0.2)
    assert abs(same["l1"].item()) == 0 and abs(same["ssim"].item()) < 1e-12 and abs(same["total"].item()) < 1e-12
    full_mask = L.loss_dict(render, gt, torch.ones(16, 16, dtype=torch.float64), 0.2)  # mask = 1 -> gt everywhere
    assert abs(full_mask["total"].item()) < 1e-12
    # constant images: mu = a, b; variances 0 -> SSIM = (2ab + c1) / (a^2 + b^2 + c1)
    a, b = 0.25, 0.75
    d = L.loss_dict(torch.full((12, 14, 3), a, dtype=torch.float64), torch.full((12, 14, 3), b, dtype=torch.float64), None, 0.5)
    s = (2 * a * b + 1e-4) / (a * a + b * b + 1e-4)
    assert abs(d["l1"].item() - 0.5) < 1e-12 and abs(d["ssim"].item() - (1 - s)) < 1e-9
    assert abs(d["total"].item() - (0.5 * 0.5 + 0.5 * (1 - s))) < 1e-9


@pytest.mark.parametrize("use_mask", [False, True])
def test_gradient_vs_finite_differences(use_mask):
    render, gt, mask = _images(14, 15, 2)
    render.requires_grad_(True)
    m = mask if use_mask else None
    L.loss_dict(render, gt, m, 0.2)["total"].backward()
    g = render.grad.clone()
    rng = np.random.default_rng(0)
    f = lambda r: L.loss_dict(r, gt, m, 0.2)["total"].item()
    base = render.detach()
    for _ in range(12):
        i, j, c = rng.integers(14), rng.integers(15), rng.integers(3)
        e = torch.zeros_like(base)
        e[i, j, c] = 1e-6
        fd = (f(base + e) - f(base - e)) / 2e-6
        assert abs(fd - g[i, j, c].item()) <= 1e-6 * max(1.0, abs(fd)) + 2e-9, (i, j, c, fd, g[i, j, c].item())
    if use_mask:
        assert float(g[mask.bool()].abs().sum()) == 0.0  # masked pixels take the ground truth: no gradient


def test_fused_loss_wrapper_validates_before_any_launch():
    """The product wrapper (easy_gaussian_splatting_b200/loss.py) rejects bad shapes and CPU tensors loudly."""
    from easy_gaussian_splatting_b200.loss import fused_l1_ssim_loss
    r, g = torch.rand(16, 16, 3), torch.rand(16, 16, 3)
    with pytest.raises(RuntimeError, match="no CPU path"):
        fused_l1_ssim_loss(r, g)
    with pytest.raises(ValueError, match="at least 11x11"):
        fused_l1_ssim_loss(r[:10], g[:10])
    with pytest.raises(ValueError, match=r"\[H,W,3\]"):
        fused_l1_ssim_loss(r[..., :2], g[..., :2])
    with pytest.raises(ValueError, match="mask must be"):
        fused_l1_ssim_loss(r, g, torch.zeros(16, 15))
