"""CPU tests: the per-Gaussian math the CUDA kernels run (csrc/egs_math.cuh), compiled for the host by
tests/host_harness, against the oracle.  Forward must be bit-exact (it feeds integer outputs); the
hand-derived backward must agree with fp64 autograd."""
import numpy as np
import pytest
import torch

from easy_gaussian_splatting_b200.synthetic import make_scene
from oracle import gsplat_oracle as O
from tests import host_harness as hh

CASES = [
    dict(kind="blob", N=6000, width=256, height=256, fx=274.5, seed=0, n_views=2),
    dict(kind="outdoor", N=40_000, width=979, height=546, fx=581.0, seed=2, n_views=2),
    dict(kind="object", N=20_000, width=800, height=800, fx=1111.11, seed=1, n_views=1),
    dict(kind="blob", N=3000, width=33, height=17, fx=15.0, seed=4, n_views=1),  # wide fov: clamp active
]


@pytest.mark.parametrize("cfg", CASES)
def test_projection_forward_bit_exact(cfg):
    sc = make_scene(**cfg)
    h = hh.projection_fwd(sc, 3)
    radii, m2, dep, con = O.fully_fused_projection(sc.means, sc.quats, sc.scales, sc.viewmats, sc.Ks, sc.width, sc.height)
    assert np.array_equal(h["radii"], radii.numpy())
    assert np.array_equal(h["means2d"], m2.numpy())
    assert np.array_equal(h["depths"], dep.numpy())
    assert np.array_equal(h["conics"], con.numpy())
    comp = O.fully_fused_projection(sc.means, sc.quats, sc.scales, sc.viewmats, sc.Ks, sc.width, sc.height,
                                    calc_compensations=True)[4]
    assert np.array_equal(h["compensations"], comp.numpy())  # antialiased-mode factor, bit-exact too
    tw, th = -(-sc.width // 16), -(-sc.height // 16)
    tpg, _, _ = O.isect_tiles(m2, radii, dep, 16, tw, th, sort=False)
    assert np.array_equal(h["tiles_per_gauss"], tpg.numpy())
    C = sc.viewmats.shape[0]
    campos = torch.inverse(sc.viewmats)[:, :3, 3]
    dirs = sc.means[None] - campos[:, None]
    cols = torch.clamp_min(O.spherical_harmonics(3, dirs, sc.colors[None].expand(C, -1, -1, -1), radii > 0) + 0.5, 0)
    cols = torch.where((radii > 0)[..., None], cols, torch.zeros(()))
    assert np.abs(h["colors"] - cols.numpy()).max() <= 2e-6


@pytest.mark.parametrize("cfg,deg", [(CASES[0], 3), (CASES[1], 2), (CASES[3], 1), (CASES[0], 0)])
def test_projection_sh_backward_vs_fp64_autograd(cfg, deg):
    cfg = dict(cfg)
    cfg["N"] = min(cfg["N"], 5000)
    sc = make_scene(**cfg)
    V, N = sc.viewmats.shape[0], sc.means.shape[0]
    h = hh.projection_fwd(sc, deg)
    m = sc.means.double().requires_grad_(True)
    q = sc.quats.double().requires_grad_(True)
    s = sc.scales.double().requires_grad_(True)
    sh = sc.colors.double().requires_grad_(True)
    vm, Ks = sc.viewmats.double(), sc.Ks.double()
    radii, m2, dep, con = O.fully_fused_projection(m, q, s, vm, Ks, sc.width, sc.height)
    vis = torch.from_numpy(h["radii"] > 0)
    campos = torch.inverse(vm)[:, :3, 3]
    cols = torch.clamp_min(O.spherical_harmonics(deg, m[None] - campos[:, None], sh[None].expand(V, -1, -1, -1), vis) + 0.5, 0)
    g = torch.Generator().manual_seed(1)
    v_m2 = torch.randn(V, N, 2, generator=g, dtype=torch.float64)
    v_con = torch.randn(V, N, 3, generator=g, dtype=torch.float64)
    v_col = torch.randn(V, N, 3, generator=g, dtype=torch.float64)
    visf = vis[..., None].double()
    ((m2 * v_m2 * visf).sum() + (con * v_con * visf).sum() + (cols * v_col * visf).sum()).backward()
    out = hh.projection_bwd(sc, deg, h["radii"], h["colors"], v_m2.float(), v_con.float(), v_col.float())
    for name, ref in (("v_means", m.grad), ("v_quats", q.grad), ("v_scales", s.grad), ("v_sh", sh.grad)):
        a = torch.from_numpy(out[name]).double()
        rel = ((a - ref).norm() / ref.norm().clamp_min(1e-30)).item()
        assert rel <= 2e-5, (name, rel)
    nb = (deg + 1) ** 2
    assert float(np.abs(out["v_sh"][:, nb:]).sum()) == 0.0  # inactive bands get exactly zero


@pytest.mark.parametrize("cfg", [CASES[0], CASES[3]])
def test_compensation_backward_vs_fp64_autograd(cfg):
    """rasterize_mode="antialiased": the gradient of the compensation factor through the 2D covariance
    (gsplat 1.0.0 add_blur_vjp) against fp64 autograd of sqrt(det_orig / det_blur)."""
    cfg = dict(cfg)
    cfg["N"] = min(cfg["N"], 4000)
    sc = make_scene(**cfg)
    V, N = sc.viewmats.shape[0], sc.means.shape[0]
    h = hh.projection_fwd(sc, 0)
    m = sc.means.double().requires_grad_(True)
    q = sc.quats.double().requires_grad_(True)
    s = sc.scales.double().requires_grad_(True)
    comp = O.fully_fused_projection(m, q, s, sc.viewmats.double(), sc.Ks.double(), sc.width, sc.height,
                                    calc_compensations=True)[4]
    vis = torch.from_numpy(h["radii"] > 0)
    assert 0.0 < float(comp[vis].min()) and float(comp[vis].max()) < 1.0
    g = torch.Generator().manual_seed(5)
    v_comp = torch.randn(V, N, generator=g, dtype=torch.float64)
    (comp * v_comp * vis.double()).sum().backward()
    z2, z3 = torch.zeros(V, N, 2), torch.zeros(V, N, 3)
    out = hh.projection_bwd(sc, -1, h["radii"], h["colors"], z2, z3, z3, v_comps=v_comp.float())
    for name, ref in (("v_means", m.grad), ("v_quats", q.grad), ("v_scales", s.grad)):
        a = torch.from_numpy(out[name]).double()
        rel = ((a - ref).norm() / ref.norm().clamp_min(1e-30)).item()
        assert rel <= 5e-5, (name, rel)


def test_tight_rectangles_never_drop_a_reachable_pixel():
    """The tight tile rectangle the projection kernels pack for the blend kernels' lists (egs_math.cuh:
    tighten_tile_rect / pack_tile_rect, compiled for the host) against brute force over the pixels: every tile of
    gsplat's rectangle that holds a pixel centre with alpha = o exp(-sigma) >= 1/255 must lie inside the tight
    rectangle, the tight rectangle must lie inside gsplat's, and it must actually be tighter on anisotropic splats."""
    from tests import host_harness
    rng = np.random.default_rng(0)
    n, W, H, T = 4000, 640, 368, 16
    tw, th = -(-W // T), -(-H // T)
    # covariances: random orientation, axis ratios up to 30:1, sizes from sub-pixel to a few tiles
    ang = rng.uniform(0, np.pi, n)
    s1 = np.exp(rng.uniform(np.log(0.3), np.log(40.0), n))
    s2 = s1 / np.exp(rng.uniform(0, np.log(30.0), n))
    c, s = np.cos(ang), np.sin(ang)
    cxx = c * c * s1 ** 2 + s * s * s2 ** 2 + 0.3
    cyy = s * s * s1 ** 2 + c * c * s2 ** 2 + 0.3
    cxy = c * s * (s1 ** 2 - s2 ** 2)
    det = cxx * cyy - cxy ** 2
    conics = np.stack([cyy / det, -cxy / det, cxx / det], -1).astype(np.float32)
    # gsplat's radius: ceil(3 sqrt(larger eigenvalue))
    mid = 0.5 * (cxx + cyy)
    lam = mid + np.sqrt(np.maximum(mid ** 2 - det, 0.01))
    radii = np.ceil(3.0 * np.sqrt(lam)).astype(np.int32)
    means2d = np.stack([rng.uniform(-20, W + 20, n), rng.uniform(-20, H + 20, n)], -1).astype(np.float32)
    opac = np.concatenate([rng.uniform(0.002, 1.0, n - 200), rng.uniform(0.0, 1 / 255.0, 200)]).astype(np.float32)
    classic, packed = host_harness.tight_rects(means2d, radii, conics, opac, tw, th, T)
    x0, y0 = packed[:, 0] & 0xffff, (packed[:, 0] >> 16) & 0xffff
    w, h = packed[:, 1] & 0xffff, (packed[:, 1] >> 16) & 0xffff
    cw, ch = classic[:, 2] - classic[:, 0], classic[:, 3] - classic[:, 1]
    nonempty = (w > 0) & (h > 0)
    # inside gsplat's rectangle
    assert np.all(x0[nonempty] >= classic[nonempty, 0]) and np.all((x0 + w)[nonempty] <= classic[nonempty, 2])
    assert np.all(y0[nonempty] >= classic[nonempty, 1]) and np.all((y0 + h)[nonempty] <= classic[nonempty, 3])
    # brute force: reachable pixels of every tile of gsplat's rectangle, in fp64
    dropped = 0
    reach_tiles = 0
    for i in range(n):
        if cw[i] <= 0 or ch[i] <= 0:
            continue
        px = np.arange(classic[i, 0] * T, min(classic[i, 2] * T, W)) + 0.5
        py = np.arange(classic[i, 1] * T, min(classic[i, 3] * T, H)) + 0.5
        if px.size == 0 or py.size == 0:
            continue
        dx = px[None, :] - float(means2d[i, 0])
        dy = py[:, None] - float(means2d[i, 1])
        a_, b_, c_ = (float(v) for v in conics[i])
        sigma = 0.5 * (a_ * dx * dx + c_ * dy * dy) + b_ * dx * dy
        alpha = float(opac[i]) * np.exp(-sigma)
        hit = (sigma >= 0) & (alpha >= 1.0 / 255.0)
        ys, xs = np.nonzero(hit)
        if ys.size == 0:
            continue
        tx = (px[xs] // T).astype(np.int64)
        ty = (py[ys] // T).astype(np.int64)
        reach_tiles += np.unique(ty * tw + tx).size
        inside = (tx >= x0[i]) & (tx < x0[i] + w[i]) & (ty >= y0[i]) & (ty < y0[i] + h[i])
        dropped += int((~inside).sum())
    assert dropped == 0, f"{dropped} reachable pixels lie outside their Gaussian's tight rectangle"
    classic_tiles = int((np.maximum(cw, 0) * np.maximum(ch, 0)).sum())
    tight_tiles = int((w.astype(np.int64) * h).sum())
    print(f"classic {classic_tiles} tiles, tight {tight_tiles} ({tight_tiles / classic_tiles:.3f}), exactly reachable {reach_tiles}")
    assert reach_tiles <= tight_tiles < 0.8 * classic_tiles
    # Gaussians that can never reach 1/255 are listed nowhere
    assert np.all(w[opac <= 1 / 255.0] == 0)
