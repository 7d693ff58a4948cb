"""CPU tests of the oracle itself (no GPU): golden fixtures, hand-computable cases, invariants,
finite-difference gradients.  The reference ships no tests or vectors for this path,
so these are the checks that pin the oracle's semantics (Appendix A)."""
import math
from pathlib import Path

import numpy as np
import pytest
import torch

from easy_gaussian_splatting_b200.synthetic import loss_weights, make_scene
from oracle import gsplat_oracle as O
from tests.util import oracle_run

GOLDEN = sorted(Path(__file__).parent.glob("golden/*.npz"))


@pytest.mark.parametrize("path", GOLDEN, ids=[p.stem for p in GOLDEN])
def test_oracle_matches_golden(path):
    g = np.load(path)
    t = lambda k: torch.from_numpy(g[k])
    leaves = {k: t(f"in_{k}").clone().requires_grad_(True) for k in ("means", "quats", "scales", "opacities", "colors")}
    W, H, deg = int(g["in_width"]), int(g["in_height"]), int(g["in_sh_degree"])
    rc, ra, meta = O.rasterization(leaves["means"], leaves["quats"], leaves["scales"], leaves["opacities"], leaves["colors"],
                                   t("in_viewmats"), t("in_Ks"), W, H, sh_degree=deg, packed=False, absgrad=True,
                                   backgrounds=t("in_background"),
                                   rasterize_mode="antialiased" if "in_antialiased" in g else "classic")
    ((rc * t("in_Wc")).sum() + (ra * t("in_Wa")).sum()).backward()
    if "in_antialiased" in g:
        assert torch.equal(meta["opacities"].detach(), t("opacities"))
    for k in ("radii", "tiles_per_gauss", "isect_ids", "flatten_ids", "isect_offsets", "last_ids"):
        assert torch.equal(meta[k], t(k)), k
    assert torch.equal(meta["means2d"].detach(), t("means2d"))
    assert torch.equal(meta["depths"].detach(), t("depths"))
    assert torch.equal(meta["conics"].detach(), t("conics"))
    assert torch.allclose(rc.detach(), t("render_colors"), atol=1e-6, rtol=0)
    assert torch.allclose(ra.detach(), t("render_alphas"), atol=1e-6, rtol=0)
    for k, leaf in leaves.items():
        ref = t(f"grad_{k}")
        assert (leaf.grad - ref).norm() <= 1e-5 * ref.norm() + 1e-9, k
    assert (meta["means2d"].absgrad - t("absgrad")).norm() <= 1e-5 * t("absgrad").norm() + 1e-9


def _one_gaussian(opacity=0.6, scale=0.05, z=4.0, W=64, H=48, fx=80.0):
    means = torch.tensor([[0.0, 0.0, 0.0]])
    quats = torch.tensor([[1.0, 0.0, 0.0, 0.0]])
    scales = torch.full((1, 3), scale)
    opac = torch.tensor([opacity])
    colors = torch.tensor([[0.2, 0.5, 0.9]])
    V = torch.eye(4)[None].clone()
    V[0, 2, 3] = z
    K = torch.tensor([[[fx, 0, W / 2], [0, fx, H / 2], [0, 0, 1.0]]])
    return means, quats, scales, opac, colors, V, K, W, H


def test_single_isotropic_gaussian_known_answers():
    """One isotropic Gaussian on the optical axis: radius = ceil(3 sqrt(sigma_px^2 + 0.3)), alpha at the
    pixel whose centre is nearest = o * exp(-0.5 d^2 / (sigma_px^2 + 0.3)) (SURVEY.md §8c hand-computable case)."""
    means, quats, scales, opac, colors, V, K, W, H = _one_gaussian()
    rc, ra, meta = O.rasterization(means, quats, scales, opac, colors, V, K, W, H, sh_degree=None, packed=False)
    var = (80.0 * 0.05 / 4.0) ** 2 + 0.3
    assert int(meta["radii"][0, 0]) == math.ceil(3 * math.sqrt(var))
    assert torch.allclose(meta["means2d"][0, 0], torch.tensor([W / 2, H / 2]))
    assert torch.allclose(meta["conics"][0, 0], torch.tensor([1 / var, 0.0, 1 / var]), atol=1e-6)
    d2 = 0.5  # nearest pixel centre is (0.5, 0.5) away
    alpha = 0.6 * math.exp(-0.5 * d2 / var)
    y, x = H // 2, W // 2
    assert abs(float(ra[0, y, x, 0]) - alpha) < 1e-6
    assert torch.allclose(rc[0, y, x], alpha * colors[0], atol=1e-6)
    # far away pixels are untouched
    assert float(ra[0, 0, 0, 0]) == 0.0


def test_alpha_clamp_and_thresholds():
    means, quats, scales, opac, colors, V, K, W, H = _one_gaussian(opacity=1.0, scale=0.5)
    rc, ra, meta = O.rasterization(means, quats, scales, opac, colors, V, K, W, H, sh_degree=None, packed=False)
    assert float(ra.max()) <= 0.999 + 1e-7  # alpha is clamped at 0.999
    means, quats, scales, opac, colors, V, K, W, H = _one_gaussian(opacity=1.0 / 255.0 - 1e-4)
    rc, ra, meta = O.rasterization(means, quats, scales, opac, colors, V, K, W, H, sh_degree=None, packed=False)
    assert float(ra.abs().max()) == 0.0  # below 1/255 nothing is blended


def test_culling_rules():
    means, quats, scales, opac, colors, V, K, W, H = _one_gaussian()
    for dz, visible in ((4.0, True), (0.005, False), (-1.0, False), (2e10, False)):
        V2 = V.clone()
        V2[0, 2, 3] = dz
        radii, *_ = O.fully_fused_projection(means, quats, scales, V2, K, W, H)
        assert bool(radii[0, 0] > 0) == visible, dz
    # far off-screen
    V3 = V.clone()
    V3[0, 0, 3] = 100.0
    radii, m2, dep, con = O.fully_fused_projection(means, quats, scales, V3, K, W, H)
    assert int(radii[0, 0]) == 0 and float(m2.abs().sum() + dep.abs().sum() + con.abs().sum()) == 0.0
    # degenerate intrinsics must cull, not crash (viewer fov 0 / 180 deg)
    for fx in (float("inf"), float("nan")):
        K2 = K.clone()
        K2[0, 0, 0] = fx
        K2[0, 1, 1] = fx
        radii, m2, dep, con = O.fully_fused_projection(means, quats, scales, V, K2, W, H)
        assert int(radii[0, 0]) == 0
    # fx = 0 is degenerate but well defined: everything lands on the principal point with the blur-only footprint
    K2 = K.clone()
    K2[0, 0, 0] = 0.0
    K2[0, 1, 1] = 0.0
    radii, m2, dep, con = O.fully_fused_projection(means, quats, scales, V, K2, W, H)
    assert int(radii[0, 0]) == math.ceil(3 * math.sqrt(0.3 + math.sqrt(0.01))) and torch.isfinite(con).all()


def test_binning_invariants():
    sc = make_scene("outdoor", 20_000, 320, 200, 200.0, 5, n_views=2)
    radii, m2, dep, con = O.fully_fused_projection(sc.means, sc.quats, sc.scales, sc.viewmats, sc.Ks, sc.width, sc.height)
    tw, th = 20, 13
    tpg, ids, flat = O.isect_tiles(m2, radii, dep, 16, tw, th)
    assert int(tpg.sum()) == ids.numel() == flat.numel()
    assert (tpg[radii <= 0] == 0).all()
    assert (ids[1:] >= ids[:-1]).all(), "keys sorted"
    same = ids[1:] == ids[:-1]
    assert (flat[1:][same] > flat[:-1][same]).all(), "stable tie order = ascending flat index"
    offs = O.isect_offset_encode(ids, 2, tw, th).reshape(-1)
    assert (offs[1:] >= offs[:-1]).all() and int(offs[0]) == 0
    nb = O.tile_n_bits(tw, th)
    tile_of = ((ids >> 32) >> nb) * (tw * th) + ((ids >> 32) & ((1 << nb) - 1))
    ends = torch.cat([offs[1:], torch.tensor([ids.numel()], dtype=torch.int32)])
    for t in (0, 7, 100, 259, 260, 519):
        seg = tile_of[int(offs[t]):int(ends[t])]
        assert (seg == t).all()
    # depth bits of positive floats order like the floats
    d = dep.reshape(-1)[flat.long()]
    same_tile = tile_of[1:] == tile_of[:-1]
    assert (d[1:][same_tile] >= d[:-1][same_tile]).all()


def test_render_invariants_and_permutation():
    sc = make_scene("blob", 1500, 96, 64, 100.0, 21)
    out = oracle_run(sc, backward=False)
    a = out["alphas"]
    assert float(a.min()) >= 0.0 and float(a.max()) < 1.0
    # colour = sum w rgb + T bg  with bg = 0 here  =>  colour <= alpha * max rgb
    assert (out["colors"].amax(-1) <= a[..., 0] * out["meta"]["colors"].max() + 1e-6).all()
    # permuting the Gaussians does not change the image (ties in depth are measure-zero here)
    perm = torch.randperm(1500, generator=torch.Generator().manual_seed(0))
    sc2 = make_scene("blob", 1500, 96, 64, 100.0, 21)
    for k in ("means", "quats", "scales", "opacities", "colors"):
        setattr(sc2, k, getattr(sc2, k)[perm].contiguous())
    out2 = oracle_run(sc2, backward=False)
    assert torch.allclose(out["colors"], out2["colors"], atol=1e-6)


def test_autograd_matches_finite_differences_fp64():
    """fp64 central differences on a tiny scene (SURVEY.md §7 hard parts: no external oracle)."""
    sc = make_scene("blob", 40, 32, 24, 30.0, 33)
    Wc, Wa = loss_weights(sc.seed, 1, sc.height, sc.width)
    Wc, Wa = Wc.double(), Wa.double()
    names = ("means", "quats", "scales", "opacities", "colors")
    base = {k: getattr(sc, k).double() for k in names}
    # larger splats so every parameter matters
    base["scales"] = base["scales"] * 4.0

    def loss_of(p):
        rc, ra, _ = O.rasterization(p["means"], p["quats"], p["scales"], p["opacities"], p["colors"], sc.viewmats.double(),
                                    sc.Ks.double(), sc.width, sc.height, sh_degree=3, packed=False,
                                    backgrounds=torch.full((1, 3), 0.3, dtype=torch.float64))
        return (rc * Wc).sum() + (ra * Wa).sum()

    leaves = {k: v.clone().requires_grad_(True) for k, v in base.items()}
    loss_of(leaves).backward()
    g = torch.Generator().manual_seed(1)
    eps = 1e-6
    for k in names:
        for _ in range(6):
            idx = tuple(int(torch.randint(0, s, (1,), generator=g)) for s in base[k].shape)
            hi = {n: v.clone() for n, v in base.items()}
            lo = {n: v.clone() for n, v in base.items()}
            hi[k][idx] += eps
            lo[k][idx] -= eps
            with torch.no_grad():
                fd = float(loss_of(hi) - loss_of(lo)) / (2 * eps)
            an = float(leaves[k].grad[idx])
            assert abs(fd - an) <= 1e-4 * max(1.0, abs(an)) + 1e-6, (k, idx, fd, an)


def test_absgrad_is_sum_of_abs_pixel_gradients():
    """absgrad >= |grad| elementwise, equality for a Gaussian whose per-pixel gradients share a sign."""
    sc = make_scene("blob", 300, 64, 48, 60.0, 8)
    out = oracle_run(sc)
    m2 = out["meta"]["means2d"]
    ab = out["absgrad"]
    assert (ab >= 0).all()
    vis = out["meta"]["radii"] > 0
    assert float(ab[~vis].abs().sum()) == 0.0


def test_update_statistics_matches_reference_semantics():
    """oracle.update_statistics restates /root/reference/model/gaussian.py:188-197 verbatim for C=1."""
    g = torch.Generator().manual_seed(0)
    N, W, H = 1000, 640, 480
    radii = torch.randint(0, 30, (1, N), generator=g, dtype=torch.int32)
    absg = torch.rand(1, N, 2, generator=g)
    mr, acc, cnt = torch.rand(N, generator=g) * 0.01, torch.rand(N, generator=g), torch.ones(N)
    a, b, c = mr.clone(), acc.clone(), cnt.clone()
    # verbatim restatement of the reference lines
    max_hw = max(H, W)
    r = radii.detach()[0] / max_hw
    x = absg.detach()[0]
    visible = r > 0.0
    a[visible] = torch.max(a[visible], r[visible])
    grads = torch.norm(x, dim=-1) * max_hw
    b[visible] = b[visible] + grads[visible]
    c[visible] = c[visible] + 1
    O.update_statistics(mr, acc, cnt, radii, absg, W, H)
    assert torch.equal(mr, a) and torch.equal(acc, b) and torch.equal(cnt, c)


def test_tile_window_and_subset_reproduce_the_full_render():
    """The argument behind tests/test_gpu_fullsize.py::test_window_parity_with_the_oracle_at_full_size: blending only a
    window of tiles, from only the Gaussians that reach those tiles, gives the full render's pixels and — with loss
    weights that vanish outside the window — the full render's gradients."""
    import torch
    from easy_gaussian_splatting_b200.synthetic import loss_weights, make_scene
    from oracle import gsplat_oracle as O
    sc = make_scene("blob", 1500, 112, 96, 110.0, 17, n_views=2)
    C, N, W, H = 2, 1500, sc.width, sc.height
    names = ("means", "quats", "scales", "opacities", "colors")
    win = (2, 4, 3, 6)
    y0, y1, x0, x1 = win[0] * 16, win[1] * 16, win[2] * 16, win[3] * 16
    Wc, Wa = loss_weights(sc.seed, C, H, W)
    mask = torch.zeros(1, H, W, 1)
    mask[:, y0:y1, x0:x1] = 1.0
    Wc, Wa = Wc * mask, Wa * mask
    bg = sc.background[None].expand(C, 3).contiguous()
    full = {k: getattr(sc, k).clone().requires_grad_(True) for k in names}
    rc, ra, meta = O.rasterization(*[full[k] for k in names], sc.viewmats, sc.Ks, W, H, sh_degree=3, packed=False,
                                   absgrad=True, backgrounds=bg)
    ((rc * Wc).sum() + (ra * Wa).sum()).backward()
    offs = meta["isect_offsets"].reshape(-1).tolist() + [meta["flatten_ids"].numel()]
    tw, th = meta["tile_width"], meta["tile_height"]
    members = [meta["flatten_ids"][offs[(c * th + ty) * tw + tx]:offs[(c * th + ty) * tw + tx + 1]].long() % N
               for c in range(C) for ty in range(win[0], win[1]) for tx in range(win[2], win[3])]
    sub = torch.unique(torch.cat(members))
    assert 0 < sub.numel() < N
    part = {k: getattr(sc, k)[sub].clone().requires_grad_(True) for k in names}
    rc2, ra2, meta2 = O.rasterization(*[part[k] for k in names], sc.viewmats, sc.Ks, W, H, sh_degree=3, packed=False,
                                      absgrad=True, backgrounds=bg, tile_window=win)
    ((rc2 * Wc).sum() + (ra2 * Wa).sum()).backward()
    assert torch.equal(rc2[:, y0:y1, x0:x1], rc[:, y0:y1, x0:x1]) and torch.equal(ra2[:, y0:y1, x0:x1], ra[:, y0:y1, x0:x1])
    assert float(rc2[:, :y0].abs().sum()) == 0.0  # outside the window nothing is blended
    rest = torch.ones(N, dtype=torch.bool)
    rest[sub] = False
    for k in names:
        assert torch.allclose(full[k].grad[sub], part[k].grad, rtol=1e-5, atol=1e-7), k
        assert float(full[k].grad[rest].abs().sum()) == 0.0, k
    assert torch.allclose(meta["means2d"].absgrad[:, sub], meta2["means2d"].absgrad, rtol=1e-5, atol=1e-7)
