"""-m gpu: BASELINE.json's full-size configurations, checked through size-independent properties (the CPU oracle
cannot run these sizes in test time): sortedness / stability of the binning, conservation laws, image bounds,
determinism, linearity of the backward pass in the upstream gradient, permutation invariance, and agreement of
the two binning routes."""
import pytest
import torch

from easy_gaussian_splatting_b200.synthetic import loss_weights, make_config_scene

pytestmark = pytest.mark.gpu


def _render(sc, Wc=None, Wa=None, backward=False, perm=None):
    from easy_gaussian_splatting_b200 import rasterization
    names = ("means", "quats", "scales", "opacities", "colors")
    p = {}
    for k in names:
        t = getattr(sc, k)
        if perm is not None:
            t = t[perm]
        p[k] = t.cuda().contiguous().requires_grad_(backward)
    rc, ra, meta = rasterization(p["means"], p["quats"], p["scales"], p["opacities"], p["colors"], sc.viewmats.cuda(),
                                 sc.Ks.cuda(), sc.width, sc.height, sh_degree=3, packed=False, absgrad=True,
                                 backgrounds=sc.background[None].cuda())
    grads = None
    if backward:
        ((rc * Wc).sum() + (ra * Wa).sum()).backward()
        grads = {k: p[k].grad for k in names}
        grads["absgrad"] = meta["means2d"].absgrad
    return rc, ra, meta, grads


@pytest.mark.parametrize("name", ["metric", "cfg2", "cfg3"])
def test_full_size_invariants(name):
    from easy_gaussian_splatting_b200 import stages
    sc = make_config_scene(name)
    N, W, H = sc.means.shape[0], sc.width, sc.height
    Wc, Wa = (t.cuda() for t in loss_weights(sc.seed, 1, H, W))
    stages.reset_binning_hints()  # first call of the shape: exact sizes, plain backward
    rc, ra, meta, g = _render(sc, Wc, Wa, backward=True)
    radii, tpg, ids, flat, offs = meta["radii"], meta["tiles_per_gauss"], meta["isect_ids"], meta["flatten_ids"], meta["isect_offsets"]
    n = ids.numel()
    tw, th = meta["tile_width"], meta["tile_height"]
    nbits = stages.tile_n_bits(tw, th)
    # --- binning: conservation, sortedness, stability, offsets ---
    assert int(tpg.sum()) == n == flat.numel()
    assert bool(((tpg > 0) == (radii > 0)).all())
    assert bool((ids[1:] >= ids[:-1]).all()), "isect_ids sorted"
    same = ids[1:] == ids[:-1]
    assert bool((flat[1:][same] > flat[:-1][same]).all()), "ties keep ascending flat index (stable sort)"
    tile_of = (ids >> 32) & ((1 << nbits) - 1)
    assert int(tile_of.max()) < tw * th
    o = offs.reshape(-1).long()
    assert int(o[0]) == 0 and bool((o[1:] >= o[:-1]).all()) and int(o[-1]) <= n
    counts = torch.bincount(tile_of, minlength=tw * th)
    ends = torch.cat([o[1:], torch.tensor([n], device=o.device)])
    assert torch.equal(ends - o, counts), "offsets delimit exactly each tile's entries"
    depth_bits = ids & 0xFFFFFFFF
    assert torch.equal(depth_bits, meta["depths"].reshape(-1)[flat.long()].view(torch.int32).long() & 0xFFFFFFFF)
    hist = torch.bincount(flat.long(), minlength=N)
    assert torch.equal(hist, tpg.reshape(-1).long()), "every Gaussian appears once per tile it was counted for"
    # --- the classic 64-bit route gives the same lists ---
    _, ids2, flat2 = stages.isect_tiles(meta["means2d"].detach(), radii, meta["depths"], 16, tw, th, sort=True, tiles_per_gauss=tpg)
    assert torch.equal(ids2, ids) and torch.equal(flat2, flat)
    assert torch.equal(stages.isect_offset_encode(ids2, 1, tw, th), offs)
    # --- image bounds ---
    assert bool(torch.isfinite(rc).all()) and float(ra.min()) >= 0.0 and float(ra.max()) <= 1.0
    bgmax = float(sc.background.max())
    cmax = float(meta["colors"].max())
    assert bool((rc.amax(-1) <= ra[..., 0] * cmax + (1 - ra[..., 0]) * bgmax + 1e-4).all())
    # --- gradients: finite, zero for culled Gaussians, absgrad dominates |grad| of means2d ---
    vis = radii[0] > 0
    for k, v in g.items():
        assert bool(torch.isfinite(v).all()), k
    assert float(g["means"][~vis].abs().sum()) == 0.0 and float(g["colors"][~vis].abs().sum()) == 0.0
    assert float(g["absgrad"][0][~vis].abs().sum()) == 0.0 and bool((g["absgrad"] >= 0).all())
    # --- determinism of the forward pass, near-determinism of the backward (reduction order) ---
    # (the second call of a shape may replay outlier-long lists in segments — stages.segment_policy — which changes the
    #  summation order of the gradients more than a re-run does: 5e-5 on the object scene; north_star allows 1e-3)
    rc2, ra2, meta2, g2 = _render(sc, Wc, Wa, backward=True)
    assert torch.equal(rc, rc2) and torch.equal(ra, ra2) and torch.equal(meta2["flatten_ids"], flat)
    for k in g:
        assert float((g[k] - g2[k]).norm() / g[k].norm().clamp_min(1e-30)) <= 2e-4, k
    # --- linearity of the VJP in the upstream gradient: grad(2 Wc, 2 Wa) = 2 grad(Wc, Wa), same code path as g2 ---
    _, _, _, g3 = _render(sc, 2.0 * Wc, 2.0 * Wa, backward=True)
    for k in g:
        assert float((g3[k] - 2.0 * g2[k]).norm() / (2.0 * g2[k]).norm().clamp_min(1e-30)) <= 1e-5, k


def test_full_size_permutation_invariance():
    """Re-ordering the Gaussians must not change the image (ties in depth are measure-zero for random scenes)."""
    sc = make_config_scene("metric")
    rc, ra, meta, _ = _render(sc)
    perm = torch.randperm(sc.means.shape[0], generator=torch.Generator().manual_seed(0))
    rc2, ra2, meta2, _ = _render(sc, perm=perm)
    assert float((rc - rc2).abs().max()) <= 1e-4 and float((ra - ra2).abs().max()) <= 1e-4
    assert torch.equal(meta["radii"][0][perm.cuda()], meta2["radii"][0])


def test_cfg5_forward_only_4k():
    """BASELINE config 5: 6 M Gaussians at 3840x2160, no_grad forward."""
    from easy_gaussian_splatting_b200 import rasterization
    sc = make_config_scene("cfg5").to("cuda")
    with torch.no_grad():
        rc, ra, meta = rasterization(sc.means, sc.quats, sc.scales, sc.opacities, sc.colors, sc.viewmats, sc.Ks, sc.width,
                                     sc.height, sh_degree=3, packed=False, backgrounds=sc.background[None])
    assert rc.shape == (1, 2160, 3840, 3) and bool(torch.isfinite(rc).all())
    assert float(ra.min()) >= 0.0 and float(ra.max()) <= 1.0 and float(ra.mean()) > 0.05
    assert int(meta["tiles_per_gauss"].sum()) == meta["flatten_ids"].numel()


# ------------------------------------------------------------------------------------------------------------------
# Oracle parity AT the BASELINE configurations' real N and real resolution ("What's missing" #2).
# The binning is compared bit for bit at full size (the oracle's projection / intersection / sort are vectorised and
# finish in seconds even for millions of Gaussians); the blending — the part of the oracle that is slow — is compared
# on a window of tiles of the full image: the oracle blends only those tiles (``tile_window``), from the sub-set of
# Gaussians that reach them (a tile's list only ever contains Gaussians whose rectangle hits it, and a sub-set keeps
# their relative order, so the window's lists — hence its pixels — are exactly those of the full scene), and the loss
# weights are zero outside the window, so the window's pixels are also the only source of gradient on both sides.
# ------------------------------------------------------------------------------------------------------------------
WINDOW_CASES = [
    # name, views, backward, window (ty0, ty1, tx0, tx1)
    ("metric", 1, True, (30, 34, 56, 62)),
    ("cfg2", 1, True, (23, 26, 22, 26)),
    ("cfg3", 1, True, (15, 19, 28, 34)),
    ("cfg4", 2, True, (40, 43, 70, 75)),
    ("cfg5", 1, False, (66, 70, 118, 124)),
]


@pytest.mark.parametrize("name,n_views,backward,window", WINDOW_CASES, ids=[c[0] for c in WINDOW_CASES])
def test_window_parity_with_the_oracle_at_full_size(name, n_views, backward, window):
    from easy_gaussian_splatting_b200 import rasterization
    from oracle import gsplat_oracle as O
    from tests.util import PARAMS, image_report, rel_err
    sc = make_config_scene(name, n_views=n_views)
    N, W, H, C = sc.means.shape[0], sc.width, sc.height, n_views
    ty0, ty1, tx0, tx1 = window
    y0, y1, x0, x1 = ty0 * 16, min(ty1 * 16, H), tx0 * 16, min(tx1 * 16, W)
    Wc, Wa = loss_weights(sc.seed, C, H, W)
    mask = torch.zeros(1, H, W, 1)
    mask[:, y0:y1, x0:x1] = 1.0
    Wc, Wa = Wc * mask, Wa * mask
    bg = sc.background[None].expand(C, 3).contiguous()

    # ---- this repo, full size ----
    p = {k: getattr(sc, k).cuda().requires_grad_(backward) for k in PARAMS}
    with torch.set_grad_enabled(backward):
        rc, ra, meta = rasterization(p["means"], p["quats"], p["scales"], p["opacities"], p["colors"], sc.viewmats.cuda(),
                                     sc.Ks.cuda(), W, H, sh_degree=3, packed=False, absgrad=backward, backgrounds=bg.cuda())
    if backward:
        ((rc * Wc.cuda()).sum() + (ra * Wa.cuda()).sum()).backward()
    got = {k: meta[k].detach().cpu() for k in ("radii", "tiles_per_gauss", "isect_ids", "flatten_ids", "isect_offsets", "means2d", "depths")}
    rc_win, ra_win = rc.detach()[:, y0:y1, x0:x1].cpu(), ra.detach()[:, y0:y1, x0:x1].cpu()
    grads = {k: p[k].grad.cpu() for k in PARAMS} if backward else None
    absgrad = meta["means2d"].absgrad.cpu() if backward else None
    del rc, ra, meta, p
    torch.cuda.empty_cache()

    # ---- oracle: binning at full size, bit for bit ----
    with torch.no_grad():
        radii, means2d, depths, _ = O.fully_fused_projection(sc.means, sc.quats, sc.scales, sc.viewmats, sc.Ks, W, H)
        tw, th = -(-W // 16), -(-H // 16)
        tpg, ids, flat = O.isect_tiles(means2d, radii, depths, 16, tw, th)
        offs = O.isect_offset_encode(ids, C, tw, th)
    assert torch.equal(got["radii"], radii), "radii"
    assert torch.equal(got["means2d"], means2d) and torch.equal(got["depths"], depths)
    assert torch.equal(got["tiles_per_gauss"], tpg), "tiles_per_gauss"
    assert got["isect_ids"].numel() == ids.numel() and torch.equal(got["isect_ids"], ids), "sorted keys"
    assert torch.equal(got["flatten_ids"], flat), "flatten ids"
    assert torch.equal(got["isect_offsets"], offs), "tile offsets"
    print(f"{name}: N={N} C={C} n_isects={ids.numel()} bit-exact")

    # ---- oracle: blend the window from the Gaussians that reach it ----
    o = offs.reshape(-1).tolist() + [ids.numel()]
    members = []
    for c in range(C):
        for ty in range(ty0, ty1):
            for tx in range(tx0, tx1):
                t = (c * th + ty) * tw + tx
                members.append(flat[o[t]:o[t + 1]].long() % N)
    sub = torch.unique(torch.cat(members))  # sorted: relative order kept
    assert sub.numel() > 0
    leaves = {k: getattr(sc, k)[sub].clone().requires_grad_(backward) for k in PARAMS}
    counters = {}
    with torch.set_grad_enabled(backward):
        rc_o, ra_o, meta_o = O.rasterization(leaves["means"], leaves["quats"], leaves["scales"], leaves["opacities"],
                                             leaves["colors"], sc.viewmats, sc.Ks, W, H, sh_degree=3, packed=False,
                                             absgrad=backward, backgrounds=bg, counters=counters, tile_window=window)
    border = counters["borderline"][:, y0:y1, x0:x1]
    rep_c = image_report(rc_win, rc_o.detach()[:, y0:y1, x0:x1], border)
    rep_a = image_report(ra_win, ra_o.detach()[:, y0:y1, x0:x1], border)
    print(f"{name}: window {x1 - x0}x{y1 - y0} px, {sub.numel()} Gaussians, P_eval={counters['P_eval']}", rep_c, rep_a)
    assert rep_c["max_clean"] <= 1e-4 and rep_a["max_clean"] <= 1e-4  # north_star: 1e-4 absolute
    assert rep_c["max_border"] <= 2e-2
    if not backward:
        return
    ((rc_o * Wc).sum() + (ra_o * Wa).sum()).backward()
    errs = {k: rel_err(grads[k][sub], leaves[k].grad) for k in PARAMS}
    errs["absgrad"] = rel_err(absgrad[:, sub], meta_o["means2d"].absgrad)
    print(f"{name}: grad rel errs", errs)
    assert all(e <= 1e-3 for e in errs.values()), errs  # north_star: 1e-3 relative
    rest = torch.ones(N, dtype=torch.bool)
    rest[sub] = False
    for k in PARAMS:
        assert float(grads[k][rest].abs().sum()) == 0.0, f"{k}: gradient outside the window's Gaussians must be exactly 0"
    assert float(absgrad[:, rest].abs().sum()) == 0.0
