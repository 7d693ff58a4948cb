"""-m gpu: per-stage parity of the CUDA kernels (called through the C-ABI) against the CPU oracle.
Integer / index work must be bit-exact; floating point within the tolerance written at each check."""
import math

import pytest
import torch

from easy_gaussian_splatting_b200.synthetic import make_scene
from oracle import gsplat_oracle as O
from tests.util import image_report, rel_err

pytestmark = pytest.mark.gpu

SCENES = [
    dict(kind="blob", N=10_000, width=256, height=256, fx=274.5, seed=0, n_views=1),
    dict(kind="blob", N=3_000, width=131, height=77, fx=120.0, seed=7, n_views=3),      # ragged tiles, C>1
    dict(kind="outdoor", N=60_000, width=979, height=546, fx=581.0, seed=2, n_views=2),  # many culled
    dict(kind="object", N=20_000, width=400, height=400, fx=555.0, seed=1, n_views=1, white_background=True),
]


def _stages():
    from easy_gaussian_splatting_b200 import stages
    return stages


@pytest.mark.parametrize("cfg", SCENES)
@pytest.mark.parametrize("sh_degree", [0, 3])
def test_projection_bit_exact(cfg, sh_degree):
    st = _stages()
    sc = make_scene(**cfg)
    d = sc.to("cuda")
    out = st.projection_fwd(d.means, d.quats, d.scales, d.opacities, d.colors, d.viewmats, d.Ks, sc.width, sc.height, sh_degree)
    radii, m2, dep, con = O.fully_fused_projection(sc.means, sc.quats, sc.scales, sc.viewmats, sc.Ks, sc.width, sc.height)
    assert torch.equal(out["radii"].cpu(), radii), "radii must be bit-exact"
    assert torch.equal(out["means2d"].cpu(), m2), "means2d must be bit-exact"
    assert torch.equal(out["depths"].cpu(), dep), "depths must be bit-exact"
    assert torch.equal(out["conics"].cpu(), con), "conics must be bit-exact"
    tw, th = st.tile_grid(sc.width, sc.height)
    tpg, _, _ = O.isect_tiles(m2, radii, dep, 16, tw, th, sort=False)
    assert torch.equal(out["tiles_per_gauss"].cpu(), tpg), "tiles_per_gauss must be bit-exact"
    C = sc.viewmats.shape[0]
    campos = torch.inverse(sc.viewmats)[:, :3, 3]
    dirs = sc.means[None] - campos[:, None]
    cols = torch.clamp_min(O.spherical_harmonics(sh_degree, dirs, sc.colors[None].expand(C, -1, -1, -1), radii > 0) + 0.5, 0)
    cols = torch.where((radii > 0)[..., None], cols, torch.zeros(()))
    assert (out["colors"].cpu() - cols).abs().max().item() <= 2e-6  # fp32 summation-order tolerance
    # packed records agree with the separate tensors where visible
    vis = radii > 0
    sp = out["splats"].cpu()[vis]
    assert torch.equal(sp[:, 0:2], m2[vis]) and torch.equal(sp[:, 2:5], con[vis])
    assert torch.equal(sp[:, 5], sc.opacities[None].expand(C, -1)[vis])
    assert torch.equal(sp[:, 6:9], out["colors"].cpu()[vis])


@pytest.mark.parametrize("n", [0, 1, 5, 255, 256, 2047, 2048, 2049, 100_000, 1_234_567])
def test_exclusive_scan(n):
    st = _stages()
    g = torch.Generator().manual_seed(n)
    x = torch.randint(0, 50, (n,), generator=g, dtype=torch.int32)
    out, total = st.exclusive_scan(x.cuda())
    ref = torch.cumsum(x.long(), 0) - x.long()
    assert torch.equal(out.cpu(), ref)
    assert int(total.item()) == int(x.long().sum())


@pytest.mark.parametrize("n", [1, 31, 32, 33, 4095, 4096, 4097, 50_000, 1_000_003])
@pytest.mark.parametrize("end_bit,dist", [(45, "uniform"), (41, "few_tiles"), (64, "uniform"), (8, "uniform"), (47, "dups")])
def test_radix_sort_pairs(n, end_bit, dist):
    st = _stages()
    g = torch.Generator().manual_seed(n * 131 + end_bit)
    if dist == "uniform":
        hi = torch.randint(0, 2 ** 31 - 1, (n,), generator=g, dtype=torch.int64)
        lo = torch.randint(0, 2 ** 31 - 1, (n,), generator=g, dtype=torch.int64)
        keys = (hi << 32) | (lo << 1) | (hi & 1)
    elif dist == "few_tiles":  # realistic: few distinct high parts, float-like low parts
        tile = torch.randint(0, 300, (n,), generator=g, dtype=torch.int64)
        depth = (torch.rand(n, generator=g) * 10 + 0.5).view(torch.int32).long()
        keys = (tile << 32) | depth
    else:  # heavy duplication: stability decides the value order
        keys = torch.randint(0, 7, (n,), generator=g, dtype=torch.int64) << 33
    if end_bit < 64:
        keys = keys & ((1 << end_bit) - 1)
    else:
        keys = keys & 0x7FFFFFFFFFFFFFFF  # torch.sort is signed; keep the sign bit clear
    vals = torch.arange(n, dtype=torch.int32)
    k, v = st.radix_sort_pairs(keys.cuda().clone(), vals.cuda().clone(), end_bit)
    rk, order = torch.sort(keys, stable=True)
    assert torch.equal(k.cpu(), rk), "sorted keys must be bit-exact"
    assert torch.equal(v.cpu(), vals[order]), "values must follow a STABLE sort"


@pytest.mark.parametrize("cfg", SCENES)
def test_isect_tiles_sort_offsets_bit_exact(cfg):
    """Given the ORACLE's projection outputs, binning must be bit-exact unconditionally (SURVEY B-1 layer ii)."""
    st = _stages()
    sc = make_scene(**cfg)
    radii, m2, dep, con = O.fully_fused_projection(sc.means, sc.quats, sc.scales, sc.viewmats, sc.Ks, sc.width, sc.height)
    tw, th = st.tile_grid(sc.width, sc.height)
    C = sc.viewmats.shape[0]
    tpg_o, ids_u, flat_u = O.isect_tiles(m2, radii, dep, 16, tw, th, sort=False)
    tpg_o, ids_o, flat_o = O.isect_tiles(m2, radii, dep, 16, tw, th, sort=True)
    tpg, ids, flat = st.isect_tiles(m2.cuda(), radii.cuda(), dep.cuda(), 16, tw, th, sort=False)
    assert torch.equal(tpg.cpu(), tpg_o)
    assert torch.equal(ids.cpu(), ids_u) and torch.equal(flat.cpu(), flat_u), "unsorted emission order / keys"
    tpg, ids, flat = st.isect_tiles(m2.cuda(), radii.cuda(), dep.cuda(), 16, tw, th, sort=True)
    assert torch.equal(ids.cpu(), ids_o), "sorted isect_ids"
    assert torch.equal(flat.cpu(), flat_o), "sorted flatten_ids (stable tie order)"
    offs = st.isect_offset_encode(ids, C, tw, th)
    assert torch.equal(offs.cpu(), O.isect_offset_encode(ids_o, C, tw, th))


@pytest.mark.parametrize("cfg", SCENES)
def test_isect_fast_path_bit_exact(cfg):
    """Two-level route (depth sort of visible Gaussians, then stable tile sort) == 64-bit key sort, bit for bit."""
    st = _stages()
    sc = make_scene(**cfg)
    radii, m2, dep, con = O.fully_fused_projection(sc.means, sc.quats, sc.scales, sc.viewmats, sc.Ks, sc.width, sc.height)
    tw, th = st.tile_grid(sc.width, sc.height)
    C = sc.viewmats.shape[0]
    tpg_o, ids_o, flat_o = O.isect_tiles(m2, radii, dep, 16, tw, th, sort=True)
    ids, flat, offs = st.isect_sorted(m2.cuda(), radii.cuda(), dep.cuda(), tpg_o.cuda(), 16, tw, th)
    assert torch.equal(ids.cpu(), ids_o), "sorted isect_ids"
    assert torch.equal(flat.cpu(), flat_o), "sorted flatten_ids (stable tie order)"
    assert torch.equal(offs.cpu(), O.isect_offset_encode(ids_o, C, tw, th))


def test_isect_fast_path_depth_ties_and_empty():
    st = _stages()
    # many Gaussians at exactly the same depth and overlapping tiles: the tie order must be the flat index
    N, W, H = 3000, 96, 64
    g = torch.Generator().manual_seed(0)
    m2 = torch.rand(1, N, 2, generator=g) * torch.tensor([W, H])
    radii = torch.randint(0, 12, (1, N), generator=g, dtype=torch.int32)
    dep = torch.full((1, N), 2.5)
    dep[0, ::7] = 1.25
    tw, th = st.tile_grid(W, H)
    tpg_o, ids_o, flat_o = O.isect_tiles(m2, radii, dep, 16, tw, th, sort=True)
    ids, flat, offs = st.isect_sorted(m2.cuda(), radii.cuda(), dep.cuda(), tpg_o.cuda(), 16, tw, th)
    assert torch.equal(ids.cpu(), ids_o) and torch.equal(flat.cpu(), flat_o)
    assert torch.equal(offs.cpu(), O.isect_offset_encode(ids_o, 1, tw, th))
    z = torch.zeros(2, 10, dtype=torch.int32, device="cuda")
    ids, flat, offs = st.isect_sorted(torch.zeros(2, 10, 2, device="cuda"), z, torch.zeros(2, 10, device="cuda"), z, 16, 3, 2)
    assert ids.numel() == 0 and flat.numel() == 0 and int(offs.abs().sum()) == 0 and offs.shape == (2, 2, 3)


@pytest.mark.parametrize("n", [1, 33, 4096, 4097, 300_001])
@pytest.mark.parametrize("end_bit", [1, 13, 19, 32])
def test_radix_sort_pairs_u32(n, end_bit):
    st = _stages()
    g = torch.Generator().manual_seed(n + end_bit)
    keys = torch.randint(0, 2 ** 31 - 1, (n,), generator=g, dtype=torch.int64)
    if end_bit < 32:
        keys = keys & ((1 << end_bit) - 1)
    vals = torch.arange(n, dtype=torch.int32)
    k, v = st.radix_sort_pairs_u32(keys.to(torch.int32).cuda(), vals.cuda().clone(), end_bit)
    rk, order = torch.sort(keys, stable=True)
    assert torch.equal(k.cpu().long(), rk) and torch.equal(v.cpu(), vals[order])


def test_offsets_empty_and_sparse():
    st = _stages()
    empty = torch.zeros(0, dtype=torch.int64, device="cuda")
    offs = st.isect_offset_encode(empty, 2, 5, 3)
    assert offs.shape == (2, 3, 5) and int(offs.abs().sum()) == 0
    nb = st.tile_n_bits(5, 3)
    # entries only in camera 1, tiles 2 and 14 (last)
    keys = torch.tensor([((1 << nb) | 2) << 32, ((1 << nb) | 2) << 32 | 5, ((1 << nb) | 14) << 32], dtype=torch.int64)
    offs = st.isect_offset_encode(keys.cuda(), 2, 5, 3).cpu()
    assert torch.equal(offs, O.isect_offset_encode(keys, 2, 5, 3))


def _oracle_front(sc, sh_degree=3):
    radii, m2, dep, con = O.fully_fused_projection(sc.means, sc.quats, sc.scales, sc.viewmats, sc.Ks, sc.width, sc.height)
    C = sc.viewmats.shape[0]
    campos = torch.inverse(sc.viewmats)[:, :3, 3]
    dirs = sc.means[None] - campos[:, None]
    cols = torch.clamp_min(O.spherical_harmonics(sh_degree, dirs, sc.colors[None].expand(C, -1, -1, -1), radii > 0) + 0.5, 0)
    tw, th = math.ceil(sc.width / 16), math.ceil(sc.height / 16)
    _, ids, flat = O.isect_tiles(m2, radii, dep, 16, tw, th)
    offs = O.isect_offset_encode(ids, C, tw, th)
    opac = sc.opacities[None].expand(C, -1).contiguous()
    return radii, m2, dep, con, cols, opac, ids, flat, offs


@pytest.mark.parametrize("cfg", SCENES)
@pytest.mark.parametrize("with_bg", [True, False])
def test_rasterize_fwd_bwd_vs_oracle(cfg, with_bg):
    """Blend kernels fed with the oracle's own intermediates: colours/alphas <= 1e-4 abs on non-borderline
    pixels (borderline = threshold-flip candidates, counted and bounded separately, SURVEY B-2);
    per-Gaussian gradients <= 1e-3 norm-wise relative."""
    st = _stages()
    sc = make_scene(**cfg)
    radii, m2, dep, con, cols, opac, ids, flat, offs = _oracle_front(sc)
    C = sc.viewmats.shape[0]
    bg = sc.background[None].expand(C, 3).contiguous() if with_bg else None
    leaves = [t.clone().requires_grad_(True) for t in (m2, con, cols, opac)]
    counters = {}
    rc, ra, last = O.rasterize_to_pixels(*leaves, sc.width, sc.height, 16, offs, flat, backgrounds=bg, absgrad=True, counters=counters)
    g = torch.Generator().manual_seed(5)
    Wc = torch.rand(rc.shape, generator=g)
    Wa = torch.rand(ra.shape, generator=g)
    ((rc * Wc).sum() + (ra * Wa).sum()).backward()

    splats = st.pack_splats(m2.cuda(), con.cuda(), cols.cuda(), opac.cuda(), dep.cuda())
    bgc = None if bg is None else bg.cuda()
    colors, alphas, last_ids, pc = st.rasterize_fwd(splats, offs.cuda(), flat.cuda(), bgc, sc.width, sc.height, count_pairs=True)
    colors2, alphas2, last2 = st.rasterize_fwd(splats, offs.cuda(), flat.cuda(), bgc, sc.width, sc.height)
    assert torch.equal(colors, colors2) and torch.equal(alphas, alphas2) and torch.equal(last_ids, last2)
    border = counters["borderline"]
    rep_c = image_report(colors.cpu(), rc.detach(), border)
    rep_a = image_report(alphas.cpu(), ra.detach(), border)
    print("fwd colours", rep_c, "alphas", rep_a, "pairs", pc.tolist(), counters["P_eval"], counters["P_acc"])
    assert rep_c["max_clean"] <= 1e-4 and rep_a["max_clean"] <= 1e-4
    assert rep_c["max_border"] <= 2e-2 and rep_c["n_border"] <= max(20, 0.002 * border.numel())
    same_last = (last_ids.cpu() == last) | border
    assert same_last.all(), "last_ids must match on non-borderline pixels"
    # pair counters agree with the oracle up to threshold flips
    assert abs(pc[0].item() - counters["P_eval"]) <= 1e-3 * counters["P_eval"] + 64
    assert abs(pc[1].item() - counters["P_acc"]) <= 1e-3 * counters["P_acc"] + 64

    v = st.rasterize_bwd(splats, offs.cuda(), flat.cuda(), bgc, sc.width, sc.height, alphas, last_ids, Wc.cuda(), Wa.cuda()).cpu()
    tol = 1e-3
    errs = dict(xy=rel_err(v[..., 0:2], leaves[0].grad), conic=rel_err(v[..., 2:5], leaves[1].grad),
                opac=rel_err(v[..., 5], leaves[3].grad), rgb=rel_err(v[..., 6:9], leaves[2].grad),
                absxy=rel_err(v[..., 9:11], m2_abs(leaves[0])))
    print("bwd rel errs", errs)
    assert all(e <= tol for e in errs.values()), errs
    assert float(v[..., 11].abs().max()) == 0.0


def m2_abs(leaf):
    return leaf.absgrad


def test_densify_stats_update():
    st = _stages()
    g = torch.Generator().manual_seed(3)
    C, N, W, H = 3, 5000, 640, 360
    radii = torch.randint(-1, 40, (C, N), generator=g, dtype=torch.int32).clamp_min(0)
    absg = torch.rand(C, N, 2, generator=g)
    stats = [torch.rand(N, generator=g) for _ in range(3)]
    ref = [s.clone() for s in stats]
    O.update_statistics(ref[0], ref[1], ref[2], radii, absg, W, H)
    dev = [s.cuda() for s in stats]
    st.densify_stats_update(dev[0], dev[1], dev[2], radii.cuda(), absg.cuda(), W, H)
    assert torch.equal(dev[0].cpu(), ref[0]), "max_radii"
    assert torch.allclose(dev[1].cpu(), ref[1], rtol=1e-6, atol=1e-6), "grad_norm_accum"
    assert torch.equal(dev[2].cpu(), ref[2]), "collecting_counts"


def test_fused_adam_matches_torch_adam():
    """§8f-3: same trajectory as torch.optim.Adam (the reference's optimizer, gaussian.py:389-412) on its 6 groups."""
    from easy_gaussian_splatting_b200.optim import FusedAdam
    g = torch.Generator().manual_seed(0)
    N = 4099
    shapes = dict(means=(N, 3), log_scales=(N, 3), quats=(N, 4), sh_0=(N, 1, 3), sh_rest=(N, 15, 3), logit_opacities=(N,))
    lrs = dict(means=1e-3, log_scales=1e-2, quats=1e-3, sh_0=2.5e-3, sh_rest=1.25e-4, logit_opacities=5e-2)
    init = {k: torch.randn(*s, generator=g) for k, s in shapes.items()}
    ref_p = {k: v.clone().requires_grad_(True) for k, v in init.items()}
    our_p = {k: v.clone().cuda().requires_grad_(True) for k, v in init.items()}
    # one parameter that is NOT 16-byte aligned (a view one float into a buffer): the kernel's scalar path; the odd N
    # leaves every other group with a ragged last unit of its 128-bit path
    off_by_one = torch.empty(N + 1, device="cuda")
    off_by_one[1:].copy_(init["logit_opacities"])
    our_p["logit_opacities"] = off_by_one[1:].detach().requires_grad_(True)
    assert our_p["logit_opacities"].data_ptr() % 16 == 4
    ref = torch.optim.Adam([{"params": [ref_p[k]], "lr": lrs[k], "name": k} for k in shapes], eps=1e-15)
    ours = FusedAdam([{"params": [our_p[k]], "lr": lrs[k], "name": k} for k in shapes], eps=1e-15)
    for it in range(5):
        for k in shapes:
            gr = torch.randn(*shapes[k], generator=g) * (10.0 ** (it - 2))
            ref_p[k].grad = gr.clone()
            our_p[k].grad = None if (it == 2 and k == "quats") else gr.cuda()  # a skipped parameter, like on densify steps
            if it == 2 and k == "quats":
                ref_p[k].grad = None
        ref.step()
        ours.step()
    for k in shapes:
        assert torch.allclose(our_p[k].detach().cpu(), ref_p[k].detach(), rtol=2e-5, atol=1e-7), k
        assert ours.state[our_p[k]]["step"] == int(ref.state[ref_p[k]]["step"])


def _classic_lists(proj, tw, th):
    stages = _stages()
    """Reference lists from the classic 64-bit route (emit + 6-pass sort + offset encode)."""
    C = proj["radii"].shape[0]
    _, ids, flat = stages.isect_tiles(proj["means2d"], proj["radii"], proj["depths"], 16, tw, th, sort=True,
                                      tiles_per_gauss=proj["tiles_per_gauss"])
    return ids, flat, stages.isect_offset_encode(ids, C, tw, th)


def test_async_binning_guess_too_small_too_large_and_empty():
    """The sync-free route sizes its buffers from the previous call of the same shape.  A guess that is too small must
    be detected and repaired, one that is far too large must still give exact-length, bit-identical lists, and a
    view that sees nothing must give all-zero offsets — each against the classic route."""
    stages = _stages()
    W, H = 400, 304
    tw, th = stages.tile_grid(W, H)
    small = make_scene("blob", 3_000, W, H, 380.0, 5).to("cuda")
    big = make_scene("blob", 60_000, W, H, 380.0, 6).to("cuda")
    stages.reset_binning_hints()
    seen = []
    for name, sc in (("small", small), ("big", big), ("big", big), ("small", small), ("none", small), ("small", small)):
        vm = sc.viewmats.clone()
        if name == "none":
            vm[:, 2, 3] -= 100.0  # everything behind the camera
        proj = stages.projection_fwd(sc.means, sc.quats, sc.scales, sc.opacities, sc.colors, vm, sc.Ks, W, H, 3)
        b = stages.isect_sorted_async(proj["means2d"], proj["radii"], proj["depths"], proj["tiles_per_gauss"], 16, tw, th)
        ok = b.resolve()
        seen.append((name, b.exact, ok, b.capacity, b.n_isects))
        if not ok:
            b = stages.isect_sorted_async(proj["means2d"], proj["radii"], proj["depths"], proj["tiles_per_gauss"], 16, tw, th,
                                          capacity=b.n_isects)
            assert b.resolve() and b.exact
        b.note_for_next_call()
        ids, flat, offs = _classic_lists(proj, tw, th)
        assert b.n_isects == flat.numel() and b.n_vis == int((proj["radii"] > 0).sum())
        assert torch.equal(b.flatten_ids, flat) and torch.equal(b.offsets, offs) and torch.equal(b.isect_ids(), ids), name
        assert int(b.offsets_store[-1]) == b.n_isects, "sentinel behind the offsets"
    print(seen)
    assert seen[0][1] is True, "first call of a shape has no guess: exact"
    assert seen[1][2] is False, "3 k -> 60 k Gaussians: the guess must have been too small"
    assert seen[2][1] is False and seen[2][2] is True, "same scene again: guessed, and the guess held"
    assert seen[3][2] is True and seen[3][3] > 4 * max(seen[3][4], 1), "a far too large buffer is fine"
    assert seen[4][4] == 0 and seen[5][2] is True
    torch.cuda.synchronize()
    hint = stages.binning_hint(1, tw, th, "cuda", tight=False)
    assert hint["n_isects"] == seen[5][4] and hint["max_tile_len"] == int(torch.diff(torch.cat([offs.reshape(-1), offs.new_tensor([flat.numel()])])).max())


def test_sentinel_mode_equals_exact_length_mode():
    """egs_rasterize_* with n_isects = -capacity (length read behind the offsets) == with the exact length."""
    stages = _stages()
    sc = make_scene(**SCENES[3]).to("cuda")
    W, H = sc.width, sc.height
    tw, th = stages.tile_grid(W, H)
    proj = stages.projection_fwd(sc.means, sc.quats, sc.scales, sc.opacities, sc.colors, sc.viewmats, sc.Ks, W, H, 3)
    stages.reset_binning_hints()
    b = stages.isect_sorted_async(proj["means2d"], proj["radii"], proj["depths"], proj["tiles_per_gauss"], 16, tw, th,
                                  capacity=None)
    b.resolve()
    b.note_for_next_call()
    b2 = stages.isect_sorted_async(proj["means2d"], proj["radii"], proj["depths"], proj["tiles_per_gauss"], 16, tw, th)
    assert b2.resolve() and not b2.exact and b2.capacity > b2.n_isects
    bg = sc.background[None]
    ref = stages.rasterize_fwd(proj["splats"], b.offsets, b.flatten_ids, bg, W, H)
    out = stages.rasterize_fwd(proj["splats"], b2.offsets, b2.flat_cap, bg, W, H, n_isects=b2.raster_n)
    for a, r in zip(out, ref):
        assert torch.equal(a, r)
    g = torch.Generator(device="cuda").manual_seed(0)
    vc, va = torch.rand(1, H, W, 3, device="cuda", generator=g), torch.rand(1, H, W, 1, device="cuda", generator=g)
    v_ref = stages.rasterize_bwd(proj["splats"], b.offsets, b.flatten_ids, bg, W, H, ref[1], ref[2], vc, va)
    v_out = stages.rasterize_bwd(proj["splats"], b2.offsets, b2.flat_cap, bg, W, H, out[1], out[2], vc, va, n_isects=b2.raster_n)
    assert float((v_out - v_ref).norm() / v_ref.norm()) <= 1e-5
    for seg, seg_min in ((128, 0), (128, 600), (512, 100000)):
        for order in (None, b2.tile_order):
            rc, ra, last, ck = stages.rasterize_fwd_checkpointed(proj["splats"], b2.offsets, b2.flat_cap, bg, W, H, seg,
                                                                 seg_min_len=seg_min, n_isects=b2.raster_n, tile_order=order)
            assert torch.equal(rc, ref[0]) and torch.equal(last, ref[2])
            v_seg = stages.rasterize_bwd_segmented(proj["splats"], b2.offsets, b2.flat_cap, bg, W, H, rc, ra, last, vc, va, ck, seg,
                                                   seg_min_len=seg_min, n_isects=b2.raster_n, tile_order=order)
            assert float((v_seg - v_ref).norm() / v_ref.norm()) <= 2e-5, (seg, seg_min)
    # the launch order egs_isect_sorted emits: a permutation of the tiles, longest lists first by length class (highest
    # set bit of the length); rendering in that order changes nothing
    order = b2.tile_order.long()
    n_tiles = tw * th
    assert torch.equal(torch.sort(order).values, torch.arange(n_tiles, device="cuda"))
    lens = torch.diff(torch.cat([b2.offsets.reshape(-1), b2.offsets_store[-1:]])).long()
    cls = ((lens[:, None] >> torch.arange(32, device="cuda")[None, :]) > 0).sum(1)  # bit length: 0 for an empty tile
    assert bool((cls[order][1:] <= cls[order][:-1]).all()), "length classes in descending order"
    out_o = stages.rasterize_fwd(proj["splats"], b2.offsets, b2.flat_cap, bg, W, H, n_isects=b2.raster_n, tile_order=b2.tile_order)
    for a, r in zip(out_o, ref):
        assert torch.equal(a, r)
    v_o = stages.rasterize_bwd(proj["splats"], b2.offsets, b2.flat_cap, bg, W, H, out_o[1], out_o[2], vc, va, n_isects=b2.raster_n,
                               tile_order=b2.tile_order)
    assert float((v_o - v_ref).norm() / v_ref.norm()) <= 1e-5


@pytest.mark.parametrize("cfg", SCENES)
def test_tight_lists_hold_a_subset_and_render_identically(cfg):
    """The blend kernels' own lists (isect_sorted_async(splats=...)): every Gaussian only in the tiles that hold a
    pixel it can reach.  They must be a sub-list of gsplat's (same order), and forward / backward through them must
    give the same pixels bit for bit and the same gradients."""
    stages = _stages()
    sc = make_scene(**cfg).to("cuda")
    W, H, C = sc.width, sc.height, sc.viewmats.shape[0]
    tw, th = stages.tile_grid(W, H)
    proj = stages.projection_fwd(sc.means, sc.quats, sc.scales, sc.opacities, sc.colors, sc.viewmats, sc.Ks, W, H, 3)
    args = (proj["means2d"], proj["radii"], proj["depths"], proj["tiles_per_gauss"], 16, tw, th)
    stages.reset_binning_hints()
    cl = stages.isect_sorted_async(*args)
    tg = stages.isect_sorted_async(*args, tight_rects=proj["tight_rects"])
    assert cl.resolve() and tg.resolve()
    n_t = C * tw * th
    key = lambda b: (torch.repeat_interleave(torch.arange(n_t, device="cuda"),
                                             torch.diff(torch.cat([b.offsets.reshape(-1), b.offsets_store[-1:]])).long()) * (C * sc.means.shape[0])
                     + b.flatten_ids.long())
    kc, kt = key(cl), key(tg)
    print(cfg["kind"], "classic", cl.n_isects, "tight", tg.n_isects, round(tg.n_isects / max(cl.n_isects, 1), 3))
    assert tg.n_isects < cl.n_isects
    pos = torch.searchsorted(torch.sort(kc).values, kt)
    assert bool((torch.sort(kc).values[pos.clamp_max(kc.numel() - 1)] == kt).all()), "every tight entry is a classic entry"
    # same relative order inside every tile: the tight list is the classic list with entries removed
    keep = torch.isin(kc, kt)
    assert torch.equal(kc[keep], kt)
    bg = sc.background[None].expand(C, 3).contiguous()
    ref = stages.rasterize_fwd(proj["splats"], cl.offsets, cl.flatten_ids, bg, W, H)
    out = stages.rasterize_fwd(proj["splats"], tg.offsets, tg.flat_cap, bg, W, H, n_isects=tg.raster_n, tile_order=tg.tile_order)
    assert torch.equal(out[0], ref[0]) and torch.equal(out[1], ref[1])
    hit = ref[1][..., 0] > 0
    assert torch.equal(tg.flat_cap[out[2][hit].long()], cl.flatten_ids[ref[2][hit].long()]), "same last blended Gaussian per pixel"
    g = torch.Generator(device="cuda").manual_seed(0)
    vc, va = torch.rand(C, H, W, 3, device="cuda", generator=g), torch.rand(C, H, W, 1, device="cuda", generator=g)
    v_ref = stages.rasterize_bwd(proj["splats"], cl.offsets, cl.flatten_ids, bg, W, H, ref[1], ref[2], vc, va)
    v_out = stages.rasterize_bwd(proj["splats"], tg.offsets, tg.flat_cap, bg, W, H, out[1], out[2], vc, va, n_isects=tg.raster_n,
                                 tile_order=tg.tile_order)
    assert float((v_out - v_ref).norm() / v_ref.norm()) <= 1e-5


@pytest.mark.parametrize("C,N,tw,th", [(2, 40_000, 37, 23), (1, 1500, 120, 68), (3, 5000, 1, 1), (1, 70_000, 5, 300),
                                        (2, 3000, 300, 200)])  # the last one: 120 000 tile slots = three level-2 passes
def test_scan_emit_on_arbitrary_packed_rectangles(C, N, tw, th):
    """The single-launch scan + emission (decoupled look-back over 1024-Gaussian blocks, table / redux owner search)
    on rectangles no projection would produce together: empty ones between full-grid ones, single columns and rows,
    visible Gaussians without a tile, depth ties — against a list built with torch sorts from the same rectangles."""
    st = _stages()
    g = torch.Generator().manual_seed(C * 1000 + N)
    n_tiles = tw * th
    x0 = torch.randint(0, tw, (C, N), generator=g)
    y0 = torch.randint(0, th, (C, N), generator=g)
    kind = torch.rand(C, N, generator=g)
    w = torch.minimum(torch.randint(1, 5, (C, N), generator=g), tw - x0)
    h = torch.minimum(torch.randint(1, 5, (C, N), generator=g), th - y0)
    full = kind < 0.03
    x0, y0 = torch.where(full, 0, x0), torch.where(full, 0, y0)
    w, h = torch.where(full, tw, w), torch.where(full, th, h)
    w = torch.where((kind > 0.90) & (kind <= 0.95), 1, w)      # single columns
    h = torch.where((kind > 0.95), 1, h)                       # single rows
    visible = torch.rand(C, N, generator=g) < 0.8
    empty = (torch.rand(C, N, generator=g) < 0.3) | ~visible  # visible, but reaches no pixel
    w, h = torch.where(empty, 0, w), torch.where(empty, 0, h)
    rects = torch.stack([torch.where(empty, 0, x0 | (y0 << 16)), torch.where(empty, 0, w | (h << 16))], -1).int().cuda()
    radii = visible.int().cuda()
    tpg = torch.where(visible, torch.clamp(w * h, min=1), 0).int().cuda()
    depths = (torch.randint(1, 50, (C, N), generator=g).float() * 0.25).cuda()  # many ties
    m2 = torch.zeros(C, N, 2, device="cuda")
    st.reset_binning_hints()
    b = st.isect_sorted_async(m2, radii, depths, tpg, 16, tw, th, tight_rects=rects)
    if not b.resolve():
        b = st.isect_sorted_async(m2, radii, depths, tpg, 16, tw, th, capacity=b.n_isects, tight_rects=rects)
        assert b.resolve()
    # reference: every (Gaussian, tile) entry, ordered by (camera, tile), then depth, then flat index
    cnt = (w * h).reshape(-1).cuda()
    flat = torch.repeat_interleave(torch.arange(C * N, device="cuda"), cnt)
    j = torch.arange(flat.numel(), device="cuda") - torch.repeat_interleave(torch.cumsum(cnt, 0) - cnt, cnt)
    ww = w.reshape(-1).cuda()[flat].clamp(min=1)
    key = (flat // N) * n_tiles + (y0.reshape(-1).cuda()[flat] + j // ww) * tw + x0.reshape(-1).cuda()[flat] + j % ww
    order = torch.sort(depths.reshape(-1)[flat], stable=True).indices
    order = order[torch.sort(key[order], stable=True).indices]
    assert b.n_isects == flat.numel()
    assert torch.equal(b.flatten_ids.long(), flat[order])
    counts = torch.bincount(key, minlength=C * n_tiles)
    offsets = torch.cumsum(counts, 0) - counts
    assert torch.equal(b.offsets.reshape(-1).long(), offsets)
    assert int(b.offsets_store[-1]) == flat.numel(), "sentinel behind the offsets"


def test_capacity_between_the_emitted_count_and_the_bound():
    """Buffers are checked against the BOUND (gsplat's count, known after the first kernel) first; when the bound does
    not fit, the count the tight route really emitted decides: room for it is enough, one entry less is not."""
    st = _stages()
    sc = make_scene(**SCENES[2]).to("cuda")
    W, H, C = sc.width, sc.height, sc.viewmats.shape[0]
    tw, th = st.tile_grid(W, H)
    proj = st.projection_fwd(sc.means, sc.quats, sc.scales, sc.opacities, sc.colors, sc.viewmats, sc.Ks, W, H, 3)
    args = (proj["means2d"], proj["radii"], proj["depths"], proj["tiles_per_gauss"], 16, tw, th)
    st.reset_binning_hints()
    ref = st.isect_sorted_async(*args, tight_rects=proj["tight_rects"])
    assert ref.resolve() and ref.exact and ref.n_bound == int(proj["tiles_per_gauss"].sum())
    n_t = ref.n_isects
    assert 0 < n_t < ref.n_bound
    mid = st.isect_sorted_async(*args, capacity=(n_t + ref.n_bound) // 2, tight_rects=proj["tight_rects"])
    assert mid.resolve() and mid.n_isects == n_t and mid.n_bound == ref.n_bound
    assert torch.equal(mid.flatten_ids, ref.flatten_ids) and torch.equal(mid.offsets, ref.offsets)
    assert int(mid.offsets_store[-1]) == n_t
    exact = st.isect_sorted_async(*args, capacity=n_t, tight_rects=proj["tight_rects"])
    assert exact.resolve() and torch.equal(exact.flatten_ids, ref.flatten_ids)
    short = st.isect_sorted_async(*args, capacity=n_t - 1, tight_rects=proj["tight_rects"])
    assert not short.resolve() and short.n_isects == n_t
    # the hint a call leaves carries the bound (what the next call sizes its buffers from) and, once it has arrived,
    # the emitted count
    ref.note_for_next_call()
    torch.cuda.synchronize()
    hint = st.binning_hint(C, tw, th, "cuda")
    assert hint["n_bound"] == ref.n_bound and hint["n_isects"] == n_t
    nxt = st.isect_sorted_async(*args, tight_rects=proj["tight_rects"])
    assert not nxt.exact and nxt.capacity >= ref.n_bound and nxt.resolve()
    assert torch.equal(nxt.flatten_ids, ref.flatten_ids)
