"""TEST INFRASTRUCTURE: the per-stage operators of ``easy_gaussian_splatting_b200.stages`` emulated on CPU with the
oracle (``oracle/gsplat_oracle.py``), so that the HOST side of the boundary call — ``rendering.rasterization`` with
its autograd node, meta contract and ``.absgrad`` tagging — can be driven end to end without a GPU, e.g. by the
reference's own ``GaussianModel`` (tests/test_reference_swap.py).

Same signatures, shapes, dtypes and packed record layouts as the CUDA operators (include/egs_raster.h):
  splat record    {x, y, conic_a, conic_b | conic_c, opacity, r, g | b, depth, 0, sigma_cut}
  gradient record {v_x, v_y, v_ca, v_cb | v_cc, v_opacity, v_r, v_g | v_b, |v_x|, |v_y|, 0}
Never imported by the product; installed by ``install(monkeypatch)`` only.
"""
from __future__ import annotations

import math

import torch

from easy_gaussian_splatting_b200 import rendering, stages
from oracle import gsplat_oracle as O


def _colors(means, colors, viewmats, radii, sh_degree):
    C, N = radii.shape
    if sh_degree is None:
        return colors[None].expand(C, N, 3) if colors.dim() == 2 else colors
    campos = torch.inverse(viewmats)[:, :3, 3]
    dirs = means[None, :, :] - campos[:, None, :]
    cols = O.spherical_harmonics(sh_degree, dirs, colors[None].expand(C, *colors.shape), masks=radii > 0)
    return torch.clamp_min(cols + 0.5, 0.0)


def projection_fwd(means, quats, scales, opacities, colors, viewmats, Ks, width, height, sh_degree, eps2d=0.3,
                   near_plane=0.01, far_plane=1e10, radius_clip=0.0, tile_size=16, antialiased=False):
    assert not antialiased, "the CPU emulation covers the reference's mode (classic)"
    with torch.no_grad():
        radii, means2d, depths, conics = O.fully_fused_projection(means, quats, scales, viewmats, Ks, width, height, eps2d,
                                                                  near_plane, far_plane, radius_clip)
        cols = _colors(means, colors, viewmats, radii, sh_degree)
        cols = torch.where((radii > 0)[..., None], cols, torch.zeros(()))
        C, N = radii.shape
        tw, th = stages.tile_grid(width, height, tile_size)
        tpg, _, _ = O.isect_tiles(means2d, radii, depths, tile_size, tw, th, sort=False)
        op = opacities[None].expand(C, N)
        cut = torch.log(255.0 * op)
        z = torch.zeros(C, N)
        splats = torch.stack([means2d[..., 0], means2d[..., 1], conics[..., 0], conics[..., 1], conics[..., 2], op,
                              cols[..., 0], cols[..., 1], cols[..., 2], depths, z, cut], -1)
        splats = torch.where((radii > 0)[..., None], splats, torch.zeros(()))
    return {"radii": radii, "means2d": means2d.contiguous(), "depths": depths.contiguous(), "conics": conics.contiguous(),
            "colors": cols.contiguous(), "tiles_per_gauss": tpg, "tight_rects": None, "splats": splats.contiguous()}


def isect_tiles(means2d, radii, depths, tile_size, tile_width, tile_height, sort=True, tiles_per_gauss=None, n_isects=None):
    return O.isect_tiles(means2d, radii, depths, tile_size, tile_width, tile_height, sort=sort)


def isect_offset_encode(isect_ids, C, tile_width, tile_height):
    return O.isect_offset_encode(isect_ids, C, tile_width, tile_height)


def isect_sorted_async(means2d, radii, depths, tiles_per_gauss, tile_size, tile_width, tile_height, capacity=None, tight_rects=None):
    C = radii.shape[0]
    _, ids, flat = O.isect_tiles(means2d, radii, depths, tile_size, tile_width, tile_height, sort=True)
    offsets = O.isect_offset_encode(ids, C, tile_width, tile_height)
    return stages.ResolvedIsects(ids, flat, offsets)


def _unpack(splats):
    return splats[..., 0:2], splats[..., 2:5], splats[..., 6:9], splats[..., 5]


def rasterize_fwd(splats, isect_offsets, flatten_ids, backgrounds, width, height, count_pairs=False, n_isects=None, tile_order=None):
    with torch.no_grad():
        m2, cn, cl, op = _unpack(splats)
        rc, ra, last = O.rasterize_to_pixels(m2, cn, cl, op, width, height, 16, isect_offsets, flatten_ids,
                                             backgrounds=backgrounds)
    return rc, ra, last


def rasterize_bwd(splats, isect_offsets, flatten_ids, backgrounds, width, height, render_alphas, last_ids,
                  v_render_colors, v_render_alphas, n_isects=None, tile_order=None):
    with torch.enable_grad():
        leaves = [t.detach().clone().requires_grad_(True) for t in _unpack(splats)]
        rc, ra, _ = O.rasterize_to_pixels(*leaves, width, height, 16, isect_offsets, flatten_ids,
                                          backgrounds=None if backgrounds is None else backgrounds.detach(), absgrad=True)
        ((rc * v_render_colors).sum() + (ra * v_render_alphas).sum()).backward()
    m2, cn, cl, op = leaves
    absg = m2.absgrad
    zero = lambda t: torch.zeros_like(splats[..., 0]) if t is None else t
    g_m2, g_cn, g_cl, g_op = (t.grad for t in leaves)
    v = torch.zeros_like(splats)
    if g_m2 is not None:
        v[..., 0:2] = g_m2
        v[..., 2:5] = g_cn
        v[..., 5] = zero(g_op)
        v[..., 6:9] = g_cl
        v[..., 9:11] = absg
    return v


def projection_bwd(means, quats, scales, colors, viewmats, Ks, width, height, sh_degree, eps2d, radii, colors_rgb,
                   v_splats, v_means2d_extra=None, want_absgrad=False, antialiased_opacities=None, opacities=None):
    assert antialiased_opacities is None
    with torch.enable_grad():
        lm, lq, ls, lc = (t.detach().clone().requires_grad_(True) for t in (means, quats, scales, colors))
        r2, means2d, depths, conics = O.fully_fused_projection(lm, lq, ls, viewmats.detach(), Ks.detach(), width, height, eps2d)
        assert torch.equal(r2, radii)
        cols = _colors(lm, lc, viewmats.detach(), radii, sh_degree)
        v_xy = v_splats[..., 0:2] if v_means2d_extra is None else v_splats[..., 0:2] + v_means2d_extra
        total = (means2d * v_xy).sum() + (conics * v_splats[..., 2:5]).sum() + (cols * v_splats[..., 6:9]).sum()
        total.backward()
    outs = []
    for src, leaf in ((means, lm), (quats, lq), (scales, ls)):
        buf = stages._grad_buffer(src)
        buf.copy_(leaf.grad if leaf.grad is not None else torch.zeros_like(src))
        outs.append(buf)
    v_op = stages._grad_buffer(opacities) if opacities is not None else torch.empty(means.shape[0])
    v_op.copy_(v_splats[..., 5].sum(0))
    v_col = stages._grad_buffer(colors)
    v_col.copy_(lc.grad if lc.grad is not None else torch.zeros_like(colors))
    outs += [v_op, v_col]
    if want_absgrad:
        outs.append(v_splats[..., 9:11].clone())
    return tuple(outs)


def densify_stats_update(max_radii, grad_norm_accum, collecting_counts, radii, absgrad, width, height):
    O.update_statistics(max_radii, grad_norm_accum, collecting_counts, radii, absgrad, width, height)


def install(monkeypatch) -> None:
    """Route ``stages`` to the emulation and let ``rendering`` accept CPU tensors (its own check refuses them:
    the product has no CPU path)."""
    for name in ("projection_fwd", "isect_sorted_async", "isect_tiles", "isect_offset_encode", "rasterize_fwd", "rasterize_bwd",
                 "projection_bwd", "densify_stats_update"):
        monkeypatch.setattr(stages, name, globals()[name])
    monkeypatch.setattr(stages, "binning_hint", lambda *a, **k: None)
    monkeypatch.setattr(rendering, "_check_inputs", lambda *a, **k: None)
    monkeypatch.delenv("EGS_BWD_SEGMENT", raising=False)
