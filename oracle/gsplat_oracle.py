"""CPU oracle for the differentiable Gaussian rasterizer hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``easy_gaussian_splatting_b200/`` may
import this module; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, and there only
as the checker / the CPU comparator.

PARITY UNPINNED.  The reference (li199603/easy_gaussian_splatting) contains no
rasterizer code: the path is the third-party call
``gsplat.rendering.rasterization`` at ``/root/reference/model/gaussian.py:353-367``
(dependency ``gsplat>=1.0.0``, ``/root/reference/requirements.txt:1``; README
recommends ``==1.0.0``, ``/root/reference/README.md:17``).  gsplat is neither
vendored in the reference nor installed / installable in this image and the
reference ships no tests, golden vectors or fixtures.  This file is therefore a
pure-PyTorch restatement of the *published gsplat 1.0.0 algorithm* (SURVEY.md
Appendix A), anchored on the reference's call contract
(``model/gaussian.py:351-374``) and on how it consumes the outputs
(``model/gaussian.py:188-197``).  "Parity with the reference" everywhere in this
repo means parity with this frozen restatement.

Design rules
------------
* Everything runs on CPU, any float dtype (fp32 for parity, fp64 for
  finite-difference / analytic gradient checks).
* The projection path uses ONLY explicitly ordered scalar-style elementwise
  ops (no einsum/bmm/norm/sum), so its fp32 results are a well defined sequence
  of IEEE-754 round-to-nearest operations that the CUDA kernels reproduce bit
  for bit (compiled with -fmad=false).  See ``CANONICAL OP ORDER`` comments.
* Gradients come from torch autograd (mathematically equal to gsplat's
  hand-written backward, A-7/A-8), absgrad from an explicit per-pair
  intermediate (A-9).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

ALPHA_MIN = 1.0 / 255.0
ALPHA_MAX = 0.999
T_MIN = 1e-4
BORDER_REL = 2e-5  # relative width of the 'borderline' band around the two thresholds

# --------------------------------------------------------------------------------------
# A-1  projection (follows gsplat 1.0.0 fully_fused_projection_fwd)
# --------------------------------------------------------------------------------------


def _sqrt_rn(x: Tensor) -> Tensor:
    """Correctly rounded square root.  torch's CPU fp32 sqrt goes through a vector-math
    library and is off by one ulp for ~0.6 % of inputs (measured in this image), which would
    make the oracle's "canonical fp32 op order" ill-defined.  sqrt evaluated in fp64 and rounded
    once to fp32 IS the IEEE-754 correctly rounded fp32 result (53 >= 2*24 + 2 bits), i.e. what
    CUDA's sqrtf / __fsqrt_rn and the host's sqrtf return."""
    if x.dtype == torch.float32:
        return torch.sqrt(x.double()).float()
    return torch.sqrt(x)


def quat_to_rotmat_rows(quats: Tensor):
    """wxyz quaternion (unnormalised) -> 9 rotation entries, row major.

    Same formula as the reference's ``normalized_quat_to_rotmat``
    (/root/reference/model/utils.py:31-49).  CANONICAL OP ORDER:
    s = ((w*w + x*x) + y*y) + z*z ; inv = 1/sqrt(s) ; q *= inv.
    """
    w, x, y, z = quats.unbind(-1)
    s = ((w * w + x * x) + y * y) + z * z
    inv = 1.0 / _sqrt_rn(s)
    w, x, y, z = w * inv, x * inv, y * inv, z * inv
    x2, y2, z2 = x * x, y * y, z * z
    xy, xz, yz = x * y, x * z, y * z
    wx, wy, wz = w * x, w * y, w * z
    R00 = 1.0 - 2.0 * (y2 + z2)
    R01 = 2.0 * (xy - wz)
    R02 = 2.0 * (xz + wy)
    R10 = 2.0 * (xy + wz)
    R11 = 1.0 - 2.0 * (x2 + z2)
    R12 = 2.0 * (yz - wx)
    R20 = 2.0 * (xz - wy)
    R21 = 2.0 * (yz + wx)
    R22 = 1.0 - 2.0 * (x2 + y2)
    return (R00, R01, R02, R10, R11, R12, R20, R21, R22)


def quat_scale_to_covar(quats: Tensor, scales: Tensor):
    """Sigma = (R diag(s)) (R diag(s))^T, returned as the 6 unique entries
    (c00, c01, c02, c11, c12, c22).  CANONICAL OP ORDER:
    M_ij = R_ij * s_j ; c_ij = (M_i0*M_j0 + M_i1*M_j1) + M_i2*M_j2."""
    R = quat_to_rotmat_rows(quats)
    s0, s1, s2 = scales.unbind(-1)
    M = [
        [R[0] * s0, R[1] * s1, R[2] * s2],
        [R[3] * s0, R[4] * s1, R[5] * s2],
        [R[6] * s0, R[7] * s1, R[8] * s2],
    ]

    def dot(i, j):
        return (M[i][0] * M[j][0] + M[i][1] * M[j][1]) + M[i][2] * M[j][2]

    return dot(0, 0), dot(0, 1), dot(0, 2), dot(1, 1), dot(1, 2), dot(2, 2)


def _project_elementwise(px, py, pz, cov6, V, fx, fy, cx, cy, width, height, eps2d):
    """Shared elementwise body. px.. are broadcastable tensors, V is a dict of the 12
    view-matrix entries (r00..r22, t0..t2) broadcastable against them."""
    c00, c01, c02, c11, c12, c22 = cov6
    # world -> camera.  CANONICAL: ((r0*px + r1*py) + r2*pz) + t
    x = ((V["r00"] * px + V["r01"] * py) + V["r02"] * pz) + V["t0"]
    y = ((V["r10"] * px + V["r11"] * py) + V["r12"] * pz) + V["t1"]
    z = ((V["r20"] * px + V["r21"] * py) + V["r22"] * pz) + V["t2"]
    # covariance world -> camera: A = W*Sigma, Sc = A*W^T.
    S = [[c00, c01, c02], [c01, c11, c12], [c02, c12, c22]]
    Wm = [[V["r00"], V["r01"], V["r02"]], [V["r10"], V["r11"], V["r12"]], [V["r20"], V["r21"], V["r22"]]]
    A = [[(Wm[i][0] * S[0][j] + Wm[i][1] * S[1][j]) + Wm[i][2] * S[2][j] for j in range(3)] for i in range(3)]

    def sc(i, j):
        return (A[i][0] * Wm[j][0] + A[i][1] * Wm[j][1]) + A[i][2] * Wm[j][2]

    k00, k01, k02, k11, k12, k22 = sc(0, 0), sc(0, 1), sc(0, 2), sc(1, 1), sc(1, 2), sc(2, 2)
    # perspective (gsplat 1.0.0 persp_proj): tan_fov = 0.5*W/fx, lim = 1.3*tan_fov
    tan_fovx = (0.5 * width) / fx
    tan_fovy = (0.5 * height) / fy
    lim_x = 1.3 * tan_fovx
    lim_y = 1.3 * tan_fovy
    rz = 1.0 / z
    rz2 = rz * rz
    tx = z * torch.minimum(lim_x, torch.maximum(-lim_x, x * rz))
    ty = z * torch.minimum(lim_y, torch.maximum(-lim_y, y * rz))
    J00 = fx * rz
    J02 = -((fx * tx) * rz2)
    J11 = fy * rz
    J12 = -((fy * ty) * rz2)
    # B = J * Sc (2x3), cov2d = B * J^T
    B00 = J00 * k00 + J02 * k02
    B01 = J00 * k01 + J02 * k12
    B02 = J00 * k02 + J02 * k22
    B11 = J11 * k11 + J12 * k12
    B12 = J11 * k12 + J12 * k22
    a = B00 * J00 + B02 * J02
    b = B01 * J11 + B02 * J12
    d = B11 * J11 + B12 * J12
    m2x = (fx * x) * rz + cx
    m2y = (fy * y) * rz + cy
    # add_blur (det_orig feeds the antialiased-mode compensation factor)
    det_orig = a * d - b * b
    a = a + eps2d
    d = d + eps2d
    det = a * d - b * b
    return x, y, z, a, b, d, det, m2x, m2y, det_orig


class _CompensationSqrt(torch.autograd.Function):
    """compensation = sqrt(max(0, r)), r = det_orig / det_blur, with gsplat 1.0.0's add_blur_vjp guard:
    d comp / d r = 0.5 / (comp + 1e-6)  (rasterize_mode="antialiased")."""

    @staticmethod
    def forward(ctx, r):
        comp = _sqrt_rn(torch.clamp_min(r, 0.0))
        ctx.save_for_backward(comp)
        return comp

    @staticmethod
    def backward(ctx, v):
        (comp,) = ctx.saved_tensors
        return v * 0.5 / (comp + 1e-6)


def _view_entries(viewmats: Tensor, Ks: Tensor, shape_suffix=(1,)):
    """[C,4,4],[C,3,3] -> dict of [C,1] tensors."""
    V = {}
    for i in range(3):
        for j in range(3):
            V[f"r{i}{j}"] = viewmats[:, i, j].reshape(-1, *shape_suffix)
        V[f"t{i}"] = viewmats[:, i, 3].reshape(-1, *shape_suffix)
    fx = Ks[:, 0, 0].reshape(-1, *shape_suffix)
    fy = Ks[:, 1, 1].reshape(-1, *shape_suffix)
    cx = Ks[:, 0, 2].reshape(-1, *shape_suffix)
    cy = Ks[:, 1, 2].reshape(-1, *shape_suffix)
    return V, fx, fy, cx, cy


def fully_fused_projection(
    means: Tensor,  # [N,3]
    quats: Tensor,  # [N,4] wxyz, unnormalised
    scales: Tensor,  # [N,3]
    viewmats: Tensor,  # [C,4,4] world->camera
    Ks: Tensor,  # [C,3,3]
    width: int,
    height: int,
    eps2d: float = 0.3,
    near_plane: float = 0.01,
    far_plane: float = 1e10,
    radius_clip: float = 0.0,
    calc_compensations: bool = False,
):
    """Returns radii[C,N] int32, means2d[C,N,2], depths[C,N], conics[C,N,3]
    (+ compensations[C,N] = sqrt(max(0, det(cov2d) / det(cov2d + eps2d I))) when calc_compensations).

    Culled entries (radii == 0) have all float outputs exactly 0 (upstream leaves
    them uninitialised).  Differentiable w.r.t. means, quats,
    scales (and viewmats) through autograd; the cull decisions are taken from a
    no-grad elementwise pass whose fp32 values the CUDA kernel reproduces exactly.
    """
    C, N = viewmats.shape[0], means.shape[0]
    dt = means.dtype
    V, fx, fy, cx, cy = _view_entries(viewmats, Ks)
    wdt = torch.tensor(float(width), dtype=dt)
    hdt = torch.tensor(float(height), dtype=dt)

    def body(mn, qt, sc, Vd, fx_, fy_, cx_, cy_):
        cov6 = quat_scale_to_covar(qt, sc)
        px, py, pz = mn.unbind(-1)
        return _project_elementwise(px, py, pz, cov6, Vd, fx_, fy_, cx_, cy_, wdt, hdt, eps2d)

    with torch.no_grad():
        x, y, z, a, b, d, det, m2x, m2y, det_orig = body(means.detach(), quats.detach(), scales.detach(),
                                              {k: v.detach() for k, v in V.items()},
                                              fx.detach(), fy.detach(), cx.detach(), cy.detach())
        valid = (z >= near_plane) & (z <= far_plane)  # cull if z < near or z > far
        valid &= det > 0
        mid = 0.5 * (a + d)
        v1 = mid + _sqrt_rn(torch.clamp_min(mid * mid - det, 0.01))
        radius = torch.ceil(3.0 * _sqrt_rn(v1))
        valid &= radius > radius_clip
        valid &= ~((m2x + radius <= 0) | (m2x - radius >= width) | (m2y + radius <= 0) | (m2y - radius >= height))
        valid &= torch.isfinite(radius) & torch.isfinite(m2x) & torch.isfinite(m2y)
        # NaN anywhere -> comparisons above are False in the right direction except the
        # negated off-screen test; the isfinite guard makes NaN/inf a cull (viewer fx=inf/0 case).
        radius = torch.where(valid, radius, torch.zeros_like(radius))
        # clamp to int32 range before the cast
        radii = radius.clamp(max=2147483520.0).to(torch.int32)
        radii = radii.expand(C, N).contiguous()
        valid = valid.expand(C, N)

    need_grad = torch.is_grad_enabled() and any(t.requires_grad for t in (means, quats, scales, viewmats, Ks))
    means2d = torch.zeros(C, N, 2, dtype=dt)
    depths = torch.zeros(C, N, dtype=dt)
    conics = torch.zeros(C, N, 3, dtype=dt)
    if not need_grad:
        inv_det = 1.0 / det
        ca, cb, cc = d * inv_det, -(b * inv_det), a * inv_det
        zero = torch.zeros((), dtype=dt)
        means2d = torch.stack([torch.where(valid, m2x, zero), torch.where(valid, m2y, zero)], -1)
        depths = torch.where(valid, z, zero).expand(C, N).contiguous()
        conics = torch.stack([torch.where(valid, ca, zero), torch.where(valid, cb, zero), torch.where(valid, cc, zero)], -1)
        if calc_compensations:
            comp = torch.where(valid, _sqrt_rn(torch.clamp_min(det_orig / det, 0.0)), zero).expand(C, N).contiguous()
            return radii, means2d.contiguous(), depths, conics.contiguous(), comp
        return radii, means2d.contiguous(), depths, conics.contiguous()

    # differentiable recomputation on the valid subset only (keeps NaNs of culled
    # entries out of the backward pass)
    ci, ni = valid.nonzero(as_tuple=True)
    Vs = {k: v[ci, 0] for k, v in V.items()}
    x, y, z, a, b, d, det, m2x, m2y, det_orig = body(means[ni], quats[ni], scales[ni], Vs, fx[ci, 0], fy[ci, 0], cx[ci, 0], cy[ci, 0])
    inv_det = 1.0 / det
    ca, cb, cc = d * inv_det, -(b * inv_det), a * inv_det
    means2d = means2d.index_put((ci, ni), torch.stack([m2x, m2y], -1))
    depths = depths.index_put((ci, ni), z)
    conics = conics.index_put((ci, ni), torch.stack([ca, cb, cc], -1))
    if calc_compensations:
        comp = torch.zeros(C, N, dtype=dt).index_put((ci, ni), _CompensationSqrt.apply(det_orig / det))
        return radii, means2d, depths, conics, comp
    return radii, means2d, depths, conics


# --------------------------------------------------------------------------------------
# A-2  spherical harmonics (gsplat 1.0.0 sh_coeffs_to_color_fast; standard real SH basis,
#       C0 is the constant at /root/reference/model/utils.py:15)
# --------------------------------------------------------------------------------------

SH_C0 = 0.2820947917738781
SH_C1 = 0.48860251190292


def sh_basis(degree: int, dirs: Tensor) -> Tensor:
    """dirs [...,3] (unnormalised) -> basis [..., (degree+1)^2]."""
    out = [torch.full(dirs.shape[:-1], SH_C0, dtype=dirs.dtype)]
    if degree >= 1:
        inorm = 1.0 / torch.sqrt((dirs * dirs).sum(-1))
        x, y, z = (dirs * inorm[..., None]).unbind(-1)
        out += [-SH_C1 * y, SH_C1 * z, -SH_C1 * x]
    if degree >= 2:
        z2 = z * z
        fTmp0B = -1.092548430592079 * z
        fC1 = x * x - y * y
        fS1 = 2.0 * x * y
        out += [0.5462742152960395 * fS1, fTmp0B * y, 0.9461746957575601 * z2 - 0.3153915652525201,
                fTmp0B * x, 0.5462742152960395 * fC1]
    if degree >= 3:
        fTmp0C = -2.285228997322329 * z2 + 0.4570457994644658
        fTmp1B = 1.445305721320277 * z
        fC2 = x * fC1 - y * fS1
        fS2 = x * fS1 + y * fC1
        out += [-0.5900435899266435 * fS2, fTmp1B * fS1, fTmp0C * y,
                z * (1.865881662950577 * z2 - 1.119528997770346), fTmp0C * x, fTmp1B * fC1,
                -0.5900435899266435 * fC2]
    if degree >= 4:
        raise NotImplementedError("oracle restates degrees 0..3 (the reference uses sh_degree <= 3)")
    return torch.stack(out, -1)


def spherical_harmonics(degree: int, dirs: Tensor, coeffs: Tensor, masks: Optional[Tensor] = None) -> Tensor:
    """dirs [...,3], coeffs [...,K,3] -> colors [...,3]; masked-out entries are 0."""
    nb = (degree + 1) ** 2
    basis = sh_basis(degree, dirs)  # [..., nb]
    col = (basis[..., None] * coeffs[..., :nb, :]).sum(-2)
    if masks is not None:
        col = torch.where(masks[..., None], col, torch.zeros((), dtype=col.dtype))
    return col


# --------------------------------------------------------------------------------------
# A-3..A-5  tile intersection, sort, offsets (integer / bit work; must be bit exact)
# --------------------------------------------------------------------------------------


def tile_n_bits(tile_width: int, tile_height: int) -> int:
    return int(math.floor(math.log2(tile_width * tile_height))) + 1


def isect_tiles(means2d: Tensor, radii: Tensor, depths: Tensor, tile_size: int, tile_width: int,
                tile_height: int, sort: bool = True):
    """-> tiles_per_gauss[C,N] int32, isect_ids[n] int64, flatten_ids[n] int32.

    key = cam << (32 + tile_n_bits) | tile << 32 | int32_bits(depth); value = c*N + n.
    With sort=True the pairs are stably sorted ascending by key (A-4)."""
    C, N = radii.shape
    m = means2d.to(torch.float32)
    r = radii.to(torch.float32)
    tsz = float(tile_size)
    tx, ty, tr = m[..., 0] / tsz, m[..., 1] / tsz, r / tsz
    vis = radii > 0

    def clampi(v, hi):
        v = torch.nan_to_num(v, nan=0.0, posinf=float(hi), neginf=0.0)
        return v.clamp(0, hi).to(torch.int64)

    xmin = clampi(torch.floor(tx - tr), tile_width)
    ymin = clampi(torch.floor(ty - tr), tile_height)
    xmax = clampi(torch.ceil(tx + tr), tile_width)
    ymax = clampi(torch.ceil(ty + tr), tile_height)
    cnt = torch.where(vis, (xmax - xmin) * (ymax - ymin), torch.zeros((), dtype=torch.int64))
    tiles_per_gauss = cnt.to(torch.int32)
    flat_cnt = cnt.reshape(-1)
    total = int(flat_cnt.sum())
    nbits = tile_n_bits(tile_width, tile_height)
    if total == 0:
        return tiles_per_gauss, torch.zeros(0, dtype=torch.int64), torch.zeros(0, dtype=torch.int32)
    owner = torch.repeat_interleave(torch.arange(C * N), flat_cnt)  # flat index of each isect
    start = torch.cumsum(flat_cnt, 0) - flat_cnt
    j = torch.arange(total) - start[owner]
    w = (xmax - xmin).reshape(-1)[owner]
    tyy = ymin.reshape(-1)[owner] + j // w
    txx = xmin.reshape(-1)[owner] + j % w
    tile_id = tyy * tile_width + txx
    cam = owner // N
    depth_bits = depths.to(torch.float32).contiguous().view(torch.int32).reshape(-1)[owner].to(torch.int64)
    keys = (cam << (32 + nbits)) | (tile_id << 32) | (depth_bits & 0xFFFFFFFF)
    vals = owner.to(torch.int32)
    if sort:
        keys, order = torch.sort(keys, stable=True)
        vals = vals[order]
    return tiles_per_gauss, keys, vals


def isect_offset_encode(isect_ids: Tensor, C: int, tile_width: int, tile_height: int) -> Tensor:
    """offsets[c,ty,tx] = first sorted index whose (cam,tile) >= that tile (A-5)."""
    n_tiles = tile_width * tile_height
    nbits = tile_n_bits(tile_width, tile_height)
    hi = isect_ids >> 32
    cam = hi >> nbits
    tile = hi & ((1 << nbits) - 1)
    lin = cam * n_tiles + tile
    offs = torch.searchsorted(lin.contiguous(), torch.arange(C * n_tiles), right=False)
    return offs.to(torch.int32).reshape(C, tile_height, tile_width)


# --------------------------------------------------------------------------------------
# A-6 / A-7 / A-9  alpha blending, forward and (autograd) backward, per tile
# --------------------------------------------------------------------------------------


def _blend_tile(xy, conic, opac, rgb, px, py, bg, want_delta=False):
    """xy [G,2], conic [G,3], opac [G], rgb [G,3]; px,py [P].  Sequential front-to-back
    semantics of A-6 expressed with cumprod (a sequential product on CPU)."""
    dx = xy[None, :, 0] - px[:, None]
    dy = xy[None, :, 1] - py[:, None]
    if want_delta:
        dx = dx.clone().requires_grad_(True) if not dx.requires_grad else dx
        dy = dy.clone().requires_grad_(True) if not dy.requires_grad else dy
        dx.retain_grad()
        dy.retain_grad()
    sigma = 0.5 * (conic[None, :, 0] * dx * dx + conic[None, :, 2] * dy * dy) + conic[None, :, 1] * dx * dy
    alpha = torch.clamp_max(opac[None, :] * torch.exp(-sigma), ALPHA_MAX)
    acc = (sigma >= 0) & (alpha >= ALPHA_MIN)
    one = torch.ones((), dtype=xy.dtype)
    fac = torch.where(acc, 1.0 - alpha, one)
    with torch.no_grad():
        Tincl = torch.cumprod(fac, 1)
        stop = acc & (Tincl <= T_MIN)
        nstop = torch.cumsum(stop.to(torch.int32), 1)
        alive = nstop == 0  # strictly before the terminating Gaussian
        evaluated = (nstop - stop.to(torch.int32)) == 0  # up to and including it
        use = acc & alive
        # pairs whose accept / terminate decision sits on a threshold: two correct
        # exp implementations may legitimately decide them differently
        near = evaluated & ((alpha * 255.0 - 1.0).abs() < BORDER_REL)
        near |= evaluated & acc & ((Tincl - T_MIN).abs() < BORDER_REL * T_MIN)
        borderline = near.any(1)
    fac2 = torch.where(use, 1.0 - alpha, one)
    Tincl2 = torch.cumprod(fac2, 1)
    Texcl = torch.cat([torch.ones(Tincl2.shape[0], 1, dtype=xy.dtype), Tincl2[:, :-1]], 1)
    wgt = torch.where(use, alpha * Texcl, torch.zeros((), dtype=xy.dtype))
    Tfin = Tincl2[:, -1] if Tincl2.shape[1] > 0 else torch.ones(px.shape[0], dtype=xy.dtype)
    col = wgt @ rgb if rgb.shape[0] > 0 else torch.zeros(px.shape[0], 3, dtype=xy.dtype)
    if bg is not None:
        col = col + Tfin[:, None] * bg[None, :]
    G = xy.shape[0]
    idx = torch.arange(G)
    last = torch.where(use, idx[None, :], torch.full((), -1, dtype=torch.int64)).amax(1) if G > 0 else \
        torch.full((px.shape[0],), -1, dtype=torch.int64)
    return col, 1.0 - Tfin, last, (dx, dy), int(evaluated.sum()), int(use.sum()), borderline


def _tile_pixels(ty, tx, tile_size, width, height, dtype):
    ys = torch.arange(ty * tile_size, min((ty + 1) * tile_size, height))
    xs = torch.arange(tx * tile_size, min((tx + 1) * tile_size, width))
    yy, xx = torch.meshgrid(ys, xs, indexing="ij")
    return yy.reshape(-1), xx.reshape(-1), xx.reshape(-1).to(dtype) + 0.5, yy.reshape(-1).to(dtype) + 0.5


class _Rasterize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means2d, conics, colors, opacities, backgrounds, width, height, tile_size,
                isect_offsets, flatten_ids, absgrad, counters, tile_window=None):
        C, N = means2d.shape[:2]
        dt = means2d.dtype
        th, tw = isect_offsets.shape[1:]
        n_isects = flatten_ids.shape[0]
        out_c = torch.zeros(C, height, width, 3, dtype=dt)
        out_a = torch.zeros(C, height, width, 1, dtype=dt)
        last_ids = torch.zeros(C, height, width, dtype=torch.int32)
        offs = isect_offsets.reshape(-1).tolist() + [n_isects]
        m2 = means2d.reshape(C * N, 2)
        cn = conics.reshape(C * N, 3)
        cl = colors.reshape(C * N, 3)
        op = opacities.reshape(C * N)
        p_eval = p_acc = 0
        border = torch.zeros(C, height, width, dtype=torch.bool)
        ty0, ty1, tx0, tx1 = (0, th, 0, tw) if tile_window is None else tile_window
        for c in range(C):
            bg = backgrounds[c] if backgrounds is not None else None
            for ty in range(ty0, ty1):
                for tx in range(tx0, tx1):
                    t = (c * th + ty) * tw + tx
                    s, e = offs[t], offs[t + 1]
                    yy, xx, px, py = _tile_pixels(ty, tx, tile_size, width, height, dt)
                    g = flatten_ids[s:e].long()
                    col, alp, last, _, ne, na, bl = _blend_tile(m2[g], cn[g], op[g], cl[g], px, py, bg)
                    border[c, yy, xx] = bl
                    p_eval += ne
                    p_acc += na
                    out_c[c, yy, xx] = col
                    out_a[c, yy, xx, 0] = alp
                    last_ids[c, yy, xx] = torch.where(last >= 0, last + s, torch.zeros((), dtype=torch.int64)).to(torch.int32)
        if counters is not None:
            counters["P_eval"] = counters.get("P_eval", 0) + p_eval
            counters["P_acc"] = counters.get("P_acc", 0) + p_acc
            counters["borderline"] = border
        ctx.save_for_backward(means2d, conics, colors, opacities, backgrounds, isect_offsets, flatten_ids)
        ctx.meta = (width, height, tile_size, absgrad, (ty0, ty1, tx0, tx1))
        ctx.mark_non_differentiable(last_ids)
        return out_c, out_a, last_ids

    @staticmethod
    def backward(ctx, v_c, v_a, _v_last):
        means2d, conics, colors, opacities, backgrounds, isect_offsets, flatten_ids = ctx.saved_tensors
        width, height, tile_size, absgrad, (ty0, ty1, tx0, tx1) = ctx.meta
        C, N = means2d.shape[:2]
        dt = means2d.dtype
        th, tw = isect_offsets.shape[1:]
        n_isects = flatten_ids.shape[0]
        offs = isect_offsets.reshape(-1).tolist() + [n_isects]
        m2 = means2d.detach().reshape(C * N, 2)
        cn = conics.detach().reshape(C * N, 3)
        cl = colors.detach().reshape(C * N, 3)
        op = opacities.detach().reshape(C * N)
        g_m2 = torch.zeros_like(m2)
        g_abs = torch.zeros_like(m2)
        g_cn = torch.zeros_like(cn)
        g_cl = torch.zeros_like(cl)
        g_op = torch.zeros_like(op)
        g_bg = torch.zeros_like(backgrounds) if backgrounds is not None else None
        for c in range(C):
            for ty in range(ty0, ty1):
                for tx in range(tx0, tx1):
                    t = (c * th + ty) * tw + tx
                    s, e = offs[t], offs[t + 1]
                    yy, xx, px, py = _tile_pixels(ty, tx, tile_size, width, height, dt)
                    vc, va = v_c[c, yy, xx], v_a[c, yy, xx, 0]
                    if g_bg is not None and e == s:
                        g_bg[c] += vc.sum(0)  # T = 1 everywhere
                        continue
                    if e == s:
                        continue
                    g = flatten_ids[s:e].long()
                    with torch.enable_grad():
                        lx = m2[g].clone().requires_grad_(True)
                        lc = cn[g].clone().requires_grad_(True)
                        lr = cl[g].clone().requires_grad_(True)
                        lo = op[g].clone().requires_grad_(True)
                        lb = backgrounds[c].detach().clone().requires_grad_(True) if backgrounds is not None else None
                        col, alp, _, (dx, dy), _, _, _ = _blend_tile(lx, lc, lo, lr, px, py, lb, want_delta=True)
                        loss = (col * vc).sum() + (alp * va).sum()
                        loss.backward()
                    g_m2.index_add_(0, g, lx.grad)
                    g_cn.index_add_(0, g, lc.grad)
                    g_cl.index_add_(0, g, lr.grad)
                    g_op.index_add_(0, g, lo.grad)
                    if g_bg is not None:
                        g_bg[c] += lb.grad
                    if absgrad:
                        ab = torch.stack([dx.grad.abs().sum(0), dy.grad.abs().sum(0)], -1)
                        g_abs.index_add_(0, g, ab)
        if absgrad:
            means2d.absgrad = g_abs.reshape(C, N, 2)
        return (g_m2.reshape(C, N, 2), g_cn.reshape(C, N, 3), g_cl.reshape(C, N, 3), g_op.reshape(C, N),
                g_bg, None, None, None, None, None, None, None, None)


def rasterize_to_pixels(means2d, conics, colors, opacities, width, height, tile_size, isect_offsets,
                        flatten_ids, backgrounds=None, absgrad=False, counters=None, tile_window=None):
    """means2d[C,N,2] conics[C,N,3] colors[C,N,3] opacities[C,N] -> colors[C,H,W,3], alphas[C,H,W,1], last_ids.
    tile_window = (ty0, ty1, tx0, tx1): blend only the tiles of that window (pixels outside stay 0 and receive no
    gradient) — bounds the CPU work of full-size parity checks; the arithmetic per tile is unchanged."""
    return _Rasterize.apply(means2d, conics, colors, opacities, backgrounds, width, height, tile_size,
                            isect_offsets, flatten_ids, absgrad, counters, tile_window)


# --------------------------------------------------------------------------------------
# the boundary: gsplat.rendering.rasterization (signature per SURVEY.md section 8b;
# reference call site /root/reference/model/gaussian.py:353-367)
# --------------------------------------------------------------------------------------


def rasterization(
    means: Tensor, quats: Tensor, scales: Tensor, opacities: Tensor, colors: Tensor,
    viewmats: Tensor, Ks: Tensor, width: int, height: int,
    near_plane: float = 0.01, far_plane: float = 1e10, radius_clip: float = 0.0, eps2d: float = 0.3,
    sh_degree: Optional[int] = None, packed: bool = True, tile_size: int = 16,
    backgrounds: Optional[Tensor] = None, render_mode: str = "RGB", sparse_grad: bool = False,
    absgrad: bool = False, rasterize_mode: str = "classic", channel_chunk: int = 32,
    counters: Optional[dict] = None, tile_window: Optional[Tuple[int, int, int, int]] = None,
) -> Tuple[Tensor, Tensor, Dict]:
    """Oracle for the whole path.  ``packed`` only changes upstream's memory layout, not
    results, so the oracle computes the dense (packed=False) form for either value."""
    if render_mode != "RGB" or rasterize_mode not in ("classic", "antialiased") or sparse_grad:
        raise NotImplementedError("oracle covers render_mode='RGB', rasterize_mode='classic'/'antialiased', sparse_grad=False")
    N, C = means.shape[0], viewmats.shape[0]
    proj = fully_fused_projection(means, quats, scales, viewmats, Ks, width, height, eps2d, near_plane, far_plane,
                                  radius_clip, calc_compensations=rasterize_mode == "antialiased")
    radii, means2d, depths, conics = proj[:4]
    opac = opacities[None, :].expand(C, N)
    if rasterize_mode == "antialiased":
        opac = opac * proj[4]  # gsplat 1.0.0 rendering.py: opacities = opacities * compensations
    if sh_degree is None:
        cols = colors[None].expand(C, N, colors.shape[-1]) if colors.dim() == 2 else colors
    else:
        campos = torch.inverse(viewmats)[:, :3, 3]  # [C,3]
        dirs = means[None, :, :] - campos[:, None, :]
        shs = colors[None].expand(C, *colors.shape) if colors.dim() == 3 else colors
        cols = spherical_harmonics(sh_degree, dirs, shs, masks=radii > 0)
        cols = torch.clamp_min(cols + 0.5, 0.0)
    tw = math.ceil(width / tile_size)
    th = math.ceil(height / tile_size)
    tiles_per_gauss, isect_ids, flatten_ids = isect_tiles(means2d.detach(), radii, depths.detach(), tile_size, tw, th)
    isect_offsets = isect_offset_encode(isect_ids, C, tw, th)
    if counters is not None:
        counters["N_vis"] = int((radii > 0).sum())
        counters["n_isects"] = int(flatten_ids.shape[0])
    render_colors, render_alphas, last_ids = rasterize_to_pixels(
        means2d, conics, cols, opac, width, height, tile_size, isect_offsets, flatten_ids,
        backgrounds=backgrounds, absgrad=absgrad, counters=counters, tile_window=tile_window)
    meta = {
        "camera_ids": None, "gaussian_ids": None, "radii": radii, "means2d": means2d, "depths": depths,
        "conics": conics, "opacities": opac, "colors": cols, "tile_width": tw, "tile_height": th,
        "tiles_per_gauss": tiles_per_gauss, "isect_ids": isect_ids, "flatten_ids": flatten_ids,
        "isect_offsets": isect_offsets, "last_ids": last_ids, "width": width, "height": height,
        "tile_size": tile_size, "n_cameras": C,
    }
    return render_colors, render_alphas, meta


# --------------------------------------------------------------------------------------
# consumer semantics: densification statistics (/root/reference/model/gaussian.py:188-197),
# generalised to C views (per-view contributions)
# --------------------------------------------------------------------------------------


def update_statistics(max_radii: Tensor, grad_norm_accum: Tensor, collecting_counts: Tensor,
                      radii: Tensor, absgrad: Tensor, width: int, height: int) -> None:
    """In-place.  radii[C,N] int32, absgrad[C,N,2]; every view c contributes exactly what
    gaussian.py:188-197 does for view 0."""
    max_hw = max(height, width)
    for c in range(radii.shape[0]):
        r = radii[c].to(max_radii.dtype) / max_hw
        visible = r > 0.0
        max_radii[visible] = torch.max(max_radii[visible], r[visible])
        grads = torch.norm(absgrad[c], dim=-1) * max_hw
        grad_norm_accum[visible] = grad_norm_accum[visible] + grads[visible]
        collecting_counts[visible] = collecting_counts[visible] + 1
