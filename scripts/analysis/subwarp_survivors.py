"""CPU analysis (no GPU): how many (warp, Gaussian) survivors the blend kernels' exact rectangle culling leaves for
different sub-rectangle granularities, on a BASELINE scene.  Decides whether per-half-warp / per-quarter-warp survivor
lists are worth building (blend.cu).  Ignores early termination (it scales every variant alike)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from easy_gaussian_splatting_b200.synthetic import make_config_scene
from oracle import gsplat_oracle as O

name = sys.argv[1] if len(sys.argv) > 1 else "metric"
sc = make_config_scene(name)
W, H = sc.width, sc.height
with torch.no_grad():
    radii, means2d, depths, conics = O.fully_fused_projection(sc.means, sc.quats, sc.scales, sc.viewmats, sc.Ks, W, H)
    tw, th = -(-W // 16), -(-H // 16)
    tpg, ids, flat = O.isect_tiles(means2d, radii, depths, 16, tw, th)
nb = O.tile_n_bits(tw, th)
tile = ((ids >> 32) & ((1 << nb) - 1)).numpy()
g = flat.long().numpy()
m = means2d[0].numpy()[g]
cn = conics[0].numpy()[g]
op = sc.opacities.numpy()[g]
cut = np.log(255.0 * op)
tx, ty = tile % tw, tile // tw


def touches(rx_lo, rx_hi, ry_lo, ry_hi):
    a, b, c = cn[:, 0], cn[:, 1], cn[:, 2]
    dxl, dxh, dyl, dyh = rx_lo - m[:, 0], rx_hi - m[:, 0], ry_lo - m[:, 1], ry_hi - m[:, 1]
    inside = (dxl <= 0) & (dxh >= 0) & (dyl <= 0) & (dyh >= 0)
    nb_c, nb_a = -b / c, -b / a
    best = np.full(len(a), np.inf, dtype=np.float32)
    for dx in (dxl, dxh):
        dy = np.minimum(np.maximum(nb_c * dx, dyl), dyh)
        best = np.minimum(best, 0.5 * (a * dx * dx + c * dy * dy) + b * dx * dy)
    for dy in (dyl, dyh):
        dx = np.minimum(np.maximum(nb_a * dy, dxl), dxh)
        best = np.minimum(best, 0.5 * (a * dx * dx + c * dy * dy) + b * dx * dy)
    return (cut > 0) & (inside | (best <= cut))


def count(wx, wy, w, h):
    """survivors of the sub-rectangle at pixel offset (wx, wy) of size w x h inside each tile"""
    x0 = tx * 16 + wx + 0.5
    y0 = ty * 16 + wy + 0.5
    return touches(x0, x0 + (w - 1), y0, y0 + (h - 1))

n = len(g)
print(f"{name}: n_isects={n}")
tot_warp = 0
for wy in (0, 8):
    full = count(0, wy, 16, 8)
    halves = [count(hx, wy, 8, 8) for hx in (0, 8)]
    quads = [count(qx, wy + qy, 8, 4) for qy in (0, 4) for qx in (0, 8)]
    quads_v = [count(qx, wy, 4, 8) for qx in (0, 4, 8, 12)]
    # per tile maxima: trips = max over sub-lists of their lengths, per (tile, warp)
    def per_tile(mask):
        return np.bincount(tile, weights=mask.astype(np.float64), minlength=tw * th)
    f = per_tile(full)
    h_ = np.stack([per_tile(x) for x in halves])
    q_ = np.stack([per_tile(x) for x in quads])
    qv = np.stack([per_tile(x) for x in quads_v])
    print(f" warp y={wy}: survivors {full.sum()/n:.3f} of entries | halves sum {sum(x.sum() for x in halves)/full.sum():.3f}x, "
          f"trips(max) {h_.max(0).sum()/f.sum():.3f} | quarters(8x4) sum {sum(x.sum() for x in quads)/full.sum():.3f}x, trips(max) {q_.max(0).sum()/f.sum():.3f}"
          f" | quarters(4x8) trips {qv.max(0).sum()/f.sum():.3f}")
    # per 64-entry batch maxima are what the kernel would really see (lists are compacted per batch)
