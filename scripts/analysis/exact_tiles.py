"""CPU analysis: of the entries of the TIGHT rectangles (scripts/analysis/tight_rects.py), how many tiles really hold a
point with sigma <= cut (exact ellipse / tile-box test on the continuous box of pixel centres)?"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from easy_gaussian_splatting_b200.synthetic import make_config_scene
from oracle import gsplat_oracle as O

for name in sys.argv[1:] or ["metric", "cfg2"]:
    sc = make_config_scene(name)
    W, H = sc.width, sc.height
    with torch.no_grad():
        radii, m2, d, conics = O.fully_fused_projection(sc.means, sc.quats, sc.scales, sc.viewmats[:1], sc.Ks[:1], W, H)
    vis = (radii[0] > 0).numpy()
    x, y = m2[0, :, 0].numpy()[vis].astype(np.float64), m2[0, :, 1].numpy()[vis].astype(np.float64)
    r = radii[0].numpy()[vis].astype(np.float64)
    a, b, c = (conics[0, :, i].numpy()[vis].astype(np.float64) for i in range(3))
    o = sc.opacities.numpy()[vis].astype(np.float64)
    tw, th = -(-W // 16), -(-H // 16)
    cx0 = np.clip(np.floor((x - r) / 16), 0, tw); cx1 = np.clip(np.ceil((x + r) / 16), 0, tw)
    cy0 = np.clip(np.floor((y - r) / 16), 0, th); cy1 = np.clip(np.ceil((y + r) / 16), 0, th)
    cut = np.log(255.0 * o)
    det = a * c - b * b
    hx = np.sqrt(np.maximum(2 * cut * c / det, 0)); hy = np.sqrt(np.maximum(2 * cut * a / det, 0))
    px0 = np.ceil(x - hx - 0.5); px1 = np.floor(x + hx - 0.5)
    py0 = np.ceil(y - hy - 0.5); py1 = np.floor(y + hy - 0.5)
    tx0 = np.maximum(np.clip(np.floor(px0 / 16), 0, tw), cx0); tx1 = np.minimum(np.clip(np.floor(px1 / 16) + 1, 0, tw), cx1)
    ty0 = np.maximum(np.clip(np.floor(py0 / 16), 0, th), cy0); ty1 = np.minimum(np.clip(np.floor(py1 / 16) + 1, 0, th), cy1)
    w = np.maximum(tx1 - tx0, 0).astype(np.int64); h = np.maximum(ty1 - ty0, 0).astype(np.int64)
    w = np.where(cut > 0, w, 0); cnt = w * h
    n = int(cnt.sum())
    owner = np.repeat(np.arange(cnt.size), cnt)
    start = np.cumsum(cnt) - cnt
    j = np.arange(n) - start[owner]
    ww = np.maximum(w[owner], 1)
    ty = ty0[owner] + j // ww; tx = tx0[owner] + j % ww
    # box of pixel centres of the tile, relative to the mean
    bx0 = tx * 16 + 0.5 - x[owner]; bx1 = np.minimum(tx * 16 + 15.5, W - 0.5) - x[owner]
    by0 = ty * 16 + 0.5 - y[owner]; by1 = np.minimum(ty * 16 + 15.5, H - 0.5) - y[owner]
    A, B, Cc = a[owner], b[owner], c[owner]
    q = lambda dx, dy: 0.5 * (A * dx * dx + Cc * dy * dy) + B * dx * dy
    inside = (bx0 <= 0) & (bx1 >= 0) & (by0 <= 0) & (by1 >= 0)
    best = np.full(n, np.inf)
    for dxe in (bx0, bx1):  # vertical edges: minimise over dy
        dy = np.clip(-B * dxe / Cc, by0, by1)
        best = np.minimum(best, q(dxe, dy))
    for dye in (by0, by1):
        dx = np.clip(-B * dye / A, bx0, bx1)
        best = np.minimum(best, q(dx, dye))
    best = np.where(inside, 0.0, best)
    keep = best <= cut[owner]
    print(f"{name}: tight entries {n}, exact {int(keep.sum())} = {keep.sum() / n:.3f} of tight; per-size: ", end="")
    for lo, hi in ((1, 1), (2, 4), (5, 16), (17, 1 << 30)):
        sel = (cnt[owner] >= lo) & (cnt[owner] <= hi)
        print(f"[{lo}-{hi}] share {sel.sum() / n:.2f} keep {keep[sel].sum() / max(sel.sum(), 1):.2f}; ", end="")
    print()
