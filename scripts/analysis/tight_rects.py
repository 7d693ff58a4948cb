"""CPU analysis: how many (tile, Gaussian) entries remain when a Gaussian's tile rectangle is the classic 3-sigma square
intersected with the axis-aligned extent of {alpha >= 1/255} = {sigma <= ln(255 o)} (hx = sqrt(2 cut cov_xx))."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from easy_gaussian_splatting_b200.synthetic import make_config_scene
from oracle import gsplat_oracle as O

for name in sys.argv[1:] or ["metric", "cfg2", "cfg3"]:
    sc = make_config_scene(name)
    W, H = sc.width, sc.height
    with torch.no_grad():
        radii, m2, d, conics = O.fully_fused_projection(sc.means, sc.quats, sc.scales, sc.viewmats, sc.Ks, W, H)
    vis = (radii[0] > 0).numpy()
    x, y = m2[0, :, 0].numpy()[vis], m2[0, :, 1].numpy()[vis]
    r = radii[0].numpy()[vis].astype(np.float32)
    a, b, c = (conics[0, :, i].numpy()[vis] for i in range(3))
    o = sc.opacities.numpy()[vis]
    tw, th = -(-W // 16), -(-H // 16)
    def rect(x0, x1, y0, y1):
        tx0 = np.clip(np.floor(x0 / 16), 0, tw); tx1 = np.clip(np.ceil(x1 / 16), 0, tw)
        ty0 = np.clip(np.floor(y0 / 16), 0, th); ty1 = np.clip(np.ceil(y1 / 16), 0, th)
        return tx0, tx1, ty0, ty1
    cx0, cx1, cy0, cy1 = rect(x - r, x + r, y - r, y + r)
    classic = ((cx1 - cx0) * (cy1 - cy0)).sum()
    cut = np.log(255.0 * o)
    det = a * c - b * b
    hx = np.sqrt(np.maximum(2 * cut * c / det, 0)) * (1 + 1e-4) + 1e-3
    hy = np.sqrt(np.maximum(2 * cut * a / det, 0)) * (1 + 1e-4) + 1e-3
    # tiles that contain a pixel centre p + 0.5 within [x - hx, x + hx]
    px0 = np.ceil(x - hx - 0.5); px1 = np.floor(x + hx - 0.5)
    py0 = np.ceil(y - hy - 0.5); py1 = np.floor(y + hy - 0.5)
    tx0 = np.clip(np.floor(px0 / 16), 0, tw); tx1 = np.clip(np.floor(px1 / 16) + 1, 0, tw)
    ty0 = np.clip(np.floor(py0 / 16), 0, th); ty1 = np.clip(np.floor(py1 / 16) + 1, 0, th)
    tx0, tx1 = np.maximum(tx0, cx0), np.minimum(tx1, cx1)
    ty0, ty1 = np.maximum(ty0, cy0), np.minimum(ty1, cy1)
    wq = np.maximum(tx1 - tx0, 0) * np.maximum(ty1 - ty0, 0)
    wq = np.where(cut > 0, wq, 0)
    print(f"{name}: visible {vis.sum()}, classic entries {int(classic)}, tight {int(wq.sum())} = {wq.sum() / classic:.3f}")
