python - <<'PY'
import sys; sys.path.insert(0, ".")
import torch
from easy_gaussian_splatting_b200 import stages
from easy_gaussian_splatting_b200.synthetic import make_scene
sc = make_scene("object", 20000, 400, 400, 555.0, 1, white_background=True).to("cuda")
W, H = 400, 400
tw, th = stages.tile_grid(W, H)
proj = stages.projection_fwd(sc.means, sc.quats, sc.scales, sc.opacities, sc.colors, sc.viewmats, sc.Ks, W, H, 3)
stages.reset_binning_hints()
for it in range(2):
    b = stages.isect_sorted_async(proj["means2d"], proj["radii"], proj["depths"], proj["tiles_per_gauss"], 16, tw, th)
    b.resolve(); b.note_for_next_call()
    order = b.tile_order.long()
    lens = torch.diff(torch.cat([b.offsets.reshape(-1), b.offsets_store[-1:]])).long()
    cls = torch.where(lens > 0, torch.floor(torch.log2(lens.clamp_min(1).double())).long() + 1, torch.zeros_like(lens))
    co = cls[order]
    bad = (co[1:] > co[:-1]).nonzero().flatten()
    print("iter", it, "exact", b.exact, "n_tiles", order.numel(), "violations", bad.numel(), "classes hist", torch.bincount(cls).tolist())
    if bad.numel():
        i = int(bad[0]); print(" first at", i, co[max(0,i-3):i+4].tolist(), order[max(0,i-3):i+4].tolist(), lens[order[max(0,i-3):i+4]].tolist())
PY
