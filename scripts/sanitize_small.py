"""Small end-to-end workload for compute-sanitizer runs (memcheck / racecheck / synccheck / initcheck): every kernel
family of the library on scenes small enough for the tools' 10-100x slowdown.

    compute-sanitizer --tool memcheck python scripts/sanitize_small.py
    torchrun --nproc-per-node 2 ... scripts/sanitize_small.py --collective     (2 GPUs: both all-reduce kernels)
"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

from easy_gaussian_splatting_b200 import rasterization, rasterization_from_parameters, stages
from easy_gaussian_splatting_b200.distributed import DensifyStats, FlatGradBucket
from easy_gaussian_splatting_b200.loss import fused_l1_ssim_loss
from easy_gaussian_splatting_b200.optim import FusedAdam
from easy_gaussian_splatting_b200.synthetic import loss_weights, make_scene

NAMES = ("means", "quats", "scales", "opacities", "colors")


def render_twice(sc, label, **kw):
    """Two calls of the same shape: the first sizes its buffers exactly (one wait), the second from the first's counts
    (device-side counts, sentinel behind the offsets, possibly the segmented backward)."""
    C = sc.viewmats.shape[0]
    Wc, Wa = (t.cuda() for t in loss_weights(1, C, sc.height, sc.width))
    bg = sc.background[None].expand(C, 3).contiguous()
    for it in range(2):
        p = [getattr(sc, k).clone().requires_grad_(True) for k in NAMES]
        rc, ra, meta = rasterization(*p, sc.viewmats, sc.Ks, sc.width, sc.height, sh_degree=3, packed=False, absgrad=True,
                                     backgrounds=bg, **kw)
        ((rc * Wc).sum() + (ra * Wa).sum()).backward()
        _ = meta["isect_ids"]
        torch.cuda.synchronize()
    print("ok", label, float(rc.mean()), float(p[0].grad.abs().mean()), "n_isects", meta["flatten_ids"].numel())
    return p, meta, rc


def pile_scene(n=4000, seed=13):
    sc = make_scene("blob", n, 160, 96, 200.0, seed)
    g = torch.Generator().manual_seed(seed)
    sc.means = (torch.randn(n, 3, generator=g) * torch.tensor([0.03, 0.03, 0.3])).contiguous()
    sc.scales = torch.exp(torch.log(torch.tensor(0.01)) + 0.3 * torch.randn(n, 3, generator=g)).contiguous()
    sc.opacities = (0.004 + 0.05 * torch.rand(n, generator=g)).contiguous()
    return sc


def single_gpu():
    stages.reset_binning_hints()
    blob = make_scene("blob", 3001, 131, 77, 120.0, 7, n_views=2).to("cuda")  # N % 4 != 0, ragged tiles, C = 2
    p, meta, rc = render_twice(blob, "blob C=2")
    render_twice(make_scene("outdoor", 20000, 320, 200, 200.0, 2).to("cuda"), "outdoor")
    render_twice(blob, "blob antialiased", rasterize_mode="antialiased")
    pile = pile_scene().to("cuda")
    render_twice(pile, "pile (automatic segmented replay on the second call)")
    os.environ["EGS_BWD_SEGMENT"] = "64"
    render_twice(pile, "pile EGS_BWD_SEGMENT=64")
    del os.environ["EGS_BWD_SEGMENT"]
    # raw-parameter entry point, gradients written straight into a flat bucket, statistics, fused Adam
    raw = [blob.means, blob.quats, torch.log(blob.scales), torch.logit(blob.opacities), blob.colors[:, :1].contiguous(),
           blob.colors[:, 1:].contiguous()]
    raw = [t.clone().requires_grad_(True) for t in raw]
    bucket = FlatGradBucket(raw)
    stats = DensifyStats(blob.means.shape[0], "cuda")
    opt = FusedAdam([{"params": [t], "lr": 1e-3} for t in raw])
    bg = blob.background[None].expand(2, 3).contiguous()
    with bucket.direct():
        rc, ra, meta = rasterization_from_parameters(*raw, blob.viewmats, blob.Ks, blob.width, blob.height, 3, backgrounds=bg, absgrad=True)
        gt = torch.rand_like(rc)
        total, _, _ = fused_l1_ssim_loss(torch.clamp(rc, 0, 1), gt, (torch.rand(2, blob.height, blob.width, device="cuda") < 0.1).float(), 0.2)
        total.sum().backward()
    stats.update_local(meta["radii"], meta["means2d"].absgrad, blob.width, blob.height)
    opt.step()
    torch.cuda.synchronize()
    print("ok raw + loss + stats + adam", float(total.sum()), float(stats.buf.sum()))
    # the classic route's operators (64-bit keys)
    tw, th = stages.tile_grid(blob.width, blob.height)
    proj = stages.projection_fwd(blob.means, blob.quats, blob.scales, blob.opacities, blob.colors, blob.viewmats, blob.Ks,
                                 blob.width, blob.height, 3)
    _, ids, flat = stages.isect_tiles(proj["means2d"], proj["radii"], proj["depths"], 16, tw, th, sort=True, tiles_per_gauss=proj["tiles_per_gauss"])
    offs = stages.isect_offset_encode(ids, 2, tw, th)
    torch.cuda.synchronize()
    print("ok classic route", ids.numel(), int(offs.max()))


def collective():
    import torch.distributed as dist
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    n = 5001
    params = [torch.zeros(n, 3, device="cuda", requires_grad=True), torch.zeros(n, 16, 3, device="cuda", requires_grad=True)]
    bucket = FlatGradBucket(params, stats_size=n)
    stats = DensifyStats(n, "cuda", bucket=bucket)
    for step in range(3):
        bucket.flat[:bucket.grad_floats].fill_(float(rank + 1))
        stats.begin_step()
        stats.step_buf[0].fill_(1.0)
        stats.step_buf[2].fill_(0.1 * (rank + 1))
        bucket.all_reduce()
        stats.all_reduce()
    torch.cuda.synchronize()
    ok = float(bucket.views[0][0, 0]) == world * (world + 1) / 2 and abs(float(stats.max_radii[0]) - 0.1 * world) < 1e-6
    print(f"rank {rank}: {bucket.exchange} | {stats.exchange} | ok={ok}")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    collective() if "--collective" in sys.argv else single_gpu()
