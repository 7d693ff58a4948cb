#!/usr/bin/env python
"""torchrun helper: time FlatGradBucket.all_reduce() (gradients + densify-stat rows of N Gaussians) against NCCL."""
import os, sys
sys.path.insert(0, ".")
import torch, torch.distributed as dist
from easy_gaussian_splatting_b200.distributed import DensifyStats, FlatGradBucket
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
params = [torch.zeros(N, w, device=dev, requires_grad=True) for w in (3, 4, 3, 1, 48)]
bucket = FlatGradBucket(params, stats_size=N)
stats = DensifyStats(N, dev, bucket=bucket)
def timeit(fn, it=40):
    for _ in range(5): fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / it], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)
ours = timeit(bucket.all_reduce)
flat2 = torch.zeros_like(bucket.flat)
nccl = timeit(lambda: dist.all_reduce(flat2))
if rank == 0:
    mb = bucket.flat.numel() * 4 / 1e6
    print(f"{os.environ.get('EGS_RASTER_LIB', 'base').split('/')[-1]}: {bucket.exchange}: {ours:.3f} ms for {mb:.0f} MB ({mb / ours:.0f} GB/s algorithmic), NCCL {nccl:.3f} ms")
dist.barrier(); dist.destroy_process_group()
