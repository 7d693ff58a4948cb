"""CPU tests (gloo, world_size 2) of the view-sharded step's host logic: the flat gradient bucket and the
densify-statistics exchange must reproduce what a single process accumulates over all views."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from easy_gaussian_splatting_b200.distributed import DensifyStats, FlatGradBucket, shard_views


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_view(v, shapes, N):
    """Deterministic per-view 'gradients' and statistics, standing in for one rasterization fwd+bwd."""
    g = torch.Generator().manual_seed(1000 + v)
    grads = [torch.randn(*s, generator=g) for s in shapes]
    radii = torch.randint(0, 20, (1, N), generator=g, dtype=torch.int32)
    absgrad = torch.rand(1, N, 2, generator=g)
    return grads, radii, absgrad


def _stats_update_cpu(stats, radii, absgrad, W, H):
    # same per-view semantics as the CUDA kernel / gaussian.py:188-197 (CPU tensors in this test)
    max_hw = max(W, H)
    r = radii[0].float() / max_hw
    vis = r > 0
    t = stats.step_buf  # rows: grad_norm_accum, collecting_counts, max_radii (the step's delta when view-sharded)
    t[2][vis] = torch.max(t[2][vis], r[vis])
    t[0][vis] += absgrad[0].norm(dim=-1)[vis] * max_hw
    t[1][vis] += 1


SHAPES = [(50, 3), (50, 4), (50, 3), (50,), (50, 16, 3)]
N_VIEWS, N, W, H = 6, 50, 640, 480


def _worker(rank, world, port, out_q, rank_attach=True):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    params = [torch.zeros(*s, requires_grad=True) for s in SHAPES]
    attach = rank_attach  # statistics riding in the gradient bucket, or exchanged on their own
    bucket = FlatGradBucket(params, stats_size=N if attach else 0)
    stats = DensifyStats(N, "cpu", bucket=bucket if attach else None)
    stats.buf[0] = 1.0  # pre-existing accumulations, identical on all replicas
    stats.buf[2] = 0.004
    bucket.zero_()
    stats.begin_step()
    for v in shard_views(N_VIEWS, rank, world):
        grads, radii, absgrad = _fake_view(v, SHAPES, N)
        for p, g in zip(params, grads):
            p.grad += g  # autograd accumulates in place into the bucket views
        _stats_update_cpu(stats, radii, absgrad, W, H)
    if attach:
        with pytest.raises(RuntimeError):
            stats.all_reduce()  # the bucket has not been reduced yet
    bucket.all_reduce()
    stats.all_reduce()
    out_q.put((rank, bucket.flat.clone(), stats.buf.clone(), [p.grad.data_ptr() == v.data_ptr() for p, v in zip(params, bucket.views)]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("attach", [True, False])
def test_view_sharded_exchange_matches_single_process(attach):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, attach)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process reference over all views
    params = [torch.zeros(*s, requires_grad=True) for s in SHAPES]
    bucket = FlatGradBucket(params)
    stats = DensifyStats(N, "cpu")
    assert stats.delta is None and stats.exchange.startswith("none")
    stats.buf[0] = 1.0
    stats.buf[2] = 0.004
    for v in range(N_VIEWS):
        grads, radii, absgrad = _fake_view(v, SHAPES, N)
        for p, g in zip(params, grads):
            p.grad += g
        _stats_update_cpu(stats, radii, absgrad, W, H)
    for rank, flat, sbuf, aliased in results:
        assert all(aliased), "param.grad must alias the flat bucket (no pack/unpack copies)"
        n = sum(-(-p.numel() // 4) * 4 for p in params)  # gradient part (the 2-rank bucket is padded and may carry stats rows)
        assert torch.allclose(flat[:n], bucket.flat[:n], atol=1e-5), f"rank {rank} gradients"
        assert torch.allclose(sbuf[:2], stats.buf[:2], atol=1e-4), f"rank {rank} SUM statistics"
        assert torch.equal(sbuf[2], stats.buf[2]), f"rank {rank} MAX statistics"
    assert torch.equal(results[0][1], results[1][1]), "replicas must end bit-identical"


def test_shard_views_partition():
    for world in (1, 2, 4, 8):
        allv = sorted(v for r in range(world) for v in shard_views(64, r, world))
        assert allv == list(range(64))
        assert all(len(shard_views(64, r, world)) == 64 // world for r in range(world))


def test_flat_bucket_survives_optimizer_zero_grad():
    params = [torch.zeros(4, 3, requires_grad=True), torch.zeros(4, requires_grad=True)]
    bucket = FlatGradBucket(params)
    opt = torch.optim.Adam(params, lr=0.1)
    (params[0].sum() + params[1].sum()).backward()
    assert float(bucket.flat.sum()) == 16.0
    opt.step()
    opt.zero_grad(set_to_none=True)
    bucket.zero_()
    assert all(p.grad is not None and p.grad.data_ptr() == v.data_ptr() for p, v in zip(params, bucket.views))


def test_direct_gradient_routes_on_cpu():
    """FlatGradBucket.direct(): gradients that reach a parameter by another route than the fused backward (here plain
    autograd on CPU tensors) are copied into the bucket and .grad is re-pointed at the bucket view; the registry of
    gradient targets never captures a tensor whose registration is stale."""
    import torch
    from easy_gaussian_splatting_b200 import stages
    from easy_gaussian_splatting_b200.distributed import FlatGradBucket
    params = [torch.randn(5, 3, requires_grad=True), torch.randn(7, requires_grad=True)]
    bucket = FlatGradBucket(params)
    assert bucket.exchange == "NCCL all-reduce" and bucket.flat.numel() % 4 == 0
    with bucket.direct():
        assert all(p.grad is None for p in params)
        ((params[0] * 2.0).sum() + (params[1] * 3.0).sum()).backward()
    assert not stages._GRAD_TARGETS  # registration removed
    for p, v, k in zip(params, bucket.views, (2.0, 3.0)):
        assert p.grad.data_ptr() == v.data_ptr() and torch.equal(p.grad, torch.full_like(p, k))
    # a parameter without gradient in a direct step ends with a zero bucket slice
    with bucket.direct():
        (params[0] * 1.0).sum().backward()
    assert torch.equal(bucket.views[1], torch.zeros(7)) and torch.equal(bucket.views[0], torch.ones(5, 3))
    # _grad_buffer: a registered target is used ONCE while the owner lives (a second backward on the same parameter
    # inside the window must not overwrite the first gradient), and is ignored once the owner is gone
    owner = torch.zeros(4, 3)
    target = torch.zeros(12)
    stages.register_grad_target(owner, target)
    out = stages._grad_buffer(owner)
    assert out.data_ptr() == target.data_ptr() and out.shape == owner.shape
    assert stages._grad_buffer(owner).data_ptr() != target.data_ptr(), "registration must be one-shot"
    assert not stages._GRAD_TARGETS
    stages.register_grad_target(owner, target)
    key = stages._target_key(owner)
    del owner, out
    import gc
    gc.collect()
    impostor = torch.zeros(4, 3)
    stages._GRAD_TARGETS[stages._target_key(impostor)] = stages._GRAD_TARGETS.pop(key)  # same address, dead owner
    assert stages._grad_buffer(impostor).data_ptr() != target.data_ptr()
    # a target that is not 16-byte aligned is never handed to the kernels (they store float4s)
    owner2 = torch.zeros(4, 3)
    store = torch.zeros(16)
    stages.register_grad_target(owner2, store[1:13])
    assert stages._grad_buffer(owner2).data_ptr() != store[1:13].data_ptr()
    # a strided input gets a dense row-major gradient buffer
    assert stages._grad_buffer(torch.zeros(3, 4).t()).is_contiguous()
    stages.clear_grad_targets()
    with pytest.raises(ValueError):
        stages.register_grad_target(torch.zeros(3), torch.zeros(4))


def test_bucket_slices_are_16_byte_aligned_for_any_n():
    """after densify / prune N is arbitrary; every parameter's slice of the flat bucket must still start on
    a 16-byte boundary because the fused backward stores float4s into it."""
    for n in (1, 5, 50, 1001, 4097):
        params = [torch.zeros(*s, requires_grad=True) for s in ((n, 3), (n, 4), (n, 3), (n,), (n, 1, 3), (n, 15, 3))]
        bucket = FlatGradBucket(params)
        base = bucket.flat.data_ptr()
        for p, v in zip(params, bucket.views):
            assert (v.data_ptr() - base) % 16 == 0 and v.shape == p.shape and p.grad.data_ptr() == v.data_ptr()
        assert bucket.flat.numel() % 4 == 0


def test_all_reduce_async_contract_without_a_group():
    params = [torch.zeros(4, 3, requires_grad=True)]
    bucket = FlatGradBucket(params, symmetric=True)  # no process group: must not try to rendezvous
    assert bucket.exchange == "NCCL all-reduce"
    assert bucket.all_reduce() is None
    work = bucket.all_reduce(async_op=True)
    assert work.wait() and work.is_completed()
