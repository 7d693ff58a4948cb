#!/usr/bin/env python
"""Benchmark of the differentiable Gaussian rasterizer hot path (BASELINE.json metric:
"fwd+bwd raster Mpix/s at 1M Gaussians 1080p").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload metric|cfg2|...]

One "step" = every rank renders its `views_per_rank` camera views of the replicated synthetic
scene through `rasterization()` forward + backward of the linear functional
L = sum(colors*Wc) + sum(alphas*Wa); for N > 1 the step ends with the NCCL
all-reduce of the flat parameter-gradient bucket and of the densification statistics (the path's
one real exchange step).  value = total pixels rendered by all ranks / time, Mpix/s.

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU restatement of the reference's
rasterizer (oracle/, pure PyTorch — gsplat itself is not installable here) on a
bounded crop of the same workload on the host cores.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

METRIC = "fwd+bwd raster Mpix/s at 1M Gaussians 1080p"
UNIT = "Mpix/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="metric")
    ap.add_argument("--views-per-rank", type=int, default=4)
    ap.add_argument("--n-gaussians", type=int, default=None, help="override N (debug only; marks the line invalid)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-stage-timing", action="store_true")
    ap.add_argument("--ref-crop", type=int, default=8, help="reference arm: crop = 1/ref_crop of W and of H")
    ap.add_argument("--train-step", action="store_true",
                    help="cfg4-style full training step: --total-views views per step split across the ranks (strong "
                         "scaling), gradient + densify-stat all-reduce, fused Adam")
    ap.add_argument("--total-views", type=int, default=64)
    ap.add_argument("--no-train-step", action="store_true",
                    help="skip the extra cfg4 training-step measurement that the default run appends as line['train_step']")
    ap.add_argument("--no-view-pipelining", action="store_true",
                    help="render the views of a step strictly one after the other on one stream")
    ap.add_argument("--activations", default="pre", choices=["pre", "torch", "folded"],
                    help="pre: activated tensors are the inputs (the metric's definition); torch: raw parameters, "
                         "exp/sigmoid/cat per view in torch like GaussianModel's properties; folded: raw parameters "
                         "through rasterization_from_parameters (activations folded into the kernels)")
    ap.add_argument("--views-per-call", type=int, default=4,
                    help="views batched into one rasterization() call (viewmats [C,4,4]); 1 = one call per view as in the reference's loop")
    ap.add_argument("--forward-only", action="store_true", help="no_grad forward renders only (cfg5-style latency runs; not the headline metric)")
    ap.add_argument("--no-call-pattern", action="store_true",
                    help="skip line['call_pattern']: the reference's own one-view-per-call loop (strictly sequential, one "
                         "stream) on the metric scene and BASELINE configs 2, 3, 5, and the config-1 CPU timing")
    ap.add_argument("--no-exchange-check", action="store_true",
                    help="N > 1: skip the pre-step that checks the view-sharded step against one GPU rendering all views")
    ap.add_argument("--quick", action="store_true", help="device-timed value only: no e2e, stage timing, sub-runs or CPU leg")
    ap.add_argument("--shard", default="strided", choices=["strided", "contiguous"],
                    help="N > 1: rank r renders views {v : v mod R == r} (distributed.shard_views, the default) or the "
                         "contiguous block r*V .. r*V+V-1 (diagnostic)")
    ap.add_argument("--no-exchange", action="store_true", help="N > 1 diagnostic: skip the exchange step (marks the line invalid)")
    return ap.parse_args()


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm_gbs=float(d["hbm_gbs"]), source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, source="fallback (B200_PROFILING.md)")


# ------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle on a bounded crop of the workload
# ------------------------------------------------------------------------------------------------
def cpu_sample_scene(workload: str, crop: int, n_override=None):
    """Same scene and camera; only a centred (W/crop x H/crop) window of the image is rendered
    (principal point shifted accordingly), which bounds the CPU work."""
    from easy_gaussian_splatting_b200.synthetic import CONFIGS, make_config_scene
    sc = make_config_scene(workload, n_views=1, N=n_override)
    cw, ch = max(16, sc.width // crop), max(16, sc.height // crop)
    x0, y0 = (sc.width - cw) // 2, (sc.height - ch) // 2
    Ks = sc.Ks.clone()
    Ks[:, 0, 2] -= x0
    Ks[:, 1, 2] -= y0
    sc.Ks, sc.width, sc.height = Ks, cw, ch
    return sc, f"{cw}x{ch} centre crop of the {CONFIGS[workload]['width']}x{CONFIGS[workload]['height']} view, all {sc.means.shape[0]} Gaussians, 1 view"


def run_cpu_oracle(sc, steps: int, warmup: int):
    """fwd+bwd of the oracle; returns (median seconds per step, Mpix/s)."""
    from easy_gaussian_splatting_b200.synthetic import loss_weights
    from oracle import gsplat_oracle as O
    Wc, Wa = loss_weights(sc.seed, 1, sc.height, sc.width)
    times = []
    for it in range(warmup + steps):
        leaves = [t.detach().clone().requires_grad_(True) for t in (sc.means, sc.quats, sc.scales, sc.opacities, sc.colors)]
        t0 = time.perf_counter()
        rc, ra, meta = O.rasterization(*leaves, sc.viewmats, sc.Ks, sc.width, sc.height, sh_degree=3, packed=False,
                                       absgrad=True, backgrounds=sc.background[None])
        ((rc * Wc).sum() + (ra * Wa).sum()).backward()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    med = statistics.median(times)
    return med, sc.width * sc.height / med / 1e6, sum(times)


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # only rank 0 runs the CPU comparator
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sc, sample = cpu_sample_scene(args.workload, args.ref_crop, args.n_gaussians)
    med, mpix, total = run_cpu_oracle(sc, args.steps, max(args.warmup, 1) if args.steps > 1 else args.warmup)
    line = {
        "metric": METRIC, "value": mpix, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": med * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "sample": sample,
                   "note": "CPU restatement of gsplat 1.0.0 semantics in pure PyTorch (oracle/); gsplat itself is "
                           "not installable in this image"},
        "cpu_baseline": {"value": mpix, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": mpix, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def linear_functional(rc, ra, Wc, Wa):
    """L = sum(colors * Wc) + sum(alphas * Wa) as two dot products: one fused read of each pair of
    tensors forward, one multiply per tensor backward — the benchmark's own scaffolding should not weigh on the
    rasterizer's timed region more than it has to (the elementwise product + sum form costs two more passes)."""
    return torch.dot(rc.reshape(-1), Wc.reshape(-1)) + torch.dot(ra.reshape(-1), Wa.reshape(-1))


def workload_name(args):
    from easy_gaussian_splatting_b200.synthetic import CONFIGS
    c = dict(CONFIGS[args.workload])
    if args.n_gaussians:
        c["N"] = args.n_gaussians
    return f"{args.workload}: {c['N']} Gaussians ({c['kind']} scene, seed {c['seed']}), {c['width']}x{c['height']}, SH degree 3, absgrad, packed=False"


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def ours(args):
    import torch.distributed as dist
    from easy_gaussian_splatting_b200 import _lib, rasterization, stages
    from easy_gaussian_splatting_b200.distributed import DensifyStats, FlatGradBucket
    from easy_gaussian_splatting_b200.synthetic import CONFIGS, loss_weights, make_config_scene

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl ours needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    V = args.views_per_rank
    if args.train_step:
        if args.total_views % world:
            raise SystemExit("--total-views must be divisible by the number of GPUs")
        V = args.total_views // world
    cfg = CONFIGS[args.workload]
    W, H = cfg["width"], cfg["height"]

    # identical replicated scene on every rank; rank r renders views {v : v mod R == r} of a world*V orbit
    # (distributed.shard_views)
    from easy_gaussian_splatting_b200.distributed import shard_views
    sc_cpu = make_config_scene(args.workload, n_views=world * V, N=args.n_gaussians)
    N = sc_cpu.means.shape[0]
    my_views = shard_views(world * V, rank, world) if args.shard == "strided" else list(range(rank * V, rank * V + V))
    names = ("means", "quats", "scales", "opacities", "colors")
    if args.activations == "pre":
        params = [getattr(sc_cpu, k).to(dev).requires_grad_(True) for k in names]
    else:  # the reference's six raw parameter tensors (model/gaussian.py:32-54)
        names = ("means", "quats", "log_scales", "logit_opacities", "sh_0", "sh_rest")
        raw = [sc_cpu.means, sc_cpu.quats, torch.log(sc_cpu.scales), torch.logit(sc_cpu.opacities),
               sc_cpu.colors[:, :1].contiguous(), sc_cpu.colors[:, 1:].contiguous()]
        params = [t.to(dev).requires_grad_(True) for t in raw]
    # N > 1: the per-step densify-stat deltas live behind the gradients in the same (symmetric) buffer and travel in
    # the same exchange launch
    bucket = FlatGradBucket(params, stats_size=N if world > 1 else 0)
    stats = DensifyStats(N, dev, bucket=bucket if world > 1 else None)
    optimizer = None
    if args.train_step:
        from easy_gaussian_splatting_b200.optim import FusedAdam
        # the reference's learning rates (configs/*.yaml: means_lr_init 1e-3, log_scales 1e-2, quats 1e-3, sh_0 2.5e-3,
        # sh_rest 1.25e-4, logit_opacities 5e-2) and Adam's default eps, as build_optimizers sets them
        # (model/gaussian.py:389-412).  The objective of the training step is the reference's L1 + SSIM loss against
        # target images (below), which has a minimum, so the scene stays put while it is being timed.
        lrs = dict(means=1e-3, quats=1e-3, scales=1e-2, opacities=5e-2, colors=2.5e-3,
                   log_scales=1e-2, logit_opacities=5e-2, sh_0=2.5e-3, sh_rest=1.25e-4)
        optimizer = FusedAdam([{"params": [p_], "lr": lrs[k], "name": k} for k, p_ in zip(names, params)])
    # The step's V views are rendered by V / C calls of C views each (viewmats [C,4,4], the call's own batching):
    # every kernel of a call then works on C views' worth of tiles, which fills the 148 SMs where one view's
    # longest tiles would leave most of them idle.  C = 1 is the reference's one-view-per-call loop.
    C = 1 if args.forward_only else max(1, min(args.views_per_call, V))  # forward-only runs report per-frame latency
    if V % C:
        raise SystemExit("--views-per-call must divide the number of views per rank")
    n_calls = V // C
    bg = sc_cpu.background[None].expand(C, 3).contiguous().to(dev)
    Wc_cpu, Wa_cpu = loss_weights(sc_cpu.seed, C, H, W)  # one "target image" worth of loss weights per view
    # pinned host copies of the per-step inputs (cameras + loss weights = the target images of a training step)
    pin = lambda t: t.contiguous().pin_memory()
    host_views = [(pin(sc_cpu.viewmats[my_views[g * C:g * C + C]]), pin(sc_cpu.Ks[my_views[g * C:g * C + C]]))
                  for g in range(n_calls)]
    host_Wc, host_Wa = pin(Wc_cpu), pin(Wa_cpu)
    dev_views = [(a.to(dev), b.to(dev)) for a, b in host_views]
    dev_Wc, dev_Wa = host_Wc.to(dev), host_Wa.to(dev)
    h2d_bytes = n_calls * (host_views[0][0].numel() * 4 + host_views[0][1].numel() * 4 + host_Wc.numel() * 4 + host_Wa.numel() * 4)
    d2h_bytes = 4

    # View pipelining (easy_gaussian_splatting_b200/training.py): consecutive calls alternate between two CUDA
    # streams so that the forward pass of call i+1 overlaps the backward pass of call i.
    from easy_gaussian_splatting_b200.training import ViewPipeline
    pipelined = not args.no_view_pipelining and not args.forward_only and n_calls > 1
    pipe = ViewPipeline(dev, enabled=pipelined)
    view_streams = pipe.streams

    render = None
    if args.activations == "torch":
        def render(viewmat, K):  # what GaussianModel.forward does through its properties (gaussian.py:97-107,353-367)
            m, q, ls, lo, s0, sr = params
            return rasterization(m, q, torch.exp(ls), torch.sigmoid(lo), torch.cat([s0, sr], dim=1), viewmat, K, W, H,
                                 sh_degree=3, packed=False, absgrad=True, backgrounds=bg)
    elif args.activations == "folded":
        from easy_gaussian_splatting_b200 import rasterization_from_parameters

        def render(viewmat, K):
            m, q, ls, lo, s0, sr = params
            return rasterization_from_parameters(m, q, ls, lo, s0, sr, viewmat, K, W, H, 3, backgrounds=bg, absgrad=True)

    targets = None
    if args.train_step:
        # Target images of the training step: every view's own render of the initial scene, dimmed (0.9 x + 0.05), so
        # the L1 + SSIM objective has non-zero, coherent gradients and a minimum next to the initial scene.
        from easy_gaussian_splatting_b200.loss import fused_l1_ssim_loss
        targets = []
        with torch.no_grad():
            for vm, K_ in dev_views:
                act = params if args.activations == "pre" else None
                if act is None:
                    m, q, ls, lo, s0, sr = params
                    act = [m, q, torch.exp(ls), torch.sigmoid(lo), torch.cat([s0, sr], dim=1)]
                rc0 = rasterization(*[a.detach() for a in act], vm, K_, W, H, sh_degree=3, packed=False, backgrounds=bg)[0]
                targets.append(rc0.clamp_(0.0, 1.0).mul_(0.9).add_(0.05))
        inv_views = 1.0 / float(world * V)

    def loss_of(slot, Wc, Wa):
        if targets is None:
            return lambda rc, ra: linear_functional(rc, ra, Wc, Wa)
        tgt = targets[slot % n_calls]
        # gaussian.py:368 clamp, :422-445 LossComputer (fused kernels, csrc/loss.cu); mean over the step's views
        return lambda rc, ra: fused_l1_ssim_loss(torch.clamp(rc, 0.0, 1.0), tgt, None, 0.2)[0].sum() * inv_views

    def one_view(viewmat, K, Wc, Wa, want_loss, slot=0):
        if not args.forward_only:
            loss = pipe.render_backward(slot, params, viewmat, K, W, H, loss_of(slot, Wc, Wa),
                                        sh_degree=3, backgrounds=bg, absgrad=True, render=render,
                                        after_backward=lambda meta: stats.update_local(meta["radii"], meta["means2d"].absgrad, W, H))
            return loss if want_loss else None
        if args.forward_only:
            with torch.no_grad():
                rc, ra, meta = rasterization(params[0], params[1], params[2], params[3], params[4], viewmat, K, W, H,
                                             sh_degree=3, packed=False, absgrad=False, backgrounds=bg)
            return rc.sum() if want_loss else None

    fork_streams, join_streams = pipe.fork, pipe.join

    # one backward per step: gradients are written straight into the bucket (FlatGradBucket.begin_direct)
    direct = n_calls == 1 and not args.forward_only

    def step_resident(marks=None):
        """marks: a list that receives one CUDA event per phase boundary (start, rendered, exchanged, stepped)"""
        def mark():
            if marks is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                marks.append(ev)
        mark()
        if direct:
            bucket.begin_direct()
        else:
            bucket.zero_()
        stats.begin_step()
        fork_streams()
        for i, (vm, K) in enumerate(dev_views):
            one_view(vm, K, dev_Wc, dev_Wa, False, slot=i)
        join_streams()
        if direct:
            bucket.end_direct()
        mark()
        if world > 1 and not args.no_exchange:
            bucket.all_reduce()
            stats.all_reduce()
        mark()
        if optimizer is not None:
            optimizer.step()
        mark()

    # e2e: per-view inputs travel host -> device inside the timed region, double buffered on a copy stream so
    # the copy of view i+1 overlaps the rendering of view i; the step's loss is read back once per step.
    copy_stream = torch.cuda.Stream(device=dev)
    stage = [dict(vm=torch.empty_like(dev_views[0][0]), K=torch.empty_like(dev_views[0][1]),
                  Wc=torch.empty_like(dev_Wc), Wa=torch.empty_like(dev_Wa),
                  ready=torch.cuda.Event(), free=torch.cuda.Event()) for _ in range(2)]

    def stage_view(k):
        """k-th call since the start (call k % n_calls of its step) -> staging buffer k % 2, on the copy stream"""
        b = stage[k % 2]
        i = k % n_calls
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(b["free"])  # the view that used this buffer two views ago has finished
            b["vm"].copy_(host_views[i][0], non_blocking=True)
            b["K"].copy_(host_views[i][1], non_blocking=True)
            b["Wc"].copy_(host_Wc, non_blocking=True)
            b["Wa"].copy_(host_Wa, non_blocking=True)
            b["ready"].record(copy_stream)

    e2e_state = {"next": 0, "staged": -1}

    def step_e2e():
        if direct:
            bucket.begin_direct()
        else:
            bucket.zero_()
        stats.begin_step()
        main = torch.cuda.current_stream(dev)
        total = torch.zeros((), device=dev)
        k0 = e2e_state["next"]
        if e2e_state["staged"] < k0:  # very first step: nothing in flight yet
            for b in stage:
                b["free"].record(main)
            stage_view(k0)
            e2e_state["staged"] = k0
        fork_streams()
        losses = []
        for k in range(k0, k0 + n_calls):
            b = stage[k % 2]
            # the copy of the NEXT view (the first view of the next step included) overlaps this view's rendering
            stage_view(k + 1)
            e2e_state["staged"] = k + 1
            run_on = view_streams[k % 2] if pipelined else main
            run_on.wait_event(b["ready"])
            losses.append(one_view(b["vm"], b["K"], b["Wc"], b["Wa"], True, slot=k))
            b["free"].record(run_on)
        join_streams()
        if direct:
            bucket.end_direct()
        for l_ in losses:
            total += l_.detach()
        e2e_state["next"] = k0 + n_calls
        if world > 1:
            bucket.all_reduce()
            stats.all_reduce()
        if optimizer is not None:
            optimizer.step()
        return total.item()  # device -> host read of the step's result (train.py:107-108 reads the loss)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    lead = {}

    def lead_in_steps(fn):
        """Steps that keep the GPU busy for ~0.35 s; the SAME number on every rank (the step contains collectives)."""
        if fn not in lead:
            torch.cuda.synchronize()
            t0 = time.time()
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            n = max(3, int(math.ceil(0.35 / max((time.time() - t0) / 3, 1e-4))))
            if world > 1:
                t = torch.tensor([n], dtype=torch.int64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                n = int(t.item())
            lead[fn] = min(n, 2000)
        return lead[fn]

    def timed(fn, steps, warmup, sample_clocks=False):
        for _ in range(warmup):
            fn()
        barrier()
        sampler = ClockSampler(local_rank) if sample_clocks else None
        if sampler:
            # nvidia-smi needs ~0.1 s before its first line and prints one every 100 ms, while the timed region of a
            # default run lasts ~0.1 s: keep the GPU under the SAME load (untimed steps of the same function) until
            # the sampler has produced a line, then time; samples taken during the timed region and in the loaded
            # lead-in both count as "under load"
            sampler.start()
            for _ in range(lead_in_steps(fn)):
                fn()
            barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if sampler:
            for _ in range(lead_in_steps(fn) // 2):  # a short run: keep the load up for one more sample
                fn()
            torch.cuda.synchronize()
        clocks = sampler.stop() if sampler else None
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, clocks

    W_ = max(args.warmup, 3)
    for _ in range(W_):  # warm-up before the allocator snapshot (timed() warms up again: >= 3 untimed steps either way)
        step_resident()
    ms0 = torch.cuda.memory_stats(dev)
    ms_total, clocks = timed(step_resident, args.steps, W_, sample_clocks=True)
    ms1 = torch.cuda.memory_stats(dev)
    # kernels of this library per step: counted at the launch sites (egs_kernel_launch_count) over `steps` more
    # steps of the same function, outside the timed region so that reading the counter cannot perturb it
    lc0 = lib.egs_kernel_launch_count()
    for _ in range(args.steps):
        step_resident()
    torch.cuda.synchronize()
    own_launches = int(lib.egs_kernel_launch_count() - lc0)
    alloc_stats = {"cudaMalloc_calls_in_timed_region": ms1.get("num_device_alloc", 0) - ms0.get("num_device_alloc", 0),
                   "cudaFree_calls_in_timed_region": ms1.get("num_device_free", 0) - ms0.get("num_device_free", 0),
                   "alloc_retries": ms1.get("num_alloc_retries", 0),
                   "reserved_gb": ms1.get("reserved_bytes.all.peak", 0) / 1e9}
    # N > 1: where one step spends its time on every rank (CUDA events at the phase boundaries of `steps` more steps).
    # A rank that finishes rendering early waits for the slowest one INSIDE the exchange kernel, so its "exchange" is
    # the transfer plus that wait; the transfer itself is what the slowest renderer sees.
    step_phases = None
    if world > 1 and not args.no_exchange:
        acc = [0.0, 0.0, 0.0]
        for _ in range(args.steps):
            marks = []
            step_resident(marks)
            torch.cuda.synchronize()
            for i in range(3):
                acc[i] += marks[i].elapsed_time(marks[i + 1]) / args.steps
        mine = {"render_ms": acc[0], "exchange_ms": acc[1], "optimizer_ms": acc[2]}
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        slowest = max(range(world), key=lambda r: gathered[r]["render_ms"])
        step_phases = {"per_rank": gathered, "slowest_renderer": slowest,
                       "transfer_ms_seen_by_the_slowest_renderer": gathered[slowest]["exchange_ms"],
                       "what": "mean over the steps, each step synchronised at its end (so the phases of one step do not overlap the next)"}
    ms_step = ms_total / args.steps
    pixels_per_step = world * V * W * H
    value = pixels_per_step / (ms_step * 1e-3) / 1e6
    if args.quick:
        ms_e2e_total, e2e_value = float("nan"), None
    else:
        ms_e2e_total, _ = timed(step_e2e, args.steps, 2)
        e2e_value = pixels_per_step / (ms_e2e_total / args.steps * 1e-3) / 1e6

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W_,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(args), "views_per_rank_per_step": V, "views_per_step": world * V, "views_per_call": C,
                   "view_pipelining": pipelined, "activations": args.activations,
                   "l2": "inputs larger than L2 (236 MB parameters + 96 MB splat/gradient records per view vs 126 MB L2)",
                   "gradients": "written straight into the flat bucket by the fused backward (one backward per step)" if direct else "accumulated into the flat bucket by autograd",
                   "exchange": "none (1 GPU)" if world == 1 else f"{bucket.exchange} of the flat 236 B/Gaussian gradient bucket + {stats.exchange} of the 12 B/Gaussian densify stats each step"},
        "clocks": clocks, "allocator": alloc_stats,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                "ms_per_step": ms_e2e_total / args.steps,
                "what": "rasterization() fwd+bwd per call with its views' cameras and loss weights (one target-image-sized buffer per view) copied "
                        "from pinned host memory inside the timed region (double buffered on a copy stream) and the step's loss "
                        "read back with .item(); Gaussian parameters stay resident (they are the model, "
                        "/root/reference/train.py:97-108)"},
        "gpu_launches": own_launches,
        "gpu_launches_what": f"this library's kernels enqueued by {args.steps} steps, counted at the launch sites "
                             "(egs_kernel_launch_count); torch's own elementwise kernels of the loss functional are not included",
    }
    if step_phases is not None:
        line["step_phases"] = step_phases
    if args.quick:
        line.pop("e2e")
    if args.n_gaussians:
        line["invalid"] = "N overridden (debug run)"
    if args.no_exchange and world > 1:
        line["invalid"] = "exchange skipped (diagnostic run)"
    if args.train_step:
        line["scaling"] = "strong"
        line["train_step"] = {"it_per_s": 1e3 / ms_step, "views_per_step": world * V,
                              "optimizer": "fused Adam (csrc/adam.cu), the reference's learning rates and eps",
                              "loss": "the reference's L1 + SSIM (lambda 0.2) against per-view target images, fused kernels (csrc/loss.cu)",
                              "what": "fwd, clamp, L1+SSIM loss, bwd of every view, densify-stat update, gradient + densify-stat exchange, Adam step"}
    if args.forward_only:
        line["invalid"] = "forward-only latency run, not the fwd+bwd metric"
        line["config"]["mode"] = "no_grad forward only"

    if rank == 0 and not args.no_stage_timing and not args.quick and args.activations == "pre":
        line.update(stage_rooflines(lib, stages, params, dev_views[0], bg, dev_Wc, dev_Wa, W, H, dev))
    return line


def add_cpu_baseline(args, line):
    """rank 0, N = 1: the oracle on a bounded crop of the same workload, all host cores (run last: it leaves the
    host's thread pool busy, which slows the launch-bound parts of whatever is timed after it).  The crop the oracle
    renders is rendered by the CUDA path too and compared — `parity` in the line comes from this very run."""
    from easy_gaussian_splatting_b200 import rasterization
    from easy_gaussian_splatting_b200.synthetic import loss_weights
    from oracle import gsplat_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sc, sample = cpu_sample_scene(args.workload, args.ref_crop, args.n_gaussians)
    med, mpix, _ = run_cpu_oracle(sc, 3, 1)
    line["cpu_baseline"] = {"value": mpix, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                            "ms_per_sample": med * 1e3}
    # parity of the CUDA path with the oracle on that sample (same inputs, same loss functional)
    names = ("means", "quats", "scales", "opacities", "colors")
    Wc, Wa = loss_weights(sc.seed, 1, sc.height, sc.width)
    counters = {}
    lo = [getattr(sc, k).detach().clone().requires_grad_(True) for k in names]
    rc_o, ra_o, meta_o = O.rasterization(*lo, sc.viewmats, sc.Ks, sc.width, sc.height, sh_degree=3, packed=False,
                                         absgrad=True, backgrounds=sc.background[None], counters=counters)
    ((rc_o * Wc).sum() + (ra_o * Wa).sum()).backward()
    lg = [getattr(sc, k).detach().clone().cuda().requires_grad_(True) for k in names]
    rc, ra, meta = rasterization(*lg, sc.viewmats.cuda(), sc.Ks.cuda(), sc.width, sc.height, sh_degree=3, packed=False,
                                 absgrad=True, backgrounds=sc.background[None].cuda())
    ((rc * Wc.cuda()).sum() + (ra * Wa.cuda()).sum()).backward()
    border = counters["borderline"]
    err = torch.maximum((rc.detach().cpu() - rc_o.detach()).abs().amax(-1), (ra.detach().cpu() - ra_o.detach()).abs().amax(-1))
    rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
    grad_rel = {k: rel(g.grad.cpu(), o.grad) for k, g, o in zip(names, lg, lo)}
    grad_rel["absgrad"] = rel(meta["means2d"].absgrad.cpu(), meta_o["means2d"].absgrad)
    line["parity"] = {
        "against": "oracle/gsplat_oracle.py (CPU restatement of gsplat 1.0.0; parity UNPINNED: gsplat itself is not installable here)",
        "sample": sample,
        "bit_exact": {k: bool(torch.equal(meta[k].cpu(), meta_o[k])) for k in ("radii", "tiles_per_gauss", "isect_ids", "flatten_ids", "isect_offsets")},
        "max_clean": float(err[~border].max()) if (~border).any() else 0.0,
        "n_border": int(border.sum()), "max_border": float(err[border].max()) if border.any() else 0.0,
        "grad_rel": grad_rel, "grad_rel_max": max(grad_rel.values()),
        "tolerances": {"image_abs": 1e-4, "grad_rel": 1e-3},
    }


def cfg1_cpu_timing():
    """BASELINE config 1: 10 k Gaussians, one 256 x 256 camera, SH 3, fwd + bwd through the CPU
    oracle, 8 threads (or all cores if fewer), forward and backward timed separately, median of 5 after 1 warm-up."""
    from easy_gaussian_splatting_b200.synthetic import loss_weights, make_config_scene
    from oracle import gsplat_oracle as O
    threads = min(8, os.cpu_count() or 1)
    torch.set_num_threads(threads)
    sc = make_config_scene("cfg1")
    Wc, Wa = loss_weights(sc.seed, 1, sc.height, sc.width)
    tf, tb = [], []
    for it in range(6):
        leaves = [getattr(sc, k).detach().clone().requires_grad_(True) for k in ("means", "quats", "scales", "opacities", "colors")]
        t0 = time.perf_counter()
        rc, ra, _ = O.rasterization(*leaves, sc.viewmats, sc.Ks, sc.width, sc.height, sh_degree=3, packed=False,
                                    absgrad=True, backgrounds=sc.background[None])
        t1 = time.perf_counter()
        ((rc * Wc).sum() + (ra * Wa).sum()).backward()
        t2 = time.perf_counter()
        if it >= 1:
            tf.append(t1 - t0)
            tb.append(t2 - t1)
    f, b = statistics.median(tf), statistics.median(tb)
    return {"workload": "cfg1: 10000 Gaussians (blob, seed 0), 256x256, SH 3, CPU oracle", "threads": threads,
            "cpu_count": os.cpu_count(), "fwd_ms": f * 1e3, "bwd_ms": b * 1e3,
            "mpix_per_s": sc.width * sc.height / (f + b) / 1e6,
            "label": "CPU restatement of gsplat 1.0.0 semantics (gsplat itself unavailable)"}


def sequential_views(workload, dev, n_views=4, steps=5, warmup=3, forward_only=False):
    """The reference's own call pattern (model/gaussian.py:353-367 from train.py:97-104 / eval.py:41): ONE camera per
    rasterization() call, calls strictly one after the other on one stream — forward, loss functional, backward,
    densify-stat update, then the next view; fresh gradients per view as after optimizer.zero_grad().  No view
    batching, no cross-call stream overlap, no direct-to-bucket gradients."""
    from easy_gaussian_splatting_b200 import rasterization, stages
    from easy_gaussian_splatting_b200.distributed import DensifyStats
    from easy_gaussian_splatting_b200.synthetic import CONFIGS, loss_weights, make_config_scene
    sc = make_config_scene(workload, n_views=n_views)
    W, H, N = sc.width, sc.height, sc.means.shape[0]
    params = [getattr(sc, k).to(dev).requires_grad_(not forward_only) for k in ("means", "quats", "scales", "opacities", "colors")]
    vms, Ks = sc.viewmats.to(dev), sc.Ks.to(dev)
    bg = sc.background[None].to(dev)
    Wc, Wa = (t.to(dev) for t in loss_weights(sc.seed, 1, H, W))
    stats = DensifyStats(N, dev)
    frame_ms = []

    def view(v, timed_frames=False):
        if forward_only:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            with torch.no_grad():
                rc, _, _ = rasterization(*params, vms[v:v + 1], Ks[v:v + 1], W, H, sh_degree=3, packed=False, backgrounds=bg)
            e1.record()
            if timed_frames:
                frame_ms.append((e0, e1))
            return
        for p_ in params:
            p_.grad = None
        rc, ra, meta = rasterization(*params, vms[v:v + 1], Ks[v:v + 1], W, H, sh_degree=3, packed=False, absgrad=True, backgrounds=bg)
        linear_functional(rc, ra, Wc, Wa).backward()
        stats.update_local(meta["radii"], meta["means2d"].absgrad, W, H)

    for _ in range(warmup):
        for v in range(n_views):
            view(v)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        for v in range(n_views):
            view(v, timed_frames=True)
    e1.record()
    torch.cuda.synchronize()
    ms_view = e0.elapsed_time(e1) / (steps * n_views)
    c = CONFIGS[workload]
    out = {"workload": f"{workload}: {N} Gaussians ({c['kind']}), {W}x{H}", "views_per_call": 1, "views_timed": steps * n_views,
           "mode": "no_grad forward" if forward_only else "fwd+bwd+stats", "ms_per_view": ms_view,
           "mpix_per_s": W * H / (ms_view * 1e-3) / 1e6}
    if forward_only:
        lat = sorted(a.elapsed_time(b) for a, b in frame_ms)
        out["p50_frame_ms"] = lat[len(lat) // 2]
        out["p90_frame_ms"] = lat[(len(lat) * 9) // 10]
    del params, stats
    torch.cuda.empty_cache()
    return out


def call_pattern_runs(dev):
    """line['call_pattern']: BASELINE.json's configurations the way the reference calls the rasterizer."""
    out = {"what": "one camera per call, calls strictly sequential on one stream (the reference's loop); device-timed with "
                   "CUDA events over all timed views, inputs resident"}
    out["metric_c1"] = sequential_views("metric", dev)
    out["cfg2_c1"] = sequential_views("cfg2", dev, n_views=4, steps=8)
    out["cfg3_c1"] = sequential_views("cfg3", dev, n_views=2, steps=5)
    out["cfg5_fwd"] = sequential_views("cfg5", dev, n_views=6, steps=5, forward_only=True)
    return out


def exchange_check(dev, rank, world):
    """N > 1 pre-step (SURVEY.md section 4 '2-process NCCL test', B-3): the view-sharded step — real kernels, gradients
    written straight into the symmetric-memory bucket, our all-reduce kernels, densify-stat exchange, fused Adam — must
    give every rank the gradients and statistics of ONE GPU rendering all views, and leave the replicas bit-identical."""
    import torch.distributed as dist
    from easy_gaussian_splatting_b200 import rasterization
    from easy_gaussian_splatting_b200.distributed import DensifyStats, FlatGradBucket, shard_views
    from easy_gaussian_splatting_b200.optim import FusedAdam
    from easy_gaussian_splatting_b200.synthetic import loss_weights, make_scene
    N, W, H, per_rank = 120_001, 640, 368, 2
    V = per_rank * world
    sc = make_scene("outdoor", N, W, H, 554.0, 77, n_views=V)
    names = ("means", "quats", "scales", "opacities", "colors")
    lrs = dict(means=1e-3, quats=1e-3, scales=1e-2, opacities=5e-2, colors=2.5e-3)
    Wc, Wa = (t.to(dev) for t in loss_weights(sc.seed, V, H, W))
    vms, Ks = sc.viewmats.to(dev), sc.Ks.to(dev)

    def run(view_ids, distributed):
        params = [getattr(sc, k).to(dev).requires_grad_(True) for k in names]
        bucket = FlatGradBucket(params, symmetric=None if distributed else False, stats_size=N if distributed else 0)
        stats = DensifyStats(N, dev, bucket=bucket if distributed else None, distributed=distributed)
        opt = FusedAdam([{"params": [p_], "lr": lrs[k], "name": k} for k, p_ in zip(names, params)])
        idx = torch.tensor(view_ids, device=dev)
        bg = sc.background[None].expand(len(view_ids), 3).contiguous().to(dev)
        stats.begin_step()
        with bucket.direct():
            rc, ra, meta = rasterization(*params, vms[idx], Ks[idx], W, H, sh_degree=3, packed=False, absgrad=True, backgrounds=bg)
            ((rc * Wc[idx]).sum() + (ra * Wa[idx]).sum()).backward()
        stats.update_local(meta["radii"], meta["means2d"].absgrad, W, H)
        if distributed:
            bucket.all_reduce()
            stats.all_reduce()
        grads = [p_.grad.detach().clone() for p_ in params]
        opt.step()
        torch.cuda.synchronize()
        return grads, stats.buf.clone(), [p_.detach().clone() for p_ in params], bucket.exchange, stats.exchange

    g_d, s_d, p_d, ex_g, ex_s = run(shard_views(V, rank, world), True)
    g_1, s_1, p_1, _, _ = run(list(range(V)), False)
    rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
    grad_rel = {k: rel(a, b) for k, a, b in zip(names, g_d, g_1)}
    res = {
        "grad_rel_max": max(grad_rel.values()),
        "accum_rel": rel(s_d[0], s_1[0]),
        "counts_equal": float(torch.equal(s_d[1], s_1[1])),
        "max_radii_equal": float(torch.equal(s_d[2], s_1[2])),
        "param_rel_after_adam": max(rel(a, b) for a, b in zip(p_d, p_1)),
    }
    # replicas bit-identical after the step: compare every rank's parameter bits with rank 0's
    same = 1.0
    for p_ in p_d + [s_d]:
        ref = p_.clone()
        dist.broadcast(ref, src=0)
        same = min(same, float(torch.equal(ref, p_)))
    res["replicas_bit_identical"] = same
    # worst case over ranks
    keys = sorted(res)
    t = torch.tensor([res[k] if k.endswith(("equal", "identical")) else -res[k] for k in keys], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    out = {k: (bool(v) if k.endswith(("equal", "identical")) else -v) for k, v in zip(keys, t.tolist())}
    out.update({"scene": f"{N} Gaussians (outdoor), {W}x{H}, {V} views = {per_rank} per rank", "grad_exchange": ex_g, "stat_exchange": ex_s,
                "tolerances": {"grad_rel": 1e-5, "accum_rel": 1e-5}, "per_tensor_rank0": grad_rel})
    out["ok"] = bool(out["grad_rel_max"] <= 1e-5 and out["accum_rel"] <= 1e-5 and out["counts_equal"] and out["max_radii_equal"]
                     and out["replicas_bit_identical"])
    return out


def stage_rooflines(lib, stages, params, view, bg, Wc, Wa, W, H, dev, reps=10):
    """Times every kernel family in isolation with CUDA events on the launching stream and relates it
    to its algorithmic work (BASELINE.md §4 constants)."""
    import ctypes
    pk = peaks()
    means, quats, scales, opac, colors = [p.detach() for p in params]
    vm, K = view  # the C views of one call
    Cn = vm.shape[0]
    N = means.shape[0] * Cn  # (camera, Gaussian) pairs one launch works on
    tw, th = stages.tile_grid(W, H)

    def tm(fn, reps=reps, inner=4):
        """median over `reps` of (CUDA-event time of `inner` back-to-back calls) / inner — back to back so the
        host-side launch work of call i+1 overlaps the device work of call i, as it does inside a real step"""
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(inner):
                fn()
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1) / inner)
        return statistics.median(ts)

    # FP32 peak probe (dependent FMA chains, 8 per thread)
    out = torch.zeros(4, device=dev)
    flops = ctypes.c_double(0)
    st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    def probe():
        lib.egs_probe_fp32_fma(148 * 16, 4096, ctypes.c_void_p(out.data_ptr()), ctypes.byref(flops), st)
    probe_ms = tm(probe, 5)
    fp32_peak = flops.value / (probe_ms * 1e-3) / 1e12

    proj = stages.projection_fwd(means, quats, scales, opac, colors, vm, K, W, H, 3)
    n_vis = int((proj["radii"] > 0).sum())
    tpg, ids_u, flat_u = stages.isect_tiles(proj["means2d"], proj["radii"], proj["depths"], 16, tw, th, sort=False,
                                            tiles_per_gauss=proj["tiles_per_gauss"])
    n_isects = ids_u.numel()
    nbits = stages.tile_n_bits(tw, th) + stages.camera_n_bits(Cn)
    ids, flat = stages.radix_sort_pairs(ids_u.clone(), flat_u.clone(), 32 + nbits)
    offs = stages.isect_offset_encode(ids, Cn, tw, th)
    rc, ra, last, pairs = stages.rasterize_fwd(proj["splats"], offs, flat, bg, W, H, count_pairs=True)
    p_eval, p_acc = int(pairs[0]), int(pairs[1])
    v_splats = stages.rasterize_bwd(proj["splats"], offs, flat, bg, W, H, ra, last, Wc, Wa)

    # what rasterization() runs: the tight lists (a Gaussian only in the tiles it can reach with alpha >= 1/255), sized
    # from the previous call, tiles taken longest list first
    def product_binning():
        b_ = stages.isect_sorted_async(proj["means2d"], proj["radii"], proj["depths"], proj["tiles_per_gauss"], 16, tw, th,
                                       tight_rects=proj["tight_rects"])
        if not b_.resolve():
            b_ = stages.isect_sorted_async(proj["means2d"], proj["radii"], proj["depths"], proj["tiles_per_gauss"], 16, tw, th,
                                           capacity=b_.n_isects, tight_rects=proj["tight_rects"])
            b_.resolve()
        b_.note_for_next_call()
        return b_
    pb = product_binning()
    n_isects_tight = pb.n_isects
    _, ra_p, last_p = stages.rasterize_fwd(proj["splats"], pb.offsets, pb.flat_cap, bg, W, H, n_isects=pb.raster_n, tile_order=pb.tile_order)

    t = {}
    t["projection_sh_fwd"] = tm(lambda: stages.projection_fwd(means, quats, scales, opac, colors, vm, K, W, H, 3))
    # product path of g3-g5: two-level route on tight rectangles (its time includes reading the count back)
    t["binning_fast_path"] = tm(product_binning)
    t["rasterize_fwd"] = tm(lambda: stages.rasterize_fwd(proj["splats"], pb.offsets, pb.flat_cap, bg, W, H, n_isects=pb.raster_n,
                                                         tile_order=pb.tile_order))
    zero_ms = tm(lambda: torch.zeros_like(v_splats))
    t["rasterize_bwd"] = tm(lambda: stages.rasterize_bwd(proj["splats"], pb.offsets, pb.flat_cap, bg, W, H, ra_p, last_p, Wc, Wa,
                                                         n_isects=pb.raster_n, tile_order=pb.tile_order)) - zero_ms
    t["projection_sh_bwd"] = tm(lambda: stages.projection_bwd(means, quats, scales, colors, vm, K, W, H, 3, 0.3,
                                                               proj["radii"], proj["colors"], v_splats))
    # standalone C-ABI operators of the classic route (64-bit keys), not on the product path any more
    u = {}
    u["scan"] = tm(lambda: stages.exclusive_scan(proj["tiles_per_gauss"]))
    u["emit_u64"] = tm(lambda: stages.isect_tiles(proj["means2d"], proj["radii"], proj["depths"], 16, tw, th, sort=False,
                                                  tiles_per_gauss=proj["tiles_per_gauss"], n_isects=n_isects)) - u["scan"]
    ka, va = ids_u.clone(), flat_u.clone()

    def sort_once():
        ka.copy_(ids_u)
        va.copy_(flat_u)
        stages.radix_sort_pairs(ka, va, 32 + nbits)
    copy_ms = tm(lambda: (ka.copy_(ids_u), va.copy_(flat_u)))
    u["radix_sort_u64"] = tm(sort_once) - copy_ms
    u["offset_encode"] = tm(lambda: stages.isect_offset_encode(ids, Cn, tw, th))
    # §8f-4: fused L1 + SSIM loss (LossComputer drop-in) on one call's rendered views against random targets
    from easy_gaussian_splatting_b200.loss import fused_l1_ssim_loss
    if H >= 11 and W >= 11:
        img = rc.detach().clamp(0, 1).requires_grad_(True)
        gt_img = torch.rand_like(img)
        msk = (torch.rand(Cn, H, W, device=dev) < 0.1).float()
        u["l1_ssim_loss_fwd"] = tm(lambda: fused_l1_ssim_loss(img, gt_img, msk, 0.2))
        tot = fused_l1_ssim_loss(img, gt_img, msk, 0.2)[0].sum()
        u["l1_ssim_loss_bwd"] = tm(lambda: torch.autograd.grad(tot, img, retain_graph=True))
    passes = math.ceil((32 + nbits) / 8)
    sort_bytes = (8 + 24 * passes) * n_isects
    # The two-level route the product runs (stages.isect_sorted_async), stage by stage, per launch of Cn views:
    #   visible keys  : tiles_per_gauss read (4 B) per (camera, Gaussian); depth read (4) + level-1 pair written (key + 4)
    #                   per visible Gaussian — one launch, chained with decoupled look-back
    #   level-1 sort  : histogram read + p1 passes of read + write over the level-1 pairs
    #   scan + emit   : order (4) + packed tight rectangle (8) read per visible Gaussian, 8 B pair written per
    #                   intersection — one launch, the tile-count scan rides in it
    #   level-2 sort  : 4 B histogram read + p2 passes of 16 B per intersection
    #   tile offsets  : 4 B key read per intersection + 4 B per tile
    #   ("intersection" = an entry of the TIGHT lists the blend kernels work from; n_isects stays gsplat's count)
    key1 = stages.LEVEL1_KEY_BYTES
    p1 = math.ceil(stages.level1_end_bit(Cn) / 8)
    p2 = math.ceil(max(1, int(Cn * tw * th - 1).bit_length()) / 8)
    two_level_bytes = (4 * N + (4 + key1 + 4) * n_vis + (key1 + 2 * (key1 + 4) * p1) * n_vis + (4 + 8) * n_vis
                       + 8 * n_isects_tight + (4 + 16 * p2) * n_isects_tight + 4 * n_isects_tight + 4 * Cn * tw * th)
    work = {  # algorithmic bytes (HBM-bound stages) or flops (blend) per launch, BASELINE.md section 4 (frozen)
        "projection_sh_fwd": ("hbm", 68 * N + 204 * n_vis),
        "binning_fast_path": ("hbm", two_level_bytes),
        "rasterize_fwd": ("fp32", 16 * p_eval + 10 * p_acc),
        "rasterize_bwd": ("fp32", 16 * p_eval + 54 * p_acc),
        "projection_sh_bwd": ("hbm", 108 * N + 12 * N + 408 * n_vis + 192 * (N - n_vis)),
    }
    work_u = {
        "scan": ("hbm", 12 * N),
        "emit_u64": ("hbm", 32 * N + 12 * n_isects),
        "radix_sort_u64": ("hbm", sort_bytes),
        "offset_encode": ("hbm", 8 * n_isects + 4 * Cn * tw * th),
        # render 12 + target 12 + mask 4 read, 3 derivative maps (36) written | maps 36 + images 28 read, gradient 12 written
        "l1_ssim_loss_fwd": ("hbm", 64 * Cn * H * W),
        "l1_ssim_loss_bwd": ("hbm", 76 * Cn * H * W),
    }

    def line(ms, bound, w):
        ms = max(ms, 1e-6)
        if bound == "hbm":
            ach = w / (ms * 1e-3) / 1e9
            return {"ms": ms, "bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                    "frac": ach / pk["hbm_gbs"], "algorithmic_bytes": w}
        ach = w / (ms * 1e-3) / 1e12
        return {"ms": ms, "bound": "fp32", "achieved": ach, "peak": fp32_peak, "unit": "TFLOP/s",
                "frac": ach / fp32_peak, "algorithmic_flops": w}

    stages_out = {k: line(t[k], *work[k]) for k in work}
    classic_bytes = 44 * N + 12 * n_isects + sort_bytes + 8 * n_isects + 4 * Cn * tw * th
    stages_out["binning_fast_path"]["note"] = ("algorithmic bytes of the two-level route the product runs (level-1 sort of the visible "
                                               f"Gaussians on depth, {p1} passes; emission; level-2 sort on the (camera, tile) index, {p2} passes; "
                                               "offsets) — includes the route's one host sync when it has one")
    stages_out["binning_fast_path"]["classic_route"] = {
        "algorithmic_bytes": classic_bytes, "frac_if_it_were_the_work": classic_bytes / (max(t["binning_fast_path"], 1e-6) * 1e-3) / 1e9 / pk["hbm_gbs"],
        "what": "SURVEY.md section 8d's frozen figure for gsplat's route (emit 64-bit keys, 6-pass LSD sort, offset encode = 172 B/isect), for comparison only"}
    standalone = {k: line(u[k], *work_u[k]) for k in work_u if k in u}
    dominant = max(t, key=lambda k: t[k])
    roof = dict(stages_out[dominant])
    roof["kernel"] = dominant
    roof["traffic"] = None
    tj = ROOT / "profiles" / "traffic.json"
    if tj.exists():  # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture
        tk = json.loads(tj.read_text())
        for name, e in tk["kernels"].items():
            if name.startswith(dominant.replace("rasterize_bwd", "rasterize_bwd_kernel").replace("rasterize_fwd", "rasterize_fwd_kernel")):
                roof["traffic"] = e["dram_bytes_per_launch"]
                roof["traffic_source"] = tk["source"]
    roof["peak_source"] = pk["source"] if roof["bound"] == "hbm" else \
        f"measured live: dependent-FMA probe kernel, {fp32_peak:.1f} TFLOP/s (theoretical 148 SM x 128 lanes x 2 x 1.965 GHz = 74.4)"
    return {"roofline": roof, "stages": stages_out, "stages_standalone": standalone,
            "scene_stats": {"views_per_launch": Cn, "N": N // Cn, "N_vis": n_vis, "n_isects": n_isects, "isects_per_visible": n_isects / max(n_vis, 1),
                            "n_isects_tight": n_isects_tight,
                            "P_eval": p_eval, "P_acc": p_acc, "fp32_peak_tflops_measured": fp32_peak,
                            "sum_stage_ms": sum(t.values())}}


def main():
    args = parse_args()
    if args.impl == "reference":
        reference_arm(args)
        return
    import copy
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    line = ours(args)
    extras = args.workload == "metric" and not (args.train_step or args.forward_only or args.n_gaussians or args.quick)
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1 and extras and not args.no_exchange_check:
        # the view-sharded step against one GPU rendering every view (real kernels, real exchange)
        line["exchange_check"] = exchange_check(dev, rank, world)
    if extras and not args.no_train_step:
        # second half of BASELINE.json's metric ("train it/s at 1/2/4/8 GPU"): BASELINE config 4, the view-sharded
        # training step — 3 M Gaussians, 1920x1080, 64 views per step split across the ranks (strong scaling),
        # fwd + L1/SSIM loss + bwd of every view, gradient + densify-stat exchange, fused Adam at the reference's rates
        a2 = copy.copy(args)
        a2.train_step, a2.workload, a2.total_views, a2.views_per_call = True, "cfg4", 64, 8
        # raw parameters (log-scales, logit-opacities, sh_0 / sh_rest) as the reference holds and optimises them
        # (model/gaussian.py:32-54, 389-412), activations folded into the projection kernels (section 8f-2)
        # (10 timed steps after 8 warm-up steps: at 8 GPUs a step lasts 18 ms, and one cudaMalloc of the caching allocator
        #  still settling would otherwise move the figure by 2 %)
        a2.steps, a2.warmup, a2.no_stage_timing, a2.no_cpu_baseline, a2.activations, a2.quick = 10, 8, True, True, "folded", True
        l2 = ours(a2)
        line["train_step"] = dict(l2["train_step"], ms_per_step=l2["ms_per_step"], mpix_per_s=l2["value"], scaling="strong",
                                  steps=a2.steps, workload=l2["config"]["workload"], views_per_call=l2["config"]["views_per_call"],
                                  exchange=l2["config"]["exchange"], allocator=l2["allocator"], gpu_launches=l2["gpu_launches"])
        if "step_phases" in l2:
            line["train_step"]["step_phases"] = l2["step_phases"]
    if rank == 0:
        if world == 1 and extras and not args.no_call_pattern:
            line["call_pattern"] = call_pattern_runs(dev)
        if world == 1 and not args.no_cpu_baseline and not args.quick:
            add_cpu_baseline(args, line)
            if extras and not args.no_call_pattern:
                line["call_pattern"]["cfg1_cpu"] = cfg1_cpu_timing()
        print(json.dumps(line), flush=True)
    if dist.is_available() and dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
