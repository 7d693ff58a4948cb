#!/bin/bash
# Runs on the GPU box: the headline bench (with cpu baseline), the reference arm, and the other BASELINE configs.
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_metric.log 2>&1; echo "metric rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.log 2>&1; echo "reference rc=$?"
timeout 600 python bench.py --workload cfg2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_cfg2.log 2>&1; echo "cfg2 rc=$?"
timeout 600 python bench.py --workload cfg3 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_cfg3.log 2>&1; echo "cfg3 rc=$?"
timeout 600 python bench.py --workload cfg5 --steps 10 --warmup 3 --no-cpu-baseline --forward-only --no-stage-timing > gpurun_out/bench_cfg5_fwd.log 2>&1; echo "cfg5 rc=$?"
timeout 600 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline --views-per-rank 8 > gpurun_out/bench_cfg4_1gpu.log 2>&1; echo "cfg4 rc=$?"
for f in metric reference cfg2 cfg3 cfg5_fwd cfg4_1gpu; do python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_$f.log").read().strip().splitlines()[-1])
    print("$f", "value", round(d["value"],2), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],2), d.get("scene_stats",{}))
    if "stages" in d: print("   " + "  ".join(f"{k}={s['ms']:.3f}({s['frac']:.2f})" for k, s in d["stages"].items()))
    if "cpu_baseline" in d: print("   cpu", d["cpu_baseline"])
except Exception as e:
    print("$f parse failed", e); print(open("gpurun_out/bench_$f.log").read()[-1500:])
PY
done
timeout 600 python bench.py --views-per-call 1 --steps 20 --warmup 5 --no-cpu-baseline --no-train-step > gpurun_out/bench_metric_c1.log 2>&1; echo "metric C=1 rc=$?"
timeout 600 python bench.py --activations folded --steps 20 --warmup 5 --no-cpu-baseline --no-train-step > gpurun_out/bench_metric_folded.log 2>&1; echo "folded rc=$?"
timeout 600 python bench.py --activations torch --steps 20 --warmup 5 --no-cpu-baseline --no-train-step > gpurun_out/bench_metric_torch.log 2>&1; echo "torch rc=$?"
timeout 600 python bench.py --workload cfg4 --train-step --total-views 64 --views-per-call 8 --steps 5 --warmup 3 --no-cpu-baseline --no-stage-timing > gpurun_out/bench_cfg4_train.log 2>&1; echo "train rc=$?"
python scripts/show_bench.py gpurun_out/bench_metric_c1.log gpurun_out/bench_metric_folded.log gpurun_out/bench_metric_torch.log gpurun_out/bench_cfg4_train.log
