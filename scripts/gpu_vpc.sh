#!/bin/bash
# view-batching sweep: bench.py at several --views-per-call values and workloads
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/vpc_$tag.log 2>&1; echo "[$tag] rc=$?";
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/vpc_$tag.log").read().strip().splitlines()[-1])
    print("  value", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), d.get("train_step",{}).get("it_per_s"))
    if "stages" in d: print("   " + "  ".join(f"{k}={s['ms']:.3f}({s['frac']:.2f})" for k, s in d["stages"].items()))
except Exception as e:
    print("  parse failed", e); print(open("gpurun_out/vpc_$tag.log").read()[-1500:])
PY
}
run metric_c4
run metric_c1 --views-per-call 1
run metric_c2 --views-per-call 2
run metric_v8c8 --views-per-rank 8 --views-per-call 8
run cfg2_c4 --workload cfg2
run cfg3_c4 --workload cfg3
run cfg4_c4 --workload cfg4 --views-per-rank 8
run cfg4_train_c8 --workload cfg4 --train-step --total-views 64 --views-per-call 8 --steps 3 --no-stage-timing
run metric_folded_c4 --activations folded
