#!/bin/bash
# GPU box: backward-segment sweep.  usage: gpu_seg.sh "<bench args>" seg...
mkdir -p gpurun_out
args="$1"; shift
for sgm in "$@"; do
  EGS_BWD_SEGMENT=$sgm python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stage-timing --no-train-step $args > gpurun_out/seg_$sgm.log 2>&1 || tail -3 gpurun_out/seg_$sgm.log
  python - <<PY
import json
d = json.loads(open("gpurun_out/seg_$sgm.log").read().strip().splitlines()[-1])
print("segment $sgm [$args]: value %.1f  ms/step %.3f  e2e %.1f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
PY
done
