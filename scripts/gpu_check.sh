#!/bin/bash
# Runs on the GPU box: full GPU test-suite, then bench for each blend variant given in $VARIANTS.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 200 --timeout-method=thread -p no:cacheprovider -x > gpurun_out/tests.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/tests.log
for v in ${VARIANTS:-22}; do
  EGS_BLEND_VARIANT=$v timeout 300 python bench.py --steps ${STEPS:-10} --warmup 3 --no-cpu-baseline > gpurun_out/bench_$v.log 2>&1
  echo "bench variant $v rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_$v.log").read().strip().splitlines()[-1])
    print("  value", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1))
    for k, s in d["stages"].items(): print("   ", k, round(s["ms"],4), "ms frac", round(s["frac"],3))
except Exception as e:
    print("  parse failed", e); print(open("gpurun_out/bench_$v.log").read()[-2000:])
PY
done
