#!/bin/bash
# Runs on the GPU box: full GPU test-suite, then bench for each "ENV=VAL,ENV=VAL" configuration in $CONFIGS
mkdir -p gpurun_out
if [ -z "${SKIP_TESTS:-}" ]; then
timeout 900 python -m pytest tests -m gpu -q --timeout 200 --timeout-method=thread -p no:cacheprovider -x > gpurun_out/tests.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/tests.log
fi
i=0
for cfg in ${CONFIGS:-default}; do
  i=$((i+1))
  envs=$(echo "$cfg" | tr ',' ' ')
  [ "$cfg" = "default" ] && envs=""
  env $envs timeout 300 python bench.py --steps ${STEPS:-10} --warmup 3 --no-cpu-baseline ${BENCH_ARGS:-} > gpurun_out/bench_$i.log 2>&1
  echo "bench [$cfg] rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_$i.log").read().strip().splitlines()[-1])
    print("  value", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1))
    print("   " + "  ".join(f"{k}={s['ms']:.3f}({s['frac']:.2f})" for k, s in d["stages"].items()))
except Exception as e:
    print("  parse failed", e); print(open("gpurun_out/bench_$i.log").read()[-2000:])
PY
done
