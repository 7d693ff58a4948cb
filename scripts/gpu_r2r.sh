#!/bin/bash
# Round 2, call R: single-launch visible keys and scan+emit (decoupled look-back), packed tight rectangles:
# GPU test-suite, default bench line, launch list of the same command.
T=${1:-r2r}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 --timeout-method=thread -p no:cacheprovider -rf --durations=8 > gpurun_out/${T}_tests.log 2>&1
echo "tests rc=$?"; tail -25 gpurun_out/${T}_tests.log
timeout 900 python bench.py > gpurun_out/${T}_bench.log 2> gpurun_out/${T}_bench.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/${T}_bench.err
python scripts/show_bench.py gpurun_out/${T}_bench.log 2>/dev/null | head -40 | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${T}.csv python bench.py --steps 2 --warmup 3 --quick --no-call-pattern --no-exchange-check > gpurun_out/launches_${T}.log 2>&1
echo "ncu rc=$?"
