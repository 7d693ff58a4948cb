#!/bin/bash
# GPU box: full GPU suite (all failures reported), then the default bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -p no:cacheprovider --durations=8 > gpurun_out/tests.log 2>&1
echo "tests rc=$?"; tail -25 gpurun_out/tests.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_default.log 2>&1
echo "bench rc=$?"; tail -1 gpurun_out/bench_default.log | cut -c1-1500
