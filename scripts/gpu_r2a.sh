#!/bin/bash
# Round 2, call A: full GPU test-suite (new parity / swap / alignment tests included) and the default bench line.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --timeout-method=thread -p no:cacheprovider -rf --durations=15 > gpurun_out/r2a_tests.log 2>&1
echo "tests rc=$?"; tail -30 gpurun_out/r2a_tests.log
timeout 900 python bench.py > gpurun_out/r2a_bench.log 2> gpurun_out/r2a_bench.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/r2a_bench.err
python scripts/show_bench.py gpurun_out/r2a_bench.log 2>/dev/null | head -60
