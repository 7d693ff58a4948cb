#!/bin/bash
# Round 2, call J: full GPU suite, default bench line, ncu evidence of the current state
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --timeout-method=thread -p no:cacheprovider -rf > gpurun_out/r2j_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r2j_tests.log
timeout 900 python bench.py > gpurun_out/r2j_bench.log 2> gpurun_out/r2j_bench.err
echo "bench rc=$?"; tail -c 500 gpurun_out/r2j_bench.err
python scripts/show_bench.py gpurun_out/r2j_bench.log 2>/dev/null | cut -c1-330
bash scripts/profile_r2.sh r2k
