#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --timeout-method=thread -p no:cacheprovider -rf -x > gpurun_out/r2o_tests.log 2>&1
echo "tests rc=$?"; tail -6 gpurun_out/r2o_tests.log | cut -c1-250
timeout 900 python bench.py > gpurun_out/r2o_bench.log 2> gpurun_out/r2o_bench.err
echo "bench rc=$?"; tail -c 600 gpurun_out/r2o_bench.err
python scripts/show_bench.py gpurun_out/r2o_bench.log 2>/dev/null | cut -c1-330
BENCH="python bench.py --steps 1 --warmup 1 --quick --no-train-step"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2o.csv $BENCH > gpurun_out/launches_r2o.log 2>&1
