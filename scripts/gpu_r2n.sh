#!/bin/bash
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 1 --quick --no-train-step"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2n.csv $BENCH > gpurun_out/launches_r2n.log 2>&1
echo "launch list rc=$?"
