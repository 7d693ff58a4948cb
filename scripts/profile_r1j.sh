#!/bin/bash
# GPU box: ncu evidence for the end-of-round state (4 views per call, bench default).
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-stage-timing --no-train-step"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1j.csv $BENCH > gpurun_out/launches_r1j.log 2>&1
echo "launch list rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"egs::" -s 96 -c 24 -f -o gpurun_out/prof_r1j $BENCH > gpurun_out/prof_r1j.log 2>&1
echo "full capture rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:l1_ssim -s 6 -c 8 -f -o gpurun_out/prof_r1j_loss python scripts/time_loss.py > gpurun_out/prof_r1j_loss.log 2>&1
echo "loss capture rc=$?"
ls -la gpurun_out/*.ncu-rep
