#!/bin/bash
# Runs on the GPU box (under gpurun): ncu launch list + one full capture of the hot kernels.
# Usage: scripts/profile_gpu.sh <tag> [kernel-regex]
set -u
TAG=${1:-r1}
KREGEX=${2:-"rasterize|onesweep|projection|histogram|emit"}
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 1 --views-per-rank 1 --no-cpu-baseline --no-stage-timing"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv $BENCH > gpurun_out/launches_${TAG}.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"${KREGEX}" -s ${SKIP:-17} -c ${COUNT:-14} -f -o gpurun_out/prof_${TAG} $BENCH > gpurun_out/prof_${TAG}.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out/
