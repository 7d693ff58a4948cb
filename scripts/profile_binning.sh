#!/bin/bash
# GPU box: ncu --set full of the binning kernels of one late step of the default bench (4 views per call).
TAG=${1:-r2v}
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --quick --no-call-pattern --no-exchange-check"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"^(visible|radix|isect|tile_)" -s 80 -c 16 -f -o gpurun_out/prof_$TAG $BENCH > gpurun_out/prof_$TAG.log 2>&1
echo "full capture rc=$?"; ls -la gpurun_out/prof_$TAG.ncu-rep
