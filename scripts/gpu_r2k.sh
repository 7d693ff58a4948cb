#!/bin/bash
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 1 --quick --no-train-step"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"^(projection|visible|radix|gscan|scan_|isect|rasterize|densify|tile_len)" -s 108 -c 27 -f -o gpurun_out/prof_r2k $BENCH > gpurun_out/prof_r2k.log 2>&1
echo "full capture rc=$?"; ls -la gpurun_out/prof_r2k.ncu-rep
