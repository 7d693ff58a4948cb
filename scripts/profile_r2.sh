#!/bin/bash
# GPU box: ncu evidence for the round-2 state (4 views per call, bench default): launch list of one step + full capture
# of every kernel of the library in one timed step.
mkdir -p gpurun_out
TAG=${1:-r2k}
BENCH="python bench.py --steps 1 --warmup 1 --quick --no-train-step"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv $BENCH > gpurun_out/launches_$TAG.log 2>&1
echo "launch list rc=$?"
# steps before the timed one: 3 + 1 warm-up, the clock sampler's lead-in (count varies) -> capture from the END: take the
# last full step by skipping a generous number of launches and keeping 2 steps' worth
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"^(projection|visible|radix|gscan|scan_|isect|rasterize|densify|tile_len)" -s 108 -c 27 -f -o gpurun_out/prof_$TAG $BENCH > gpurun_out/prof_$TAG.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out/prof_$TAG.ncu-rep
