#!/bin/bash
# Round 2, call B: parity of the quarter-list blend kernels, A/B against the round-1 kernels, launch list of the
# one-view-per-call pattern on cfg2, train step with the reference's learning rates on raw parameters, ncu of the blend kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stages.py tests/test_gpu_rasterization.py tests/test_gpu_fullsize.py -m gpu -q --timeout 600 -p no:cacheprovider -x -k "rasterize or parity or segment or window or golden or long_tile or isect or sort" > gpurun_out/r2b_tests.log 2>&1
echo "tests rc=$?"; tail -8 gpurun_out/r2b_tests.log
STEPS=10 BENCH_ARGS="--no-call-pattern" bash scripts/gpu_ab.sh base bwd12 r1blend sort12 sort16
timeout 300 python bench.py --train-step --workload cfg4 --total-views 64 --views-per-call 8 --activations folded --quick --steps 3 --warmup 3 > gpurun_out/r2b_train.log 2>&1; tail -c 900 gpurun_out/r2b_train.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 700 -c 140 --csv --log-file gpurun_out/r2b_c1_cfg2.csv python bench.py --workload cfg2 --views-per-call 1 --no-view-pipelining --quick --steps 2 --warmup 3 > gpurun_out/r2b_c1_cfg2.log 2>&1
tail -c 300 gpurun_out/r2b_c1_cfg2.log
BENCH="python bench.py --steps 1 --warmup 1 --quick --no-train-step"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"rasterize_" -s 8 -c 2 -f -o gpurun_out/prof_r2b $BENCH > gpurun_out/prof_r2b.log 2>&1
echo "full capture rc=$?"
EGS_RASTER_LIB=$PWD/easy_gaussian_splatting_b200/_C/variants/r1blend.so timeout 900 ncu --set full --clock-control none --import-source on -k regex:"rasterize_" -s 8 -c 2 -f -o gpurun_out/prof_r2b_r1 $BENCH > gpurun_out/prof_r2b_r1.log 2>&1
ls -la gpurun_out/*.ncu-rep
