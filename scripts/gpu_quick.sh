#!/bin/bash
# quick loop on the GPU box: blend/stage tests, then bench at one view per call and at the default batching
mkdir -p gpurun_out
if [ -z "${SKIP_TESTS:-}" ]; then
python -m pytest tests/test_gpu_rasterization.py tests/test_gpu_stages.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -3
fi
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --views-per-call 1 ${BENCH_ARGS:-} > gpurun_out/b1.log 2>&1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline ${BENCH_ARGS:-} > gpurun_out/b4.log 2>&1
python scripts/show_bench.py gpurun_out/b1.log gpurun_out/b4.log
