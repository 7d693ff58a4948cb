#!/bin/bash
# Round 2, call D: full GPU suite on the sync-free binning route, default bench line, sort-tile A/B, sanitizer passes.
mkdir -p gpurun_out profiles/sanitizer
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --timeout-method=thread -p no:cacheprovider -rf > gpurun_out/r2d_tests.log 2>&1
echo "tests rc=$?"; tail -12 gpurun_out/r2d_tests.log
timeout 900 python bench.py > gpurun_out/r2d_bench.log 2> gpurun_out/r2d_bench.err
echo "bench rc=$?"; tail -c 800 gpurun_out/r2d_bench.err
python scripts/show_bench.py gpurun_out/r2d_bench.log 2>/dev/null | cut -c1-420
STEPS=10 BENCH_ARGS="--no-call-pattern" bash scripts/gpu_ab.sh sort12 2>&1 | head -3 | cut -c1-300
for tool in memcheck racecheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool python scripts/sanitize_small.py > profiles/sanitizer/r2_$tool.txt 2>&1
  echo "$tool rc=$? $(grep -c '^ok' profiles/sanitizer/r2_$tool.txt) ok-lines; $(grep 'ERROR SUMMARY' profiles/sanitizer/r2_$tool.txt)"
done
CUDA_MODULE_LOADING=EAGER timeout 900 compute-sanitizer --tool memcheck python scripts/sanitize_small.py > profiles/sanitizer/r2_memcheck_eager_loading.txt 2>&1
echo "memcheck eager: $(grep 'ERROR SUMMARY' profiles/sanitizer/r2_memcheck_eager_loading.txt)"
cp profiles/sanitizer/r2_*.txt gpurun_out/
