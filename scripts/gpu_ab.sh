#!/bin/bash
# GPU box: A/B of library variants (scripts/build_variant.py).  Usage: gpu_ab.sh base v1 v2 ...   (base = in-tree library)
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = "base" ]; then unset EGS_RASTER_LIB; else export EGS_RASTER_LIB=$PWD/easy_gaussian_splatting_b200/_C/variants/$v.so; fi
  python bench.py --steps ${STEPS:-10} --warmup 3 --no-cpu-baseline --no-train-step ${BENCH_ARGS:-} > gpurun_out/ab_$v.log 2>&1 || tail -5 gpurun_out/ab_$v.log
done
python scripts/show_bench.py $(for v in "$@"; do echo gpurun_out/ab_$v.log; done)
