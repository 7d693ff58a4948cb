#!/bin/bash
# A/B of library variants (scripts/build_variant.py): stage + rasterization tests on the default build, then the
# device-timed bench value and an ncu launch list per variant.   usage: gpu_ab.sh TAG variant [variant ...]
T=$1; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 900 python -m pytest tests/test_gpu_stages.py tests/test_gpu_rasterization.py -m gpu -q -x --timeout 600 --timeout-method=thread -p no:cacheprovider -rf > gpurun_out/${T}_tests.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/${T}_tests.log
for V in default "$@"; do
  if [ "$V" = default ]; then unset EGS_RASTER_LIB; else export EGS_RASTER_LIB=$PWD/easy_gaussian_splatting_b200/_C/variants/$V.so; fi
  timeout 600 python bench.py --quick --steps 30 --warmup 5 --no-call-pattern --no-exchange-check > gpurun_out/${T}_$V.log 2> gpurun_out/${T}_$V.err
  echo "$V bench rc=$?"; python -c "
import json,sys
for ln in open('gpurun_out/${T}_$V.log'):
    if ln.startswith('{'):
        d=json.loads(ln); print('   $V', round(d['value'],1), 'Mpix/s', round(d['ms_per_step'],4), 'ms/step')"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${T}_$V.csv python bench.py --steps 2 --warmup 3 --quick --no-call-pattern --no-exchange-check > gpurun_out/launches_${T}_$V.log 2>&1
  python scripts/show_launches.py gpurun_out/launches_${T}_$V.csv | grep -v "tile_\|scan_hist"
done
