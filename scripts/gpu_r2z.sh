#!/bin/bash
# Round 2, call Z: histogram scan folded into the histogram kernel, tile schedule in two launches.
T=${1:-r2z}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stages.py tests/test_gpu_rasterization.py -m gpu -q -x --timeout 600 --timeout-method=thread -p no:cacheprovider -rf > gpurun_out/${T}_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/${T}_tests.log
timeout 600 python bench.py --quick --steps 30 --warmup 5 --no-call-pattern --no-exchange-check > gpurun_out/${T}_quick.log 2> gpurun_out/${T}_quick.err
python -c "
import json
for ln in open('gpurun_out/${T}_quick.log'):
    if ln.startswith('{'):
        d=json.loads(ln); print('   quick', round(d['value'],1), 'Mpix/s', round(d['ms_per_step'],4), 'ms/step', d.get('gpu_launches'))"
for W in metric cfg2; do python scripts/c1_breakdown.py $W 2>&1 | tail -1; done
