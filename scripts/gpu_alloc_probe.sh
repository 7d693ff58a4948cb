#!/bin/bash
# do cudaMalloc calls inside the timed region of the cfg4 training step go away with more warm-up?
for WU in 5 15; do
  timeout 300 python bench.py --train-step --workload cfg4 --total-views 64 --views-per-call 8 --activations folded --quick --steps 5 --warmup $WU --no-call-pattern --no-exchange-check --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('warmup $WU:', round(d['ms_per_step'],2), 'ms/step', d['allocator'])"
done
