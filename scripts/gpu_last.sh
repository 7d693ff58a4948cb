#!/bin/bash
timeout 200 python -m pytest tests/test_gpu_stages.py -m gpu -q -x -k "async or capacity or sentinel or tight or arbitrary" --timeout 180 -p no:cacheprovider 2>&1 | tail -2
timeout 120 python bench.py --quick --steps 20 --warmup 5 --no-call-pattern --no-exchange-check --no-train-step 2>/dev/null | python -c "
import json,sys
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('quick', round(d['value'],1), round(d['ms_per_step'],4), d['allocator'])"
