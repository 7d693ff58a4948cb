#!/bin/bash
# Round 2, call L: longest-first tile order against grid order (batched step and one view per call), stage tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stages.py tests/test_gpu_rasterization.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -3
for e in 1 0 1 0; do
  EGS_TILE_ORDER=$e timeout 300 python bench.py --steps 20 --warmup 5 --quick --no-train-step > gpurun_out/r2l_order$e.log 2>/dev/null
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2l_order$e.log").read().strip().splitlines()[-1])
print("EGS_TILE_ORDER=$e batched: value", round(d["value"],1), "ms/step", round(d["ms_per_step"],3))
PY
done
for e in 1 0; do for wl in metric cfg2 cfg3; do
  echo "EGS_TILE_ORDER=$e $wl C=1: $(EGS_TILE_ORDER=$e python scripts/seq_views.py $wl 4 6 2>&1 | tail -1 | cut -c90-200)"
done; done
echo "EGS_TILE_ORDER=1 cfg5 fwd: $(EGS_TILE_ORDER=1 python scripts/seq_views.py cfg5 6 5 fwd | tail -1 | cut -c100-260)"
echo "EGS_TILE_ORDER=0 cfg5 fwd: $(EGS_TILE_ORDER=0 python scripts/seq_views.py cfg5 6 5 fwd | tail -1 | cut -c100-260)"
