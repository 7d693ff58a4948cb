#!/bin/bash
# Round 2, call C: what the existing long-list knobs buy in the one-view-per-call pattern (decides the automatic policy)
mkdir -p gpurun_out
for cfg in "" "EGS_LONG_TILE_THRESHOLD=1408" "EGS_BWD_SEGMENT=512" "EGS_BWD_SEGMENT=1024" "EGS_LONG_TILE_THRESHOLD=2048"; do
  for wl in cfg2 metric; do
    echo "[$cfg] $wl: $(env $cfg python scripts/seq_views.py $wl 4 6 2>&1 | tail -1 | cut -c1-220)"
  done
done 2>&1 | tee gpurun_out/r2c_knobs.log
echo "[fwd] cfg2: $(python scripts/seq_views.py cfg2 4 6 fwd | tail -1 | cut -c1-250)" | tee -a gpurun_out/r2c_knobs.log
echo "[fwd LONG=1408] cfg2: $(EGS_LONG_TILE_THRESHOLD=1408 python scripts/seq_views.py cfg2 4 6 fwd | tail -1 | cut -c1-250)" | tee -a gpurun_out/r2c_knobs.log
