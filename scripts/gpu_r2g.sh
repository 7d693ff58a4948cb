#!/bin/bash
# Round 2, call G (4 GPUs): where the 4-GPU step time goes — sharding order x exchange on/off
mkdir -p gpurun_out
i=0
for args in "--shard strided" "--shard contiguous" "--shard strided --no-exchange" "--shard contiguous --no-exchange"; do
  i=$((i+1))
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $((29560+i)) bench.py --gpus 4 --steps 20 --warmup 5 --quick $args > gpurun_out/r2g_$i.log 2>/dev/null
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2g_$i.log").read().strip().splitlines()[-1])
print("$args", "value", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), d["config"]["exchange"][:60])
PY
done
