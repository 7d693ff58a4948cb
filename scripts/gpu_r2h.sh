#!/bin/bash
# Round 2, call H (4 GPUs): chunked backward test on one GPU, then the overlapped exchange against the plain one.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_rasterization.py -m gpu -q -p no:cacheprovider -x -k "chunked or direct" 2>&1 | tail -3
i=0
for env in "EGS_EXCHANGE_OVERLAP=1" "EGS_EXCHANGE_OVERLAP=0" "EGS_EXCHANGE_OVERLAP=1"; do
  i=$((i+1))
  env $env timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $((29570+i)) bench.py --gpus 4 --steps 20 --warmup 5 --quick > gpurun_out/r2h_$i.log 2>gpurun_out/r2h_$i.err || tail -5 gpurun_out/r2h_$i.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2h_$i.log").read().strip().splitlines()[-1])
print("$env", "value", round(d["value"],1), "ms/step", round(d["ms_per_step"],3))
PY
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29579 bench.py --gpus 4 --steps 10 --warmup 3 --no-train-step > gpurun_out/r2h_full.log 2>gpurun_out/r2h_full.err
python scripts/show_bench.py gpurun_out/r2h_full.log | grep -v "standalone\|projection_sh\|roofline" | cut -c1-600
