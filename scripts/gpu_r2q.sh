#!/bin/bash
# Round 2, call Q (N GPUs): default bench line at N ranks on the end-of-round code
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2959$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2q_bench$N.log 2> gpurun_out/r2q_bench$N.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r2q_bench$N.err
python scripts/show_bench.py gpurun_out/r2q_bench$N.log | grep -v "standalone\|projection_sh\|roofline" | cut -c1-500
