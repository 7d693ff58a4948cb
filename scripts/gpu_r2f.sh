#!/bin/bash
# Round 2, call F (4 GPUs): the NVLS (multimem) exchange with its MAX tail, exchange_check and the default line at 4 ranks.
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 scripts/sanitize_small.py --collective 2>&1 | grep -v "OMP\|\*\*\*\*" | tail -6
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2f_bench4.log 2> gpurun_out/r2f_bench4.err
echo "bench rc=$?"; tail -c 400 gpurun_out/r2f_bench4.err
python scripts/show_bench.py gpurun_out/r2f_bench4.log | cut -c1-800 | grep -v "standalone\|projection_sh\|roofline"
