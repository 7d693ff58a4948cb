#!/bin/bash
# Round 2, call E (2 GPUs): view-sharded step against one GPU (exchange_check), default 2-GPU bench line, sanitizer on
# both all-reduce kernels.
mkdir -p gpurun_out profiles/sanitizer
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2e_bench2.log 2> gpurun_out/r2e_bench2.err
echo "bench rc=$?"; tail -c 600 gpurun_out/r2e_bench2.err
python scripts/show_bench.py gpurun_out/r2e_bench2.log | cut -c1-700
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 --no-python compute-sanitizer --tool memcheck python scripts/sanitize_small.py --collective > profiles/sanitizer/r2_memcheck_collective_2gpu.txt 2>&1
echo "collective memcheck rc=$?"; grep -h "ok=\|ERROR SUMMARY" profiles/sanitizer/r2_memcheck_collective_2gpu.txt | head
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 --no-python compute-sanitizer --tool racecheck python scripts/sanitize_small.py --collective > profiles/sanitizer/r2_racecheck_collective_2gpu.txt 2>&1
echo "collective racecheck rc=$?"; grep -h "ok=\|RACECHECK SUMMARY" profiles/sanitizer/r2_racecheck_collective_2gpu.txt | head
cp profiles/sanitizer/r2_*collective* gpurun_out/
