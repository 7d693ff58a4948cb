#!/bin/bash
# Round 2, call I (4 GPUs): tuning of the NVLS all-reduce kernel (floats in flight per thread, grid size)
i=0
for v in base ar_u1c8 ar_u4 ar_u8 ar_u4c4 ar_u4c16; do
  i=$((i+1))
  if [ "$v" = "base" ]; then unset EGS_RASTER_LIB; else export EGS_RASTER_LIB=$PWD/easy_gaussian_splatting_b200/_C/variants/$v.so; fi
  timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $((29580+i)) scripts/time_exchange.py 2>/dev/null | tail -1
done
