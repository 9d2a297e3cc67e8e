#!/bin/bash
# which block layouts fail with one rank per block (AA)?
mkdir -p gpurun_out
port=29800
for cfg in "2 2,1,1" "2 1,2,1" "2 1,1,2" "4 2,2,1" "4 2,1,2" "4 1,2,2" "8 2,2,2"; do
  set -- $cfg
  port=$((port+1))
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$1 --master-addr 127.0.0.1 --master-port $port tests/_gpu_block_worker.py AA $2 2>/dev/null | grep RESULT
done > gpurun_out/r2f_blocks.txt 2>&1
cat gpurun_out/r2f_blocks.txt
