#!/bin/bash
# round 2, call i (4 GPUs): binary Shan-Chen with the interior-first step order
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29712 tools/bench_sc.py 2>/dev/null | grep '^{' > gpurun_out/r2i_sc_n4.json; cat gpurun_out/r2i_sc_n4.json
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29713 tools/bench_sc.py 2>/dev/null | grep '^{' > gpurun_out/r2i_sc_n2.json; cat gpurun_out/r2i_sc_n2.json
timeout 300 python -m pytest tests/test_gpu_multi_rank.py -q -k sc3d 2>&1 | tail -3
