#!/bin/bash
# round 2, call l (2 GPUs): final code -- multi-rank / block tests on real peers, N=2 bench with parity
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi_rank.py tests/test_gpu_block_decomposition.py tests/test_gpu_shan_chen_multi.py -q 2>&1 | tail -4 > gpurun_out/r2l_tests.txt; cat gpurun_out/r2l_tests.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29741 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2l_bench_n2.json 2> gpurun_out/r2l_bench_n2.err
echo "bench rc=$?"; grep '^{' gpurun_out/r2l_bench_n2.json | cut -c1-200
