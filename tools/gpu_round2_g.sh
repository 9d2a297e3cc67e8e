#!/bin/bash
# round 2, call g (8 GPUs): one rank per block on real peer links, N=8 bench (weak + strong + parity), multi-rank tests,
# binary Shan-Chen on 1 / 4 GPUs
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/r2g_ngpu.txt
for ap in AA AB; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 2990$((RANDOM % 9)) tests/_gpu_block_worker.py $ap 2,2,2 2>/dev/null | grep RESULT
done > gpurun_out/r2g_blocks.txt 2>&1
cat gpurun_out/r2g_blocks.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2g_bench_n8.json 2> gpurun_out/r2g_bench_n8.err
echo "bench rc=$?"; grep '^{' gpurun_out/r2g_bench_n8.json | cut -c1-300; tail -3 gpurun_out/r2g_bench_n8.err
timeout 300 python -m pytest tests/test_gpu_multi_rank.py -q -x 2>&1 | tail -4 > gpurun_out/r2g_multi_rank.txt; cat gpurun_out/r2g_multi_rank.txt
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29612 tools/bench_sc.py 2>/dev/null | grep '^{' > gpurun_out/r2g_sc_n4.json; cat gpurun_out/r2g_sc_n4.json
timeout 200 python tools/bench_sc.py 2>/dev/null | grep '^{' > gpurun_out/r2g_sc_n1.json; cat gpurun_out/r2g_sc_n1.json
