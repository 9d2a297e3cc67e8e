#!/bin/bash
# round 2, call b (2 GPUs): multi-rank parity on real peer links, N=2 bench (weak + strong + parity), kernel experiments
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2b_gpus.txt
timeout 900 python -m pytest tests/test_gpu_multi_rank.py tests/test_gpu_shan_chen_multi.py tests/test_gpu_multi_subdomain.py -q -x 2>&1 | tail -15 > gpurun_out/r2b_multi_tests.txt
tail -3 gpurun_out/r2b_multi_tests.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2b_bench_n2.json 2> gpurun_out/r2b_bench_n2.err
echo "bench rc=$?"; cut -c1-600 gpurun_out/r2b_bench_n2.json; tail -5 gpurun_out/r2b_bench_n2.err
timeout 120 tools/bin/exp_ab 256 > gpurun_out/r2b_exp_ab.txt 2>&1; cat gpurun_out/r2b_exp_ab.txt
timeout 120 tools/bin/exp_sc 256 > gpurun_out/r2b_exp_sc.txt 2>&1; cat gpurun_out/r2b_exp_sc.txt
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 --deselect tests/test_gpu_multi_rank.py --deselect tests/test_gpu_shan_chen_multi.py --deselect tests/test_gpu_multi_subdomain.py --deselect tests/test_gpu_physics.py --deselect tests/test_gpu_parity.py 2>&1 | tail -25 > gpurun_out/r2b_tests.txt
tail -4 gpurun_out/r2b_tests.txt
