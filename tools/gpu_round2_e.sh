#!/bin/bash
# round 2, call e (1 GPU): multi-step v2, tile-shaped Shan-Chen kernel, block decompositions, plugin replay
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_step_replay.py tests/test_gpu_plugin_replay.py tests/test_gpu_block_decomposition.py tests/test_gpu_shan_chen.py tests/test_gpu_shan_chen_multi.py tests/test_gpu_multi_subdomain.py tests/test_gpu_checkpoint.py -q --maxfail=10 --durations=6 2>&1 | tail -30 > gpurun_out/r2e_tests.txt
tail -12 gpurun_out/r2e_tests.txt
timeout 300 python tools/bench_configs.py C1 C1b C5 > gpurun_out/r2e_bench_configs.txt 2>&1; cat gpurun_out/r2e_bench_configs.txt
SLF_SC_LINEAR=1 timeout 200 python tools/bench_configs.py C5 > gpurun_out/r2e_bench_c5_linear.txt 2>&1; cat gpurun_out/r2e_bench_c5_linear.txt
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; cut -c1-400 gpurun_out/r2e_bench.json
