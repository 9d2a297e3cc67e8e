#!/bin/bash
# round 2, call d (1 GPU): new tests, per-config numbers, SC tile experiment, ncu of the Poiseuille kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_step_replay.py tests/test_gpu_plugin_replay.py tests/test_gpu_pbc_halo_kernels.py tests/test_gpu_parity.py -q --maxfail=10 --durations=8 2>&1 | tail -30 > gpurun_out/r2d_tests_a.txt
tail -4 gpurun_out/r2d_tests_a.txt
timeout 900 python -m pytest tests/test_gpu_full_size.py -q --maxfail=10 --durations=10 2>&1 | tail -30 > gpurun_out/r2d_tests_b.txt
tail -4 gpurun_out/r2d_tests_b.txt
timeout 600 python tools/bench_configs.py > gpurun_out/r2d_bench_configs.txt 2>&1; cat gpurun_out/r2d_bench_configs.txt
SLF_NO_MULTISTEP=1 timeout 200 python tools/bench_configs.py C1 > gpurun_out/r2d_bench_c1_graph.txt 2>&1; cat gpurun_out/r2d_bench_c1_graph.txt
timeout 200 tools/bin/exp_sc 256 > gpurun_out/r2d_exp_sc.txt 2>&1; tail -12 gpurun_out/r2d_exp_sc.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:collide_bulk_aa -s 30 -c 2 -o gpurun_out/prof_r2d_c3 -f python tools/bench_configs.py C3 > gpurun_out/r2d_ncu_c3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:collide_bulk_ab -s 30 -c 2 -o gpurun_out/prof_r2d_ab -f python tools/bench_configs.py C2ab > gpurun_out/r2d_ncu_ab.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
