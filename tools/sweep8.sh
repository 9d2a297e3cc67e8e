#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/sweep8_tests.txt
cat gpurun_out/sweep8_tests.txt
timeout 300 python tools/bench_configs.py C5 > gpurun_out/sweep8_sc.txt 2>&1; cat gpurun_out/sweep8_sc.txt
timeout 600 python tools/bench_output.py > gpurun_out/bench_output.txt 2>&1; tail -3 gpurun_out/bench_output.txt
