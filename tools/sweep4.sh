#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/sweep4.txt
timeout 600 python -m pytest tests/test_gpu_shan_chen.py tests/test_gpu_shan_chen_multi.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/sweep4_tests.txt
cat gpurun_out/sweep4_tests.txt
for m in 4 5 8; do
  echo "== sc fused minb $m" >> gpurun_out/sweep4.txt
  SLF_B200_LIB=$PWD/tools/bin/libslf_sc$m.so timeout 300 python tools/bench_configs.py C5 >> gpurun_out/sweep4.txt 2>&1
done
echo "== default (6)" >> gpurun_out/sweep4.txt
timeout 300 python tools/bench_configs.py C5 >> gpurun_out/sweep4.txt 2>&1
cat gpurun_out/sweep4.txt
