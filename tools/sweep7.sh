#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/sweep7.txt
timeout 600 python -m pytest tests/test_gpu_shan_chen.py tests/test_gpu_shan_chen_multi.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/sweep7_tests.txt
cat gpurun_out/sweep7_tests.txt
echo "== two-phase fused, default (8 blocks/SM)" >> gpurun_out/sweep7.txt
timeout 300 python tools/bench_configs.py C5 >> gpurun_out/sweep7.txt 2>&1
for m in 6 10; do
  echo "== two-phase fused minb $m" >> gpurun_out/sweep7.txt
  SLF_B200_LIB=$PWD/tools/bin/libslf_sc$m.so timeout 300 python tools/bench_configs.py C5 >> gpurun_out/sweep7.txt 2>&1
done
cat gpurun_out/sweep7.txt
