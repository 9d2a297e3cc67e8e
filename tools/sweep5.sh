#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/sweep5.txt
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/sweep5_tests.txt
cat gpurun_out/sweep5_tests.txt
echo "== address tables OFF (previous kernels)" >> gpurun_out/sweep5.txt
SLF_B200_LIB=$PWD/tools/bin/libslf_notab.so timeout 300 python tools/bench_configs.py C1b C2 C2m C2q C3 C3f >> gpurun_out/sweep5.txt 2>&1
echo "== address tables ON (default build)" >> gpurun_out/sweep5.txt
timeout 300 python tools/bench_configs.py C1b C2 C2m C2q C3 C3f C5 >> gpurun_out/sweep5.txt 2>&1
for m in 4 5 8 10; do
  echo "== sc fused minb $m" >> gpurun_out/sweep5.txt
  SLF_B200_LIB=$PWD/tools/bin/libslf_sc$m.so timeout 300 python tools/bench_configs.py C5 >> gpurun_out/sweep5.txt 2>&1
done
cat gpurun_out/sweep5.txt
