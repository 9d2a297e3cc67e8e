#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/sweep3.txt
for m in 11 12; do
  echo "== minb $m" >> gpurun_out/sweep3.txt
  SLF_B200_LIB=$PWD/tools/bin/libslf_minb$m.so timeout 300 python tools/bench_configs.py C2 C2m C2q C3 >> gpurun_out/sweep3.txt 2>&1
done
echo "== default (mrt10 q27_8)" >> gpurun_out/sweep3.txt
timeout 300 python tools/bench_configs.py C2 C2m C2q C3 C3f >> gpurun_out/sweep3.txt 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k collide_bulk -s 30 -c 2 -o gpurun_out/prof_r1d_ab -f python tools/bench_configs.py C2ab > gpurun_out/ncu_r1d_ab.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'sc_' -s 30 -c 3 -o gpurun_out/prof_r1d_sc -f python tools/bench_configs.py C5 > gpurun_out/ncu_r1d_sc.log 2>&1
cat gpurun_out/sweep3.txt
