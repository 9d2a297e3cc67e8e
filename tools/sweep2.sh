#!/bin/bash
# register-cap sweep for the D3Q19-MRT / D3Q27 AA bulk kernels + ncu captures of the non-headline kernels
mkdir -p gpurun_out
for m in 6 8 10; do
  echo "== minb $m" >> gpurun_out/sweep2.txt
  SLF_B200_LIB=$PWD/tools/bin/libslf_minb$m.so timeout 300 python tools/bench_configs.py C2m C2q C3 >> gpurun_out/sweep2.txt 2>&1
done
echo "== default" >> gpurun_out/sweep2.txt
timeout 300 python tools/bench_configs.py C2m C2q C3 C2ab >> gpurun_out/sweep2.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'collide_bulk|sc_collide|sc_prepare' -s 30 -c 6 -o gpurun_out/prof_r1d_variants -f python tools/bench_configs.py C2m C2ab C5 > gpurun_out/ncu_r1d.log 2>&1
tail -30 gpurun_out/sweep2.txt
