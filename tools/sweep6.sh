#!/bin/bash
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'sc_collide_fused' -s 10 -c 2 -o gpurun_out/prof_r1e_scfused -f python tools/bench_configs.py C5 > gpurun_out/ncu_r1e_sc.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 60 --csv --log-file gpurun_out/launches_r1e_sc.csv python tools/bench_configs.py C5 > gpurun_out/ncu_r1e_sc2.log 2>&1
tail -3 gpurun_out/ncu_r1e_sc.log
