#!/bin/bash
# round 2, call j (1 GPU): launch list + full capture of the bench kernel (profiles/), AB cache-operator variants,
# Shan-Chen launch list, the new semaphore tests, smoke
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_semaphores.py tests/test_gpu_zoo_node_types.py -q 2>&1 | tail -3
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 120 tools/bin/exp_ab 256 > gpurun_out/r2j_exp_ab.txt 2>&1; cat gpurun_out/r2j_exp_ab.txt
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r2j.csv python bench.py --steps 6 --warmup 3 --no_cpu_baseline --no_parity > gpurun_out/r2j_ncu_a.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:collide_bulk_aa_tab -s 8 -c 2 -o gpurun_out/prof_bulk_r2j -f python bench.py --steps 6 --warmup 3 --no_cpu_baseline --no_parity --no_e2e > gpurun_out/r2j_ncu_b.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 200 -c 120 --csv --log-file gpurun_out/launches_r2j_sc.csv python tools/bench_configs.py C5 > gpurun_out/r2j_ncu_c.log 2>&1
ls -la gpurun_out/launches_r2j*.csv gpurun_out/prof_bulk_r2j.ncu-rep
