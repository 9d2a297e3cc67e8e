#!/bin/bash
# round-1 final profile set: bench line, launch list and full capture of the top kernel, output path cost
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_r1f.json 2> gpurun_out/bench_r1f.err; tail -1 gpurun_out/bench_r1f.json
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1f.csv python bench.py --steps 6 --warmup 3 > gpurun_out/ncu_r1f_a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:collide_bulk -s 8 -c 2 -o gpurun_out/prof_bulk_r1f -f python bench.py --steps 6 --warmup 3 > gpurun_out/ncu_r1f_b.log 2>&1
timeout 600 python tools/bench_output.py > gpurun_out/bench_output.txt 2>&1; tail -2 gpurun_out/bench_output.txt
