#!/bin/bash
# round 2, call h (1 GPU): the whole GPU suite (timing!), smoke, segment-list kernel on the pipe, bench
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 --durations=12 ) 2>&1 | tail -40 > gpurun_out/r2h_tests.txt
tail -8 gpurun_out/r2h_tests.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 200 python tools/bench_configs.py C3 C3f > gpurun_out/r2h_bench_c3_seg.txt 2>&1; cat gpurun_out/r2h_bench_c3_seg.txt
SLF_NO_SEGMENTS=1 timeout 200 python tools/bench_configs.py C3 C3f > gpurun_out/r2h_bench_c3_dense.txt 2>&1; cat gpurun_out/r2h_bench_c3_dense.txt
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; cut -c1-300 gpurun_out/r2h_bench.json
