#!/bin/bash
# round 2, call m (1 GPU): the final code -- whole GPU suite, smoke, bench
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 --durations=8 ) 2>&1 | tail -24 > gpurun_out/r2m_tests.txt
tail -7 gpurun_out/r2m_tests.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; cut -c1-260 gpurun_out/r2m_bench.json
