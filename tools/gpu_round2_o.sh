#!/bin/bash
# round 2, call o (1 GPU, last minutes of the budget): the new node types (copy, Yu outflow, TMS) on the GPU, the
# re-recorded plugin traces (ABI v6), then the rest of the GPU suite except the full-size cases, and smoke()
mkdir -p gpurun_out
( time timeout 100 python -m pytest tests/test_gpu_zoo_node_types.py tests/test_gpu_plugin_replay.py -m gpu -q --maxfail=10 ) 2>&1 | tail -25 > gpurun_out/r2o_new_types.txt
tail -12 gpurun_out/r2o_new_types.txt
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 170 python -m pytest tests -m gpu -q --maxfail=10 --deselect tests/test_gpu_full_size.py --deselect tests/test_gpu_zoo_node_types.py --deselect tests/test_gpu_plugin_replay.py ) 2>&1 | tail -25 > gpurun_out/r2o_rest.txt
tail -8 gpurun_out/r2o_rest.txt
