#!/bin/bash
# round 2, call p (1 GPU, the last ~2 minutes of the budget): the slip nodes on the GPU + one replayed plugin trace
mkdir -p gpurun_out
( time timeout 75 python -m pytest tests/test_gpu_zoo_node_types.py tests/test_gpu_plugin_replay.py -m gpu -q --maxfail=10 -k "slip or ldc2d_mrt_AB" ) 2>&1 | tail -25 > gpurun_out/r2p_slip.txt
tail -12 gpurun_out/r2p_slip.txt
