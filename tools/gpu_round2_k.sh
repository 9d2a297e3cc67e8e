#!/bin/bash
# round 2, call k (4 GPUs): does honouring the node priorities in the replayed step graph remove the ~55 us per step?
mkdir -p gpurun_out
run() {
  timeout 400 env $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus 4 --steps 40 --warmup 6 --no_e2e --no_parity --no_cpu_baseline 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); s=d['strong']; print('$1', 'weak', round(d['value']), d['ms_per_step'], 'strong', round(s['value']), s['ms_per_step'], 'speedup', s.get('speedup_vs_one_gpu'), 'one_gpu', s.get('one_gpu_value'))"
}
{
run "SLF_X=1" 29731
run "SLF_GRAPH_NO_NODE_PRIORITY=1" 29732
} > gpurun_out/r2k_priority.txt 2>&1
cat gpurun_out/r2k_priority.txt
