#!/bin/bash
# round 2, call c (2 GPUs): why is weak scaling at N=2 below R1's?  step scheme x graph replay
mkdir -p gpurun_out
run() {
  timeout 300 env $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus 2 --steps 40 --warmup 6 --no_strong --no_e2e --no_parity --no_cpu_baseline 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', d['value'], d['ms_per_step'], d['gpu_launches'])"
}
{
run "SLF_STEP_SCHEME=concurrent" 29521
run "SLF_STEP_SCHEME=serial" 29522
run "SLF_STEP_SCHEME=concurrent SLF_NO_STEP_GRAPHS=1" 29523
run "SLF_STEP_SCHEME=serial SLF_NO_STEP_GRAPHS=1" 29524
run "SLF_STEP_SCHEME=concurrent" 29525
} > gpurun_out/r2c_schemes.txt 2>&1
cat gpurun_out/r2c_schemes.txt
