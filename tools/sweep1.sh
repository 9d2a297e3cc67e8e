mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8
for hint in 0 1 2; do for bs in 128 256 512; do
  echo "hint=$hint bs=$bs $(SLF_HINT=$hint SLF_BULK_BS=$bs timeout 300 python bench.py --size 256 --steps 50 --warmup 5 --no_cpu_baseline --no_e2e 2>&1 | tail -1 | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(round(d["value"]), round(d["roofline"]["frac"],4), d["roofline"]["ms_per_launch"])')"
done; done | tee gpurun_out/sweep1.txt
