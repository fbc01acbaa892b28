mkdir -p gpurun_out/r02p
for i in 1 2 3; do
  python bench.py --steps 30 --warmup 5 --no-train-step --no-gpu-baseline --no-cpu-baseline --no-config3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('early', d['ms_per_step'], d['clocks']['sm_mhz'])"
  AITB_QUERY_FORK_LATE=1 python bench.py --steps 30 --warmup 5 --no-train-step --no-gpu-baseline --no-cpu-baseline --no-config3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('late ', d['ms_per_step'], d['clocks']['sm_mhz'])"
done
