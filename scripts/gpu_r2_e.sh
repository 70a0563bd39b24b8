#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for ab in 0; do
  PBSO_TC_ABLATE=$ab timeout 300 python bench.py --steps 10 --warmup 3 --no-realtime --no-cpu-baseline > gpurun_out/bench_e.json 2> gpurun_out/bench_e.err
  tail -c 300 gpurun_out/bench_e.err
  python -c "import sys,json; l=json.loads(open('gpurun_out/bench_e.json').read()); print('kernel_ms',l['roofline']['kernel_ms'], 'step', l['ms_per_step'], 'e2e', l['e2e'], l['clocks'], l['config']['mix_abs_sum'])"
done
for ab in 2 4 8 15; do
  echo "ablate=$ab"
  PBSO_TC_ABLATE=$ab timeout 120 python bench.py --steps 5 --warmup 3 --no-realtime --no-cpu-baseline 2>/dev/null | python -c "import sys,json; l=json.loads(sys.stdin.read()); print(l['roofline']['kernel_ms'], l['ms_per_step'], l['clocks']['sm_mhz'], l['clocks']['power_w_max'])"
done
