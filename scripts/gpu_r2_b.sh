#!/bin/bash
# ablation of k_batch_tc roles: bit0 B-gen, bit1 A-gen, bit2 epilogue, bit3 MMA
mkdir -p gpurun_out
for ab in 0 1 2 4 8 3 5 6 7 9 10 12 15; do
  echo "ablate=$ab"
  PBSO_TC_ABLATE=$ab timeout 120 python bench.py --steps 5 --warmup 3 --no-realtime --no-cpu-baseline 2>/dev/null | python -c "import sys,json; l=json.loads(sys.stdin.read()); print(l['roofline']['kernel_ms'], l['ms_per_step'], l['clocks']['sm_mhz'], l['clocks']['power_w_max'])"
done
