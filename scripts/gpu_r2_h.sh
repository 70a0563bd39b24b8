#!/bin/bash
mkdir -p gpurun_out
timeout 60 python scripts/tc_debug.py 2>&1 | tail -9
run() { timeout 120 python bench.py --steps 5 --warmup 3 --no-realtime --no-cpu-baseline 2>/dev/null | python -c "import sys,json; l=json.loads(sys.stdin.read()); print(l['roofline']['kernel_ms'], l['ms_per_step'], l['clocks']['sm_mhz'], l['clocks']['power_w_max'], l['config']['mix_abs_sum'])"; }
echo "cl=2"; run
echo "cl=4"; PBSO_TC_PAIR=4 run
echo "cl=1"; PBSO_TC_PAIR=1 run
export PBSO_TC_GAIN=1
for ab in 2 4 8 14; do echo "cl=2 ablate=$ab"; PBSO_TC_ABLATE=$ab run; done
for ab in 14; do echo "cl=4 ablate=$ab"; PBSO_TC_PAIR=4 PBSO_TC_ABLATE=$ab run; done
