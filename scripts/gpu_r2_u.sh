#!/bin/bash
echo "--- pair=2 debug"
timeout 90 python scripts/tc_debug.py 2>&1 | tail -9
echo "rc=$?"
run() { timeout 120 python bench.py --steps 8 --warmup 3 --no-realtime --no-cpu-baseline --no-kernels --no-parity 2>/dev/null | python -c "import sys,json; l=json.loads(sys.stdin.read()); print(l['roofline']['kernel_ms'], l['ms_per_step'], l['clocks']['sm_mhz'], l['clocks']['power_w_max'], l['config']['mix_abs_sum'])"; }
for pr in 2 1 2 1; do echo "pair=$pr"; PBSO_TC_PAIR=$pr run; done
PBSO_TC_GAIN=1 PBSO_TC_PROF=1 timeout 120 python bench.py --steps 1 --warmup 3 --no-realtime --no-cpu-baseline --no-kernels --no-parity 2>&1 | grep "tc prof" | tail -25
