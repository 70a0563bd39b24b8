#!/bin/bash
run() { timeout 120 python bench.py --steps 5 --warmup 3 --no-realtime --no-cpu-baseline --no-kernels --no-parity 2>/dev/null | python -c "import sys,json; l=json.loads(sys.stdin.read()); print(l['roofline']['kernel_ms'], l['ms_per_step'], l['clocks']['sm_mhz'], l['clocks']['power_w_max'], l['config']['mix_abs_sum'])"; }
export PBSO_TC_GAIN=1
for ab in 0 5; do echo "ablate=$ab"; PBSO_TC_ABLATE=$ab run; done
