#!/bin/bash
timeout 60 python scripts/tc_debug.py 2>&1 | tail -2
timeout 300 python bench.py --steps 10 --warmup 3 --no-realtime --no-cpu-baseline --no-kernels 2>/dev/null | python -c "import sys,json; l=json.loads(sys.stdin.read()); print(l['roofline']['kernel_ms'], l['ms_per_step'], l['e2e']['ms_per_step'], l['clocks']['sm_mhz'], l['clocks']['power_w_max'], l['config']['mix_abs_sum'], l['parity'])"
timeout 300 python bench.py --steps 10 --warmup 3 --no-realtime --no-cpu-baseline --no-kernels 2>/dev/null | python -c "import sys,json; l=json.loads(sys.stdin.read()); print(l['roofline']['kernel_ms'], l['ms_per_step'], l['e2e']['ms_per_step'], l['clocks']['sm_mhz'], l['clocks']['power_w_max'], l['config']['mix_abs_sum'], l['parity'])"
