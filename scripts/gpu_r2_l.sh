#!/bin/bash
run() { timeout 120 python bench.py --steps 5 --warmup 3 --no-realtime --no-cpu-baseline 2>/dev/null | python -c "import sys,json; l=json.loads(sys.stdin.read()); print(l['roofline']['kernel_ms'], l['ms_per_step'], l['clocks']['sm_mhz'], l['clocks']['power_w_max'], l['config']['mix_abs_sum'])"; }
echo "full"; run; run
echo "f64"; timeout 300 python bench.py --steps 1 --warmup 1 --no-realtime --no-cpu-baseline --precision f64 2>/dev/null | python -c "import sys,json; l=json.loads(sys.stdin.read()); print(l['ms_per_step'], l['config']['mix_abs_sum'])"
timeout 900 python -m pytest tests -m gpu -x -q -k "batch" 2>&1 | tail -5
