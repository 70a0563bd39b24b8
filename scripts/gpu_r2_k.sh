#!/bin/bash
mkdir -p gpurun_out
timeout 60 python scripts/tc_debug.py 2>&1 | tail -9
run() { timeout 120 python bench.py --steps 5 --warmup 3 --no-realtime --no-cpu-baseline 2>/dev/null | python -c "import sys,json; l=json.loads(sys.stdin.read()); print(l['roofline']['kernel_ms'], l['ms_per_step'], l['clocks']['sm_mhz'], l['clocks']['power_w_max'], l['config']['mix_abs_sum'])"; }
echo "full"; run
export PBSO_TC_GAIN=1
for ab in 1 4 8 13; do echo "ablate=$ab"; PBSO_TC_ABLATE=$ab run; done
timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed.sum --clock-control none -k regex:k_batch_tc -s 2 -c 1 python bench.py --steps 1 --warmup 3 --no-realtime --no-cpu-baseline 2>&1 | grep -E "gpu__time|sm__pipe|l1tex|smsp__" | head
