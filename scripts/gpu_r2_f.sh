#!/bin/bash
mkdir -p gpurun_out
run() { timeout 120 python bench.py --steps 5 --warmup 3 --no-realtime --no-cpu-baseline 2>/dev/null | python -c "import sys,json; l=json.loads(sys.stdin.read()); print(l['roofline']['kernel_ms'], l['ms_per_step'], l['clocks']['sm_mhz'], l['clocks']['power_w_max'])"; }
export PBSO_TC_GAIN=1
for ab in 0 2 4 8 14; do echo "ablate=$ab"; PBSO_TC_ABLATE=$ab run; done
for w in 1 2 4 16 32; do echo "window=$w"; PBSO_TC_WINDOW=$w run; done
echo "pair=1"; PBSO_TC_PAIR=1 run
echo "flush=2"; PBSO_TC_FLUSH=2 run
echo "flush=64"; PBSO_TC_FLUSH=64 run
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_batch_tc -s 2 -c 1 python bench.py --steps 1 --warmup 3 --no-realtime --no-cpu-baseline 2>&1 | grep -E "k_batch_tc|dram__|lts__|sm__pipe|gpu__time|l1tex|smsp__" | head -20
