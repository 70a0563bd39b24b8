#!/bin/bash
mkdir -p gpurun_out
export PBSO_TC_GAIN=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_batch_tc -s 2 -c 1 -o gpurun_out/r2_tc_v1 python bench.py --steps 1 --warmup 3 --no-realtime --no-cpu-baseline > gpurun_out/ncu_v1.log 2>&1
tail -3 gpurun_out/ncu_v1.log
ls -la gpurun_out/*.ncu-rep
