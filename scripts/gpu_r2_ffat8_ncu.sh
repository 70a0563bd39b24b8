#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ffat --csv --log-file gpurun_out/ffat8_launches.csv python scripts/ffat8_once.py > gpurun_out/ncu_ffat8.log 2>&1
grep -o '"k_ffat[^"]*".*' gpurun_out/ffat8_launches.csv | awk -F'","' '{print $1, $NF}' | head -20
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ffat_gather_q8x4 -c 1 -o gpurun_out/r2_ffat_gather_u8 -f python scripts/ffat8_once.py >> gpurun_out/ncu_ffat8.log 2>&1
tail -2 gpurun_out/ncu_ffat8.log
