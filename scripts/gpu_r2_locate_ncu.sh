#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ffat_locate -s 1 -c 1 -o gpurun_out/r2_ffat_locate -f python scripts/ffat8_once.py > gpurun_out/ncu_locate.log 2>&1
tail -2 gpurun_out/ncu_locate.log
