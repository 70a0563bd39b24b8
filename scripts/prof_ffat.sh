#!/bin/bash
# ncu capture of the FFAT kernels (many-listener case) -- run under gpurun
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_ffat -c 12 -o gpurun_out/r1_ffat python scripts/bench_kernels.py --quick --ffat-only > gpurun_out/ncu_ffat.log 2>&1
tail -3 gpurun_out/ncu_ffat.log
