#!/bin/bash
# K3 iteration loop (run under gpurun): FFAT tests, many-listener micro-benchmark (texel tiles vs per-listener gather), one ncu capture.
mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests -m gpu -x -q -k "ffat or moving or fit" 2>&1 | tail -15
timeout 300 python scripts/bench_kernels.py --ffat-only > gpurun_out/ffat_only.json 2> gpurun_out/ffat_only.err; tail -c 300 gpurun_out/ffat_only.err; cat gpurun_out/ffat_only.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_ffat_tiles|k_ffat_locate" -s 6 -c 2 -o gpurun_out/r1_ffat_tiles python scripts/bench_kernels.py --ffat-only > gpurun_out/ncu_ffat.log 2>&1
tail -3 gpurun_out/ncu_ffat.log
