#!/bin/bash
# K6 iteration loop (run under gpurun): fit tests, A/B of the TMA-staged and direct-gather kernels, one ncu capture.
mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests -m gpu -x -q -k "fit or moving" 2>&1 | tail -15
timeout 300 python scripts/bench_kernels.py --fit-only > gpurun_out/fit_direct.json 2> gpurun_out/fit_direct.err; tail -c 300 gpurun_out/fit_direct.err; cat gpurun_out/fit_direct.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fit_solve -s 4 -c 2 -o gpurun_out/r1_fit python scripts/bench_kernels.py --fit-only > gpurun_out/ncu_fit.log 2>&1
tail -3 gpurun_out/ncu_fit.log
