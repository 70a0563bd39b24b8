#!/bin/bash
# GPU-box run for the FFAT-fit widening (run under gpurun): GPU tests, kernel micro-benchmarks incl. K6, one full ncu
# capture of k_fit_solve, the default bench line.
mkdir -p gpurun_out
set -x
nproc; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 python scripts/bench_kernels.py > gpurun_out/kernels.json 2> gpurun_out/kernels.err; tail -c 300 gpurun_out/kernels.err; cat gpurun_out/kernels.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fit_solve -s 4 -c 2 -o gpurun_out/r1_fit python scripts/bench_kernels.py --fit-only > gpurun_out/ncu_fit.log 2>&1
tail -3 gpurun_out/ncu_fit.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 400 gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json
ls -la gpurun_out
