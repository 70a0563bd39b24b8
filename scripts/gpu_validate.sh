#!/bin/bash
# GPU-box validation (run under gpurun): GPU tests, default bench line, ncu launch list of the default (tc3x) bench.
mkdir -p gpurun_out
set -x
nproc; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 400 gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>&1; cat gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r1_launches_tc3x.csv python bench.py --steps 2 --warmup 3 --no-realtime --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -c 300 gpurun_out/ncu_bench.log
timeout 300 python scripts/bench_kernels.py > gpurun_out/kernels.json 2> gpurun_out/kernels.err; tail -c 300 gpurun_out/kernels.err
ls -la gpurun_out
