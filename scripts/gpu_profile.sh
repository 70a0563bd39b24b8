#!/bin/bash
# GPU-box profiling recipe (run under gpurun): GPU tests, ncu launch list, one full capture of the synthesis kernel.
mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
# launch list (cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 2 --warmup 3 --objects 592 --no-realtime --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -c 600 gpurun_out/ncu_bench.log
# full capture of the synthesis kernel
ncu --set full --clock-control none --import-source on -k regex:k_batch_pow -s 2 -c 1 -o gpurun_out/r1_batch_pow python bench.py --steps 1 --warmup 3 --objects 592 --no-realtime --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
