#!/bin/bash
# Full GPU-box validation (run under gpurun): GPU tests, kernel micro-benchmarks (K3-K6), ncu captures of the K3 tile
# kernel and the K6 fit kernel, the default bench line, the reference arm, and the ncu launch list of the bench.
mkdir -p gpurun_out
set -x
nproc; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 python scripts/bench_kernels.py > gpurun_out/kernels.json 2> gpurun_out/kernels.err; tail -c 300 gpurun_out/kernels.err; cat gpurun_out/kernels.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fit_solve -s 4 -c 1 -o gpurun_out/r1_fit python scripts/bench_kernels.py --fit-only > gpurun_out/ncu_fit.log 2>&1
tail -2 gpurun_out/ncu_fit.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_ffat_tiles|k_ffat_locate" -s 6 -c 2 -o gpurun_out/r1_ffat_tiles python scripts/bench_kernels.py --ffat-only > gpurun_out/ncu_ffat.log 2>&1
tail -2 gpurun_out/ncu_ffat.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 400 gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>&1; cat gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-realtime --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -c 300 gpurun_out/ncu_bench.log
ls -la gpurun_out
