#!/bin/bash
# Round-2 verification on one B200: GPU tests, smoke, the default bench line, the reference arm, the ncu launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 1500 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -2 gpurun_out/bench_default.err; head -c 1500 gpurun_out/bench_default.json; echo
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; head -c 900 gpurun_out/bench_reference.json; echo
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-kernels --no-parity > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log | cut -c1-300
