#!/bin/bash
# round 2, call A: regression of the GPU suite after the ADVICE fixes + tcgen05 probes
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 300 python scripts/probe_tc.py > gpurun_out/probe_tc.json 2> gpurun_out/probe_tc.err; tail -c 600 gpurun_out/probe_tc.err; cat gpurun_out/probe_tc.json
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
