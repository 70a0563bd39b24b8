#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py > gpurun_out/bench_m.json 2> gpurun_out/bench_m.err; tail -c 500 gpurun_out/bench_m.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench_m.json').read())
print({k:l[k] for k in ('value','ms_per_step','e2e','parity','gpu_launches','clocks')})
print(l['roofline'])
print(json.dumps(l.get('kernels'))[:3000])
PY
export PBSO_TC_GAIN=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_batch_tc -s 2 -c 1 -o gpurun_out/r2_tc_v2 python bench.py --steps 1 --warmup 3 --no-realtime --no-cpu-baseline --no-kernels --no-parity > gpurun_out/ncu_v2.log 2>&1
tail -2 gpurun_out/ncu_v2.log | cut -c1-300
