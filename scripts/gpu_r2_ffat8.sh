#!/bin/bash
# 8-bit FFAT view: parity tests + K3 micro-benchmarks
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_dropin.py -x -q -m gpu -k "ffat or drop" 2>&1 | tail -5
timeout 300 python scripts/bench_kernels.py --ffat-only > gpurun_out/k3_q8.json 2> gpurun_out/k3_q8.err; tail -3 gpurun_out/k3_q8.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/k3_q8.json'))
print(json.dumps(d["K3_ffat_eval"],indent=1))
PY
