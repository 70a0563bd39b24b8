#!/bin/bash
# K5 (batched projection on tensor cores): parity tests, the storm pipeline, micro-benchmark
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "project or storm or dense" 2>&1 | tail -5
timeout 300 python scripts/bench_kernels.py > gpurun_out/k5.json 2> gpurun_out/k5.err; tail -3 gpurun_out/k5.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/k5.json'))
for r in d["K5_project_tc"]["runs"]: print(r["B"], round(r["ms"],4), "ms", round(r["tflops"],1), "TF alg, frac", round(r["frac_of_tf32_peak"],3), r["parity_col_rel_l2_vs_fp64"])
PY
