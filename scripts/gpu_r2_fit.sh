#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_headless.py -x -q -m gpu -k "ffat_fit or fit_tool" 2>&1 | tail -4
for V in staged gather; do
unset PBSO_FIT_NOSTAGE
if [ $V = gather ]; then export PBSO_FIT_NOSTAGE=1; fi
echo "--- $V"
timeout 300 python scripts/bench_kernels.py --fit-only > gpurun_out/k6_$V.json 2> gpurun_out/k6.err; tail -3 gpurun_out/k6.err
python - <<PY
import json
d=json.load(open('gpurun_out/k6_$V.json'))
for r in d["K6_ffat_fit"]["runs"]: print(r["layout"], r["power_scaling"], r["deferred_scale"], round(r["us"],1), "us frac", round(r["frac_of_hbm"],3), "sector frac", round(r["sector_frac_of_hbm"],3), r["parity_max_rel_vs_reference_layout"])
print(d["K6_ffat_fit"]["parity_max_rel_device_entry_vs_host_entry"])
PY
done
