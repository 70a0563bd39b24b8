#!/bin/bash
mkdir -p gpurun_out
for V in "0 0 8" "1 0 8" "1 1 8" "2 0 8" "2 1 8" "1 0 16" "1 1 16" "2 0 32"; do
set -- $V
export PBSO_FIT_PF=$1 PBSO_FIT_MPB=$3
if [ $2 = 1 ]; then export PBSO_FIT_PF_L1=1; else unset PBSO_FIT_PF_L1; fi
echo "--- prefetch distance $1, L1=$2, mpb max $3"
timeout 300 python scripts/bench_kernels.py --fit-only > gpurun_out/k6_v.json 2> gpurun_out/k6.err; tail -3 gpurun_out/k6.err
python - <<PY
import json
d=json.load(open('gpurun_out/k6_v.json'))
print([ (r["layout"][:3], int(r["power_scaling"]), int(r["deferred_scale"]), round(r["us"],1), r["parity_max_rel_vs_reference_layout"]) for r in d["K6_ffat_fit"]["runs"]])
PY
done
