#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "storm or projection or transfer_table" 2>&1 | tail -6
timeout 600 python - <<'PY'
import sys, json, os
sys.path.insert(0, os.getcwd())
import bench, openpbso_b200 as pbso
from openpbso_b200 import synth
pk = pbso.measure_tc_peak(0, 1, 128)[0]
print(json.dumps(bench.contact_storm(pbso, synth, pk, 100), indent=1))
PY
