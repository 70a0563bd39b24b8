#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ffat" 2>&1 | tail -3
for i in 1 2; do
timeout 300 python scripts/bench_kernels.py --ffat-only 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print([ (r['L'], round(r['us'],1), round(r['frac_of_hbm'],3), r.get('parity_max_rel_vs_per_listener_gather')) for r in d['K3_ffat_eval']['runs']], [ (r['L'], round(r['us'],1)) for r in d['K3_ffat_eval']['compressed_view_u8']['runs']])"
done
