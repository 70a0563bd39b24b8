#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ffat" 2>&1 | tail -3
for NS in 0 1; do for U in 1 2 4; do for W in 1 4; do
echo "NOSPLIT=$NS U=$U WAVES=$W"; if [ $NS = 1 ]; then export PBSO_FFAT_Q8_NOSPLIT=1; else unset PBSO_FFAT_Q8_NOSPLIT; fi; PBSO_FFAT_Q8_U=$U PBSO_FFAT_Q8_WAVES=$W timeout 300 python scripts/bench_kernels.py --ffat-only 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print([ (r['L'], round(r['us'],1), r['parity_max_rel_vs_the_doubles_of_compressed_Psi']) for r in d['K3_ffat_eval']['compressed_view_u8']['runs']])"
done; done; done
