#!/bin/bash
# the bench job at other shapes (parity against the FP64 kernel at each): many modes per object, a mode count that is not a
# multiple of 16, few large objects (the mode-block shape), a longer render
mkdir -p gpurun_out
run() { timeout 600 python bench.py --steps 3 --warmup 3 --no-kernels --no-realtime --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', '->', round(d['ms_per_step'],2), 'ms', '%.3e' % d['value'], 'frac', round(d['roofline']['frac'],3), 'parity', '%.2e' % d['parity']['rel_l2'], '%.2e' % d['parity']['max_abs'])"; }
if [ "$1" != "--shapes-only" ]; then timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "batch" 2>&1 | tail -3; fi
run --objects 4096 --modes 512
run --objects 1024 --modes 2048
run --objects 4096 --modes 500
run --objects 64 --modes 8192
run --objects 8 --modes 65536
run --objects 2048 --modes 512 --buffers 3446
run --objects 8192 --modes 256
