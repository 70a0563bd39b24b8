#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_dropin.py tests/test_headless.py -x -q -m gpu -k "ffat or legacy or drop or fit" 2>&1 | tail -12
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
