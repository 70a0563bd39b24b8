#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_dropin.py tests/test_headless.py -x -q -m gpu -k "ffat or legacy or drop or fit" 2>&1 | tail -12
