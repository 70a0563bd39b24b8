#!/bin/bash
mkdir -p gpurun_out
PBSO_TC_GAIN=1 PBSO_TC_PAIR=1 timeout 60 python scripts/tc_debug.py 4 2>&1 | tail -8
echo "--- pair"
PBSO_TC_GAIN=1 timeout 60 python scripts/tc_debug.py 2>&1 | tail -10
echo "--- calibrated"
timeout 60 python scripts/tc_debug.py 2>&1 | tail -10
