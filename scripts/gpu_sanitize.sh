#!/bin/bash
# compute-sanitizer over the kernels added this round (run under gpurun): memcheck on the FFAT / fit / listener tests, smoke().
mkdir -p gpurun_out
set -x
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "ffat or fit or listeners" > gpurun_out/sanitize_memcheck.log 2>&1; echo memcheck rc=$?
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds|misaligned" gpurun_out/sanitize_memcheck.log | head -20
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
