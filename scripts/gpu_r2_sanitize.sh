#!/bin/bash
# compute-sanitizer memcheck over the kernels added or changed in round 2 (run under gpurun): the tensor-core batch path (pairs,
# stems, any buffer size, event heads, stateful ranges), the storm pipeline, the 8-bit FFAT view, K6 flags, the legacy loader.
mkdir -p gpurun_out
timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
  -k "stateful or any_buffer_size or stems or many_events or object_batches or longer_render or storm or compress or legacy or ffat_fit_many or transfer_table_grows" \
  > gpurun_out/r2_sanitize_memcheck.log 2>&1; echo memcheck rc=$?
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds|misaligned" gpurun_out/r2_sanitize_memcheck.log | head -20
tail -3 gpurun_out/r2_sanitize_memcheck.log
