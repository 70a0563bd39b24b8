#!/bin/bash
# ncu --set full capture of k_batch_tc<2> on a full cfg5 launch (one GPU) -> gpurun_out/r2_tc_final.ncu-rep + a metrics summary.
# PBSO_TC_GAIN=1 skips the gain calibration (its small launches would be captured instead of the bench-size one).
mkdir -p gpurun_out
export PBSO_TC_GAIN=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_batch_tc -s 2 -c 1 -o gpurun_out/r2_tc_final -f \
    python bench.py --steps 1 --warmup 3 --no-realtime --no-cpu-baseline --no-kernels --no-parity > gpurun_out/ncu_tc_final.log 2>&1
tail -2 gpurun_out/ncu_tc_final.log | cut -c1-300
ncu -i gpurun_out/r2_tc_final.ncu-rep --page raw --csv > gpurun_out/r2_tc_final_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows = list(csv.reader(open('gpurun_out/r2_tc_final_raw.csv')))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'launch__cluster_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'lts__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.avg', 'sm__cycles_elapsed.avg.per_second',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio']
with open('gpurun_out/r2_k_batch_tc_final_metrics.txt', 'w') as f:
    for w in want:
        if w in hdr:
            i = hdr.index(w); f.write('%-90s %-16s %s\n' % (w, units[i], vals[i]))
print(open('gpurun_out/r2_k_batch_tc_final_metrics.txt').read())
PY
