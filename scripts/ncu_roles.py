"""Summarise an ncu report of k_batch_tc by warp role (regions between USETMAXREG markers) and headline metrics.
usage: python scripts/ncu_roles.py gpurun_out/x.ncu-rep"""
import csv, subprocess, sys, io
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
stalls = [h for h in hdr if h.startswith('stall_') and 'Not' not in h]
marks = [i for i, r in enumerate(data) if 'USETMAXREG' in r[ci['Source']]] + [len(data)]
def agg(lo, hi, name):
    n = 0; inst = 0; a = {h: 0 for h in stalls}; byop = {}
    for r in data[lo:hi]:
        ns = int(r[ci['# Samples']]); n += ns; inst += int(r[ci['Instructions Executed']])
        for h in stalls: a[h] += int(r[ci[h]])
        toks = r[ci['Source']].strip().split(); op = toks[1] if toks[0].startswith('@') else toks[0]
        d = byop.setdefault(op, [0, 0]); d[0] += ns; d[1] += int(r[ci['Instructions Executed']])
    print(name, "samples", n, "warp-instr", inst, sorted(a.items(), key=lambda x: -x[1])[:6])
    for op, (s, e) in sorted(byop.items(), key=lambda x: -x[1][0])[:12]: print("   ", op.ljust(30), s, e)
for k in range(len(marks) - 1):
    agg(marks[k], marks[k + 1], "region %d (%s)" % (k, data[marks[k]][ci['Source']].strip()))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
keys = ["gpu__time_duration.sum", "smsp__issue_active.avg.pct", "l1tex__data_pipe_lsu_wavefronts.avg.pct", "sm__pipe_tensor_cycles_active.avg.pct",
        "smsp__inst_executed.sum ", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum ", "dram__bytes_write.sum ", "sm__cycles_active.avg ",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ", "sm__inst_executed_pipe_tmem", "lts__t_sectors_op_red.sum ", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum "]
for h, u, v in zip(rows[0], rows[1], rows[2]):
    if any(k in h + " " for k in keys) and v not in ("0", ""): print(h, u, v)
