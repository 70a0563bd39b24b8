"""Small TC3X-vs-FP64 cases on the GPU box with per-case errors (debugging aid for batch_tc.cu)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import openpbso_b200 as pbso
from openpbso_b200 import synth
H = synth.H
cases = [(1, 16, 64, "low_damping"), (2, 32, 64, "low_damping"), (5, 80, 24, "low_damping"), (3, 300, 12, "high_damping"),
         (2, 16, 150, "low_damping"), (7, 33, 70, "high_damping"), (40, 64, 200, "low_damping"), (300, 128, 130, "high_damping")]
if len(sys.argv) > 1:
    cases = cases[:int(sys.argv[1])]
for n_obj, n_modes, n_buf, mat in cases:
    w = synth.batch_workload(n_obj, n_modes, n_buf, 11, mat, first_second_bufs=max(1, n_buf // 2))
    br = pbso.BatchRenderer(H, w["a"], w["b"]); br.set_transfer(w["trans"])
    br.set_impulses(np.arange(n_obj), w["imp_buf"], w["space"])
    y64 = br.render_mix(256, n_buf, pbso.PREC_F64)
    t0 = time.time()
    ytc = br.render_mix(256, n_buf, pbso.PREC_TC3X)
    dt = time.time() - t0
    full = np.max(np.abs(y64))
    err = np.abs(ytc - y64)
    i = int(np.argmax(err))
    g = float(np.dot(ytc, y64) / np.dot(y64, y64))
    print("case %s: rel-L2 %.3e max-abs %.3e at sample %d (tile %d row-in-mtile %d) fitted gain-1 %.2e nan=%d  %.3fs gain=%.9f" % (
        (n_obj, n_modes, n_buf, mat), np.linalg.norm(ytc - y64) / np.linalg.norm(y64), err.max() / full, i, i // 128, (i // 128) % 128,
        g - 1, int(np.isnan(ytc).sum()), dt, pbso.tc_gain()), flush=True)
