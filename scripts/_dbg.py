import os, sys
import numpy as np
sys.path.insert(0, "/root/repo")
import openpbso_b200 as pbso
from openpbso_b200 import synth
for n_obj in (8, 48):
    n_modes, n_buf = 512, 1723
    w = synth.batch_workload(n_obj, n_modes, n_buf, 1005)
    br = pbso.BatchRenderer(synth.H, w["a"], w["b"]); br.set_transfer(w["trans"])
    br.set_impulses(np.arange(n_obj), w["imp_buf"], w["space"])
    y64 = br.render_mix(256, n_buf, pbso.PREC_F64)
    for rep in range(2):
        y = br.render_mix(256, n_buf, pbso.PREC_TC3X)
        e = (y - y64).reshape(-1, 128)
        full = np.max(np.abs(y64))
        per_tile = np.max(np.abs(e), axis=1) / full
        bad = np.nonzero(per_tile > 1e-5)[0]
        print(n_obj, rep, "rel", np.linalg.norm(y - y64) / np.linalg.norm(y64), "bad tiles", len(bad), bad[:20], "M-tiles", sorted(set(bad // 128))[:30])
        if len(bad):
            t = bad[0]; print("  tile", t, "cols bad", np.nonzero(np.abs(e[t]) / full > 1e-5)[0][:40])
