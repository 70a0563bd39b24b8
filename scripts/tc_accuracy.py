"""Accuracy of the tensor-core batch render (PBSO_PREC_TC3X) on a cfg5 slice against the FP64 kernel:
rel-L2, max-abs / full scale, and the fitted gain error (accumulator truncation is a coherent bias).
Variants are selected with PBSO_TC_SPLIT / PBSO_TC_CHAIN (read once per process)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import openpbso_b200 as pbso
from openpbso_b200 import synth

n_obj, n_modes, n_buf = 96, 512, 1723
w = synth.batch_workload(n_obj, n_modes, n_buf, 1005)
br = pbso.BatchRenderer(synth.H, w["a"], w["b"]); br.set_transfer(w["trans"])
br.set_impulses(np.arange(n_obj), w["imp_buf"], w["space"])
y64 = br.render_mix(256, n_buf, pbso.PREC_F64)
y = br.render_mix(256, n_buf, pbso.PREC_TC3X)
gain = float(np.dot(y, y64) / np.dot(y64, y64))
full = np.max(np.abs(y64))
r = y - gain * y64
print("split=%s chain=%s: rel-L2 %.3e max-abs %.3e gain-1 %.3e; after removing the gain: rel-L2 %.3e max-abs %.3e" % (
    os.environ.get("PBSO_TC_SPLIT", "1"), os.environ.get("PBSO_TC_CHAIN", "2"),
    np.linalg.norm(y - y64) / np.linalg.norm(y64), np.max(np.abs(y - y64)) / full, gain - 1.0,
    np.linalg.norm(r) / np.linalg.norm(y64), np.max(np.abs(r)) / full))
y32 = br.render_mix(256, n_buf, pbso.PREC_F32_TILED)
print("f32_tiled for comparison: rel-L2 %.3e max-abs %.3e" % (np.linalg.norm(y32 - y64) / np.linalg.norm(y64), np.max(np.abs(y32 - y64)) / full))
