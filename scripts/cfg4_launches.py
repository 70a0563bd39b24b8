"""cfg4 (64 moving listeners, 1024 modes) for a few buffers -- run under `ncu --metrics gpu__time_duration.sum` to list the
per-buffer launches, or plain to print the host-side latency split."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import openpbso_b200 as pbso
from openpbso_b200 import synth

N, L, BUF = 1024, 64, 256
f = synth.mode_frequencies(N, 1004)
a, b = synth.ab_from_material(f, synth.MATERIALS["low_damping"])
it = pbso.ModalIntegrator(N, synth.H, a, b)
fm = pbso.FFATMaps.from_dicts(synth.ffat_maps(f, 2000))
pos = synth.listeners(L, 1004)
sp = np.random.default_rng(0).standard_normal(N); tm = np.zeros(BUF); tm[0] = 1.0
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
t_set = np.empty(n); t_ren = np.empty(n)
for i in range(n):
    t0 = time.perf_counter(); it.set_transfer_ffat(fm, pos); t1 = time.perf_counter()
    it.render_buffer(sp if i % 8 == 0 else np.zeros(N), tm, want_qnorm=False); t2 = time.perf_counter()
    t_set[i] = t1 - t0; t_ren[i] = t2 - t1
print("set_transfer_ffat p50 %.1f us, render_buffer p50 %.1f us" % (np.median(t_set[n // 10:]) * 1e6, np.median(t_ren[n // 10:]) * 1e6))
