"""cfg2 latency shape: percentiles with and without listener jumps, and the split between the two host calls."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import openpbso_b200 as pbso
from openpbso_b200 import synth

N, BUF = 1024, 256
f = synth.mode_frequencies(N, 1002)
a, b = synth.ab_from_material(f, synth.MATERIALS["low_damping"])
it = pbso.ModalIntegrator(N, synth.H, a, b)
fm = pbso.FFATMaps.from_dicts(synth.ffat_maps(f, 2000))
ls = synth.listeners(400, 1002)
it.set_transfer_ffat(fm, ls[:1])
zero = np.zeros(N); tm = np.zeros(BUF)
for _ in range(300): it.render_buffer(zero, tm)
for jumps in (False, True):
    n = 10000; lat = np.empty(n); jl = []
    for i in range(n):
        t0 = time.perf_counter()
        if jumps and i % 50 == 0:
            it.set_transfer_ffat(fm, ls[i // 50:i // 50 + 1])
        it.render_buffer(zero, tm)
        lat[i] = time.perf_counter() - t0
        if jumps and i % 50 == 0: jl.append(lat[i])
    us = lat * 1e6
    print("jumps=%s" % jumps, {p: round(float(np.percentile(us, p)), 1) for p in (50, 90, 95, 98, 99, 99.9)}, "max %.0f" % us.max(),
          ("jump buffers p50 %.1f" % (np.median(jl) * 1e6)) if jl else "")

# the same loop with the interpreter's garbage collector off and the C entry called directly on preallocated buffers:
# separates the library's latency from the Python harness around it
import gc, ctypes as C
from openpbso_b200 import _capi as capi
L = pbso.lib()
y = np.empty(BUF); qn = np.empty(N)
args = (it._h, capi.dp(zero), capi.dp(tm), BUF, capi.dp(y), capi.dp(qn))
for label, off in (("gc on, direct C call", False), ("gc off, direct C call", True)):
    if off: gc.disable()
    n = 10000; lat = np.empty(n)
    for i in range(n):
        t0 = time.perf_counter()
        L.pbso_render_buffer(*args)
        lat[i] = time.perf_counter() - t0
    if off: gc.enable()
    us = lat * 1e6
    print(label, {p: round(float(np.percentile(us, p)), 1) for p in (50, 90, 95, 98, 99, 99.9)}, "max %.0f" % us.max())
