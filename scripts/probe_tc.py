"""tcgen05 probes on the GPU box: pair self-test (operand placement of cta_group::2) and bare-MMA peaks for every
(kind, cta_group, N) with and without concurrent shared-memory stores.  Prints one JSON object."""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import openpbso_b200 as pbso

out = {"selftest": {}, "peak": []}
for kind, name in ((0, "tf32"), (1, "f16")):
    try:
        out["selftest"][name] = pbso.tc_selftest(kind)
    except Exception as e:  # noqa
        out["selftest"][name] = "error: %s" % e
for kind, name in ((0, "tf32"), (1, "f16")):
    for cg in (1, 2):
        for n in (128, 256):
            for stress in (0, 1):
                try:
                    t, cyc, wf = pbso.measure_tc_peak(kind, cg, n, stress)
                    out["peak"].append(dict(kind=name, cta_group=cg, n=n, stress=stress, tflops=round(t, 1), cycles_per_mma=round(cyc, 2),
                                            stress_wavefronts_per_cycle=round(wf, 3)))
                except Exception as e:  # noqa
                    out["peak"].append(dict(kind=name, cta_group=cg, n=n, stress=stress, error=str(e)))
print(json.dumps(out, indent=1))
