"""GPU-box experiment: parity + throughput of every pole-power kernel variant (PBSO_POW_VARIANT)."""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
check = r'''
import numpy as np, sys
sys.path.insert(0, %r)
import openpbso_b200 as pbso
from openpbso_b200 import synth
w = synth.batch_workload(24, 512, 300, 1005)
br = pbso.BatchRenderer(synth.H, w["a"], w["b"]); br.set_transfer(w["trans"]); br.set_impulses(np.arange(24), w["imp_buf"] %% 200, w["space"])
y64 = br.render_mix(256, 300, pbso.PREC_F64); y32 = br.render_mix(256, 300, pbso.PREC_F32_TILED)
print("parity rel-L2 %%.2e max-abs %%.2e" %% (np.linalg.norm(y32-y64)/np.linalg.norm(y64), np.max(np.abs(y32-y64))/np.max(np.abs(y64))))
''' % ROOT
for v in sys.argv[1:] or ["0", "1", "2", "3", "4", "5"]:
    env = dict(os.environ, PBSO_POW_VARIANT=v)
    out = subprocess.run([sys.executable, "-c", check], env=env, capture_output=True, text=True)
    par = out.stdout.strip() or out.stderr.strip()[-300:]
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "3", "--warmup", "3", "--objects", "1184",
                          "--no-realtime", "--no-cpu-baseline"], env=env, capture_output=True, text=True)
    try:
        j = json.loads(out.stdout.strip().splitlines()[-1])
        print("variant %s: %.3e mode-samples/s, frac %.3f, kernel %.2f ms | %s" % (v, j["value"], j["roofline"]["frac"], j["roofline"]["kernel_ms"], par), flush=True)
    except Exception as e:
        print("variant %s failed: %s %s" % (v, out.stdout[-300:], out.stderr[-500:]), flush=True)
