"""One evaluation of the 8-bit FFAT view at the HUD-sphere size (1024 maps x 10242 listeners), for ncu."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import openpbso_b200 as pbso
from openpbso_b200 import synth
Mf, L = 1024, int(os.environ.get("L", "10242"))
dicts = synth.ffat_maps(synth.mode_frequencies(Mf, 1004), 2000)
fm = pbso.FFATMaps.from_dicts(dicts); fm.Compress()
pos = torch.from_numpy(synth.listeners(L, 5)).cuda(); o = torch.empty(L, Mf, device="cuda", dtype=torch.float64)
for view in (1, 0, 1, 0):
    pbso._capi.check(pbso.lib().pbso_ffat_eval_device_view(fm._h, Mf, C.c_void_p(pos.data_ptr()), L, view, C.c_void_p(o.data_ptr()), None))
torch.cuda.synchronize()
