"""Generates tests/golden/legacy_fatcube/ -- .fatcube files in the LEGACY form (libigl's igl::serialize of the FFAT_Map<T,3> object)
written by the REFERENCE'S OWN FFAT_Map<double,3>::Save (ffat_solver.h:1066-1068) with libigl's own igl/serialize.h, both compiled in
place from /root/reference into oracle/_ref -- and legacy_eval.npz, what the reference's own legacy loader
(FFAT_Map<double,3>::LoadAll, igl::deserialize) + |GetMapVal| gives for them.  Run HERE:

    python tests/golden/make_golden_legacy.py

Two kinds of file: mode-0..2 are complete maps (constructor + Solve: three shells, every member valid; non-cubic shells of different
sizes); mode-3 comes from a protobuf .fatcube loaded by FFAT_Map_Serialize::Load and saved in the legacy form (members the protobuf
form does not keep are empty or uninitialised in it, exactly as the reference leaves them)."""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc, fatcube          # noqa: E402
from openpbso_b200 import synth                    # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "legacy_fatcube")


def main():
    assert orc.ref() is not None, "oracle/_ref is not built: /root/reference missing?"
    os.makedirs(OUT, exist_ok=True)
    w = synth.ffat_fit_workload(3, 61, half_cells=((1, 2, 2), (2, 3, 3), (3, 3, 4)), cell_size=0.25)
    for m in range(3):
        rows = orc.ref_ffat_legacy_fit_save(m, w["cell_size"], w["V"], w["n_elements"], w["k"][m], w["pressure"][m], m == 1,
                                            os.path.join(OUT, "mode-%d.fatcube" % m))
        assert rows > 0
    # a protobuf map re-saved in the legacy form, re-keyed to mode id 3
    src = fatcube.load(os.path.join(ROOT, "tests", "golden", "fatcube", "mode-2.fatcube")); src["modeid"] = 3
    tmp = os.path.join(OUT, "_tmp.pb"); fatcube.save(tmp, src)
    assert orc.ref_ffat_legacy_from_fatcube(tmp, os.path.join(OUT, "mode-3.fatcube")) == 3
    os.remove(tmp)
    rng = np.random.default_rng(62)
    axes = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]], dtype=float)
    pos = np.concatenate([synth.listeners(150, 62), 4.0 * axes, 3.0 * np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], dtype=float),
                          0.2 * (rng.random((10, 3)) * 2 - 1)])
    out = orc.ref_ffat_eval_legacy(OUT, pos)
    assert out is not None and out.shape == (len(pos), 4)
    np.savez_compressed(os.path.join(os.path.dirname(OUT), "legacy_eval.npz"), pos=pos, out=out)
    print("wrote", sorted(os.listdir(OUT)), out.shape, "non-finite:", int(np.sum(~np.isfinite(out))))


if __name__ == "__main__":
    main()
