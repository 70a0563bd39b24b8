"""Generates tests/golden/ffat_eval.npz with the REFERENCE'S OWN run-time FFAT code: FFAT_Map_Serialize::LoadAll +
|FFAT_Map<double,3>::GetMapVal| (Intersect / Interpolate / GetDataQuadStride / Reconstruct), compiled in place from
/root/reference into oracle/_ref.  Run HERE (the build container):

    python tests/golden/make_golden_ffat_eval.py

Two cases:
  files_*   the committed tests/golden/fatcube maps (three maps of different sizes: the per-map kernel) at 260 probe positions
            incl. face axes, cube edges / corners (tie-breaking) and points inside the box
  shared_*  24 maps sharing one 6 x 8 x 8 geometry (written as .fatcube, loaded by the reference) at 2200 listener positions:
            few of them exercise the fused small-L kernel, all of them the texel-stationary kernel
"""
import os
import sys
import tempfile
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc, fatcube          # noqa: E402
from openpbso_b200 import synth                    # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def probes(n, seed, half=1.5):
    rng = np.random.default_rng(seed)
    p = [synth.listeners(n, seed)]
    axes = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]], dtype=float)
    p.append(4.0 * axes)                                                        # face axes
    corners = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], dtype=float)
    p.append(3.0 * corners)                                                     # cube diagonals: three-way ties
    edges = np.array([[sx, sy, 0.3] for sx in (-1, 1) for sy in (-1, 1)] + [[0.2, sy, sz] for sy in (-1, 1) for sz in (-1, 1)], dtype=float)
    p.append(5.0 * edges)                                                       # two-way ties
    p.append(0.6 * half * (rng.random((20, 3)) * 2 - 1))                        # inside the box: the ray leaves it backwards
    return np.concatenate(p)


def main():
    assert orc.ref() is not None, "oracle/_ref is not built: /root/reference missing?"
    fdir = os.path.join(OUT, "fatcube")
    pos_files = probes(218, 31)
    files = orc.ref_ffat_eval(fdir, pos_files)
    freqs = synth.mode_frequencies(24, 1004)
    maps = synth.ffat_maps(freqs, 2000, n=8)
    pos_shared = np.concatenate([probes(38, 33), synth.listeners(2200 - 80, 34)])
    with tempfile.TemporaryDirectory() as d:
        for m in maps:
            fatcube.save(os.path.join(d, "mode-%03d.fatcube" % m["modeid"]), m)
        shared = orc.ref_ffat_eval(d, pos_shared)
    assert files is not None and shared is not None and shared.shape == (2200, 24)
    np.savez_compressed(os.path.join(OUT, "ffat_eval.npz"), files_pos=pos_files, files_out=files,
                        shared_pos=pos_shared, shared_out=shared, shared_freqs=freqs)
    print("wrote ffat_eval.npz:", files.shape, shared.shape, "non-finite in files:", int(np.sum(~np.isfinite(files))))


if __name__ == "__main__":
    main()
