"""Generates tests/golden/ffat_fit.npz with the REFERENCE'S OWN code: the cube-map mesh from
FFAT_Map<double,1>::CubemapMesh and Psi from FFAT_Map<double,3>(modeId, cellSize, V, N_elements) + Solve(k, p, powerScaling),
both compiled in place from /root/reference into oracle/_ref (oracle/Makefile target `ref`).  Run HERE (the build
container, where /root/reference exists):

    python tests/golden/make_golden_fit.py

The fixture holds the inputs (three nested shells of 4x6x8 / 6x10x12 / 10x12x16 cells, cell 0.2, four modes with seeded
complex pressures incl. unread odd entries) and the reference's Psi with and without power scaling, its map centre, and the
geometry its FFAT_Map_Serialize::Save writes for shell 2.  tests use it where /root/reference is absent (the GPU box).
"""
import os
import sys
import tempfile
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc, fatcube          # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    assert orc.ref() is not None, "oracle/_ref is not built: /root/reference missing?"
    cell = 0.2
    grid_low = np.array([-2.0, -2.0, -2.0]); dim = (20, 20, 20)
    boxes = [((8, 7, 6), (11, 12, 13)), ((7, 5, 4), (12, 14, 15)), ((5, 4, 2), (14, 15, 17))]   # cells lo..hi inclusive, all centred on 0
    Vs, nes = [], []
    for lo, hi in boxes:
        V, ne, _ = orc.ref_cubemap_mesh(lo, hi, cell, grid_low, dim)
        Vs.append(V); nes.append(ne)
    V = np.concatenate(Vs); n_elements = np.stack(nes)
    n_total = int(sum(int(a) * int(b) for ne in nes for a, b in ne))
    rng = np.random.default_rng(20261017)
    n_maps = 4
    k = np.array([1.5, 7.25, 31.0, 120.0])
    centres = V.reshape(-1, 4, 3).mean(axis=1)
    r = np.linalg.norm(centres, axis=1)
    P = np.empty((n_maps, 2 * n_total), dtype=np.complex128)
    for m in range(n_maps):
        amp = np.abs(rng.standard_normal(n_total)) + 0.2
        p = amp * np.exp(-1j * k[m] * r) / (k[m] * r) * (1.0 + 0.1 * rng.standard_normal(n_total))
        P[m, 0::2] = p
        P[m, 1::2] = rng.standard_normal(n_total) + 1j * rng.standard_normal(n_total)      # must never be read
    psi = {}; centre = None; saved = None
    for scaling in (False, True):
        rows = []
        for m in range(n_maps):
            with tempfile.TemporaryDirectory() as d:
                f = os.path.join(d, "m.fatcube")
                ps, c = orc.ref_ffat_fit(m, cell, V, n_elements, k[m], P[m], scaling, save_to=f)
                rows.append(ps); centre = c
                if saved is None:
                    saved = fatcube.load(f)
        psi[scaling] = np.array(rows)
    np.savez_compressed(os.path.join(OUT, "ffat_fit.npz"), cell_size=cell, V=V, n_elements=n_elements, k=k, pressure=P,
                        psi=psi[False], psi_scaled=psi[True], centre=centre,
                        shell2_lowcorners=np.asarray(saved["lowcorners"]), shell2_bboxlow=saved["bboxlow"], shell2_bboxtop=saved["bboxtop"],
                        shell2_strides=saved["strides"], shell2_n_elements=np.asarray(saved["n_elements"]))
    print("wrote ffat_fit.npz:", V.shape, n_elements.tolist(), psi[False].shape)


if __name__ == "__main__":
    main()
