"""The C++ header mirror (include/openpbso/) used the way the reference's tool uses the reference headers:
tests/cpp/drop_in_main.cpp is compiled against it, linked to libpbso_b200.so and driven by a script; every
buffer it produces is compared with the CPU oracle's ModalSolver restatement."""
import os
import subprocess
import numpy as np
import pytest
from openpbso_b200 import synth
from conftest import ROOT

BUF = 256


@pytest.fixture(scope="module")
def drop_in_exe(pbso, tmp_path_factory):
    out = str(tmp_path_factory.mktemp("cpp") / "drop_in_main")
    inc = os.path.join(ROOT, "include", "openpbso")
    libdir = os.path.join(ROOT, "openpbso_b200")
    cmd = ["/usr/bin/g++", "-std=c++17", "-O2", "-Wall", "-Wno-sign-compare", "-I" + os.path.join(inc, "eigen_shim"), "-I" + inc,
           os.path.join(ROOT, "tests", "cpp", "drop_in_main.cpp"), "-L" + libdir, "-lpbso_b200", "-Wl,-rpath," + libdir, "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return out


def _write_case(d, orc, seed=31):
    """64 modes (4 above the audible threshold), 120 vertices, 60 FFAT maps in one shared geometry."""
    from oracle import fatcube
    M, V = 64, 120
    mat = synth.MATERIALS["high_damping"]
    freqs = synth.mode_frequencies(M, seed, 100.0, 19000.0)
    w2 = synth.omega_squared(freqs, mat["density"])
    U = synth.mode_shapes(M, 3 * V, seed + 1) * 1e6
    os.makedirs(os.path.join(d, "ffat"), exist_ok=True)
    with open(os.path.join(d, "material.txt"), "w") as f:
        f.write("# density E nu alpha beta\n%r %r %r %r %r\n" % (mat["density"], mat["youngsModulus"], mat["poissonRatio"], mat["alpha"], mat["beta"]))
    orc.modes_write(os.path.join(d, "object.modes"), w2, U)
    thr = 15000.0
    open(os.path.join(d, "ffat", "freq_threshold.txt"), "w").write("%r\n" % thr)
    N = int(np.sum(freqs <= thr))
    maps = synth.ffat_maps(freqs[:N], 2000, n=8)
    for m in maps:
        fatcube.save(os.path.join(d, "ffat", "mode-%03d.fatcube" % m["modeid"]), m)
    return dict(M=M, V=V, mat=mat, w2=w2, U=U, N=N, maps=maps, thr=thr)


SCRIPT = [
    ("listener", [0.5, 4.0, -2.0]), ("point", [3, 0.0, 0.6, 0.8]), ("none", []), ("gauss", [900.0, 17, 1.0, 0.0, 0.0]),
    ("face", [1, 5, 9, 0.2, 0.3, 0.5, 0.0, 0.0, 1.0]), ("listener", [6.0, 1.0, 1.0]), ("none", []), ("clear", []), ("none", []),
    ("ar_start", [40, 0.6, 0.0, 0.8]), ("none", []), ("arprm", [0.7, 0.2, 0.002, 0.1]), ("ar_data", [41, 0.0, 1.0, 0.0]),
    ("ar_end", [42, 1.0, 0.0, 0.0]), ("unit_transfer", []), ("point", [100, 0.0, 0.0, 1.0]), ("use_transfer", []),
    ("listener", [-3.0, -3.0, 5.0]), ("none", []), ("none", []),
]


def _oracle_run(case, orc):
    a, b = orc.build_ab(case["mat"]["density"], case["w2"], case["mat"]["alpha"], case["mat"]["beta"], case["N"])
    s = orc.Solver(orc.Integrator(synth.H, a, b), BUF)
    U, N = case["U"], case["N"]
    outs = []
    for kind, x in SCRIPT:
        if kind == "listener": s.enqueue_trans(orc.ffat_eval(case["maps"], np.array(x))[0])
        elif kind == "point": s.enqueue_force(orc.project_vertex(U, int(x[0]), x[1:], N), orc.POINT)
        elif kind == "gauss": s.enqueue_force(orc.project_vertex(U, int(x[1]), x[2:], N), orc.GAUSSIAN, width_us=x[0])
        elif kind == "face": s.enqueue_force(orc.project_face(U, [int(v) for v in x[:3]], x[3:6], x[6:], N), orc.POINT)
        elif kind == "clear": s.enqueue_force(np.zeros(N), orc.POINT, flags=orc.F_CLEAR)
        elif kind == "ar_start": s.enqueue_force(orc.project_vertex(U, int(x[0]), x[1:], N), orc.AR, flags=orc.F_SUSTAIN_START)
        elif kind == "ar_data": s.enqueue_force(orc.project_vertex(U, int(x[0]), x[1:], N), orc.AR)
        elif kind == "ar_end": s.enqueue_force(orc.project_vertex(U, int(x[0]), x[1:], N), orc.AR, flags=orc.F_SUSTAIN_END)
        elif kind == "arprm": s.enqueue_arprm(*x)
        elif kind == "unit_transfer": s.set_use_transfer(False)
        elif kind == "use_transfer": s.set_use_transfer(True)
        outs.append(s.step())
    return outs


def _write_script(d):
    with open(os.path.join(d, "script.txt"), "w") as f:
        for kind, x in SCRIPT:
            f.write(" ".join([kind] + [repr(v) for v in x]) + "\n")


def test_header_mirror_compiles_and_refuses_to_run_without_gpu(pbso, orc, drop_in_exe, tmp_path):
    if pbso.device_count() > 0:
        pytest.skip("a CUDA device is present")
    d = str(tmp_path); _write_case(d, orc); _write_script(d)
    r = subprocess.run([drop_in_exe, d, os.path.join(d, "out.bin")], capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU path" in r.stderr


@pytest.mark.gpu
def test_drop_in_caller_matches_oracle(pbso, orc, drop_in_exe, tmp_path):
    d = str(tmp_path); case = _write_case(d, orc); _write_script(d)
    r = subprocess.run([drop_in_exe, d, os.path.join(d, "out.bin")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    raw = open(os.path.join(d, "out.bin"), "rb").read()
    N = int(np.frombuffer(raw[:4], dtype=np.int32)[0]); off = 4
    assert N == case["N"] == 60 or N == case["N"]                 # numModesAudible cull via freq_threshold.txt
    ref = _oracle_run(case, orc)
    full = max(np.max(np.abs(r_[0])) for r_ in ref if r_ is not None)
    n_buf = 0
    for (kind, _), want in zip(SCRIPT, ref):
        produced = int(np.frombuffer(raw[off:off + 4], dtype=np.int32)[0]); off += 4
        assert bool(produced) == (want is not None), kind       # clearAllForces yields no buffer
        if not produced:
            continue
        y = np.frombuffer(raw[off:off + 8 * BUF]); off += 8 * BUF
        qn = np.frombuffer(raw[off:off + 8 * N]); off += 8 * N
        assert np.max(np.abs(y - want[0])) <= 1e-9 * full, (kind, n_buf)
        assert np.allclose(qn, want[1], rtol=1e-8, atol=1e-9 * np.max(want[1]) + 1e-300), (kind, n_buf)
        n_buf += 1
    tr = np.frombuffer(raw[off:off + 8 * N])
    assert np.allclose(tr, orc.ffat_eval(case["maps"], np.array([2.0, -3.0, 4.0]))[0], rtol=1e-12)
    assert n_buf == len(SCRIPT) - 1


def test_ffat_construction_mirror_compiles_and_refuses_to_run_without_gpu(pbso, tmp_path):
    """tests/cpp/ffat_fit_main.cpp -- FFAT_Map<double,3>(modeId, cellSize, V, N_elements), Solve, ReadNElementsFile,
    ReadComplexVector, FFAT_Map_Serialize::Save/Load/Check, GetMapVal through the header mirror -- builds on a machine
    without a GPU, and there the constructor throws instead of falling back to a CPU path."""
    from openpbso_b200 import synth
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    inc = os.path.join(root, "include", "openpbso"); libdir = os.path.join(root, "openpbso_b200")
    exe = str(tmp_path / "ffat_fit_main")
    r = subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-Wall", "-Wno-sign-compare", "-I" + os.path.join(inc, "eigen_shim"), "-I" + inc,
                        os.path.join(root, "tests", "cpp", "ffat_fit_main.cpp"), "-L" + libdir, "-lpbso_b200", "-Wl,-rpath," + libdir, "-o", exe],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    if pbso.device_count() > 0:
        pytest.skip("a CUDA device is present")
    w = synth.ffat_fit_workload(1, 5, half_cells=(1, 2, 3), cell_size=0.5)
    nfile = str(tmp_path / "n.txt")
    with open(nfile, "w") as f:
        for shell in w["n_elements"]:
            f.write(" ".join("%d %d" % (a, b) for a, b in shell) + "\n")
    vfile = str(tmp_path / "V.f64"); np.ascontiguousarray(w["V"]).tofile(vfile)
    pfile = str(tmp_path / "p.bin")
    with open(pfile, "wb") as f:
        f.write(np.int32(2 * w["pressure"].shape[1]).tobytes()); f.write(np.ascontiguousarray(w["pressure"][0]).view(np.float64).tobytes())
    r = subprocess.run([exe, nfile, vfile, "0.5", "0", repr(float(w["k"][0])), pfile, "1", "0", str(tmp_path / "o.fatcube"), "1", "2", "3"],
                       capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU path" in r.stderr
