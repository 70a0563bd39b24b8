"""the headless driver tools/pbso_render.cpp -- the reference tool's path conventions
(tools/real_time_modal_sound.cpp:478-501, .meta :389-397), BuildSolver, impulse script, WAV output with the
/1e10 stereo convention (:207-210).  CPU tests cover the host-only parts (OBJ reader, area-weighted vertex
normals, WAV writer, argument / path resolution, refusal to run without a GPU); the GPU test renders a script
and compares every sample with the CPU oracle."""
import os
import struct
import subprocess
import numpy as np
import pytest
from openpbso_b200 import synth
from conftest import ROOT, assert_waveform_parity

INC = os.path.join(ROOT, "include", "openpbso")
LIBDIR = os.path.join(ROOT, "openpbso_b200")
GXX = ["/usr/bin/g++", "-std=c++17", "-O2", "-pthread", "-Wall", "-Wno-sign-compare", "-I" + os.path.join(INC, "eigen_shim"), "-I" + INC]


@pytest.fixture(scope="module")
def render_exe(pbso, tmp_path_factory):
    out = str(tmp_path_factory.mktemp("tools") / "pbso_render")
    r = subprocess.run(GXX + [os.path.join(ROOT, "tools", "pbso_render.cpp"), "-L" + LIBDIR, "-lpbso_b200",
                              "-Wl,-rpath," + LIBDIR, "-o", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return out


@pytest.fixture(scope="module")
def fit_exe(pbso, tmp_path_factory):
    out = str(tmp_path_factory.mktemp("tools") / "pbso_fit_ffat")
    r = subprocess.run(GXX + [os.path.join(ROOT, "tools", "pbso_fit_ffat.cpp"), "-L" + LIBDIR, "-lpbso_b200",
                              "-Wl,-rpath," + LIBDIR, "-o", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return out


def _write_fit_inputs(tmp_path, w, binary):
    nfile = str(tmp_path / "n_elements.txt")
    with open(nfile, "w") as f:
        for shell in w["n_elements"]:
            f.write(" ".join("%d %d" % (a, b) for a, b in shell) + "\n")
    vfile = str(tmp_path / "V.f64"); np.ascontiguousarray(w["V"]).tofile(vfile)
    kfile = str(tmp_path / "k.txt")
    with open(kfile, "w") as f:
        for k in w["k"]:
            f.write("%.17g\n" % k)
    for m, P in enumerate(w["pressure"]):
        if binary:
            with open(str(tmp_path / ("p-%d.bin" % (m + 3))), "wb") as f:
                f.write(np.int32(2 * len(P)).tobytes()); f.write(np.ascontiguousarray(P).view(np.float64).tobytes())
        else:
            with open(str(tmp_path / ("p-%d.txt" % (m + 3))), "w") as f:
                for z in P:
                    f.write("%.17g %.17g\n" % (z.real, z.imag))
    return nfile, vfile, kfile


def test_fit_tool_arguments(fit_exe, tmp_path):
    r = subprocess.run([fit_exe], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr
    r = subprocess.run([fit_exe, "-n", str(tmp_path / "none.txt"), "-v", "x", "-c", "0.1", "-k", "k", "-p", "p%d", "-o", str(tmp_path / "o")],
                       capture_output=True, text=True)
    assert r.returncode == 3 and "cannot open" in r.stderr


def test_fit_tool_refuses_to_run_without_gpu(pbso, fit_exe, tmp_path):
    if pbso.device_count() > 0:
        pytest.skip("a CUDA device is present")
    w = synth.ffat_fit_workload(1, 5, half_cells=(1, 2, 3), cell_size=0.5)
    nfile, vfile, kfile = _write_fit_inputs(tmp_path, w, True)
    r = subprocess.run([fit_exe, "-n", nfile, "-v", vfile, "-c", "0.5", "-k", kfile, "-p", str(tmp_path / "p-%d.bin"), "-o", str(tmp_path / "o"),
                        "-first", "3", "-b"], capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU path" in r.stderr          # no fallback: the fit runs on the B200 or not at all
    assert not os.path.exists(str(tmp_path / "o"))


@pytest.mark.gpu
@pytest.mark.parametrize("binary,scaling", [(True, False), (False, True)])
def test_fit_tool_writes_the_maps_the_oracle_fits(pbso, orc, fit_exe, tmp_path, binary, scaling):
    """pbso_fit_ffat end to end: files in the reference's formats -> one batched fit on the GPU -> a .fatcube directory that
    LoadAll reads back; Psi and geometry against the oracle's constructor + Solve."""
    w = synth.ffat_fit_workload(4, 61, half_cells=((2, 3, 4), (3, 5, 6), (5, 6, 8)), cell_size=0.2, noise=0.02)
    nfile, vfile, kfile = _write_fit_inputs(tmp_path, w, binary)
    out = str(tmp_path / "maps")
    cmd = [fit_exe, "-n", nfile, "-v", vfile, "-c", repr(w["cell_size"]), "-k", kfile, "-p", str(tmp_path / ("p-%d." + ("bin" if binary else "txt"))),
           "-o", out, "-first", "3"] + (["-b"] if binary else []) + (["-s"] if scaling else [])
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "4 modes, 3 shells" in r.stdout
    fit = orc.ffat_fit_geometry(w["cell_size"], w["V"], w["n_elements"])
    want, _ = orc.ffat_fit_solve(fit, w["k"], w["pressure"], scaling)
    fm = pbso.FFATMaps.LoadAll(out)
    assert list(fm.mode_ids()) == [3, 4, 5, 6]
    for m in range(4):
        d = fm.get_map(3 + m)
        assert d["k"] == w["k"][m] and np.allclose(d["psi"], want[m], rtol=1e-12, atol=0)
        g, ig = fit["geom"][2], fit["igeom"][2]
        assert np.array_equal(np.asarray(d["lowcorners"]).ravel(), g[1:19]) and np.array_equal(d["bboxlow"], g[22:25])
        assert np.array_equal(d["strides"], ig[12:]) and np.array_equal(np.asarray(d["n_elements"]).ravel(), ig[:12])
    # -compress -legacy: the same maps through FFAT_Map::Compress, written in the igl::serialize form; the oracle's independent
    # reader finds _compressed_Psi = its own quantisation of the fitted Psi
    from oracle import fatcube
    out2 = str(tmp_path / "maps_cl")
    r = subprocess.run(cmd[:cmd.index("-o") + 1] + [out2] + cmd[cmd.index("-o") + 2:] + ["-compress", "-legacy"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for m in range(4):
        raw = open(os.path.join(out2, "%d.fatcube" % (3 + m)), "rb").read()
        assert fatcube.is_legacy(raw)
        d = fatcube.decode_legacy(raw)
        plain = fm.get_map(3 + m)
        q, amp, _ = orc.ffat_quantise(plain)
        assert d["is_compressed"] and d["modeid"] == 3 + m and np.array_equal(d["psi"], orc.ffat_dequantise(plain, q, amp))
    short = str(tmp_path / ("p-3." + ("bin" if binary else "txt")))
    with open(short, "wb") as f:                                  # a truncated pressure file: Solve's size assert
        f.write(np.int32(0).tobytes() if binary else b"")
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 4 and "wrong size" in r.stderr


@pytest.fixture(scope="module")
def units_exe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("tools") / "headless_units")
    r = subprocess.run(GXX + [os.path.join(ROOT, "tests", "cpp", "headless_units.cpp"), "-o", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return out


def _bumpy_sphere(n_lat=7, n_lon=10, seed=5):
    """A closed triangle mesh with unequal face areas (so area weighting matters): a perturbed UV sphere."""
    rng = np.random.default_rng(seed)
    V = [[0.0, 0.0, 1.0]]
    for i in range(1, n_lat):
        th = np.pi * i / n_lat
        for j in range(n_lon):
            ph = 2 * np.pi * j / n_lon
            r = 1.0 + 0.15 * rng.standard_normal()
            V.append([r * np.sin(th) * np.cos(ph), r * np.sin(th) * np.sin(ph), r * np.cos(th)])
    V.append([0.0, 0.0, -1.0])
    F = []
    ring = lambda i, j: 1 + (i - 1) * n_lon + (j % n_lon)
    for j in range(n_lon):
        F.append([0, ring(1, j), ring(1, j + 1)])
        F.append([len(V) - 1, ring(n_lat - 1, j + 1), ring(n_lat - 1, j)])
    for i in range(1, n_lat - 1):
        for j in range(n_lon):
            F.append([ring(i, j), ring(i + 1, j), ring(i + 1, j + 1)])
            F.append([ring(i, j), ring(i + 1, j + 1), ring(i, j + 1)])
    return np.array(V), np.array(F, dtype=np.int32)


def _write_obj(path, V, F, slashes=False):
    with open(path, "w") as f:
        f.write("# test mesh\n")
        for v in V:
            f.write("v %r %r %r\n" % tuple(float(x) for x in v))
        for t in F:
            f.write("f " + " ".join(("%d/%d/%d" % (i + 1, i + 1, i + 1)) if slashes else str(i + 1) for i in t) + "\n")


def _read_float_wav(path):
    raw = open(path, "rb").read()
    assert raw[:4] == b"RIFF" and raw[8:12] == b"WAVE"
    assert struct.unpack("<I", raw[4:8])[0] == len(raw) - 8
    off, fmt, data = 12, None, None
    while off < len(raw):
        tag, size = raw[off:off + 4], struct.unpack("<I", raw[off + 4:off + 8])[0]
        body = raw[off + 8:off + 8 + size]
        if tag == b"fmt ": fmt = struct.unpack("<HHIIHH", body[:16])
        if tag == b"data": data = np.frombuffer(body, dtype="<f4")
        off += 8 + size + (size & 1)
    return fmt, data


def test_obj_reader_and_vertex_normals_match_oracle(orc, units_exe, tmp_path):
    V, F = _bumpy_sphere()
    for slashes in (False, True):
        obj = str(tmp_path / ("m%d.obj" % slashes)); _write_obj(obj, V, F, slashes)
        out = str(tmp_path / "n.bin")
        assert subprocess.run([units_exe, "normals", obj, out]).returncode == 0
        raw = open(out, "rb").read()
        nv, nf = struct.unpack("<ii", raw[:8])
        assert (nv, nf) == (len(V), len(F))
        N = np.frombuffer(raw[8:]).reshape(nv, 3)
        Vo, Fo = orc.read_obj(obj)
        assert np.array_equal(Fo, F)
        assert np.allclose(N, orc.per_vertex_normals(Vo, Fo), rtol=0, atol=1e-15)
        assert np.allclose(np.linalg.norm(N, axis=1), 1.0)


def test_wav_writer_follows_the_callback_convention(units_exe, tmp_path):
    rng = np.random.default_rng(3)
    y = rng.standard_normal(5 * 256 + 17) * 0.4e10
    src = str(tmp_path / "y.f64"); y.tofile(src)
    out = str(tmp_path / "y.wav")
    assert subprocess.run([units_exe, "wav", src, "256", "0.5", out]).returncode == 0
    fmt, data = _read_float_wav(out)
    assert fmt == (3, 2, synth.SAMPLE_RATE if hasattr(synth, "SAMPLE_RATE") else 44100, 44100 * 8, 8, 32)
    want = (y / 1e10 * 0.5).astype(np.float32)          # tools/real_time_modal_sound.cpp:208
    assert data.size == 2 * y.size
    assert np.array_equal(data[0::2], want) and np.array_equal(data[1::2], want)
    from scipy.io import wavfile
    rate, arr = wavfile.read(out)
    assert rate == 44100 and arr.shape == (y.size, 2) and arr.dtype == np.float32


def test_driver_arguments_and_path_conventions(render_exe, tmp_path):
    r = subprocess.run([render_exe], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr
    script = str(tmp_path / "s.txt"); open(script, "w").write("run 1\n")
    # -d DIR without -name: the object name is parsed from the first *.tet.obj (tools/...cpp:484-486)
    d = tmp_path / "data"; d.mkdir()
    r = subprocess.run([render_exe, "-d", str(d), "-script", script], capture_output=True, text=True)
    assert r.returncode == 2 and "no *.tet.obj" in r.stderr
    (d / "bell.tet.obj").write_text("v 0 0 0\n")
    r = subprocess.run([render_exe, "-d", str(d), "-script", script], capture_output=True, text=True)
    assert "object name: bell" in r.stdout and r.returncode == 3 and "bell_material.txt" in r.stderr
    r = subprocess.run([render_exe, "-d", str(d), "-name", "gong", "-script", script], capture_output=True, text=True)
    assert "object name: gong" in r.stdout and "gong_material.txt" in r.stderr
    # .meta: four lines obj / modes / material / ffat (tools/...cpp:389-397, assets/meta/*.meta)
    meta = tmp_path / "x.meta"; meta.write_text("/a/o.tet.obj\n/a/o_surf.modes\n/a/glass.txt\n/a/ffat_maps_new\n")
    r = subprocess.run([render_exe, "-meta", str(meta), "-script", script], capture_output=True, text=True)
    assert r.returncode == 3 and "/a/glass.txt" in r.stderr
    r = subprocess.run([render_exe, "-meta", str(meta), "-script", script, "-buf", "100"], capture_output=True, text=True)
    assert r.returncode == 2 and "-buf" in r.stderr


def _write_object(d, name, orc, seed=77):
    from oracle import fatcube
    V, F = _bumpy_sphere()
    M = 48
    mat = synth.MATERIALS["high_damping"]
    freqs = synth.mode_frequencies(M, seed, 150.0, 16000.0)
    w2 = synth.omega_squared(freqs, mat["density"])
    U = synth.mode_shapes(M, 3 * len(V), seed + 1) * 1e6
    _write_obj(os.path.join(d, name + ".tet.obj"), V, F)
    with open(os.path.join(d, name + "_material.txt"), "w") as f:
        f.write("%r %r %r %r %r\n" % (mat["density"], mat["youngsModulus"], mat["poissonRatio"], mat["alpha"], mat["beta"]))
    orc.modes_write(os.path.join(d, name + "_surf.modes"), w2, U)
    fdir = os.path.join(d, name + "_ffat_maps"); os.makedirs(fdir, exist_ok=True)
    maps = synth.ffat_maps(freqs, 2000, n=8)
    for m in maps:
        fatcube.save(os.path.join(fdir, "mode-%03d.fatcube" % m["modeid"]), m)
    return dict(V=V, F=F, M=M, mat=mat, w2=w2, U=U, maps=maps)


SCRIPT = """# listener, a click, a few buffers, a gaussian tap, a moved listener
listener 0.4 3.0 -2.5
hit 12
run 3
gauss 700 30 0.0 0.6 0.8
run 2
listener -5.0 1.0 2.0
face 1 5 9 0.2 0.3 0.5 0 0 1
run 4
clear
run 2
until 0.1
"""


def _oracle_script(case, orc, BUF):
    a, b = orc.build_ab(case["mat"]["density"], case["w2"], case["mat"]["alpha"], case["mat"]["beta"], case["M"])
    s = orc.Solver(orc.Integrator(synth.H, a, b), BUF)
    U, N = case["U"], case["M"]
    VN = orc.per_vertex_normals(case["V"], case["F"])
    out = []
    def run(n):
        for _ in range(n):
            r = s.step()
            if r is not None: out.append(r[0])
    s.enqueue_trans(orc.ffat_eval(case["maps"], np.array([0.4, 3.0, -2.5]))[0])
    s.enqueue_force(orc.project_vertex(U, 12, VN[12], N), orc.POINT); run(3)
    s.enqueue_force(orc.project_vertex(U, 30, [0.0, 0.6, 0.8], N), orc.GAUSSIAN, width_us=700.0); run(2)
    s.enqueue_trans(orc.ffat_eval(case["maps"], np.array([-5.0, 1.0, 2.0]))[0])
    s.enqueue_force(orc.project_face(U, [1, 5, 9], [0.2, 0.3, 0.5], [0, 0, 1], N), orc.POINT); run(4)
    s.enqueue_force(np.zeros(N), orc.POINT, flags=orc.F_CLEAR); run(2)
    while len(out) * BUF < 0.1 * 44100: run(1)
    return np.concatenate(out)


def test_driver_refuses_to_run_without_gpu(pbso, orc, render_exe, tmp_path):
    if pbso.device_count() > 0:
        pytest.skip("a CUDA device is present")
    d = str(tmp_path); _write_object(d, "bell", orc)
    script = os.path.join(d, "s.txt"); open(script, "w").write(SCRIPT)
    r = subprocess.run([render_exe, "-d", d, "-script", script, "-buf", "256"], capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU path" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("BUF", [256, 513])
def test_driver_renders_script_like_the_oracle(pbso, orc, render_exe, tmp_path, BUF):
    d = str(tmp_path); case = _write_object(d, "bell", orc)
    script = os.path.join(d, "s.txt"); open(script, "w").write(SCRIPT)
    wav, raw = os.path.join(d, "out.wav"), os.path.join(d, "out.f64")
    r = subprocess.run([render_exe, "-d", d, "-script", script, "-buf", str(BUF), "-o", wav, "-raw", raw, "-stats"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "modes audible: 48 of 48" in r.stdout and "step latency" in r.stdout
    y = np.fromfile(raw)
    ref = _oracle_script(case, orc, BUF)
    assert y.size == ref.size
    assert np.max(np.abs(y - ref)) <= 1e-9 * np.max(np.abs(ref))
    assert_waveform_parity(y, ref)
    fmt, data = _read_float_wav(wav)
    assert fmt[:3] == (3, 2, 44100)
    want = (y / 1e10).astype(np.float32)
    assert np.array_equal(data[0::2], want) and np.array_equal(data[1::2], want)


SCRIPT_LONG = """# every command of the driver: impulses in a row (one is dequeued per buffer), a sustained autoregressive force with a
# parameter change, gaussian taps overlapping an impulse, listener moves in mid-render, the unit transfer, clearAllForces
listener 0.4 3.0 -2.5
hit 12
hit 7
point 3 0.0 1.0 0.0
run 9
gauss 1500 30 0.0 0.6 0.8
run 1
hit 20
run 6
listener -5.0 1.0 2.0
run 2
listener 2.0 2.0 2.0
listener 9.0 9.0 9.0
face 1 5 9 0.2 0.3 0.5 0 0 1
run 5
ar_start 5 0.0 0.0 1.0
run 2
arprm 0.7 0.2 0.002 0.1
ar_data 6 1.0 0.0 0.0
run 3
ar_end
run 7
unit_transfer
hit 2
run 4
use_transfer
listener 1.0 -4.0 0.5
hit 9
run 3
clear
run 2
until 0.45
"""


@pytest.mark.gpu
@pytest.mark.parametrize("BUF,prec", [(256, "tc3x"), (513, "tc3x"), (256, "f32"), (513, "f64")])
def test_driver_batch_mode(pbso, orc, render_exe, tmp_path, BUF, prec):
    """pbso_render -batch: the script planned on the host and rendered on the batch path (tools/offline_render.h) -- impulse
    stretches as stateful batch ranges, transfer swaps as range boundaries, Gaussian / autoregressive buffers on the
    per-buffer path in between.  The short script against the oracle's solver, the long one (every command) against the
    tool's own real-time path, which the test above pins to the oracle."""
    d = str(tmp_path); case = _write_object(d, "bell", orc)
    tol = dict(rel=1e-9, mx=1e-9) if prec == "f64" else {}
    for name, text in (("s", SCRIPT), ("l", SCRIPT_LONG)):
        script = os.path.join(d, name + ".txt"); open(script, "w").write(text)
        live, off = os.path.join(d, name + "_live.f64"), os.path.join(d, name + "_off.f64")
        r = subprocess.run([render_exe, "-d", d, "-script", script, "-buf", str(BUF), "-raw", live], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        r = subprocess.run([render_exe, "-d", d, "-script", script, "-buf", str(BUF), "-raw", off, "-batch", "-prec", prec, "-block", "32"],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert "offline:" in r.stdout and "range(s)" in r.stdout
        y_live, y_off = np.fromfile(live), np.fromfile(off)
        assert y_live.size == y_off.size and np.abs(y_live).max() > 0
        assert_waveform_parity(y_off, y_live, **tol)
        if name == "s":
            assert_waveform_parity(y_off, _oracle_script(case, orc, BUF), **tol)


@pytest.mark.gpu
def test_fit_tool_feeds_the_render_tool(pbso, orc, fit_exe, render_exe, tmp_path):
    """The two headless tools chained: shell pressures -> pbso_fit_ffat -> .fatcube directory -> pbso_render (-m -s -t -p),
    against the oracle's own chain (fit -> GetMapVal -> ModalSolver::step)."""
    d = str(tmp_path); case = _write_object(d, "bell", orc)
    M = case["M"]
    w = synth.ffat_fit_workload(M, 67, half_cells=(3, 5, 6), cell_size=0.25)
    mat = case["mat"]
    freqs = np.sqrt(case["w2"] / mat["density"]) / (2 * np.pi)
    w["k"] = 2 * np.pi * freqs / synth.SPEED_OF_SOUND                        # the object's own wavenumbers
    fitdir = tmp_path / "fit"; fitdir.mkdir()
    nfile, vfile, kfile = _write_fit_inputs(fitdir, w, True)
    maps_dir = str(tmp_path / "fitted_maps")
    # the helper names the pressure files from mode id 3; this object's modes are 0..M-1
    for m in range(M):
        os.rename(str(fitdir / ("p-%d.bin" % (m + 3))), str(fitdir / ("q-%d.bin" % m)))
    r = subprocess.run([fit_exe, "-n", nfile, "-v", vfile, "-c", repr(w["cell_size"]), "-k", kfile, "-p", str(fitdir / "q-%d.bin"),
                        "-o", maps_dir, "-b", "-s"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    script = os.path.join(d, "s.txt"); open(script, "w").write(SCRIPT)
    raw = os.path.join(d, "out.f64")
    r = subprocess.run([render_exe, "-m", os.path.join(d, "bell.tet.obj"), "-s", os.path.join(d, "bell_surf.modes"),
                        "-t", os.path.join(d, "bell_material.txt"), "-p", maps_dir, "-script", script, "-buf", "256", "-raw", raw],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    fit = orc.ffat_fit_geometry(w["cell_size"], w["V"], w["n_elements"])
    psi, _ = orc.ffat_fit_solve(fit, w["k"], w["pressure"], True)
    g, ig = fit["geom"][2], fit["igeom"][2]
    case["maps"] = [dict(cellsize=g[0], lowcorners=g[1:19].reshape(6, 3), center1=g[19:22], bboxlow=g[22:25], bboxtop=g[25:28],
                         center=g[28:31], k=w["k"][m], n_elements=ig[:12].reshape(6, 2), strides=ig[12:], psi=psi[m], modeid=m) for m in range(M)]
    y = np.fromfile(raw); ref = _oracle_script(case, orc, 256)
    assert y.size == ref.size
    assert np.max(np.abs(y - ref)) <= 1e-9 * np.max(np.abs(ref))
