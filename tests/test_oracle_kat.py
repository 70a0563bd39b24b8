"""Known-answer tests that anchor the CPU oracle (the reference ships no tests: SURVEY.md section 4)."""
import os
import numpy as np
import pytest
from openpbso_b200 import synth

H = synth.H


def test_impulse_response_closed_form(orc):
    """q[k] = c3 eps^k sin((k+1) theta)/sin(theta) for a unit impulse at k=0 (modal_integrator.h:95-97,109)."""
    mat = synth.MATERIALS["low_damping"]
    f = np.array([80.0, 440.0, 3000.0, 18000.0])
    a, b = orc.build_ab(mat["density"], synth.omega_squared(f, mat["density"]), mat["alpha"], mat["beta"])
    c1, c2, c3 = orc.coeffs(H, a, b)
    eps = np.exp(-a / 2 * H); theta = H * np.sqrt(b - a * a / 4)
    assert np.allclose(c1, 2 * eps * np.cos(theta), rtol=1e-15)
    assert np.allclose(c2, -eps ** 2, rtol=1e-15)
    integ = orc.Integrator(H, a, b)
    n = 44100
    q = np.empty((n, len(f)))
    q[0] = integ.step(np.ones(len(f)))
    for k in range(1, n):
        q[k] = integ.step()
    kk = np.arange(n)[:, None]
    closed = c3 * eps ** kk * np.sin((kk + 1) * theta) / np.sin(theta)
    rel = np.linalg.norm(q - closed, axis=0) / np.linalg.norm(closed, axis=0)
    assert np.all(rel < 5e-9), rel


def test_c3_independent_formula(orc):
    a = np.array([1.5, 40.0, 300.0]); b = np.array([4e5, 9e7, 1.2e10])
    _, _, c3 = orc.coeffs(H, a, b)
    eps = np.exp(-a * H / 2); wd = np.sqrt(b - a * a / 4); th = H * wd; w = np.sqrt(b); g = np.arcsin(a / (2 * w))
    ref = 2 * (eps * np.cos(th + g) - eps ** 2 * np.cos(2 * th + g)) / (3 * w * wd) * 1e9
    assert np.allclose(c3, ref, rtol=1e-13)


def test_step_without_force_equals_zero_force(orc):
    a = np.array([2.0, 3.0]); b = np.array([1e6, 4e7])
    i1 = orc.Integrator(H, a, b); i2 = orc.Integrator(H, a, b)
    i1.step(np.array([1.0, -2.0])); i2.step(np.array([1.0, -2.0]))
    for _ in range(10):
        assert np.array_equal(i1.step(), i2.step(np.zeros(2)))


def _one_map(n=8, R=1.5, seed=5, k=2.0, center=(0.1, -0.2, 0.3)):
    g = synth.ffat_geometry(R, n, center)
    d = dict(g); d.update(k=k, psi=np.random.default_rng(seed).uniform(0.2, 3.0, 6 * n * n), modeid=0)
    return d


def test_ffat_texel_centre_identity(orc):
    """At a texel centre GetMapVal = |Psi[idx] / (k r)| exactly (weights (1,0,0,0))."""
    m = _one_map()
    pts = synth.texel_centres(m)
    c = m["center"]
    for idx in [0, 7, 63, 64, 100, 6 * 64 - 1, 200, 333]:
        p = c + 2.5 * (pts[idx] - c)
        val = orc.ffat_eval([m], p)[0, 0]
        r = np.linalg.norm(p - c)
        assert val == pytest.approx(abs(m["psi"][idx] / (m["k"] * r)), rel=1e-12)


def test_ffat_inverse_r_law_and_continuity(orc):
    m = _one_map()
    rng = np.random.default_rng(3)
    d = synth.unit_vectors(64, 9)
    c = m["center"]
    v1 = orc.ffat_eval([m], c + 4.0 * d)[:, 0]
    v2 = orc.ffat_eval([m], c + 8.0 * d)[:, 0]
    assert np.allclose(v1, 2.0 * v2, rtol=1e-12)
    # continuity inside a face: a 1e-9 step moves the value by O(1e-9)
    p = c + 5.0 * d
    dv = orc.ffat_eval([m], p + 1e-9 * rng.standard_normal(p.shape))[:, 0] - orc.ffat_eval([m], p)[:, 0]
    assert np.max(np.abs(dv)) < 1e-6


def test_ffat_face_selection_and_indices(orc):
    m = _one_map(n=4, R=1.0, center=(0, 0, 0))
    for f, p in enumerate([(5, .1, .2), (-5, .1, .2), (.1, 5, .2), (.1, -5, .2), (.1, .2, 5), (.1, .2, -5)]):
        surf, ind = orc.ffat_intersect(m, np.array(p, dtype=float))
        assert ind[0] == f
        dk = f // 2
        assert abs(abs(surf[dk]) - 1.0) < 1e-12
        idx, co = orc.ffat_interpolate(m, surf, ind)
        assert co.sum() == pytest.approx(1.0, abs=1e-15) and np.all(co >= 0)
        assert np.all(idx[:, 0] == f) and np.all(idx[:, 1:] >= 0) and np.all(idx[:, 1:] < 4)


def test_num_modes_audible_and_cache_quirk(orc):
    rho = 2600.0
    f = np.array([100.0, 1000.0, 5000.0, 15000.0, 21000.0])
    w2 = synth.omega_squared(f, rho)
    assert orc.num_modes_audible(w2, rho, 20000.0) == 4
    assert orc.num_modes_audible(w2, rho, 50.0) == 0
    assert orc.num_modes_audible(w2, rho, 30000.0) == 5
    assert orc.num_modes_audible(np.array([]), rho, 1000.0) == 0
    cache = np.array([-1, 22100., -1.])
    assert orc.num_modes_audible(w2, rho, 6000.0, cache) == 3 and cache[0] == 3
    # early-return branches do not refresh the cache (ModeData.h:131-136) ...
    assert orc.num_modes_audible(w2, rho, 30000.0, cache) == 5 and cache[0] == 3
    # ... so a repeated query with the cached key is served from it
    assert orc.num_modes_audible(w2[:2], rho, 6000.0, cache) == 3


def test_material_and_modes_files(orc, tmp_path):
    p = tmp_path / "mat.txt"
    p.write_text("# comment\n# another\n2600 6.2e10 0.2 1.0 1e-7\n")
    m = orc.material_read(str(p))
    assert m == dict(density=2600.0, youngsModulus=6.2e10, poissonRatio=0.2, alpha=1.0, beta=1e-7)
    assert orc.material_read(str(tmp_path / "missing.txt")) is None
    U = synth.mode_shapes(5, 12, 1); w2 = np.arange(1.0, 6.0)
    orc.modes_write(str(tmp_path / "x.modes"), w2, U)
    raw = open(tmp_path / "x.modes", "rb").read()
    assert np.frombuffer(raw[:8], dtype=np.int32).tolist() == [12, 5]          # nDOF, nModes (ModeData.h:66-68)
    w2b, Ub = orc.modes_read(str(tmp_path / "x.modes"))
    assert np.array_equal(w2, w2b) and np.array_equal(U, Ub)


def test_force_profiles(orc):
    out, alive = orc.force_profile(orc.POINT, 0.0, 256, 3)
    assert alive.tolist() == [1, 0, 0] and out[0, 0] == 1.0 and out.sum() == 1.0
    width_us = 900.0
    ws = max(1, int(width_us / 1e6 * 44100)); centre = int(4.5 * ws)
    out, alive = orc.force_profile(orc.GAUSSIAN, width_us, 256, 4)
    n_alive = int(np.ceil(10 * ws / 256))
    assert alive.tolist() == [1] * n_alive + [0] * (4 - n_alive)
    i = np.arange(256 * n_alive)
    assert np.allclose(out[:n_alive].ravel(), np.exp(-0.5 * ((i - centre) / ws) ** 2), rtol=1e-14)
    out0, alive0 = orc.force_profile(orc.GAUSSIAN, 0.0, 256, 1)
    assert alive0[0] == 0 and not out0.any()
    ar, alive = orc.force_profile(orc.AR, 0.0, 256, 2)
    assert alive.tolist() == [1, 1] and abs(ar.mean() - 0.142) < 0.01 and ar.std() < 0.05


def test_projection(orc):
    U = synth.mode_shapes(7, 30, 2)
    vn = np.array([0.6, -0.8, 0.0])
    assert np.allclose(orc.project_vertex(U, 4, vn), U[:, 12:15] @ vn, rtol=1e-14)
    vids = [1, 5, 9]; bc = np.array([0.2, 0.3, 0.5])
    ref = sum(bc[j] * (U[:, 3 * v:3 * v + 3] @ vn) for j, v in enumerate(vids))
    assert np.allclose(orc.project_face(U, vids, bc, vn), ref, rtol=1e-13)
    F = np.random.default_rng(0).standard_normal((3, 30))
    assert np.allclose(orc.project_dense(U, F), F @ U.T, rtol=1e-12)


def test_solver_quirks(orc):
    """H4 quirks: clearAllForces yields no buffer; unit transfer = 1e7; impulse lands on sample 0."""
    a = np.array([2.0, 3.0]); b = np.array([1e6, 4e7])
    s = orc.Solver(orc.Integrator(H, a, b), 64)
    s.enqueue_force(np.array([1.0, 0.5]))
    y, qn = s.step()
    _, _, c3 = orc.coeffs(H, a, b)
    assert y[0] == pytest.approx(1e7 * (c3[0] * 1.0 + c3[1] * 0.5), rel=1e-14)
    s.enqueue_force(np.zeros(2), flags=orc.F_CLEAR)
    assert s.step() is None and s.num_active() == 0
    assert s.step() is not None
    # queue capacities (readerwriterqueue.h:101): trans holds 1, force holds 1023
    assert s.enqueue_trans(np.ones(2)) and not s.enqueue_trans(np.ones(2))
    n = 0
    while s.enqueue_force(np.zeros(2)):
        n += 1
    assert n == 1023


def test_golden_cfg1_reproducible(orc, golden_dir):
    g = np.load(os.path.join(golden_dir, "cfg1_ball.npz"))
    mat = synth.MATERIALS["low_damping"]
    freqs = synth.mode_frequencies(64, 1001)
    a, b = orc.build_ab(mat["density"], synth.omega_squared(freqs, mat["density"]), mat["alpha"], mat["beta"])
    assert np.allclose(orc.project_vertex(g["u_vid"], 0, g["vn"]), g["space"], rtol=1e-14)
    trans = orc.ffat_eval(synth.ffat_maps(freqs, 2000), g["listener"])[0]
    assert np.allclose(trans, g["trans"], rtol=1e-13)
    s = orc.Solver(orc.Integrator(H, a, b), 256)
    s.enqueue_trans(trans); s.enqueue_force(g["space"] * g["scale"])
    y = np.concatenate([s.step()[0] for _ in range(173)])
    assert np.max(np.abs(y - g["y"])) / np.max(np.abs(g["y"])) < 1e-11
    assert np.max(np.abs(g["y"])) / 1e10 == pytest.approx(0.5, rel=1e-6)


# --------------------------------------------------------------------------- FFAT map construction
def test_ffat_fit_uniform_field_closed_form(orc):
    """Pressure of constant magnitude A and phase on concentric cube shells of half-widths a_s: every shell sees
    |p_s| = A at radius r_s = r_2 a_s/a_2, so the one-column least squares (ffat_solver.h:881-895) gives
    psi = A k r_2 (sum_s 1/rho_s) / (sum_s 1/rho_s^2), rho_s = a_s/a_2; and power scaling (:909-929) rescales so
    that sum |p_0|^2 = sum (psi/(k r_0))^2."""
    half = (4, 6, 8); cell = 0.125; A = 3.5; k = 7.25
    Vs, nes = zip(*[synth.cubemap_vertices((0, 0, 0), h, cell) for h in half])
    V = np.concatenate(Vs); ne = np.stack(nes)
    fit = orc.ffat_fit_geometry(cell, V, ne)
    assert fit["n_total"] == 6 * 4 * (16 + 36 + 64) and fit["n_dir"] == 6 * 256
    assert list(fit["strides"]) == [0, 6 * 64, 6 * 64 + 6 * 144]
    P = np.full((1, 2 * fit["n_total"]), A * np.exp(0.3j))
    P[0, 1::2] = 99.0                                  # the odd (second-triangle) entries must never be read (:1054)
    psi, scale, R, Pabs = orc.ffat_fit_solve(fit, [k], P, False, want_R=True)
    assert np.allclose(Pabs, A, rtol=1e-14)
    rho = np.array(half) / half[2]
    assert np.allclose(R, R[:, 2:3] * rho, rtol=1e-14)
    centres = V[4 * fit["strides"][2]:].reshape(-1, 4, 3).mean(axis=1)
    assert np.allclose(R[:, 2], np.linalg.norm(centres, axis=1), rtol=1e-14)
    want = A * k * R[:, 2] * np.sum(1 / rho) / np.sum(1 / rho ** 2)
    assert np.allclose(psi[0], want, rtol=1e-13)
    psi_s, scale_s = orc.ffat_fit_solve(fit, [k], P, True)
    s = np.sqrt(fit["n_dir"] * A * A / np.sum((psi[0] / (k * R[:, 0])) ** 2))
    assert np.isclose(scale_s[0], s, rtol=1e-13) and np.allclose(psi_s[0], psi[0] * s, rtol=1e-13)


def test_ffat_fit_single_shell_weighting(orc):
    """With the pressure zero on shells 0 and 1, only shell 2 contributes to u.b: psi = (|p_2|/(k r_2)) / sum_s 1/(k r_s)^2;
    at shell 2 the stencil sits on texel centres, so |p_2| is the sample itself."""
    half = (3, 4, 5); cell = 0.2; k = 2.0
    Vs, nes = zip(*[synth.cubemap_vertices((0, 0, 0), h, cell) for h in half])
    fit = orc.ffat_fit_geometry(cell, np.concatenate(Vs), np.stack(nes))
    rng = np.random.default_rng(4)
    P = np.zeros((1, 2 * fit["n_total"]), dtype=complex)
    vals = rng.standard_normal(fit["n_dir"]) + 1j * rng.standard_normal(fit["n_dir"])
    P[0, 2 * fit["strides"][2]::2] = vals
    psi, _, R, Pabs = orc.ffat_fit_solve(fit, [k], P, False, want_R=True)
    assert np.allclose(Pabs[0, :, 2], np.abs(vals), rtol=1e-14) and np.all(Pabs[0, :, :2] == 0)
    want = (np.abs(vals) / (k * R[:, 2])) / np.sum(1 / (k * R) ** 2, axis=1)
    assert np.allclose(psi[0], want, rtol=1e-13)


def test_ffat_fit_oracle_reproduces_the_reference_fixture(orc, golden_dir):
    """tests/golden/ffat_fit.npz was produced by the reference's own CubemapMesh + FFAT_Map<double,3> constructor + Solve
    (tests/golden/make_golden_fit.py); the restatement must reproduce it without the reference being present."""
    g = np.load(os.path.join(golden_dir, "ffat_fit.npz"))
    fit = orc.ffat_fit_geometry(float(g["cell_size"]), g["V"], g["n_elements"])
    assert np.array_equal(fit["geom"][2][1:19], g["shell2_lowcorners"].ravel())
    assert np.array_equal(fit["geom"][2][22:25], g["shell2_bboxlow"]) and np.array_equal(fit["geom"][2][25:28], g["shell2_bboxtop"])
    assert np.array_equal(fit["igeom"][2][12:], g["shell2_strides"]) and np.array_equal(fit["geom"][2][28:31], g["centre"])
    for scaling, key in ((False, "psi"), (True, "psi_scaled")):
        psi, _ = orc.ffat_fit_solve(fit, g["k"], g["pressure"], scaling)
        assert np.allclose(psi, g[key], rtol=1e-13, atol=0)


def test_ffat_eval_oracle_reproduces_the_reference_fixture(orc, golden_dir):
    """tests/golden/ffat_eval.npz holds |GetMapVal| computed by the reference's own LoadAll + GetMapVal
    (tests/golden/make_golden_ffat_eval.py) at probe positions incl. face axes, edge / corner ties and interior points."""
    from oracle import fatcube
    g = np.load(os.path.join(golden_dir, "ffat_eval.npz"))
    maps = fatcube.load_all(os.path.join(golden_dir, "fatcube"))
    got = np.concatenate([orc.ffat_eval([maps[i]], g["files_pos"]) for i in range(3)], axis=1)
    ref = g["files_out"]
    assert np.array_equal(np.isfinite(got), np.isfinite(ref))
    fin = np.isfinite(ref)
    assert np.allclose(got[fin], ref[fin], rtol=1e-13, atol=0)
    shared = synth.ffat_maps(g["shared_freqs"], 2000, n=8)
    assert np.allclose(orc.ffat_eval(shared, g["shared_pos"]), g["shared_out"], rtol=1e-13, atol=0)


def test_ffat_compress_oracle_reproduces_the_opencv_fixture(orc, golden_dir):
    """tests/golden/ffat_compress.npz holds FFAT_Map<T,3>::Compress carried out with OpenCV's own cast and JPEG codec
    (tests/golden/make_golden_ffat_compress.py): the restated quantisation must give OpenCV's bytes, and the restated
    de-quantisation of the bytes that came back from the JPEG file must give _compressed_Psi, bit for bit."""
    g = np.load(os.path.join(golden_dir, "ffat_compress.npz"))
    assert np.array_equal(orc.cv_saturate_u8(g["cast_in"]), g["cast_out"])
    maps = synth.ffat_maps(g["freqs"], 2000, n=8)
    for i, m in enumerate(maps):
        m = dict(m); m["psi"] = g["psi"][i]
        q, amp, gmax = orc.ffat_quantise(m)
        assert np.array_equal(q, g["q8_pre"][i])
        assert np.array_equal(amp, g["max_amp"][i]) and gmax == g["max_amp_global"][i]
        assert np.array_equal(orc.ffat_dequantise(m, g["q8_post"][i], amp), g["compressed_psi"][i])


def test_legacy_fatcube_oracle_reader_reproduces_the_reference_fixture(orc, golden_dir):
    """tests/golden/legacy_fatcube/*.fatcube were written by the reference's own FFAT_Map<double,3>::Save through libigl's own
    igl::serialize, and legacy_eval.npz by its own legacy LoadAll + |GetMapVal| (tests/golden/make_golden_legacy.py): the oracle's
    reader of that form + the oracle's GetMapVal must reproduce it."""
    from oracle import fatcube
    g = np.load(os.path.join(golden_dir, "legacy_eval.npz"))
    maps = [fatcube.load_any(os.path.join(golden_dir, "legacy_fatcube", "mode-%d.fatcube" % i)) for i in range(4)]
    assert [m["modeid"] for m in maps] == [0, 1, 2, 3]
    got = np.concatenate([orc.ffat_eval([m], g["pos"]) for m in maps], axis=1)
    assert np.isfinite(g["out"]).all() and np.allclose(got, g["out"], rtol=1e-13, atol=0)
