"""Parity of the CUDA path (through the C ABI) against the CPU oracle.  Needs a B200: -m gpu.

Tolerance (BASELINE.json north_star): waveform rel-L2 <= 1e-5 and max-abs <= 1e-6 of full scale.
FP64 kernels are held to far tighter bounds; they restate the reference's own arithmetic."""
import os
import numpy as np
import pytest
from openpbso_b200 import synth
from conftest import assert_waveform_parity, waveform_errors

pytestmark = pytest.mark.gpu
H = synth.H


def test_native_library_is_loaded(pbso):
    assert pbso.device_count() >= 1
    info = pbso.device_info()
    assert info["cc"][0] == 10, info          # Blackwell
    maps = open("/proc/self/maps").read()
    assert "libpbso_b200.so" in maps


# --------------------------------------------------------------------------- K2 / Step
@pytest.mark.parametrize("material", ["low_damping", "high_damping"])
def test_coefficients(pbso, orc, material):
    mat = synth.MATERIALS[material]
    f = synth.mode_frequencies(1024, 7)
    w2 = synth.omega_squared(f, mat["density"])
    it = pbso.ModalIntegrator.Build(mat["density"], w2, mat["alpha"], mat["beta"], H, 1000)
    assert it.N == 1000                                    # Build(..., N) culls (modal_integrator.h:53-57)
    a, b = orc.build_ab(mat["density"], w2, mat["alpha"], mat["beta"], 1000)
    c1, c2, c3 = orc.coeffs(H, a, b)
    g1, g2, g3 = it.coeffs()
    assert np.allclose(g1, c1, rtol=1e-14) and np.allclose(g2, c2, rtol=1e-14) and np.allclose(g3, c3, rtol=1e-12)


def test_build_rejects_bad_N(pbso):
    with pytest.raises(pbso.PbsoError) as e:
        pbso.ModalIntegrator.Build(1.0, np.ones(4), 1.0, 1e-7, H, 5)
    assert e.value.code == 1


def test_step_sequence(pbso, orc):
    mat = synth.MATERIALS["low_damping"]
    f = synth.mode_frequencies(300, 17)
    a, b = synth.ab_from_material(f, mat)
    it = pbso.ModalIntegrator(300, H, a, b); ref = orc.Integrator(H, a, b)
    rng = np.random.default_rng(1)
    for k in range(200):
        Q = rng.standard_normal(300) if k % 5 == 0 else None
        g = it.Step(Q); r = ref.step(Q)
        assert np.max(np.abs(g - r)) <= 1e-10 * np.max(np.abs(r)), k
    q1, q2 = it.get_state(); r1, r2 = ref.state()
    assert np.allclose(q1, r1, rtol=0, atol=1e-10 * np.max(np.abs(r1)))
    assert np.allclose(q2, r2, rtol=0, atol=1e-10 * np.max(np.abs(r2)))
    with pytest.raises(pbso.PbsoError):
        it.Step(np.ones(299))                              # "input force incorrect dimension" (:108)


def test_state_roundtrip_continues_render(pbso, orc):
    """get_state/set_state (checkpoint of the 3-slot ring): a restored handle continues identically."""
    f = synth.mode_frequencies(70, 3); a, b = synth.ab_from_material(f, synth.MATERIALS["high_damping"])
    A = pbso.ModalIntegrator(70, H, a, b); B = pbso.ModalIntegrator(70, H, a, b)
    sp = np.random.default_rng(0).standard_normal(70); tm = np.zeros(256); tm[0] = 1
    A.render_buffer(sp, tm)
    B.set_state(*A.get_state())
    ya, _ = A.render_buffer(np.zeros(70), np.zeros(256)); yb, _ = B.render_buffer(np.zeros(70), np.zeros(256))
    assert np.array_equal(ya, yb)


# --------------------------------------------------------------------------- K1 per buffer
def test_cfg1_golden_waveform(pbso, golden_dir):
    """ball.obj vertex 0, 64 modes, FFAT transfer, one PointForce, 173 x 256 samples."""
    g = np.load(os.path.join(golden_dir, "cfg1_ball.npz"))
    mat = synth.MATERIALS["low_damping"]
    freqs = synth.mode_frequencies(64, 1001)
    it = pbso.ModalIntegrator.Build(mat["density"], synth.omega_squared(freqs, mat["density"]), mat["alpha"], mat["beta"], H)
    # projection (K4 sparse) and transfer (K3) from the committed inputs
    K = 3 * int(g["n_vertices"])
    U = np.zeros((64, K)); U[:, :3] = g["u_vid"]
    space = pbso.ModeShapes(U).GetModalForceVertex(64, int(g["vid"]), g["vn"])
    assert np.allclose(space, g["space"], rtol=1e-13)
    trans = pbso.FFATMaps.from_dicts(synth.ffat_maps(freqs, 2000)).computeTransfer(g["listener"])[0]
    assert np.allclose(trans, g["trans"], rtol=1e-12)
    it.set_transfer(trans)
    ys = []
    for bi in range(173):
        tm = np.zeros(256)
        sp = np.zeros(64)
        if bi == 0:
            tm[0] = 1.0; sp = space * g["scale"]
        y, _ = it.render_buffer(sp, tm)
        ys.append(y[0])
    rel, mx = assert_waveform_parity(np.concatenate(ys), g["y"], rel=1e-9, mx=1e-9)


def test_force_script_buffers(pbso, orc, golden_dir):
    """Every buffer the oracle's force state machine produces (point, overlapping gaussian, clear, sustained
    AR start/update/end, transfer swap, unit transfer) re-rendered by K1 from the same (space, time)."""
    g = np.load(os.path.join(golden_dir, "script_forces.npz"))
    a, b, spaces, trans = g["a"], g["b"], g["spaces"], g["trans"]
    N = len(a)
    s = orc.Solver(orc.Integrator(H, a, b), 256)
    it = pbso.ModalIntegrator(N, H, a, b)
    n_out = 0
    for kind, arg in zip(g["kinds"], g["args"]):
        kind = str(kind); arg = int(arg)
        if kind == "point": s.enqueue_force(spaces[arg], orc.POINT)
        elif kind == "gauss": s.enqueue_force(spaces[arg], orc.GAUSSIAN, width_us=900.0)
        elif kind == "clear": s.enqueue_force(spaces[0], orc.POINT, flags=orc.F_CLEAR)
        elif kind == "ar_start": s.enqueue_force(spaces[arg], orc.AR, flags=orc.F_SUSTAIN_START)
        elif kind == "ar_data": s.enqueue_force(spaces[arg], orc.AR)
        elif kind == "ar_end": s.enqueue_force(spaces[arg], orc.AR, flags=orc.F_SUSTAIN_END)
        elif kind == "arprm": s.enqueue_arprm(0.7, 0.2, 0.002, 0.1)
        elif kind == "trans": s.enqueue_trans(trans[arg])
        elif kind == "unit_transfer": s.set_use_transfer(False)
        elif kind == "use_transfer": s.set_use_transfer(True)
        r = s.step()
        if r is None:
            continue
        assert np.allclose(r[0], g["y"][n_out], rtol=1e-9, atol=1e-9 * np.max(np.abs(g["y"])))   # golden fixture
        sp, tm = s.last_force()
        it.set_transfer(s.latest_transfer())
        y, qn = it.render_buffer(sp, tm)
        full = max(np.max(np.abs(r[0])), 1e-300)
        assert np.max(np.abs(y[0] - r[0])) <= 1e-9 * max(full, np.max(np.abs(g["y"])) * 1e-3), (kind, n_out)
        assert np.allclose(qn, r[1], rtol=1e-9, atol=1e-9 * np.max(r[1]) + 1e-300)
        n_out += 1
    assert n_out == len(g["y"]) == 22


@pytest.mark.parametrize("N,T", [(1, 1), (31, 17), (256, 256), (1000, 513), (2048, 256)])
def test_render_buffer_shapes(pbso, orc, N, T):
    """Ragged sizes: N not a multiple of the block, T not a multiple of the 32-sample tile (incl. the
    reference default FRAMES_PER_BUFFER = 513), n_transfer < N (q.head(n).dot, modal_solver.h:268)."""
    f = synth.mode_frequencies(N, N + T); a, b = synth.ab_from_material(f, synth.MATERIALS["high_damping"])
    rng = np.random.default_rng(N)
    it = pbso.ModalIntegrator(N, H, a, b); ref = orc.Integrator(H, a, b)
    nt = max(1, N - 3)
    tr = np.abs(rng.standard_normal(nt)) + 0.1
    it.set_transfer(tr)
    for rep in range(3):
        sp = rng.standard_normal(N); tm = rng.standard_normal(T)
        y, qn = it.render_buffer(sp, tm)
        q = np.array([ref.step(sp * tm[i]) for i in range(T)])
        yr = q[:, :nt] @ tr
        assert np.max(np.abs(y[0] - yr)) <= 1e-9 * np.max(np.abs(yr))
        assert np.allclose(qn, np.sqrt((q * q).sum(0)), rtol=1e-9)


def test_render_buffer_many_listeners(pbso, orc):
    """cfg4 shape: 64 listeners share one IIR pass: y_l = sum_m T[l][m] q_m."""
    N, L, T = 1024, 64, 256
    f = synth.mode_frequencies(N, 1004); a, b = synth.ab_from_material(f, synth.MATERIALS["low_damping"])
    rng = np.random.default_rng(4)
    tr = np.abs(rng.standard_normal((L, N))) + 0.1
    it = pbso.ModalIntegrator(N, H, a, b); ref = orc.Integrator(H, a, b)
    it.set_transfer(tr, L)
    for rep in range(2):
        sp = rng.standard_normal(N); tm = np.zeros(T); tm[0] = 1.0 if rep == 0 else 0.0
        y, _ = it.render_buffer(sp, tm)
        q = np.array([ref.step(sp * tm[i]) for i in range(T)])
        yr = (q @ tr.T).T
        assert y.shape == (L, T)
        assert np.max(np.abs(y - yr)) <= 1e-9 * np.max(np.abs(yr))


@pytest.mark.parametrize("N,n_tr,L,T", [(300, 257, 13, 513), (1024, 1024, 5, 256), (70, 70, 33, 64), (513, 1, 9, 31)])
def test_render_buffer_many_listeners_ragged(pbso, orc, N, n_tr, L, T):
    """More than 4 listeners take the shared-recurrence path (k_iir_history + k_modal_sum): odd transfer lengths, listener
    counts and buffer sizes off the tile sizes, state carried across buffers, qnorm."""
    f = synth.mode_frequencies(N, 1004); a, b = synth.ab_from_material(f, synth.MATERIALS["high_damping"])
    rng = np.random.default_rng(11)
    tr = np.abs(rng.standard_normal((L, n_tr))) + 0.1
    it = pbso.ModalIntegrator(N, H, a, b); ref = orc.Integrator(H, a, b)
    it.set_transfer(tr, L)
    for rep in range(3):
        sp = rng.standard_normal(N); tm = rng.standard_normal(T) if rep == 1 else np.eye(1, T, 0)[0] * (rep == 0)
        y, qn = it.render_buffer(sp, tm)
        q = np.array([ref.step(sp * tm[i]) for i in range(T)])
        yr = (q[:, :n_tr] @ tr.T).T
        assert y.shape == (L, T)
        assert np.max(np.abs(y - yr)) <= 1e-9 * max(np.max(np.abs(yr)), 1e-300)
        assert np.allclose(qn, np.sqrt((q * q).sum(axis=0)), rtol=1e-12, atol=0)


def test_moving_listeners_transfer_stays_on_device(pbso, orc):
    """cfg4: per buffer, computeTransfer for every listener (K3) lands in the integrator's transfer table without a host
    round trip, then one IIR pass renders all listeners.  Against the oracle's ffat_eval + integrator."""
    N, L, T = 200, 64, 256
    f = synth.mode_frequencies(N, 1004); a, b = synth.ab_from_material(f, synth.MATERIALS["low_damping"])
    maps = synth.ffat_maps(f, 2000, n=8)
    fm = pbso.FFATMaps.from_dicts(maps)
    it = pbso.ModalIntegrator(N, H, a, b); ref = orc.Integrator(H, a, b)
    rng = np.random.default_rng(8)
    pos = synth.listeners(L, 21)
    for rep in range(4):
        pos = pos + 0.05 * rng.standard_normal(pos.shape)                 # listeners move every buffer
        it.set_transfer_ffat(fm, pos)
        tr = orc.ffat_eval(maps, pos)                                     # [L][N]
        sp = rng.standard_normal(N); tm = np.zeros(T); tm[0] = 1.0 if rep % 2 == 0 else 0.0
        y, _ = it.render_buffer(sp, tm)
        q = np.array([ref.step(sp * tm[i]) for i in range(T)])
        yr = (q @ tr.T).T
        assert y.shape == (L, T)
        assert np.max(np.abs(y - yr)) <= 1e-9 * np.max(np.abs(yr))
    it.set_transfer_ffat(fm, pos[:3], n_transfer=150)                     # fewer listeners / modes: q.head(n).dot(transfer)
    y, _ = it.render_buffer(np.zeros(N), np.zeros(T))
    q = np.array([ref.step() for i in range(T)])
    assert np.max(np.abs(y - (q[:, :150] @ orc.ffat_eval(maps[:150], pos[:3]).T).T)) <= 1e-9 * np.max(np.abs(q))
    with pytest.raises(pbso.PbsoError) as e:                              # _ffat_maps->at(ii) past the last map
        it.set_transfer_ffat(pbso.FFATMaps.from_dicts(maps[:10]), pos)
    assert e.value.code == 5


def test_transfer_table_grows_after_a_render(pbso, orc):
    """Regression: growing the transfer table after a render with N > 256 (cross-slab scratch allocated), and
    alternating set_transfer_ffat / set_transfer, must leave the renderer's scratch buffers alone."""
    N, T = 600, 256
    f = synth.mode_frequencies(N, 77); a, b = synth.ab_from_material(f, synth.MATERIALS["low_damping"])
    maps = synth.ffat_maps(f, 2000, n=8)
    fm = pbso.FFATMaps.from_dicts(maps)
    it = pbso.ModalIntegrator(N, H, a, b); ref = orc.Integrator(H, a, b)
    rng = np.random.default_rng(77)
    pos = synth.listeners(40, 5)

    def check(tr, rep):
        sp = rng.standard_normal(N); tm = np.zeros(T); tm[0] = 1.0
        y, _ = it.render_buffer(sp, tm)
        q = np.array([ref.step(sp * tm[i]) for i in range(T)])
        yr = (q[:, :tr.shape[1]] @ tr.T).T
        assert y.shape == yr.shape, rep
        assert np.max(np.abs(y - yr)) <= 1e-9 * np.max(np.abs(yr)), rep

    tr1 = np.abs(rng.standard_normal((1, N))) + 0.1
    it.set_transfer(tr1, 1); check(tr1, "L=1")
    tr9 = np.abs(rng.standard_normal((9, N))) + 0.1
    it.set_transfer(tr9, 9); check(tr9, "grown to L=9 after a render")
    it.set_transfer_ffat(fm, pos[:3]); check(orc.ffat_eval(maps, pos[:3]), "ffat L=3")
    tr40 = np.abs(rng.standard_normal((40, N))) + 0.1
    it.set_transfer(tr40, 40); check(tr40, "grown to L=40 after set_transfer_ffat")
    it.set_transfer_ffat(fm, pos); check(orc.ffat_eval(maps, pos), "ffat L=40")
    it.set_transfer(tr9, 9); check(tr9, "shrunk to L=9")
    it.close()


# --------------------------------------------------------------------------- K3 FFAT
def test_ffat_shared_geometry(pbso, orc):
    freqs = synth.mode_frequencies(200, 1004)
    maps = synth.ffat_maps(freqs, 2000)
    pos = synth.listeners(50, 12)
    got = pbso.FFATMaps.from_dicts(maps).computeTransfer(pos)
    ref = orc.ffat_eval(maps, pos)
    assert got.shape == (50, 200)
    assert np.allclose(got, ref, rtol=1e-12)


def test_ffat_general_geometry_from_files(pbso, orc, golden_dir):
    from oracle import fatcube
    d = os.path.join(golden_dir, "fatcube")
    ref_maps = fatcube.load_all(d)
    ref_maps[1]["k"] = 0.0                                        # k = 0 on the wire: division by zero -> inf
    fm = pbso.FFATMaps.LoadAll(d)
    pos = np.concatenate([synth.listeners(40, 5, 4.0, 9.0), [[5.0, 0.0, 0.0], [0.0, 0.0, -7.0], [3.0, 3.0, 3.0]]])
    got = fm.computeTransfer(pos)
    ref = np.concatenate([orc.ffat_eval([ref_maps[i]], pos) for i in range(3)], axis=1)   # maps differ in size
    assert np.array_equal(np.isinf(got), np.isinf(ref))
    fin = np.isfinite(ref)
    assert np.allclose(got[fin], ref[fin], rtol=1e-12)
    with pytest.raises(pbso.PbsoError) as e:                      # _ffat_maps->at(3) throws (modal_solver.h:296)
        fm.computeTransfer(pos, n_modes=4)
    assert e.value.code == 5


def test_ffat_many_listeners_staged(pbso, orc):
    """Many listeners: the default coalesced gather and the opt-in kernel that stages whole maps in shared memory
    with bulk async copies; modes not a multiple of 4 and a ragged listener count exercise the tails."""
    freqs = synth.mode_frequencies(37, 77)
    maps = synth.ffat_maps(freqs, 2000, n=16)
    pos = synth.listeners(2500, 13)
    ref = orc.ffat_eval(maps, pos)
    os.environ["PBSO_FFAT_STAGED"] = "1"                      # opt-in: whole maps staged in shared memory (UBLKCP)
    try:
        got = pbso.FFATMaps.from_dicts(maps).computeTransfer(pos)
    finally:
        del os.environ["PBSO_FFAT_STAGED"]
    assert np.allclose(got, ref, rtol=1e-13)
    assert np.allclose(pbso.FFATMaps.from_dicts(maps).computeTransfer(pos), ref, rtol=1e-13)   # default path (texel tiles at this size)
    got1 = pbso.FFATMaps.from_dicts(maps).computeTransfer(pos[:100])            # small-L path, same numbers
    assert np.allclose(got1, got[:100], rtol=1e-14)


@pytest.mark.parametrize("n_modes,n_tex,L", [(37, 16, 2500), (200, 12, 3000), (256, 32, 4500), (130, 9, 2100), (66, 8, 6100)])
def test_ffat_many_listeners_texel_tiles(pbso, orc, n_modes, n_tex, L):
    """L >= 2048 and 4 L >= texels: the texel-stationary kernel (tile rows staged with bulk async copies).  Odd mode
    counts take the non-bulk / scalar-store tails, 12- and 9-texel faces leave partial tiles, 200/256 modes span two
    mode slabs, and with 8-texel faces (one tile per face) 6100 listeners put ~1000 records in a tile: four record chunks
    through the two record buffers.  Same numbers as the per-listener gather."""
    freqs = synth.mode_frequencies(n_modes, 78)
    maps = synth.ffat_maps(freqs, 2000, n=n_tex)
    pos = synth.listeners(L, 14)
    pos[:64] = 3.0 * synth.texel_centres(maps[0])[:64]          # exact texel centres / face corners: clamped stencils
    ref = orc.ffat_eval(maps, pos)
    got = pbso.FFATMaps.from_dicts(maps).computeTransfer(pos)
    assert got.shape == (L, n_modes)
    assert np.allclose(got, ref, rtol=1e-13)
    os.environ["PBSO_FFAT_GATHER"] = "1"
    try:
        gat = pbso.FFATMaps.from_dicts(maps).computeTransfer(pos)
    finally:
        del os.environ["PBSO_FFAT_GATHER"]
    assert np.allclose(gat, ref, rtol=1e-13) and np.allclose(gat, got, rtol=1e-14)


def test_ffat_full_size_properties(pbso):
    """cfg4 size (1024 modes x 64 listeners): 1/r law and texel-centre identity, no oracle needed."""
    freqs = synth.mode_frequencies(1024, 1004)
    maps = synth.ffat_maps(freqs, 2000)
    fm = pbso.FFATMaps.from_dicts(maps)
    d = synth.unit_vectors(64, 3)
    t1 = fm.computeTransfer(4.0 * d); t2 = fm.computeTransfer(8.0 * d)
    assert np.allclose(t1, 2.0 * t2, rtol=1e-12)
    pts = synth.texel_centres(maps[0])
    idx = np.random.default_rng(0).integers(0, len(pts), 64)
    t = fm.computeTransfer(3.0 * pts[idx])
    psi = np.array([m["psi"][idx] for m in maps]).T
    k = np.array([m["k"] for m in maps])
    r = np.linalg.norm(3.0 * pts[idx], axis=1)[:, None]
    assert np.allclose(t, np.abs(psi / (k * r)), rtol=1e-12)


# --------------------------------------------------------------------------- K4 projection
def test_projection_sparse_and_dense(pbso, orc):
    M, V = 96, 500
    U = synth.mode_shapes(M, 3 * V, 1003)
    md = pbso.ModeShapes(U)
    rng = np.random.default_rng(2)
    vn = synth.unit_vectors(1, 8)[0]
    assert np.allclose(md.GetModalForceVertex(M, 123, vn), orc.project_vertex(U, 123, vn), rtol=1e-13)
    assert np.allclose(md.GetModalForceVertex(40, 499, vn), orc.project_vertex(U, 499, vn, 40), rtol=1e-13)
    vids = [3, 77, 499]; bc = np.array([0.5, 0.25, 0.25])
    assert np.allclose(md.GetModalForceFace(M, vids, bc, vn), orc.project_face(U, vids, bc, vn), rtol=1e-12, atol=1e-14)
    vs = rng.integers(0, V, 33); vns = synth.unit_vectors(33, 9)
    got = md.project_vertices(M, vs, vns)
    ref = np.array([orc.project_vertex(U, v, n) for v, n in zip(vs, vns)])
    assert np.allclose(got, ref, rtol=1e-13)
    f = rng.standard_normal(3 * V)
    assert np.allclose(md.project_dense(f)[0], orc.project_dense(U, f.reshape(1, -1))[0], rtol=1e-11, atol=1e-11)
    F = rng.standard_normal((9, 3 * V))
    Y = md.project_dense(F); Yr = orc.project_dense(U, F)
    assert np.max(np.linalg.norm(Y - Yr, axis=1) / np.linalg.norm(Yr, axis=1)) < 1e-12
    for bad in (-1, V):
        with pytest.raises(pbso.PbsoError) as e:                   # std::vector::at
            md.GetModalForceVertex(M, bad, vn)
        assert e.value.code == 5
    with pytest.raises(pbso.PbsoError):
        md.GetModalForceVertex(M + 1, 0, vn)


def test_modes_file_roundtrip(pbso, orc, tmp_path):
    U = synth.mode_shapes(9, 21, 5); w2 = np.linspace(1e5, 1e9, 9)
    p = str(tmp_path / "t.modes"); orc.modes_write(p, w2, U)
    md = pbso.ModeShapes.read(p)
    assert (md.M, md.K) == (9, 21) and np.array_equal(md.omegaSquared(), w2)
    vn = np.array([0.0, 1.0, 0.0])
    assert np.array_equal(md.GetModalForceVertex(9, 2, vn), U[:, 7])
    with pytest.raises(pbso.PbsoError) as e:
        pbso.ModeShapes.read(str(tmp_path / "missing.modes"))
    assert e.value.code == 3


@pytest.mark.parametrize("M,K,B", [(128, 64, 128), (128, 256, 16), (96, 1500, 9), (300, 4100, 200), (2048, 6000, 580), (2048, 60000, 580)])
def test_projection_tensor_core_3xtf32(pbso, M, K, B):
    """K5: tcgen05 3xTF32 batched projection; column rel-L2 <= 1e-5 against float64 numpy.
    Ragged M, K (not multiples of the 128 x 32 tiles), B below and above one N tile, split-K; the last case is cfg3 itself:
    2048 modes x 20 000 vertices (K = 60 000), 580 impulses = one 256-sample buffer of the 100 k impulses/s contact storm."""
    rng = np.random.default_rng(M + K + B)
    U = rng.standard_normal((M, K)); F = rng.standard_normal((B, K))
    md = pbso.ModeShapes(U)
    Y = md.project_dense(F, precision=pbso.PREC_TF32X3)
    Yr = F @ U.T
    err = np.linalg.norm(Y - Yr, axis=1) / np.linalg.norm(Yr, axis=1)
    print("3xTF32 M=%d K=%d B=%d: max column rel-L2 %.2e" % (M, K, B, err.max()))
    assert err.max() <= 1e-5
    Y2 = md.project_dense(F[:, :], forceDim=M - 5, precision=pbso.PREC_TF32X3)      # forceDim < M (culled modes)
    assert np.allclose(Y2, Y[:, :M - 5], rtol=0, atol=1e-4 * np.abs(Yr).max())


def test_contact_storm_pipeline_vs_oracle(pbso, orc):
    """cfg3 on the device: B vertex impulses per buffer -> projection (sparse FP64 gather, or dense load vectors through
    the tensor-core GEMM) -> summed modal load -> integrator, against the oracle's GetModalForceVertex + ModalSolver::step
    fed with the summed load; state carries over the buffers, the impulse count changes per buffer."""
    M, V, T = 96, 300, 256
    f = synth.mode_frequencies(M, 1003); a, b = synth.ab_from_material(f, synth.MATERIALS["low_damping"])
    U = synth.mode_shapes(M, 3 * V, 1003)
    md = pbso.ModeShapes(U)
    rng = np.random.default_rng(1003)
    tr = np.abs(rng.standard_normal(M)) + 0.1
    outs = {}
    for prec in (pbso.PREC_F64, pbso.PREC_TF32X3):
        it = pbso.ModalIntegrator(M, H, a, b); it.set_transfer(tr)
        ref = orc.Solver(orc.Integrator(H, a, b), T); ref.enqueue_trans(tr)
        rs = np.random.default_rng(7)
        ys, yr = [], []
        for bi, B in enumerate((37, 1, 130, 580, 5)):
            vids = rs.integers(0, V, B); vn = synth.unit_vectors(B, 100 + bi)
            y, qn = md.storm_buffer(it, vids, vn, T, prec, want_qnorm=True)
            space = sum(orc.project_vertex(U, int(v), n) for v, n in zip(vids, vn))
            ref.enqueue_force(space)
            r = ref.step()
            ys.append(y[0]); yr.append(r[0])
            assert np.allclose(qn, r[1], rtol=1e-9 if prec == pbso.PREC_F64 else 2e-5)
        ys = np.concatenate(ys); yr = np.concatenate(yr)
        if prec == pbso.PREC_F64: assert_waveform_parity(ys, yr, rel=1e-9, mx=1e-9)
        else: outs["rel"], outs["mx"] = assert_waveform_parity(ys, yr)
        it.close()
    print("contact storm, tensor-core projection: rel-L2 %.2e max-abs %.2e" % (outs["rel"], outs["mx"]))
    with pytest.raises(pbso.PbsoError):
        md.storm_buffer(pbso.ModalIntegrator(M, H, a, b), [V], [[1.0, 0.0, 0.0]], T)        # vertex id out of range


# --------------------------------------------------------------------------- batch renderer
def _batch_case(n_obj, n_modes, n_buf, seed, material="low_damping", extra_events=0):
    w = synth.batch_workload(n_obj, n_modes, n_buf, seed, material, first_second_bufs=max(1, n_buf // 2))
    return w


@pytest.mark.parametrize("precision", ["f64", "f32_tiled"])
@pytest.mark.parametrize("n_obj,n_modes,n_buf,material", [(5, 80, 24, "low_damping"), (3, 300, 12, "high_damping"), (2, 16, 40, "low_damping")])
def test_batch_mix_vs_oracle(pbso, orc, precision, n_obj, n_modes, n_buf, material):
    w = _batch_case(n_obj, n_modes, n_buf, 100 + n_modes, material)
    ref = orc.batch_render(H, w["a"], w["b"], w["space"], w["trans"], w["imp_buf"], 256, n_buf)
    br = pbso.BatchRenderer(H, w["a"], w["b"])
    br.set_transfer(w["trans"])
    br.set_impulses(np.arange(n_obj), w["imp_buf"], w["space"])
    prec = pbso.PREC_F64 if precision == "f64" else pbso.PREC_F32_TILED
    mix = br.render_mix(256, n_buf, prec)
    if precision == "f64":
        assert_waveform_parity(mix, ref, rel=1e-10, mx=1e-10)
    else:
        rel, mx = assert_waveform_parity(mix, ref)
        print("f32_tiled: rel-L2 %.2e max-abs %.2e" % (rel, mx))


def test_batch_multiple_events_and_stems(pbso, orc):
    """Two impulses on one object in different buffers; stems sum to the mix."""
    n_obj, n_modes, n_buf = 4, 64, 30
    w = _batch_case(n_obj, n_modes, n_buf, 9)
    rng = np.random.default_rng(9)
    obj = np.array([0, 1, 2, 3, 1, 3]); buf = np.array([2, 0, 5, 7, 9, 8])
    space = rng.standard_normal((6, n_modes))
    br = pbso.BatchRenderer(H, w["a"], w["b"]); br.set_transfer(w["trans"]); br.set_impulses(obj, buf, space)
    ref = np.zeros(n_buf * 256)
    for o in range(n_obj):
        s = orc.Solver(orc.Integrator(H, w["a"][o], w["b"][o]), 256)
        s.enqueue_trans(w["trans"][o])
        for bi in range(n_buf):
            for e in np.nonzero((obj == o) & (buf == bi))[0]:
                s.enqueue_force(space[e])
            ref[bi * 256:(bi + 1) * 256] += s.step()[0]
    per_obj = np.zeros((n_obj, n_buf * 256))
    for o in range(n_obj):
        s = orc.Solver(orc.Integrator(H, w["a"][o], w["b"][o]), 256)
        s.enqueue_trans(w["trans"][o])
        for bi in range(n_buf):
            for e in np.nonzero((obj == o) & (buf == bi))[0]:
                s.enqueue_force(space[e])
            per_obj[o, bi * 256:(bi + 1) * 256] = s.step()[0]
    for prec, tol in ((pbso.PREC_F64, 1e-10), (pbso.PREC_F32_TILED, None), (pbso.PREC_TC3X, None)):
        mix = br.render_mix(256, n_buf, prec)
        if tol: assert_waveform_parity(mix, ref, rel=tol, mx=tol)
        else: assert_waveform_parity(mix, ref)
        stems = br.render_stems(256, n_buf, prec)
        assert_waveform_parity(stems.astype(np.float64).sum(0), ref, rel=1e-5, mx=2e-6)
        for o in range(n_obj):                               # every stem is its own object's solver loop (float32 storage)
            assert np.abs(stems[o] - per_obj[o]).max() <= 3e-6 * np.abs(per_obj).max(), (prec, o)
    with pytest.raises(pbso.PbsoError):                      # one message per object per buffer (modal_solver.h:184)
        br.set_impulses([0, 0], [3, 3], space[:2])
    with pytest.raises(pbso.PbsoError):
        br.set_impulses([9], [0], space[:1])


@pytest.mark.parametrize("n_chunks", [0, 1, 5, 22])
def test_batch_time_chunks_cfg1_golden(pbso, golden_dir, n_chunks):
    """One 64-mode object, 173 buffers: too small to fill the GPU by objects, so the render is split into
    independent time chunks whose start states come from closed-form pole powers (north_star (1)).  Checked
    against the cfg1 golden waveform for automatic and explicit chunk counts."""
    g = np.load(os.path.join(golden_dir, "cfg1_ball.npz"))
    mat = synth.MATERIALS["low_damping"]
    freqs = synth.mode_frequencies(64, 1001)
    a, b = synth.ab_from_material(freqs, mat)
    br = pbso.BatchRenderer(H, a[None, :], b[None, :])
    br.set_transfer(g["trans"][None, :])
    br.set_impulses([0], [0], (g["space"] * g["scale"])[None, :])
    y = br.render_mix(256, 173, pbso.PREC_F32_TILED, n_chunks)
    rel, mx = assert_waveform_parity(y, g["y"])
    print("chunks=%d: rel-L2 %.2e max-abs %.2e" % (n_chunks, rel, mx))


def test_batch_time_chunks_many_events(pbso):
    """Impulse stream on few objects (cfg2-like) rendered in chunks: identical to the single-chunk render."""
    n_obj, n_modes, n_buf = 3, 200, 400
    w = synth.batch_workload(n_obj, n_modes, n_buf, 21, "high_damping")
    rng = np.random.default_rng(21)
    obj = np.repeat(np.arange(n_obj), 40)
    buf = np.concatenate([rng.choice(n_buf, 40, replace=False) for _ in range(n_obj)])
    space = rng.standard_normal((len(obj), n_modes))
    br = pbso.BatchRenderer(H, w["a"], w["b"]); br.set_transfer(w["trans"]); br.set_impulses(obj, buf, space)
    y1 = br.render_mix(256, n_buf, pbso.PREC_F32_TILED, 1)
    y64 = br.render_mix(256, n_buf, pbso.PREC_F64)
    assert_waveform_parity(y1, y64)
    for nc in (0, 7, 50):
        yc = br.render_mix(256, n_buf, pbso.PREC_F32_TILED, nc)
        assert_waveform_parity(yc, y64)
        assert np.max(np.abs(yc - y1)) <= 5e-7 * np.max(np.abs(y64))


def test_batch_full_size_slice_f32_vs_f64(pbso):
    """cfg5 slice at full length (512 modes x 1723 buffers = 10 s): FP32 pole-power tiles against the FP64
    direct-form kernel, plus linearity and time-shift invariance (size-independent properties)."""
    n_obj, n_modes, n_buf = 48, 512, 1723
    w = synth.batch_workload(n_obj, n_modes, n_buf, 1005)
    br = pbso.BatchRenderer(H, w["a"], w["b"]); br.set_transfer(w["trans"])
    br.set_impulses(np.arange(n_obj), w["imp_buf"], w["space"])
    y64 = br.render_mix(256, n_buf, pbso.PREC_F64)
    y32 = br.render_mix(256, n_buf, pbso.PREC_F32_TILED)
    rel, mx = assert_waveform_parity(y32, y64)
    print("cfg5 slice: f32_tiled vs f64 rel-L2 %.2e max-abs %.2e" % (rel, mx))
    # late-time accuracy: error relative to the local envelope in the last second stays small
    tail = slice(-44100, None)
    assert np.linalg.norm(y32[tail] - y64[tail]) <= 1e-5 * np.linalg.norm(y64[tail])
    # linearity
    br.set_impulses(np.arange(n_obj), w["imp_buf"], 2.0 * w["space"])
    y2 = br.render_mix(256, n_buf, pbso.PREC_F32_TILED)
    assert np.max(np.abs(y2 - 2.0 * y32)) <= 2e-6 * np.max(np.abs(y2))
    # time shift: all impulses 5 buffers later -> same waveform delayed by 5*256 samples
    br.set_impulses(np.arange(n_obj), w["imp_buf"] + 5, w["space"])
    ys = br.render_mix(256, n_buf, pbso.PREC_F32_TILED)
    assert np.max(np.abs(ys[5 * 256:] - y32[:-5 * 256])) <= 2e-6 * np.max(np.abs(y32))
    assert not ys[:5 * 256 + int(w["imp_buf"].min()) * 256].any()


# --------------------------------------------------------------------------- K1 batch on tensor cores (3xTF32)
@pytest.mark.parametrize("n_obj,n_modes,n_buf,material", [(5, 80, 24, "low_damping"), (3, 300, 12, "high_damping"),
                                                           (2, 16, 150, "low_damping"), (7, 33, 70, "high_damping")])
def test_batch_tc3x_vs_oracle(pbso, orc, n_obj, n_modes, n_buf, material):
    """Pole-power synthesis as one tcgen05 contraction (batch_tc.cu) against the CPU oracle: ragged mode counts
    (not a multiple of the 16-mode K chunk), renders shorter and longer than one 128-tile M-tile."""
    w = _batch_case(n_obj, n_modes, n_buf, 11, material)
    br = pbso.BatchRenderer(H, w["a"], w["b"]); br.set_transfer(w["trans"])
    br.set_impulses(np.arange(n_obj), w["imp_buf"], w["space"])
    ref = orc.batch_render(H, w["a"], w["b"], w["space"], w["trans"], w["imp_buf"], 256, n_buf)
    mix = br.render_mix(256, n_buf, pbso.PREC_TC3X)
    rel, mx = assert_waveform_parity(mix, ref)
    print("tc3x: rel-L2 %.2e max-abs %.2e" % (rel, mx))


def test_batch_tc3x_many_events(pbso):
    """Impulse streams (several impulses per object, some in the same M-tile, unsorted input) against the FP64
    direct-form kernel; re-render after changing the script and the transfer (cached tables must follow)."""
    n_obj, n_modes, n_buf = 3, 200, 400
    w = synth.batch_workload(n_obj, n_modes, n_buf, 21, "high_damping")
    rng = np.random.default_rng(21)
    obj = np.repeat(np.arange(n_obj), 40)
    buf = np.concatenate([rng.choice(n_buf, 40, replace=False) for _ in range(n_obj)])
    space = rng.standard_normal((len(obj), n_modes))
    br = pbso.BatchRenderer(H, w["a"], w["b"]); br.set_transfer(w["trans"]); br.set_impulses(obj, buf, space)
    y64 = br.render_mix(256, n_buf, pbso.PREC_F64)
    ytc = br.render_mix(256, n_buf, pbso.PREC_TC3X)
    assert_waveform_parity(ytc, y64)
    br.set_impulses(obj[::3], buf[::3], space[::3]); br.set_transfer(2.0 * w["trans"])
    assert_waveform_parity(br.render_mix(256, n_buf, pbso.PREC_TC3X), br.render_mix(256, n_buf, pbso.PREC_F64))
    assert_waveform_parity(br.render_mix(128, 2 * n_buf - 3, pbso.PREC_TC3X), br.render_mix(128, 2 * n_buf - 3, pbso.PREC_F64))
    # 64-sample buffers: impulses land inside the 128-sample tiles (k_batch_event_heads renders up to the boundary)
    assert_waveform_parity(br.render_mix(64, n_buf, pbso.PREC_TC3X), br.render_mix(64, n_buf, pbso.PREC_F64))


@pytest.mark.parametrize("material", ["low_damping", "high_damping"])
def test_batch_tc3x_and_f64_vs_oracle_full_length(pbso, orc, material):
    """The headline configuration's shape side by side with the CPU oracle: 512 modes x 1723 buffers (10 s, 441 088
    samples), 4 objects, both materials.  PREC_F64 must meet the oracle to 1e-9 of full scale, PREC_TC3X the north-star
    tolerance (rel-L2 <= 1e-5, max-abs <= 1e-6 of full scale) over the whole waveform AND, where the signal is still
    inside FP32's range, over its last second alone."""
    n_obj, n_modes, n_buf = 4, 512, 1723
    w = synth.batch_workload(n_obj, n_modes, n_buf, 1005, material)
    ref = np.zeros(n_buf * 256)
    import threading
    parts = [np.zeros(n_buf * 256) for _ in range(n_obj)]
    th = [threading.Thread(target=orc.batch_render, args=(H, w["a"][o:o + 1], w["b"][o:o + 1], w["space"][o:o + 1], w["trans"][o:o + 1],
                                                         w["imp_buf"][o:o + 1], 256, n_buf, parts[o])) for o in range(n_obj)]
    for t in th: t.start()
    for t in th: t.join()
    for part in parts: ref += part
    br = pbso.BatchRenderer(H, w["a"], w["b"]); br.set_transfer(w["trans"])
    br.set_impulses(np.arange(n_obj), w["imp_buf"], w["space"])
    y64 = br.render_mix(256, n_buf, pbso.PREC_F64)
    assert_waveform_parity(y64, ref, rel=1e-9, mx=1e-9)
    ytc = br.render_mix(256, n_buf, pbso.PREC_TC3X)
    rel, mx = assert_waveform_parity(ytc, ref)
    tail = slice(-44100, None)
    if np.linalg.norm(ref[tail]) > 1e-30 * np.linalg.norm(ref):      # high damping: the last second is below FP32's range (e^-135)
        assert np.linalg.norm(ytc[tail] - ref[tail]) <= 1e-5 * np.linalg.norm(ref[tail])
    print("full length vs oracle (%s): tc3x rel-L2 %.2e max-abs %.2e" % (material, rel, mx))


def test_batch_tc3x_longer_render_after_shorter(pbso):
    """Regression (round-1 ADVICE): the cached unit list depends on the exact render length, not only on the number of
    M-tiles -- render 100 buffers, then 120 (same M-tile count), then 40, on one handle with impulses in the added range."""
    n_obj, n_modes = 6, 48
    w = synth.batch_workload(n_obj, n_modes, 120, 31, "high_damping")
    rng = np.random.default_rng(31)
    obj = np.repeat(np.arange(n_obj), 3)
    buf = np.concatenate([[3 + o, 104 + o, 117 - o] for o in range(n_obj)])
    space = rng.standard_normal((len(obj), n_modes))
    br = pbso.BatchRenderer(H, w["a"], w["b"]); br.set_transfer(w["trans"]); br.set_impulses(obj, buf, space)
    for n_buf in (100, 120, 40, 120):
        assert_waveform_parity(br.render_mix(256, n_buf, pbso.PREC_TC3X), br.render_mix(256, n_buf, pbso.PREC_F64))


@pytest.mark.parametrize("BUF", [513, 100, 129, 640, 43])
def test_batch_tc3x_any_buffer_size(pbso, orc, BUF):
    """The tensor-core path at buffer sizes that are not multiples of its 128-sample tiles -- 513 is the reference's
    default FRAMES_PER_BUFFER (modal_solver.h:100): impulses land inside a tile (their first samples come from the
    FP64 recurrence, k_batch_event_heads; from the next tile boundary on they are part of the contraction), the last tile
    is cut at the end of the render.  Mix against the oracle's solver loop, stems against the FP64 kernel, and a
    two-range render with the state handed over."""
    n_obj, n_modes = 5, 200
    n_buf = max(12, 9000 // BUF)
    w = synth.batch_workload(n_obj, n_modes, n_buf, 17, "low_damping")
    rng = np.random.default_rng(17 + BUF)
    obj = np.repeat(np.arange(n_obj), 3)
    buf = np.concatenate([rng.choice(n_buf, 3, replace=False) for _ in range(n_obj)])
    buf[0] = 0; buf[3] = n_buf - 1                                           # first sample of the render / last buffer
    space = rng.standard_normal((len(obj), n_modes))
    want = np.zeros(n_buf * BUF)
    for o in range(n_obj):
        sv = orc.Solver(orc.Integrator(H, w["a"][o], w["b"][o]), BUF); sv.enqueue_trans(w["trans"][o])
        for bi in range(n_buf):
            for e in np.nonzero((obj == o) & (buf == bi))[0]:
                sv.enqueue_force(space[e])
            want[bi * BUF:(bi + 1) * BUF] += sv.step()[0]
    br = pbso.BatchRenderer(H, w["a"], w["b"]); br.set_transfer(w["trans"]); br.set_impulses(obj, buf, space)
    y64 = br.render_mix(BUF, n_buf, pbso.PREC_F64)
    assert_waveform_parity(y64, want, rel=1e-10, mx=1e-10)
    ytc = br.render_mix(BUF, n_buf, pbso.PREC_TC3X)
    assert_waveform_parity(ytc, want)
    s64 = br.render_stems(BUF, n_buf, pbso.PREC_F64); stc = br.render_stems(BUF, n_buf, pbso.PREC_TC3X)
    assert np.abs(stc - s64).max() <= 3e-6 * np.abs(s64).max()
    # an impulse that lands inside the LAST tile of the render and is the only one: nothing reaches the contraction, every
    # sample it produces is a head sample
    b1 = pbso.BatchRenderer(H, w["a"][:1], w["b"][:1]); b1.set_transfer(w["trans"][:1])
    for nb in range(2, 7):
        if ((nb - 1) * BUF) % 128 != 0 and -(-((nb - 1) * BUF) // 128) * 128 >= nb * BUF:
            b1.set_impulses([0], [nb - 1], space[:1])
            assert_waveform_parity(b1.render_mix(BUF, nb, pbso.PREC_TC3X), b1.render_mix(BUF, nb, pbso.PREC_F64), rel=1e-9, mx=1e-9)
    n1 = n_buf // 2
    first = buf < n1
    br.set_impulses(obj[first], buf[first], space[first]); y1 = br.render_mix(BUF, n1, pbso.PREC_TC3X)
    br.set_state(*br.end_state(BUF, n1)); br.set_impulses(obj[~first], buf[~first] - n1, space[~first])
    y2 = br.render_mix(BUF, n_buf - n1, pbso.PREC_TC3X)
    assert_waveform_parity(np.concatenate([y1, y2]), want)


def test_batch_stateful_ranges(pbso, orc):
    """Stateful range renders (pbso_batch_set_state / pbso_batch_get_end_state): a script rendered as two ranges with the
    state handed over equals the one-shot render in every arithmetic; a TransMessage swapped in at the range boundary
    (modal_solver.h:249-252) is the oracle's own solver loop with enqueue_trans; the end state is the pair the oracle's
    ModalIntegrator holds after the same steps; and a range can hand over to the per-buffer path (a Gaussian force,
    forces.h:92-113) and back."""
    n_obj, n_modes, BUF, n1, n2 = 3, 300, 256, 10, 14
    w = synth.batch_workload(n_obj, n_modes, n1 + n2, 91, "low_damping")
    rng = np.random.default_rng(91)
    obj = np.repeat(np.arange(n_obj), 4)
    buf = np.concatenate([[1 + o, 7 - o, 10 + o, 19 + o] for o in range(n_obj)])
    space = rng.standard_normal((len(obj), n_modes))
    T1 = w["trans"]; T2 = T1 * rng.uniform(0.2, 3.0, T1.shape)
    br = pbso.BatchRenderer(H, w["a"], w["b"])

    def ranges(prec):
        br.set_state(None); br.set_transfer(T1)
        first = buf < n1
        br.set_impulses(obj[first], buf[first], space[first])
        y1 = br.render_mix(BUF, n1, prec)
        st = br.end_state(BUF, n1)
        br.set_state(*st); br.set_transfer(T2)
        br.set_impulses(obj[~first], buf[~first] - n1, space[~first])
        y2 = br.render_mix(BUF, n2, prec)
        end = br.end_state(BUF, n2)
        br.set_state(None)
        return np.concatenate([y1, y2]), st, end

    # the oracle's solver loop, one object at a time, with the transfer swapped by a TransMessage dequeued in step n1
    want = np.zeros((n1 + n2) * BUF); states_mid = []; states_end = []
    for o in range(n_obj):
        integ = orc.Integrator(H, w["a"][o], w["b"][o]); sv = orc.Solver(integ, BUF)
        assert sv.enqueue_trans(T1[o])
        ev = {int(b_): space[i] for i, b_ in enumerate(buf) if obj[i] == o}
        for bi in range(n1 + n2):
            if bi in ev:
                assert sv.enqueue_force(ev[bi])
            if bi == n1:
                assert sv.enqueue_trans(T2[o])
                states_mid.append(integ.state())
            y, _ = sv.step()
            want[bi * BUF:(bi + 1) * BUF] += y
        states_end.append(integ.state())
    scale = np.abs(want).max()
    for prec, tol in ((pbso.PREC_F64, 1e-11), (pbso.PREC_F32_TILED, 2e-6), (pbso.PREC_TC3X, 2e-6)):
        got, st, end = ranges(prec)
        assert np.abs(got - want).max() <= tol * scale, prec
        for o in range(n_obj):
            for k in range(2):
                ref_mid, ref_end = states_mid[o][k], states_end[o][k]
                assert np.abs(st[k][o] - ref_mid).max() <= 1e-10 * np.abs(ref_mid).max()
                assert np.abs(end[k][o] - ref_end).max() <= 1e-10 * np.abs(ref_end).max()
    # hand-over to the per-buffer path and back: impulses (batch range) -> a Gaussian force alive for 4 buffers (per-buffer
    # path from the batch's end state) -> free decay (batch range from the integrator's state), against the oracle's solver
    o = 1
    integ = orc.Integrator(H, w["a"][o], w["b"][o]); sv = orc.Solver(integ, BUF)
    assert sv.enqueue_trans(T1[o])
    ev = {int(b_): space[i] for i, b_ in enumerate(buf) if obj[i] == o and b_ < n1}
    g_space = rng.standard_normal(n_modes); want = []
    for bi in range(n1 + 4 + 6):
        if bi in ev:
            assert sv.enqueue_force(ev[bi])
        if bi == n1:
            assert sv.enqueue_force(g_space, orc.GAUSSIAN, 2000.0)
        want.append(sv.step()[0])
    want = np.concatenate(want)
    b1 = pbso.BatchRenderer(H, w["a"][o:o + 1], w["b"][o:o + 1]); b1.set_transfer(T1[o:o + 1])
    sel = (obj == o) & (buf < n1)
    b1.set_impulses(np.zeros(sel.sum(), dtype=np.int32), buf[sel], space[sel])
    got = [b1.render_mix(BUF, n1, pbso.PREC_TC3X)]
    q1, q2 = b1.end_state(BUF, n1)
    it = pbso.ModalIntegrator(n_modes, H, w["a"][o], w["b"][o]); it.set_state(q1[0], q2[0]); it.set_transfer(T1[o][None, :])
    prof, alive = orc.force_profile(orc.GAUSSIAN, 2000.0, BUF, 5)
    assert list(alive) == [1, 1, 1, 1, 0]
    for k in range(4):
        got.append(it.render_buffer(g_space, prof[k])[0][0])
    b1.set_state(*[x[None, :] for x in it.get_state()]); b1.set_impulses([], [], np.zeros((0, n_modes)))
    got.append(b1.render_mix(BUF, 6, pbso.PREC_TC3X))
    got = np.concatenate(got)
    assert np.abs(got - want).max() <= 2e-6 * np.abs(want).max()


def test_batch_tc3x_objects_with_thousands_of_modes(pbso):
    """An object with 8192 modes is 512 K chunks per unit: the FP32 register sums of the epilogue are flushed to the FP64 mix every
    128 chunks, inside a unit too (flushing per unit only cost a factor 2 in max-abs error at this size: 1.3e-6 in the bench at
    64 x 8192)."""
    n_obj, n_modes, n_buf = 3, 8192, 200
    w = synth.batch_workload(n_obj, n_modes, n_buf, 41, "low_damping", first_second_bufs=120)
    br = pbso.BatchRenderer(H, w["a"], w["b"]); br.set_transfer(w["trans"])
    br.set_impulses(np.arange(n_obj), w["imp_buf"], w["space"])
    rel, mx = assert_waveform_parity(br.render_mix(256, n_buf, pbso.PREC_TC3X), br.render_mix(256, n_buf, pbso.PREC_F64))
    assert mx <= 7e-7, mx
    stems = br.render_stems(256, n_buf, pbso.PREC_TC3X); s64 = br.render_stems(256, n_buf, pbso.PREC_F64)
    assert np.abs(stems - s64).max() <= 3e-6 * np.abs(s64).max()


def test_batch_tc3x_object_batches(pbso, monkeypatch):
    """Operand tables larger than the table budget: the renderer walks the objects in batches and rebuilds the tables per
    batch; same waveform as the FP64 kernel."""
    n_obj, n_modes, n_buf = 23, 40, 150
    w = synth.batch_workload(n_obj, n_modes, n_buf, 77, "low_damping", first_second_bufs=100)
    monkeypatch.setenv("PBSO_TC_TABLE_BYTES", str(5 * 3 * 5520))           # room for 5 objects (3 chunks of 5520 bytes each)
    br = pbso.BatchRenderer(H, w["a"], w["b"]); br.set_transfer(w["trans"])
    br.set_impulses(np.arange(n_obj), w["imp_buf"], w["space"])
    y64 = br.render_mix(256, n_buf, pbso.PREC_F64)
    for _ in range(2):
        assert_waveform_parity(br.render_mix(256, n_buf, pbso.PREC_TC3X), y64)


def test_batch_tc3x_cfg1_golden(pbso, golden_dir):
    g = np.load(os.path.join(golden_dir, "cfg1_ball.npz"))
    a, b = synth.ab_from_material(synth.mode_frequencies(64, 1001), synth.MATERIALS["low_damping"])
    br = pbso.BatchRenderer(H, a[None, :], b[None, :])
    br.set_transfer(g["trans"][None, :])
    br.set_impulses([0], [0], (g["space"] * g["scale"])[None, :])
    rel, mx = assert_waveform_parity(br.render_mix(256, 173, pbso.PREC_TC3X), g["y"])
    print("tc3x cfg1: rel-L2 %.2e max-abs %.2e" % (rel, mx))


def test_batch_tc3x_full_size_slice(pbso):
    """cfg5 slice at full length (512 modes x 10 s) against the FP64 kernel, plus linearity and time shift."""
    n_obj, n_modes, n_buf = 48, 512, 1723
    w = synth.batch_workload(n_obj, n_modes, n_buf, 1005)
    br = pbso.BatchRenderer(H, w["a"], w["b"]); br.set_transfer(w["trans"])
    br.set_impulses(np.arange(n_obj), w["imp_buf"], w["space"])
    y64 = br.render_mix(256, n_buf, pbso.PREC_F64)
    ytc = br.render_mix(256, n_buf, pbso.PREC_TC3X)
    rel, mx = assert_waveform_parity(ytc, y64)
    print("cfg5 slice: tc3x vs f64 rel-L2 %.2e max-abs %.2e" % (rel, mx))
    tail = slice(-44100, None)
    assert np.linalg.norm(ytc[tail] - y64[tail]) <= 1e-5 * np.linalg.norm(y64[tail])
    br.set_impulses(np.arange(n_obj), w["imp_buf"], 2.0 * w["space"])
    y2 = br.render_mix(256, n_buf, pbso.PREC_TC3X)
    assert np.max(np.abs(y2 - 2.0 * ytc)) <= 2e-6 * np.max(np.abs(y2))
    br.set_impulses(np.arange(n_obj), w["imp_buf"] + 5, w["space"])
    ys = br.render_mix(256, n_buf, pbso.PREC_TC3X)
    assert np.max(np.abs(ys[5 * 256:] - ytc[:-5 * 256])) <= 2e-6 * np.max(np.abs(ytc))
    assert not ys[:5 * 256 + int(w["imp_buf"].min()) * 256].any()


# --------------------------------------------------------------------------- K6: FFAT map construction
def _fit_case(orc, n_maps, seed, **kw):
    w = synth.ffat_fit_workload(n_maps, seed, **kw)
    fit = orc.ffat_fit_geometry(w["cell_size"], w["V"], w["n_elements"])
    return w, fit


@pytest.mark.parametrize("half_cells,cell,centre", [
    ((3, 5, 6), 0.25, (0.0, 0.0, 0.0)),
    ((8, 12, 16), 0.09375, (0.0, 0.0, 0.0)),
    (((2, 3, 4), (3, 5, 6), (5, 6, 8)), 0.2, (0.1, -0.05, 0.2)),            # ragged: 472 directions, non-cubic faces
    (((2, 3, 4), (3, 5, 6), (5, 6, 8), (7, 8, 9)), 0.2, (0.0, 0.0, 0.0)),   # four shells: generic kernel
    ((6, 5, 3), 0.25, (0.0, 0.0, 0.0)),                                     # shell 2 innermost
])
@pytest.mark.parametrize("scaling", [False, True])
def test_ffat_fit_vs_oracle(pbso, orc, half_cells, cell, centre, scaling):
    """FFAT_Map<T,3> constructor + Solve (+ Scaling) on the device against the restatement (itself pinned to the
    reference's own Solve in tests/test_oracle_vs_ref.py).  FP64 throughout; what differs is the order of three-
    and four-term sums and, with scaling, of the two reductions over directions."""
    w, fit = _fit_case(orc, 7, 41, half_cells=half_cells, cell_size=cell, center=centre, noise=0.05)
    ft = pbso.FFATFitter(w["cell_size"], w["V"], w["n_elements"])
    assert (ft.n_shells, ft.n_elements_total, ft.n_directions) == (len(fit["strides"]), fit["n_total"], fit["n_dir"])
    assert np.array_equal(ft.strides, fit["strides"])
    for s in range(ft.n_shells):
        g, ig = ft.shell(s)
        assert np.array_equal(g, fit["geom"][s]) and np.array_equal(ig, fit["igeom"][s])      # constructor: bit-exact
    psi, scale = ft.Solve(w["k"], w["pressure"], scaling)
    want, wscale = orc.ffat_fit_solve(fit, w["k"], w["pressure"], scaling)
    assert np.allclose(psi, want, rtol=1e-12, atol=0)
    assert np.allclose(scale, wscale, rtol=1e-12, atol=0)


def test_ffat_fit_many_modes_and_device_entry(pbso, orc):
    """Enough modes that blocks fold several modes (stencil reuse) and the host entry chunks its staging; the
    device-resident entry must give the same bits as the host entry."""
    import torch
    w, fit = _fit_case(orc, 300, 43, half_cells=(4, 6, 8), cell_size=0.1875)
    ft = pbso.FFATFitter(w["cell_size"], w["V"], w["n_elements"])
    psi, scale = ft.Solve(w["k"], w["pressure"], True)
    want, wscale = orc.ffat_fit_solve(fit, w["k"], w["pressure"], True)
    assert np.allclose(psi, want, rtol=1e-12, atol=0) and np.allclose(scale, wscale, rtol=1e-12, atol=0)
    assert ft.last_kernel_ms() > 0
    dk = torch.from_numpy(w["k"]).cuda()
    dp = torch.from_numpy(np.ascontiguousarray(w["pressure"]).view(np.float64)).cuda()
    dpsi = torch.empty(300, ft.n_directions, dtype=torch.float64, device="cuda")
    dscale = torch.empty(300, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    st = torch.cuda.Stream()              # an explicit stream: a NULL handle would select the fitter's own stream
    ft.solve_device(300, dk.data_ptr(), dp.data_ptr(), dpsi.data_ptr(), True, dscale.data_ptr(), st.cuda_stream)
    st.synchronize()
    assert np.array_equal(dpsi.cpu().numpy(), psi) and np.array_equal(dscale.cpu().numpy(), scale)
    # packed layout (one complex per quad = the even entries, the only ones Solve reads, ffat_solver.h:1054-1056): same bits,
    # host and device entries; deferred scale: Psi unscaled + the factor
    packed = np.ascontiguousarray(w["pressure"][:, 0::2])
    psi_p, scale_p = ft.Solve(w["k"], packed, True, packed=True)
    assert np.array_equal(psi_p, psi) and np.array_equal(scale_p, scale)
    dpp = torch.from_numpy(packed.view(np.float64)).cuda()
    ft.solve_device(300, dk.data_ptr(), dpp.data_ptr(), dpsi.data_ptr(), True, dscale.data_ptr(), st.cuda_stream, packed=True, defer_scale=True)
    st.synchronize()
    psi_u, _ = ft.Solve(w["k"], w["pressure"], False)
    assert np.array_equal(dpsi.cpu().numpy(), psi_u) and np.array_equal(dscale.cpu().numpy(), scale)
    assert np.array_equal(psi_u * scale[:, None], psi)


def test_ffat_fit_feeds_runtime_map(pbso, orc):
    """Solve -> run-time map set -> computeTransfer: the whole chain against the oracle's chain, and the property the
    fit exists for: on shell 2 itself |GetMapVal| reproduces a pure 1/(kr) field sampled on that shell."""
    w, fit = _fit_case(orc, 5, 47, half_cells=(4, 6, 8), cell_size=0.1875)
    ft = pbso.FFATFitter(w["cell_size"], w["V"], w["n_elements"])
    psi, _ = ft.Solve(w["k"], w["pressure"], False)
    maps = ft.to_maps(w["k"], psi)
    pos = synth.listeners(64, 3)
    got = maps.computeTransfer(pos)
    opsi, _ = orc.ffat_fit_solve(fit, w["k"], w["pressure"], False)
    g, ig = fit["geom"][2], fit["igeom"][2]
    omaps = [dict(cellsize=g[0], lowcorners=g[1:19].reshape(6, 3), center1=g[19:22], bboxlow=g[22:25], bboxtop=g[25:28],
                  center=g[28:31], k=w["k"][m], n_elements=ig[:12].reshape(6, 2), strides=ig[12:], psi=opsi[m], modeid=m) for m in range(5)]
    assert np.allclose(got, orc.ffat_eval(omaps, pos), rtol=1e-11, atol=0)
    # uniform-magnitude field: closed form of tests/test_oracle_kat.py::test_ffat_fit_uniform_field_closed_form
    A, k = 2.0, 5.0
    P = np.full((1, 2 * ft.n_elements_total), A * np.exp(0.7j))
    psi1, _ = ft.Solve([k], P, False)
    centres = w["V"][4 * fit["strides"][2]:].reshape(-1, 4, 3).mean(axis=1)
    rho = np.array([4, 6, 8]) / 8.0
    assert np.allclose(psi1[0], A * k * np.linalg.norm(centres, axis=1) * np.sum(1 / rho) / np.sum(1 / rho ** 2), rtol=1e-12)


def test_ffat_fit_rejects_bad_arguments(pbso):
    V, ne = synth.cubemap_vertices((0, 0, 0), 2, 0.5)
    with pytest.raises(pbso.PbsoError) as e:                 # N_shells must reach _shells.at(2) (ffat_solver.h:982)
        pbso.FFATFitter(0.5, np.concatenate([V, V]), np.stack([ne, ne]))
    assert e.value.code == 1
    with pytest.raises(pbso.PbsoError) as e:                 # V.block would run past the vertex list (:971)
        pbso.FFATFitter(0.5, np.concatenate([V, V]), np.stack([ne, ne, ne]))
    assert e.value.code == 1


def test_ffat_fit_header_mirror_caller(pbso, orc, tmp_path):
    """tests/cpp/ffat_fit_main.cpp: ReadNElementsFile + ReadComplexVector (text and binary) + FFAT_Map<double,3> ctor +
    Solve + FFAT_Map_Serialize::Save/Load/Check + GetMapVal through the header mirror."""
    import subprocess
    from oracle import fatcube
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    inc = os.path.join(root, "include", "openpbso"); libdir = os.path.join(root, "openpbso_b200")
    exe = str(tmp_path / "ffat_fit_main")
    r = subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-I" + os.path.join(inc, "eigen_shim"), "-I" + inc,
                        os.path.join(root, "tests", "cpp", "ffat_fit_main.cpp"), "-L" + libdir, "-lpbso_b200",
                        "-Wl,-rpath," + libdir, "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    w, fit = _fit_case(orc, 1, 53, half_cells=((2, 3, 4), (3, 5, 6), (5, 6, 8)), cell_size=0.2)
    nfile = str(tmp_path / "n_elements.txt")
    with open(nfile, "w") as f:
        for shell in w["n_elements"]:
            f.write(" ".join("%d %d" % (a, b) for a, b in shell) + "\n")
    vfile = str(tmp_path / "V.f64"); np.ascontiguousarray(w["V"]).tofile(vfile)
    P = w["pressure"][0]
    pbin = str(tmp_path / "p.bin")
    with open(pbin, "wb") as f:
        f.write(np.int32(2 * len(P)).tobytes()); f.write(np.ascontiguousarray(P).view(np.float64).tobytes())
    ptxt = str(tmp_path / "p.txt")
    with open(ptxt, "w") as f:
        for z in P:
            f.write("%.17g %.17g\n" % (z.real, z.imag))
    probe = (2.0, -1.0, 3.5)
    for scaling in (0, 1):
        want, _ = orc.ffat_fit_solve(fit, w["k"], w["pressure"], bool(scaling))
        for pfile, binary in ((pbin, 1), (ptxt, 0)):
            out = str(tmp_path / ("m%d%d.fatcube" % (scaling, binary)))
            r = subprocess.run([exe, nfile, vfile, repr(w["cell_size"]), "6", repr(float(w["k"][0])), pfile, str(binary), str(scaling),
                                out] + [repr(x) for x in probe], capture_output=True, text=True)
            assert r.returncode == 0, r.stderr
            rows, cols, v1, v2, amp, c1, c2, v3, n_legacy, l1, l2 = r.stdout.split()
            assert (int(rows), int(cols)) == (fit["n_dir"], 1) and v1 == v2 == v3 and c1 == c2
            assert int(n_legacy) == 1 and l1 == c1 and l2 == v1          # legacy Save / Load / LoadAll: the compressed map survives (both views)
            d = fatcube.load(out)
            assert d["modeid"] == 6 and np.allclose(d["psi"], want[0], rtol=1e-12, atol=0)
            g, ig = fit["geom"][2], fit["igeom"][2]
            m = dict(cellsize=g[0], lowcorners=g[1:19].reshape(6, 3), center1=g[19:22], bboxlow=g[22:25], bboxtop=g[25:28],
                     center=g[28:31], k=w["k"][0], n_elements=ig[:12].reshape(6, 2), strides=ig[12:], psi=want[0], modeid=0)
            assert np.isclose(float(v1), orc.ffat_eval([m], [probe])[0, 0], rtol=1e-11)
            md = dict(m, psi=d["psi"])                                   # Compress works on the Psi the device fitted
            q, amp6, gmax = orc.ffat_quantise(md)
            assert float(amp) == gmax
            assert np.isclose(float(c1), orc.ffat_eval([dict(md, psi=orc.ffat_dequantise(md, q, amp6))], [probe])[0, 0], rtol=1e-12)
            dc = fatcube.load(out + ".compressed")
            assert dc["is_compressed"] and np.array_equal(dc["psi"], orc.ffat_dequantise(md, q, amp6))


def test_ffat_fit_reproduces_the_reference_fixture(pbso, golden_dir):
    """K6 against Psi computed by the reference's OWN Solve (tests/golden/ffat_fit.npz, generated here with oracle/_ref):
    no oracle in between.  Non-cubic shells, four modes with wavenumbers from 1.5 to 120, unread odd entries filled with noise."""
    g = np.load(os.path.join(golden_dir, "ffat_fit.npz"))
    ft = pbso.FFATFitter(float(g["cell_size"]), g["V"], g["n_elements"])
    geom, igeom = ft.shell(2)
    assert np.array_equal(geom[1:19], g["shell2_lowcorners"].ravel()) and np.array_equal(igeom[12:], g["shell2_strides"])
    assert np.array_equal(geom[22:25], g["shell2_bboxlow"]) and np.array_equal(geom[25:28], g["shell2_bboxtop"])
    assert np.array_equal(geom[28:31], g["centre"])
    for scaling, key in ((False, "psi"), (True, "psi_scaled")):
        psi, _ = ft.Solve(g["k"], g["pressure"], scaling)
        assert np.allclose(psi, g[key], rtol=1e-12, atol=0)


def test_legacy_fatcube_directory_evaluates_like_the_reference(pbso, golden_dir):
    """K3 on maps read from the LEGACY .fatcube form (tests/golden/legacy_fatcube/, written by the reference's own
    FFAT_Map<double,3>::Save through libigl's own igl::serialize) against |GetMapVal| computed by the reference's own legacy loader
    (legacy_eval.npz): maps of different sizes, so the per-map kernel."""
    g = np.load(os.path.join(golden_dir, "legacy_eval.npz"))
    fm = pbso.FFATMaps.LoadAll(os.path.join(golden_dir, "legacy_fatcube"))
    assert np.allclose(fm.computeTransfer(g["pos"]), g["out"], rtol=1e-12, atol=0)


def test_ffat_compress_reproduces_the_opencv_fixture(pbso, orc, golden_dir, tmp_path):
    """FFAT_Map<T,3>::Compress (ffat_solver.h:1125-1178) against tests/golden/ffat_compress.npz (OpenCV's own cast and JPEG
    codec): quantise() gives OpenCV's bytes; the bytes that came back from the JPEG file, handed to set_compressed_u8(),
    give _compressed_Psi bit for bit; GetMapVal(p, getCompressed=true) read from the device's BYTE table equals the oracle
    on the stored doubles, and is bit-identical to the same kernels reading those doubles; Save/Load keeps the view."""
    g = np.load(os.path.join(golden_dir, "ffat_compress.npz"))
    maps = synth.ffat_maps(g["freqs"], 2000, n=8)
    for i, m in enumerate(maps):
        m["psi"] = g["psi"][i]
    fm = pbso.FFATMaps.from_dicts(maps)
    with pytest.raises(pbso.PbsoError):
        fm.computeTransfer(synth.listeners(4, 1), use_compressed=True)           # asserts _is_compressed (:1183-1186)
    for i in range(3):
        q, amp, gmax = fm.quantise(i)
        assert np.array_equal(q, g["q8_pre"][i]) and np.array_equal(amp, g["max_amp"][i]) and gmax == g["max_amp_global"][i]
        fm.set_compressed_u8(i, g["q8_post"][i], amp)
        q2, sc, c = fm.get_compressed(i)
        assert np.array_equal(q2, g["q8_post"][i]) and np.array_equal(c, g["compressed_psi"][i])
    cmaps = [dict(m, psi=g["compressed_psi"][i], is_compressed=True) for i, m in enumerate(maps)]
    fd = pbso.FFATMaps.from_dicts(cmaps)                                         # the same view held as doubles
    for L in (80, 700, 3000):                                                    # fused / locate + gather (twice: no byte tiles)
        pos = synth.listeners(L, 77)
        got = fm.computeTransfer(pos, use_compressed=True)
        assert np.array_equal(got, fd.computeTransfer(pos[:L], use_compressed=True)) or L >= 2048
        assert np.allclose(got, orc.ffat_eval(cmaps, pos), rtol=1e-12, atol=0)
        # _Psi is still there (Compress keeps it)
        assert np.allclose(fm.computeTransfer(pos), orc.ffat_eval(maps, pos), rtol=1e-12, atol=0)
    # a multiple of four maps and many listeners: the four-maps-per-thread byte kernel (products by (maxAmp/255)/k and 1/r
    # instead of the division: tolerance, not bits)
    m8 = synth.ffat_maps(synth.mode_frequencies(8, 1004), 2000, n=8)
    f8 = pbso.FFATMaps.from_dicts(m8); f8.Compress()
    c8 = [dict(m, psi=f8.get_compressed(i)[2]) for i, m in enumerate(m8)]
    for L in (257, 3000):
        pos = synth.listeners(L, 79)
        assert np.allclose(f8.computeTransfer(pos, use_compressed=True), orc.ffat_eval(c8, pos), rtol=1e-12, atol=0)
    # per-map geometry (the general kernel) reads bytes too: a set whose second map has another cell size
    odd = [dict(maps[0]), dict(maps[1], cellsize=maps[1]["cellsize"] * 1.0000001)]
    fo = pbso.FFATMaps.from_dicts(odd)
    fo.Compress()
    codd = [dict(m, psi=fo.get_compressed(i)[2]) for i, m in enumerate(odd)]
    pos = synth.listeners(300, 78)
    assert np.allclose(fo.computeTransfer(pos, use_compressed=True), orc.ffat_eval(codd, pos), rtol=1e-12, atol=0)
    # Save writes _compressed_Psi (ffat_map_serialize.h:149-153); a map loaded compressed has no _Psi and cannot be compressed
    fn = str(tmp_path / "m.fatcube")
    fm.Save(1, fn)
    back = pbso.FFATMaps.Load(fn)
    got = back.get_map(1)
    assert got["is_compressed"] and np.array_equal(got["psi"], g["compressed_psi"][1])
    with pytest.raises(pbso.PbsoError):
        back.Compress(1)
    # Compress() of every map without a codec in between == quantise + de-quantise of the oracle
    fa = pbso.FFATMaps.from_dicts(maps)
    gm = fa.Compress()
    for i, m in enumerate(maps):
        q, amp, gmax = orc.ffat_quantise(m)
        assert gm[i] == gmax and np.array_equal(fa.get_compressed(i)[2], orc.ffat_dequantise(m, q, amp))


def test_ffat_eval_reproduces_the_reference_fixture(pbso, golden_dir):
    """K3 against |GetMapVal| computed by the reference's OWN code (tests/golden/ffat_eval.npz): the per-map kernel on the
    committed .fatcube files, and on 24 maps sharing one geometry the fused few-listener kernel, the locate + gather pair and
    the texel-stationary kernel -- probe positions incl. face axes, edge / corner ties and points inside the box."""
    g = np.load(os.path.join(golden_dir, "ffat_eval.npz"))
    got = pbso.FFATMaps.LoadAll(os.path.join(golden_dir, "fatcube")).computeTransfer(g["files_pos"])
    ref = g["files_out"]
    assert np.array_equal(np.isfinite(got), np.isfinite(ref))
    fin = np.isfinite(ref)
    assert np.allclose(got[fin], ref[fin], rtol=1e-12, atol=0)
    fm = pbso.FFATMaps.from_dicts(synth.ffat_maps(g["shared_freqs"], 2000, n=8))
    pos, want = g["shared_pos"], g["shared_out"]
    assert np.allclose(fm.computeTransfer(pos[:80]), want[:80], rtol=1e-12, atol=0)          # fused single launch (L <= 256)
    assert np.allclose(fm.computeTransfer(pos[:700]), want[:700], rtol=1e-12, atol=0)        # locate + gather
    assert np.allclose(fm.computeTransfer(pos), want, rtol=1e-12, atol=0)                    # texel tiles (L >= 2048)
