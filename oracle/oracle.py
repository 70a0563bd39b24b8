"""ctypes front-end of the CPU oracle (oracle/pbso_oracle.cpp) -- TEST INFRASTRUCTURE ONLY.

Every wrapper names the reference file:line its C++ side restates.  `ref()` loads the optional
oracle/_ref/libpbso_ref.so (the reference's own headers compiled against the Eigen shim) used to
pin this oracle.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)


def _dp(a):
    return None if a is None else a.ctypes.data_as(c_dp)


def _ip(a):
    return None if a is None else a.ctypes.data_as(c_ip)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _cpu_stamp():
    """The oracle is built -march=native; a .so that travelled from another host must be rebuilt."""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    import hashlib
                    return hashlib.sha1(line.encode()).hexdigest()
    except OSError:
        pass
    return "unknown"


def build(force=False):
    """Compile libpbso_oracle.so (and oracle/_ref when /root/reference exists)."""
    so = os.path.join(_HERE, "libpbso_oracle.so")
    src = os.path.join(_HERE, "pbso_oracle.cpp")
    stamp_file = os.path.join(_HERE, ".oracle_cpu_stamp")
    stamp = _cpu_stamp()
    old = open(stamp_file).read() if os.path.exists(stamp_file) else ""
    if force or old != stamp or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-B", "-C", _HERE, "libpbso_oracle.so"], stdout=subprocess.DEVNULL)
        with open(stamp_file, "w") as f:
            f.write(stamp)
    if os.path.isdir("/root/reference"):
        ref_so = os.path.join(_HERE, "_ref", "libpbso_ref.so")
        bridge = os.path.join(_HERE, "ref_bridge.cpp")
        if force or not os.path.exists(ref_so) or os.path.getmtime(ref_so) < os.path.getmtime(bridge):
            subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


def lib():
    global _LIB
    if _LIB is None:
        build()
        L = C.CDLL(os.path.join(_HERE, "libpbso_oracle.so"))
        L.orc_integrator_create.restype = C.c_void_p
        L.orc_integrator_create.argtypes = [C.c_int, C.c_double, c_dp, c_dp]
        L.orc_integrator_destroy.argtypes = [C.c_void_p]
        L.orc_integrator_step.argtypes = [C.c_void_p, c_dp, c_dp]
        L.orc_integrator_get_state.argtypes = [C.c_void_p, c_dp, c_dp]
        L.orc_solver_create.restype = C.c_void_p
        L.orc_solver_create.argtypes = [C.c_int, C.c_int, C.c_void_p]
        L.orc_solver_destroy.argtypes = [C.c_void_p]
        L.orc_solver_enqueue_force.argtypes = [C.c_void_p, c_dp, C.c_int, C.c_double, C.c_int]
        L.orc_solver_enqueue_trans.argtypes = [C.c_void_p, c_dp, C.c_int]
        L.orc_solver_enqueue_arprm.argtypes = [C.c_void_p] + [C.c_double] * 4
        L.orc_solver_set_use_transfer.argtypes = [C.c_void_p, C.c_int]
        L.orc_solver_step.argtypes = [C.c_void_p, c_dp, c_dp]
        L.orc_solver_num_active.argtypes = [C.c_void_p]
        L.orc_solver_last_force.argtypes = [C.c_void_p, c_dp, c_dp]
        L.orc_solver_latest_transfer.argtypes = [C.c_void_p, c_dp]
        L.orc_build_ab.argtypes = [C.c_double, c_dp, C.c_int, C.c_double, C.c_double, c_dp, c_dp]
        L.orc_coeffs.argtypes = [C.c_int, C.c_double, c_dp, c_dp, c_dp, c_dp, c_dp]
        L.orc_force_profile.argtypes = [C.c_int, C.c_double, C.c_int, C.c_int, c_dp, c_ip]
        L.orc_project_vertex.argtypes = [C.c_int, c_dp, C.c_int, C.c_int, c_dp, c_dp]
        L.orc_project_face.argtypes = [C.c_int, c_dp, C.c_int, c_ip, c_dp, c_dp, c_dp]
        L.orc_project_dense.argtypes = [C.c_int, C.c_int, C.c_int, c_dp, c_dp, c_dp]
        L.orc_ffat_eval.argtypes = [C.c_int, c_dp, c_ip, c_dp, C.c_int, c_dp, C.c_int, c_dp]
        L.orc_ffat_intersect.argtypes = [c_dp, c_ip, c_dp, c_dp, c_ip]
        L.orc_ffat_interpolate.argtypes = [c_dp, c_ip, c_dp, c_ip, c_ip, c_dp]
        L.orc_num_modes_audible.argtypes = [c_dp, C.c_int, C.c_double, C.c_double, c_dp]
        L.orc_ffat_fit_geometry.argtypes = [C.c_double, c_dp, c_ip, C.c_int, C.c_double, c_dp, c_ip, c_ip, c_ip]
        L.orc_ffat_fit_solve.argtypes = [C.c_int, c_dp, c_ip, c_ip, C.c_int, c_dp, c_dp, C.c_int, c_dp, c_dp, c_dp, c_dp]
        L.orc_material_read.argtypes = [C.c_char_p, c_dp]
        L.orc_modes_read_header.argtypes = [C.c_char_p, c_ip, c_ip]
        L.orc_modes_read.argtypes = [C.c_char_p, c_dp, c_dp]
        L.orc_modes_write.argtypes = [C.c_char_p, C.c_int, C.c_int, c_dp, c_dp]
        L.orc_batch_render.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, c_dp, c_dp,
                                       c_dp, c_dp, c_ip, c_dp]
        _LIB = L
    return _LIB


def ref():
    """The reference's own headers compiled against the Eigen shim, or None when not built."""
    global _REF
    if _REF is None:
        build()
        p = os.path.join(_HERE, "_ref", "libpbso_ref.so")
        if not os.path.exists(p):
            return None
        R = C.CDLL(p)
        R.ref_integrator_build.restype = C.c_void_p
        R.ref_integrator_build.argtypes = [C.c_double, c_dp, C.c_int, C.c_double, C.c_double,
                                           C.c_double, C.c_int]
        R.ref_integrator_create.restype = C.c_void_p
        R.ref_integrator_create.argtypes = [C.c_int, C.c_double, c_dp, c_dp]
        R.ref_integrator_destroy.argtypes = [C.c_void_p]
        R.ref_integrator_step.argtypes = [C.c_void_p, C.c_int, c_dp, c_dp]
        R.ref_force_profile.argtypes = [C.c_int, C.c_double, C.c_int, C.c_int, c_dp, c_ip]
        R.ref_num_modes_audible.argtypes = [c_dp, C.c_int, C.c_double, C.c_double, C.c_int, c_ip]
        R.ref_modes_roundtrip.argtypes = [C.c_char_p, C.c_char_p, c_ip, c_ip, c_dp]
        R.ref_material_read.argtypes = [C.c_char_p, c_dp]
        R.ref_material_xi.restype = C.c_double
        R.ref_material_xi.argtypes = [C.c_double] * 3
        R.ref_material_omega_di.restype = C.c_double
        R.ref_material_omega_di.argtypes = [C.c_double] * 3
        R.ref_solver_create.restype = C.c_void_p
        R.ref_solver_create.argtypes = [C.c_int, C.c_int, C.c_void_p]
        R.ref_solver_destroy.argtypes = [C.c_void_p]
        R.ref_solver_enqueue_force.argtypes = [C.c_void_p, c_dp, C.c_int, C.c_double, C.c_int]
        R.ref_solver_enqueue_trans.argtypes = [C.c_void_p, c_dp, C.c_int]
        R.ref_solver_enqueue_arprm.argtypes = [C.c_void_p] + [C.c_double] * 4
        R.ref_solver_set_use_transfer.argtypes = [C.c_void_p, C.c_int]
        R.ref_solver_step.argtypes = [C.c_void_p, c_dp, c_dp]
        R.ref_solver_latest_transfer.argtypes = [C.c_void_p, c_dp]
        R.ref_solver_read_ffat_maps.argtypes = [C.c_void_p, C.c_char_p]
        R.ref_solver_compute_transfer.argtypes = [C.c_void_p, c_dp, c_dp]
        R.ref_solver_compute_transfer_enqueue.argtypes = [C.c_void_p, c_dp]
        R.ref_ffat_load_all.restype = C.c_void_p
        R.ref_ffat_load_all.argtypes = [C.c_char_p, c_ip]
        R.ref_ffat_free.argtypes = [C.c_void_p]
        R.ref_ffat_eval.argtypes = [C.c_void_p, c_dp, C.c_int, C.c_int, c_dp]
        R.ref_ffat_load_save.argtypes = [C.c_char_p, C.c_char_p, c_ip, c_dp]
        R.ref_ffat_legacy_from_fatcube.argtypes = [C.c_char_p, C.c_char_p]
        R.ref_ffat_legacy_fit_save.argtypes = [C.c_int, C.c_double, c_dp, C.c_int, c_ip, C.c_int, C.c_double, c_dp, C.c_int, C.c_char_p]
        R.ref_ffat_legacy_load_all.restype = C.c_void_p
        R.ref_ffat_legacy_load_all.argtypes = [C.c_char_p, c_ip]
        R.ref_list_dir_files.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
        R.ref_batch_render.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, c_dp, c_dp, c_dp, c_dp, c_ip, c_dp]
        R.ref_cubemap_mesh.argtypes = [c_ip, c_ip, C.c_double, c_dp, c_ip, c_dp, C.c_int, c_ip, c_ip]
        R.ref_ffat_fit.argtypes = [C.c_int, C.c_double, c_dp, C.c_int, c_ip, C.c_int, C.c_double, c_dp, C.c_int,
                                   c_dp, c_dp, C.c_char_p]
        _REF = R
    return _REF


# ---------------------------------------------------------------------------------------------
def build_ab(density, omega2, alpha, beta, N=None):
    """modal_integrator.h:47-70"""
    omega2 = _f64(omega2)
    N = len(omega2) if N is None or N < 0 else N
    a = np.empty(N); b = np.empty(N)
    lib().orc_build_ab(density, _dp(omega2), N, alpha, beta, _dp(a), _dp(b))
    return a, b


def coeffs(h, a, b):
    """modal_integrator.h:86-100"""
    a = _f64(a); b = _f64(b); N = len(a)
    c1 = np.empty(N); c2 = np.empty(N); c3 = np.empty(N)
    lib().orc_coeffs(N, h, _dp(a), _dp(b), _dp(c1), _dp(c2), _dp(c3))
    return c1, c2, c3


class Integrator:
    """modal_integrator.h:19-123"""

    def __init__(self, h, a, b):
        self.a = _f64(a); self.b = _f64(b); self.N = len(self.a); self.h = h
        self._p = lib().orc_integrator_create(self.N, h, _dp(self.a), _dp(self.b))

    def step(self, Q=None):
        out = np.empty(self.N)
        Qc = None if Q is None else _f64(Q)
        lib().orc_integrator_step(self._p, _dp(Qc), _dp(out))
        return out

    def state(self):
        q1 = np.empty(self.N); q2 = np.empty(self.N)
        lib().orc_integrator_get_state(self._p, _dp(q1), _dp(q2))
        return q1, q2

    def __del__(self):
        if getattr(self, "_p", None):
            lib().orc_integrator_destroy(self._p); self._p = None


POINT, GAUSSIAN, AR = 0, 1, 2
F_SUSTAIN_START, F_SUSTAIN_END, F_CLEAR = 1, 2, 4


class Solver:
    """modal_solver.h:100-276 (single-threaded; the sound queue is drained by the caller)."""

    def __init__(self, integrator, BUF=256):
        self.integ = integrator; self.N = integrator.N; self.BUF = BUF
        self._p = lib().orc_solver_create(self.N, BUF, integrator._p)

    def enqueue_force(self, data, ftype=POINT, width_us=0.0, flags=0):
        d = _f64(data); assert len(d) == self.N
        return bool(lib().orc_solver_enqueue_force(self._p, _dp(d), ftype, width_us, flags))

    def enqueue_trans(self, data):
        d = _f64(data)
        return bool(lib().orc_solver_enqueue_trans(self._p, _dp(d), len(d)))

    def enqueue_arprm(self, a0, a1, sigma, mu):
        return bool(lib().orc_solver_enqueue_arprm(self._p, a0, a1, sigma, mu))

    def set_use_transfer(self, use):
        lib().orc_solver_set_use_transfer(self._p, int(use))

    def step(self):
        """Returns (sound[BUF], qnorm[N]) or None when the step produced no buffer."""
        y = np.empty(self.BUF); qn = np.empty(self.N)
        if lib().orc_solver_step(self._p, _dp(y), _dp(qn)):
            return y, qn
        return None

    def num_active(self):
        return lib().orc_solver_num_active(self._p)

    def last_force(self):
        """(space[N], time[BUF]) of the last step (modal_solver.h:206-240)."""
        sp = np.empty(self.N); tm = np.empty(self.BUF)
        lib().orc_solver_last_force(self._p, _dp(sp), _dp(tm))
        return sp, tm

    def latest_transfer(self):
        t = np.empty(self.N)
        lib().orc_solver_latest_transfer(self._p, _dp(t))
        return t

    def __del__(self):
        if getattr(self, "_p", None):
            lib().orc_solver_destroy(self._p); self._p = None


class RefSolver:
    """The REFERENCE's own ModalSolver<double,BUF> + ModalIntegrator<double> (modal_solver.h,
    modal_integrator.h compiled in place into oracle/_ref), behind the same methods as Solver.
    BUF must be one of the sizes oracle/ref_bridge.cpp instantiates (64, 256, 513)."""

    def __init__(self, h, a, b, BUF=256):
        R = ref()
        assert R is not None, "oracle/_ref not built"
        a = _f64(a); b = _f64(b)
        self.N = len(a); self.BUF = BUF
        self._i = R.ref_integrator_create(self.N, h, _dp(a), _dp(b))
        self._p = R.ref_solver_create(self.N, BUF, self._i)
        assert self._p, "BUF=%d not instantiated in ref_bridge.cpp" % BUF

    def enqueue_force(self, data, ftype=POINT, width_us=0.0, flags=0):
        d = _f64(data); assert len(d) == self.N
        return bool(ref().ref_solver_enqueue_force(self._p, _dp(d), ftype, width_us, flags))

    def enqueue_trans(self, data):
        d = _f64(data)
        return bool(ref().ref_solver_enqueue_trans(self._p, _dp(d), len(d)))

    def enqueue_arprm(self, a0, a1, sigma, mu):
        return bool(ref().ref_solver_enqueue_arprm(self._p, a0, a1, sigma, mu))

    def set_use_transfer(self, use):
        ref().ref_solver_set_use_transfer(self._p, int(use))

    def step(self):
        y = np.empty(self.BUF); qn = np.empty(self.N)
        if ref().ref_solver_step(self._p, _dp(y), _dp(qn)):
            return y, qn
        return None

    def latest_transfer(self):
        t = np.empty(self.N)
        ref().ref_solver_latest_transfer(self._p, _dp(t))
        return t

    def read_ffat_maps(self, dirname):
        ref().ref_solver_read_ffat_maps(self._p, dirname.encode())

    def compute_transfer(self, pos, n):
        """computeTransfer(pos, T*) (modal_solver.h:303-315); None when no maps are loaded."""
        out = np.empty(n); pos = _f64(pos)
        return out if ref().ref_solver_compute_transfer(self._p, _dp(pos), _dp(out)) else None

    def compute_transfer_enqueue(self, pos):
        """computeTransfer(pos) (modal_solver.h:286-300): evaluates and enqueues a TransMessage."""
        pos = _f64(pos)
        return bool(ref().ref_solver_compute_transfer_enqueue(self._p, _dp(pos)))

    def __del__(self):
        R = _REF
        if R is not None and getattr(self, "_p", None):
            R.ref_solver_destroy(self._p); self._p = None
            R.ref_integrator_destroy(self._i); self._i = None


def ref_ffat_eval(dirname, pos, use_compressed=False):
    """The reference's LoadAll + |GetMapVal| over a directory of .fatcube files -> [L][N] or None."""
    R = ref(); n = C.c_int()
    h = R.ref_ffat_load_all(dirname.encode(), C.byref(n))
    pos = _f64(pos).reshape(-1, 3); L = len(pos)
    out = np.empty((L, n.value))
    ok = R.ref_ffat_eval(h, _dp(pos), L, int(use_compressed), _dp(out))
    R.ref_ffat_free(h)
    return out if ok else None


def ref_ffat_eval_legacy(dirname, pos):
    """The reference's LEGACY loader -- FFAT_Map<double,3>::LoadAll (igl::deserialize, ffat_solver.h:1069-1085) -- + |GetMapVal|
    over a directory of legacy .fatcube files -> [L][N] or None."""
    R = ref(); n = C.c_int()
    h = R.ref_ffat_legacy_load_all(dirname.encode(), C.byref(n))
    pos = _f64(pos).reshape(-1, 3); L = len(pos)
    out = np.empty((L, n.value))
    ok = R.ref_ffat_eval(h, _dp(pos), L, 0, _dp(out))
    R.ref_ffat_free(h)
    return out if ok else None


def ref_ffat_legacy_from_fatcube(in_file, out_file):
    """FFAT_Map_Serialize::Load(protobuf file) -> FFAT_Map<double,3>::Save(legacy file), the reference's own code; returns modeId."""
    return ref().ref_ffat_legacy_from_fatcube(in_file.encode(), out_file.encode())


def ref_ffat_legacy_fit_save(mode_id, cell_size, V, n_elements, k, pressure, power_scaling, out_file):
    """The reference's constructor + Solve, then FFAT_Map<double,3>::Save(legacy file): a complete legacy map, three shells."""
    V = _f64(V); ne = np.ascontiguousarray(n_elements, dtype=np.int32)
    P = np.ascontiguousarray(pressure, dtype=np.complex128).view(np.float64)
    return ref().ref_ffat_legacy_fit_save(int(mode_id), float(cell_size), _dp(V), len(V), _ip(ne), ne.shape[0], float(k), _dp(P),
                                          int(power_scaling), out_file.encode())


def force_profile(ftype, width_us, BUF, n_buf):
    """forces.h:81-128"""
    out = np.empty((n_buf, BUF)); alive = np.empty(n_buf, dtype=np.int32)
    lib().orc_force_profile(ftype, width_us, BUF, n_buf, _dp(out), _ip(alive))
    return out, alive


def project_vertex(U, vid, vn, forceDim=None):
    """tools/real_time_modal_sound.cpp:268-280; U is [nModes][nDOF] (ModeData.h:23-24)."""
    U = _f64(U); M, K = U.shape
    n = M if forceDim is None else forceDim
    vn = _f64(vn); out = np.empty(n)
    lib().orc_project_vertex(n, _dp(U), K, int(vid), _dp(vn), _dp(out))
    return out


def project_face(U, vids, coords, vn, forceDim=None):
    """tools/real_time_modal_sound.cpp:236-251"""
    U = _f64(U); M, K = U.shape
    n = M if forceDim is None else forceDim
    vids = np.ascontiguousarray(vids, dtype=np.int32)
    coords = _f64(coords); vn = _f64(vn); out = np.empty(n)
    lib().orc_project_face(n, _dp(U), K, _ip(vids), _dp(coords), _dp(vn), _dp(out))
    return out


def project_dense(U, F):
    """Y[b][m] = sum_k U[m][k] F[b][k]  (U [M][K], F [B][K]) -- dense form of the same projection (cfg3)."""
    U = _f64(U); F = _f64(F); M, K = U.shape; B = F.shape[0]
    Y = np.empty((B, M))
    lib().orc_project_dense(M, K, B, _dp(U), _dp(F), _dp(Y))
    return Y


def pack_ffat(maps):
    """maps: list of dicts as produced by oracle.fatcube.load / synth; -> (geom, igeom, Psi, D)."""
    n = len(maps); D = len(maps[0]["psi"])
    geom = np.empty((n, 32)); igeom = np.empty((n, 18), dtype=np.int32); Psi = np.empty((n, D))
    for i, m in enumerate(maps):
        assert len(m["psi"]) == D
        geom[i, 0] = m["cellsize"]
        geom[i, 1:19] = np.asarray(m["lowcorners"], dtype=np.float64).reshape(18)
        geom[i, 19:22] = m["center1"]; geom[i, 22:25] = m["bboxlow"]; geom[i, 25:28] = m["bboxtop"]
        geom[i, 28:31] = m["center"]; geom[i, 31] = m["k"]
        igeom[i, :12] = np.asarray(m["n_elements"], dtype=np.int32).reshape(12)
        igeom[i, 12:] = m["strides"]
        Psi[i] = m["psi"]
    return geom, igeom, Psi, D


def ffat_eval(maps, pos):
    """modal_solver.h:286-315 + ffat_solver.h:1180-1206: returns [L][N] (= col-major N x L)."""
    geom, igeom, Psi, D = pack_ffat(maps)
    pos = _f64(pos).reshape(-1, 3); L = len(pos)
    out = np.empty((L, len(maps)))
    lib().orc_ffat_eval(len(maps), _dp(geom), _ip(igeom), _dp(Psi), D, _dp(pos), L, _dp(out))
    return out


def cv_saturate_u8(x):
    """cv::saturate_cast<uchar>(double) = cvRound (cvtsd2si: round half to even; NaN / outside int -> INT_MIN) clamped to
    [0, 255] -- what cv::Mat::convertTo(CV_8U) applies (ffat_solver.h:1147).  [pinned: tests/golden/ffat_compress.npz
    cast_in/cast_out and q8_pre, produced by OpenCV itself]"""
    x = np.asarray(x, dtype=np.float64)
    with np.errstate(invalid="ignore"):
        bad = ~(np.abs(x) < 2147483648.0)
        r = np.clip(np.rint(np.where(bad, 0.0, x)), 0, 255)
    return np.where(bad, 0, r).astype(np.uint8)


def ffat_quantise(m):
    """FFAT_Map<T,3>::Compress up to the image (ffat_solver.h:1128-1147; ConvertToImages :1106-1122)
    -> (q8, maxAmp[6], maxAmp_global)."""
    psi = _f64(m["psi"]); q = np.zeros(len(psi), np.uint8); amp = np.empty(6); off = 0; gmax = -1.0
    for fc, (nx, ny) in enumerate(np.asarray(m["n_elements"]).reshape(6, 2)):
        A = psi[off:off + nx * ny]
        mx = A.max(); amp[fc] = mx; gmax = max(gmax, mx)                       # :1134-1139
        with np.errstate(divide="ignore", invalid="ignore"):
            q[off:off + nx * ny] = cv_saturate_u8(A * (np.float64(255) / mx))  # :1144-1147
        off += nx * ny
    return q, amp, gmax


def ffat_dequantise(m, q8, amp):
    """Compress after the image round trip (ffat_solver.h:1159-1171): _compressed_Psi = data_s (CV_64F) * (maxAmp/255.)."""
    q8 = np.asarray(q8, dtype=np.uint8); c = np.zeros(len(q8)); off = 0
    for fc, (nx, ny) in enumerate(np.asarray(m["n_elements"]).reshape(6, 2)):
        c[off:off + nx * ny] = q8[off:off + nx * ny].astype(np.float64) * (amp[fc] / 255.)
        off += nx * ny
    return c


def ffat_intersect(m, p):
    """ffat_solver.h:676-712"""
    geom, igeom, _, _ = pack_ffat([m])
    p = _f64(p); surf = np.empty(3); ind = np.empty(3, dtype=np.int32)
    lib().orc_ffat_intersect(_dp(geom), _ip(igeom), _dp(p), _dp(surf), _ip(ind))
    return surf, ind


def ffat_interpolate(m, surf, nn):
    """ffat_solver.h:736-803"""
    geom, igeom, _, _ = pack_ffat([m])
    surf = _f64(surf); nn = np.ascontiguousarray(nn, dtype=np.int32)
    idx = np.empty((4, 3), dtype=np.int32); co = np.empty(4)
    lib().orc_ffat_interpolate(_dp(geom), _ip(igeom), _dp(surf), _ip(nn), _ip(idx), _dp(co))
    return idx, co


def ffat_fit_geometry(cell_size, V, n_elements, bbox_init=0.0):
    """FFAT_Map<T,3>::FFAT_Map(modeId, cellSize, V, N_elements) (ffat_solver.h:944-989, shells :399-428).
    V [rows][3]; n_elements [S][6][2].  Returns dict(geom [S][32], igeom [S][18], strides [S], n_total, n_dir)."""
    V = _f64(V).reshape(-1, 3); ne = np.ascontiguousarray(n_elements, dtype=np.int32).reshape(-1, 12); S = len(ne)
    geom = np.empty((S, 32)); igeom = np.empty((S, 18), dtype=np.int32); st = np.empty(S, dtype=np.int32)
    cnt = np.empty(2, dtype=np.int32)
    assert V.shape[0] >= 4 * int(sum(int(r[2 * f]) * int(r[2 * f + 1]) for r in ne for f in range(6)))
    lib().orc_ffat_fit_geometry(float(cell_size), _dp(V), _ip(ne), S, float(bbox_init), _dp(geom), _ip(igeom), _ip(st), _ip(cnt))
    return dict(geom=geom, igeom=igeom, strides=st, n_total=int(cnt[0]), n_dir=int(cnt[1]), cell_size=float(cell_size))


def ffat_fit_solve(fit, k, pressure, power_scaling=False, want_R=False):
    """FFAT_Map<T,3>::Solve (ffat_solver.h:1007-1069) for a batch of modes sharing `fit`:
    k [n_maps]; pressure complex [n_maps][2*n_total].  Returns (psi [n_maps][n_dir], scale [n_maps]) (+ R, |P|)."""
    k = _f64(np.atleast_1d(k)); n_maps = len(k)
    P = np.ascontiguousarray(pressure, dtype=np.complex128).reshape(n_maps, 2 * fit["n_total"])
    S = len(fit["strides"])
    psi = np.empty((n_maps, fit["n_dir"])); scale = np.empty(n_maps)
    R = np.empty((fit["n_dir"], S)); Pabs = np.empty((n_maps, fit["n_dir"], S))
    lib().orc_ffat_fit_solve(S, _dp(fit["geom"]), _ip(fit["igeom"]), _ip(fit["strides"]), n_maps, _dp(k),
                             P.view(np.float64).ctypes.data_as(c_dp), int(bool(power_scaling)), _dp(psi), _dp(scale),
                             _dp(R), _dp(Pabs))
    return (psi, scale, R, Pabs) if want_R else (psi, scale)


def ref_ffat_fit(mode_id, cell_size, V, n_elements, k, pressure, power_scaling=False, save_to=None):
    """The reference's own FFAT_Map<double,3> ctor + Solve (+ FFAT_Map_Serialize::Save when save_to is given),
    compiled in place.  Returns (psi [n_dir], centre[3])."""
    V = _f64(V).reshape(-1, 3); ne = np.ascontiguousarray(n_elements, dtype=np.int32).reshape(-1, 12); S = len(ne)
    P = np.ascontiguousarray(pressure, dtype=np.complex128).ravel()
    n_dir = int(sum(int(ne[2][2 * f]) * int(ne[2][2 * f + 1]) for f in range(6)))
    psi = np.empty(n_dir); centre = np.empty(3)
    ref().ref_ffat_fit(int(mode_id), float(cell_size), _dp(V), V.shape[0], _ip(ne), S, float(k),
                       P.view(np.float64).ctypes.data_as(c_dp), int(bool(power_scaling)), _dp(psi), _dp(centre),
                       None if save_to is None else os.fsencode(save_to))
    return psi, centre


def ref_cubemap_mesh(bbox_low_r, bbox_top_r, cell_size, grid_low, dim):
    """The reference's own FFAT_Map<double,1>::CubemapMesh (ffat_solver.h:334-397), compiled in place.
    Returns (V [4*n_quads][3], n_elements [6][2], data_indices [2*n_quads])."""
    lo = np.ascontiguousarray(bbox_low_r, dtype=np.int32); hi = np.ascontiguousarray(bbox_top_r, dtype=np.int32)
    dm = np.ascontiguousarray(dim, dtype=np.int32); gl = _f64(grid_low)
    n = hi - lo + 1
    quads = 2 * int(n[0] * n[1] + n[1] * n[2] + n[2] * n[0])
    V = np.empty((4 * quads, 3)); ne = np.empty((6, 2), dtype=np.int32); idx = np.empty(2 * quads, dtype=np.int32)
    rows = ref().ref_cubemap_mesh(_ip(lo), _ip(hi), float(cell_size), _dp(gl), _ip(dm), _dp(V), len(V), _ip(ne), _ip(idx))
    assert rows == len(V), rows
    return V, ne, idx


def num_modes_audible(omega2, density, freq, cache=None):
    """ModeData.h:120-148; cache = [N, freqThres, density] mutated in place."""
    omega2 = _f64(omega2)
    c = np.array([-1, 22100., -1.]) if cache is None else cache
    return lib().orc_num_modes_audible(_dp(omega2), len(omega2), density, freq, _dp(c))


def material_read(path):
    """ModalMaterial.h:35-55 -> dict or None"""
    out = np.empty(5)
    if not lib().orc_material_read(path.encode(), _dp(out)):
        return None
    return dict(density=out[0], youngsModulus=out[1], poissonRatio=out[2], alpha=out[3], beta=out[4])


def modes_read(path):
    """ModeData.h:61-83 -> (omega2[nModes], U[nModes][nDOF])"""
    nd = C.c_int(); nm = C.c_int()
    if not lib().orc_modes_read_header(path.encode(), C.byref(nd), C.byref(nm)):
        raise IOError(path)
    w2 = np.empty(nm.value); U = np.empty((nm.value, nd.value))
    lib().orc_modes_read(path.encode(), _dp(w2), _dp(U))
    return w2, U


def modes_write(path, omega2, U):
    """ModeData.h:87-107"""
    w2 = _f64(omega2); U = _f64(U)
    lib().orc_modes_write(path.encode(), U.shape[1], U.shape[0], _dp(w2), _dp(U))


def batch_render(h, a, b, space, trans, imp_buf, BUF, n_buf, mix=None):
    """Reference loop (modal_solver.h:181-276) over independent objects, mixed down."""
    a = _f64(a); b = _f64(b); space = _f64(space); trans = _f64(trans)
    n_obj, N = a.shape
    imp = np.ascontiguousarray(imp_buf, dtype=np.int32)
    if mix is None:
        mix = np.zeros(n_buf * BUF)
    lib().orc_batch_render(n_obj, N, BUF, n_buf, h, _dp(a), _dp(b), _dp(space), _dp(trans),
                           _ip(imp), _dp(mix))
    return mix


def ref_batch_render(h, a, b, space, trans, imp_buf, n_buf):
    """Same job as batch_render (BUF = 256) run by the reference's own ModalSolver / ModalIntegrator."""
    a = _f64(a); b = _f64(b); space = _f64(space); trans = _f64(trans)
    n_obj, N = a.shape
    imp = np.ascontiguousarray(imp_buf, dtype=np.int32)
    mix = np.zeros(n_buf * 256)
    ref().ref_batch_render(n_obj, N, n_buf, h, _dp(a), _dp(b), _dp(space), _dp(trans), _ip(imp), _dp(mix))
    return mix


def read_obj(path):
    """Plain `v x y z` / `f i j k` OBJ reader (assets/ball.obj has no vn/vt; 1-indexed)."""
    V = []; F = []
    with open(path) as f:
        for line in f:
            t = line.split()
            if not t:
                continue
            if t[0] == "v":
                V.append([float(x) for x in t[1:4]])
            elif t[0] == "f":
                F.append([int(x.split("/")[0]) - 1 for x in t[1:4]])
    return np.array(V, dtype=np.float64), np.array(F, dtype=np.int32)


def per_vertex_normals(V, F):
    """external/libigl/include/igl/per_vertex_normals.cpp:61-68,75-82,104 (area weighting):
    N[v] += doublearea(f) * unit_normal(f) for each incident face, rows normalised."""
    e1 = V[F[:, 1]] - V[F[:, 0]]; e2 = V[F[:, 2]] - V[F[:, 0]]
    cr = np.cross(e1, e2)               # = doublearea * unit normal
    N = np.zeros_like(V)
    for j in range(3):
        np.add.at(N, F[:, j], cr)
    return N / np.linalg.norm(N, axis=1, keepdims=True)
