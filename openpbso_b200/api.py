"""Object view of the C ABI, named after the reference classes each wraps."""
import ctypes as C
import numpy as np
from . import _capi as capi
from ._capi import check, f64, dp, ip, lib


def device_count():
    n = C.c_int()
    check(lib().pbso_device_count(C.byref(n)))
    return n.value


def set_device(dev):
    check(lib().pbso_set_device(dev))


def device_info():
    sm = C.c_int(); ma = C.c_int(); mi = C.c_int(); gib = C.c_double()
    check(lib().pbso_device_info(C.byref(sm), C.byref(ma), C.byref(mi), C.byref(gib)))
    return dict(sm_count=sm.value, cc=(ma.value, mi.value), hbm_gib=gib.value)


def measure_fma_peak(kind):
    t = C.c_double(); mhz = C.c_double()
    check(lib().pbso_measure_fma_peak(kind, C.byref(t), C.byref(mhz)))
    return t.value, mhz.value


def measure_tc_peak(kind, cta_group=1, n=128, stress=0):
    """(TFLOP/s, cycles per MMA, stress wavefronts per cycle per SM) of a bare tcgen05.mma loop."""
    t = C.c_double(); cyc = C.c_double(); wf = C.c_double()
    check(lib().pbso_measure_tc_peak(kind, cta_group, n, stress, C.byref(t), C.byref(cyc), C.byref(wf)))
    return t.value, cyc.value, wf.value


def measure_tc_peak_sustained(kind, cta_group=1, n=128, min_ms=200.0):
    t = C.c_double()
    check(lib().pbso_measure_tc_peak_sustained(kind, cta_group, n, min_ms, C.byref(t)))
    return t.value


def tc_gain():
    g = C.c_double()
    check(lib().pbso_tc_gain(C.byref(g)))
    return g.value


def tc_selftest(kind):
    e = C.c_double()
    check(lib().pbso_tc_selftest(kind, C.byref(e)))
    return e.value


def measure_copy_bw(nbytes):
    g = C.c_double()
    check(lib().pbso_measure_copy_bw(nbytes, C.byref(g)))
    return g.value


def flush_l2(nbytes=256 << 20):
    check(lib().pbso_flush_l2(nbytes))


class Comm:
    """The ranks of a multi-GPU render: pbso_comm_* (NCCL from the C ABI).  `unique_id()` on rank 0, hand the 128
    bytes to every rank, then Comm(nranks, rank, id) on each (collective)."""

    @staticmethod
    def unique_id():
        buf = (C.c_ubyte * 128)()
        check(lib().pbso_comm_unique_id(buf))
        return bytes(buf)

    def __init__(self, nranks, rank, uid=None):
        self._h = C.c_void_p()
        buf = (C.c_ubyte * 128).from_buffer_copy(uid) if uid is not None else None
        check(lib().pbso_comm_init(nranks, rank, buf, C.byref(self._h)))
        self.nranks, self.rank = nranks, rank

    def shard(self, n_units):
        lo = C.c_longlong(); hi = C.c_longlong()
        check(lib().pbso_comm_shard(self._h, n_units, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def nccl_version(self):
        n = C.c_int(); r = C.c_int(); v = C.c_int()
        check(lib().pbso_comm_info(self._h, C.byref(n), C.byref(r), C.byref(v)))
        return v.value

    def reduce_audio(self, d_ptr, n, root=0, stream=None):
        """Sum of the ranks' double[n] at device pointer d_ptr onto `root` (-1: every rank), enqueued on `stream`."""
        check(lib().pbso_comm_reduce_audio(self._h, C.c_void_p(d_ptr), n, root, C.c_void_p(stream) if stream else None))

    def close(self):
        if self._h:
            lib().pbso_comm_destroy(self._h); self._h = C.c_void_p()

    def __del__(self):
        try: self.close()
        except Exception: pass


class ModalIntegrator:
    """ModalIntegrator<double> (modal_integrator.h:19-45) + the ModalSolver::step hot loop."""

    def __init__(self, N, h, a, b):
        a = f64(a); b = f64(b)
        self._h = C.c_void_p()
        check(lib().pbso_integrator_create(N, h, dp(a), dp(b), C.byref(self._h)))
        self.N = N

    @classmethod
    def Build(cls, density, omegaSquared, alpha, beta, h, N=-1):
        """modal_integrator.h:39-42"""
        w2 = f64(omegaSquared)
        self = cls.__new__(cls)
        self._h = C.c_void_p()
        check(lib().pbso_integrator_build(density, dp(w2), len(w2), alpha, beta, h, N, C.byref(self._h)))
        n = C.c_int(); check(lib().pbso_integrator_size(self._h, C.byref(n)))
        self.N = n.value
        return self

    def coeffs(self):
        c1 = np.empty(self.N); c2 = np.empty(self.N); c3 = np.empty(self.N)
        check(lib().pbso_integrator_get_coeffs(self._h, dp(c1), dp(c2), dp(c3)))
        return c1, c2, c3

    def Step(self, Q=None):
        """modal_integrator.h:43-44"""
        out = np.empty(self.N)
        Qc = None if Q is None else f64(Q)
        if Qc is not None and len(Qc) != self.N:
            raise capi.PbsoError(capi.ERR_INVALID, "input force incorrect dimension")   # :108
        check(lib().pbso_integrator_step(self._h, dp(Qc), dp(out)))
        return out

    def get_state(self):
        q1 = np.empty(self.N); q2 = np.empty(self.N)
        check(lib().pbso_integrator_get_state(self._h, dp(q1), dp(q2)))
        return q1, q2

    def set_state(self, q1, q2):
        q1 = f64(q1); q2 = f64(q2)
        check(lib().pbso_integrator_set_state(self._h, dp(q1), dp(q2)))

    def set_transfer(self, transfer, L=1):
        """transfer: [L][n_transfer] (= column-major n_transfer x L)."""
        t = f64(transfer).reshape(L, -1) if L > 0 else np.zeros((0, 0))
        self.L = L
        check(lib().pbso_integrator_set_transfer(self._h, dp(t), t.shape[1] if L > 0 else 0, L))

    def set_transfer_ffat(self, maps, pos, n_transfer=None):
        """computeTransfer for L listener positions evaluated on the device into the resident table (cfg4)."""
        pos = f64(pos).reshape(-1, 3)
        self.L = len(pos)
        check(lib().pbso_integrator_set_transfer_ffat(self._h, maps._h, self.N if n_transfer is None else n_transfer, dp(pos), self.L))

    def render_buffer(self, space, time, want_qnorm=True):
        """ModalSolver::step hot loop (modal_solver.h:261-272): returns (y[L][T], qnorm[N])."""
        space = f64(space); time = f64(time); T = len(time)
        L = getattr(self, "L", 1)
        y = np.empty((max(L, 1), T)); qn = np.empty(self.N) if want_qnorm else None
        check(lib().pbso_render_buffer(self._h, dp(space), dp(time), T, dp(y), dp(qn)))
        return (y if L > 0 else None), qn

    def close(self):
        if getattr(self, "_h", None):
            try:
                lib().pbso_integrator_destroy(self._h)
            except Exception:          # interpreter shutdown
                pass
            self._h = None

    __del__ = close


class FFATMaps:
    """std::map<int, FFAT_Map<double,3>> as produced by FFAT_Map_Serialize::LoadAll."""

    def __init__(self, handle):
        self._h = handle

    @classmethod
    def LoadAll(cls, dirname):
        h = C.c_void_p()
        rc = lib().pbso_ffat_load_dir(dirname.encode(), C.byref(h))
        self = cls(h)
        if rc != capi.OK:
            self.close()
            check(rc)
        return self

    @classmethod
    def Load(cls, filename):
        h = C.c_void_p()
        check(lib().pbso_ffat_load_file(filename.encode(), C.byref(h)))
        return cls(h)

    @classmethod
    def from_dicts(cls, maps):
        n = len(maps); D = len(maps[0]["psi"])
        geom = np.empty((n, 32)); igeom = np.empty((n, 18), dtype=np.int32); psi = np.empty((n, D))
        ids = np.empty(n, dtype=np.int32); comp = np.zeros(n, dtype=np.uint8)
        for i, m in enumerate(maps):
            geom[i, 0] = m["cellsize"]; geom[i, 1:19] = np.asarray(m["lowcorners"], dtype=np.float64).reshape(18)
            geom[i, 19:22] = m["center1"]; geom[i, 22:25] = m["bboxlow"]; geom[i, 25:28] = m["bboxtop"]
            geom[i, 28:31] = m["center"]; geom[i, 31] = m["k"]
            igeom[i, :12] = np.asarray(m["n_elements"], dtype=np.int32).reshape(12); igeom[i, 12:] = m["strides"]
            psi[i] = m["psi"]; ids[i] = m["modeid"]; comp[i] = bool(m.get("is_compressed", False))
        h = C.c_void_p()
        check(lib().pbso_ffat_create(n, ip(ids), dp(geom), ip(igeom), dp(psi), D,
                                     comp.ctypes.data_as(C.POINTER(C.c_ubyte)), C.byref(h)))
        return cls(h)

    def size(self):
        n = C.c_int(); check(lib().pbso_ffat_num_maps(self._h, C.byref(n))); return n.value

    def mode_ids(self):
        ids = np.empty(self.size(), dtype=np.int32)
        if len(ids):
            check(lib().pbso_ffat_mode_ids(self._h, ip(ids)))
        return ids

    def get_map(self, mode_id):
        geom = np.empty(32); igeom = np.empty(18, dtype=np.int32)
        n = C.c_int(); cols = C.c_int(); comp = C.c_int()
        check(lib().pbso_ffat_get_map(self._h, mode_id, dp(geom), ip(igeom), C.byref(n), C.byref(cols), C.byref(comp), None))
        psi = np.empty((cols.value, n.value))
        check(lib().pbso_ffat_get_map(self._h, mode_id, None, None, None, None, None, dp(psi)))
        return dict(cellsize=geom[0], lowcorners=geom[1:19].reshape(6, 3).copy(), center1=geom[19:22].copy(),
                    bboxlow=geom[22:25].copy(), bboxtop=geom[25:28].copy(), center=geom[28:31].copy(), k=geom[31],
                    n_elements=igeom[:12].reshape(6, 2).copy(), strides=igeom[12:].copy(),
                    psi=psi[0].copy(), psi_cols=[c.copy() for c in psi], is_compressed=bool(comp.value), modeid=mode_id)

    def Save(self, mode_id, filename):
        check(lib().pbso_ffat_save_file(self._h, mode_id, filename.encode()))

    def SaveLegacy(self, mode_id, filename):
        """FFAT_Map<T,3>::Save (ffat_solver.h:1066-1068): the igl::serialize form; LoadAll / Load read either form."""
        check(lib().pbso_ffat_save_legacy_file(self._h, mode_id, filename.encode()))

    def quantise(self, mode_id):
        """First half of FFAT_Map<T,3>::Compress (ffat_solver.h:1125-1147): per face, Psi * (255/maxAmp) saturated to 8 bits
        the way cv::Mat::convertTo(CV_8U) does -> (q8 [psi_len] uint8, maxAmp[6], maxAmp_global)."""
        n = C.c_int()
        check(lib().pbso_ffat_get_map(self._h, mode_id, None, None, C.byref(n), None, None, None))
        q = np.empty(n.value, dtype=np.uint8); amp = np.empty(6); g = C.c_double()
        check(lib().pbso_ffat_quantise(self._h, mode_id, q.ctypes.data_as(C.POINTER(C.c_ubyte)), dp(amp), C.byref(g)))
        return q, amp, g.value

    def set_compressed_u8(self, mode_id, q8, max_amp):
        """Second half of Compress (ffat_solver.h:1159-1173): _compressed_Psi = q8 * (maxAmp/255.) per face, _is_compressed =
        true.  q8 is what came back from the image round trip (or quantise()'s own bytes when there is none)."""
        q8 = np.ascontiguousarray(q8, dtype=np.uint8); amp = f64(max_amp)
        check(lib().pbso_ffat_set_compressed_u8(self._h, mode_id, q8.ctypes.data_as(C.POINTER(C.c_ubyte)), len(q8), dp(amp)))

    def Compress(self, mode_id=None):
        """FFAT_Map<T,3>::Compress for one map (returns maxAmp_global) or, mode_id None, every map (returns the array): the
        8-bit quantisation without the JPEG file round trip (quantise() / set_compressed_u8() are the two halves for a caller
        who wants an image codec in between)."""
        n = 1 if mode_id is not None else self.size()
        g = np.empty(n)
        check(lib().pbso_ffat_compress(self._h, -1 if mode_id is None else mode_id, dp(g)))
        return g[0] if mode_id is not None else g

    def get_compressed(self, mode_id):
        """-> (q8, maxAmp[6], _compressed_Psi as the doubles the reference would hold)."""
        n = C.c_int()
        check(lib().pbso_ffat_get_map(self._h, mode_id, None, None, C.byref(n), None, None, None))
        q = np.empty(n.value, dtype=np.uint8); sc = np.empty(6); c = np.empty(n.value)
        check(lib().pbso_ffat_get_compressed(self._h, mode_id, q.ctypes.data_as(C.POINTER(C.c_ubyte)), dp(sc), dp(c)))
        return q, sc, c

    def computeTransfer(self, pos, n_modes=None, use_compressed=False):
        """modal_solver.h:286-315 for L listeners: returns [L][n_modes]."""
        pos = f64(pos).reshape(-1, 3); L = len(pos)
        n = self.size() if n_modes is None else n_modes
        out = np.empty((L, n))
        check(lib().pbso_ffat_eval(self._h, n, dp(pos), L, int(use_compressed), dp(out)))
        return out

    def close(self):
        if getattr(self, "_h", None):
            try:
                lib().pbso_ffat_destroy(self._h)
            except Exception:          # interpreter shutdown
                pass
            self._h = None

    __del__ = close


class FFATFitter:
    """FFAT_Map<double,3>(modeId, cellSize, V, N_elements) + Solve(k, dirichletPressure, powerScaling)
    (reference ffat_solver.h:944-1069) for all modes of an object at once (kernel K6)."""

    def __init__(self, cell_size, V, n_elements):
        V = f64(V).reshape(-1, 3)
        ne = np.ascontiguousarray(n_elements, dtype=np.int32).reshape(-1, 12)
        h = C.c_void_p()
        check(lib().pbso_ffat_fitter_create(float(cell_size), dp(V), V.shape[0], ip(ne), len(ne), C.byref(h)))
        self._h = h
        S = C.c_int(); nt = C.c_int(); nd = C.c_int()
        check(lib().pbso_ffat_fitter_info(self._h, C.byref(S), C.byref(nt), C.byref(nd), None))
        self.n_shells, self.n_elements_total, self.n_directions = S.value, nt.value, nd.value
        self.strides = np.empty(self.n_shells, dtype=np.int32)
        check(lib().pbso_ffat_fitter_info(self._h, None, None, None, ip(self.strides)))

    def shell(self, s):
        geom = np.empty(32); igeom = np.empty(18, dtype=np.int32)
        check(lib().pbso_ffat_fitter_shell(self._h, int(s), dp(geom), ip(igeom)))
        return geom, igeom

    FIT_POWER_SCALING, FIT_DEFER_SCALE, FIT_PACKED = 1, 2, 4

    def Solve(self, k, pressure, powerScaling=False, packed=False):
        """k [n_maps]; pressure complex [n_maps][2*N_elements_total] (reference layout) or, packed=True, one complex per quad
        [n_maps][N_elements_total].  Returns (Psi [n_maps][N_directions], scale [n_maps])."""
        k = f64(np.atleast_1d(k)); n = len(k)
        P = np.ascontiguousarray(pressure, dtype=np.complex128).reshape(n, (1 if packed else 2) * self.n_elements_total)
        psi = np.empty((n, self.n_directions)); scale = np.empty(n)
        flags = (self.FIT_POWER_SCALING if powerScaling else 0) | (self.FIT_PACKED if packed else 0)
        check(lib().pbso_ffat_fitter_solve(self._h, n, dp(k), P.view(np.float64).ctypes.data_as(capi.c_dp), flags, dp(psi), dp(scale)))
        return psi, scale

    def solve_device(self, n_maps, d_k_ptr, d_pressure_ptr, d_psi_ptr, powerScaling=False, d_scale_ptr=0, stream_ptr=0, packed=False,
                     defer_scale=False):
        flags = (self.FIT_POWER_SCALING if powerScaling else 0) | (self.FIT_PACKED if packed else 0) | (self.FIT_DEFER_SCALE if defer_scale else 0)
        check(lib().pbso_ffat_fitter_solve_device(self._h, int(n_maps), C.c_void_p(d_k_ptr), C.c_void_p(d_pressure_ptr), flags,
                                                  C.c_void_p(d_psi_ptr), C.c_void_p(d_scale_ptr) if d_scale_ptr else None,
                                                  C.c_void_p(stream_ptr) if stream_ptr else None))

    def last_kernel_ms(self):
        ms = C.c_float(); check(lib().pbso_ffat_fitter_last_kernel_ms(self._h, C.byref(ms))); return ms.value

    def to_maps(self, k, psi, mode_ids=None):
        """The run-time map set (what FFAT_Map_Serialize::Save keeps: shell 2 + Psi + k) of solved modes."""
        k = f64(np.atleast_1d(k)); psi = f64(psi).reshape(len(k), self.n_directions)
        g2, ig2 = self.shell(2)
        geom = np.tile(g2, (len(k), 1)); geom[:, 31] = k
        igeom = np.tile(ig2, (len(k), 1)).astype(np.int32)
        ids = np.arange(len(k), dtype=np.int32) if mode_ids is None else np.ascontiguousarray(mode_ids, dtype=np.int32)
        h = C.c_void_p()
        check(lib().pbso_ffat_create(len(k), ip(ids), dp(geom), ip(igeom), dp(psi), self.n_directions, None, C.byref(h)))
        return FFATMaps(h)

    def close(self):
        if getattr(self, "_h", None):
            try:
                lib().pbso_ffat_fitter_destroy(self._h)
            except Exception:          # interpreter shutdown
                pass
            self._h = None

    __del__ = close


class ModeShapes:
    """ModeData<double>::_modes resident on the device + GetModalForceVertex/Face."""

    def __init__(self, U=None, _handle=None, _shape=None):
        if _handle is not None:
            self._h = _handle; self.M, self.K = _shape
            return
        U = f64(U); self.M, self.K = U.shape
        self._h = C.c_void_p()
        check(lib().pbso_modes_upload(dp(U), self.M, self.K, C.byref(self._h)))

    @classmethod
    def read(cls, filename):
        """ModeData::read (ModeData.h:61-83)"""
        h = C.c_void_p(); M = C.c_int(); K = C.c_int()
        check(lib().pbso_modes_read_file(filename.encode(), C.byref(h), C.byref(M), C.byref(K)))
        return cls(_handle=h, _shape=(M.value, K.value))

    def omegaSquared(self):
        w2 = np.empty(self.M); check(lib().pbso_modes_omega_squared(self._h, dp(w2))); return w2

    def GetModalForceVertex(self, forceDim, vid, vn):
        vn = f64(vn); out = np.empty(forceDim)
        check(lib().pbso_modes_project_vertex(self._h, forceDim, int(vid), dp(vn), dp(out)))
        return out

    def GetModalForceFace(self, forceDim, vids, coords, vn):
        vids = np.ascontiguousarray(vids, dtype=np.int32); coords = f64(coords); vn = f64(vn)
        out = np.empty(forceDim)
        check(lib().pbso_modes_project_face(self._h, forceDim, ip(vids), dp(coords), dp(vn), dp(out)))
        return out

    def project_vertices(self, forceDim, vids, vn):
        vids = np.ascontiguousarray(vids, dtype=np.int32); vn = f64(vn).reshape(-1, 3); B = len(vids)
        out = np.empty((B, forceDim))
        check(lib().pbso_modes_project_vertices(self._h, forceDim, B, ip(vids), dp(vn), dp(out)))
        return out

    def project_dense(self, F, forceDim=None, precision=capi.PREC_F64):
        """F: [B][K] dense load vectors (or one K-vector) -> Y [B][forceDim]."""
        F = f64(F)
        if F.ndim == 1:
            F = F.reshape(1, -1)
        assert F.shape[1] == self.K
        n = self.M if forceDim is None else forceDim
        Y = np.empty((F.shape[0], n))
        check(lib().pbso_modes_project_dense(self._h, n, dp(F), F.shape[0], dp(Y), precision))
        return Y

    def last_kernel_ms(self):
        ms = C.c_float(); check(lib().pbso_modes_last_kernel_ms(self._h, C.byref(ms))); return ms.value

    def storm_buffer(self, integrator, vids, vn, T, precision=capi.PREC_TF32X3, want_qnorm=False):
        """cfg3 contact storm, one buffer on the device: B vertex impulses -> projection -> summed load -> K1.
        Returns (y[L][T], qnorm or None)."""
        vids = np.ascontiguousarray(vids, dtype=np.int32); vn = f64(vn).reshape(len(vids), 3)
        L = max(getattr(integrator, "L", 1), 1)
        y = np.empty((L, T)); qn = np.empty(integrator.N) if want_qnorm else None
        check(lib().pbso_modes_storm_buffer(self._h, integrator._h, integrator.N, len(vids), ip(vids), dp(vn), T, dp(y), dp(qn), precision))
        return y, qn

    def project_dense_device(self, d_F_ptr, B, d_Y_ptr, forceDim=None, stream_ptr=0):
        n = self.M if forceDim is None else forceDim
        check(lib().pbso_modes_project_dense_device(self._h, n, C.c_void_p(d_F_ptr), B, C.c_void_p(d_Y_ptr),
                                                    C.c_void_p(stream_ptr)))

    def close(self):
        if getattr(self, "_h", None):
            try:
                lib().pbso_modes_destroy(self._h)
            except Exception:          # interpreter shutdown
                pass
            self._h = None

    __del__ = close


class BatchRenderer:
    """Many independent ModalSolver step loops rendered offline and mixed."""

    def __init__(self, h, a, b):
        a = f64(a); b = f64(b)
        self.n_obj, self.n_modes = a.shape
        self._h = C.c_void_p()
        check(lib().pbso_batch_create(self.n_obj, self.n_modes, h, dp(a), dp(b), C.byref(self._h)))

    def set_transfer(self, trans, wait=True):
        """wait=False: the copy is only enqueued (pbso_batch_set_transfer_async); `trans` must then be a contiguous float64
        array that stays alive and unchanged until the stream has passed it (ideally pinned)."""
        t = f64(trans); assert t.shape == (self.n_obj, self.n_modes)
        if wait:
            check(lib().pbso_batch_set_transfer(self._h, dp(t)))
        else:
            assert t is trans or np.shares_memory(t, trans), "async copies need the caller's own contiguous float64 buffer"
            check(lib().pbso_batch_set_transfer_async(self._h, dp(t)))

    def set_impulses(self, obj, buf, space, wait=True):
        obj = np.ascontiguousarray(obj, dtype=np.int32); buf = np.ascontiguousarray(buf, dtype=np.int32)
        sp = f64(space).reshape(len(obj), self.n_modes)
        if wait:
            check(lib().pbso_batch_set_impulses(self._h, len(obj), ip(obj), ip(buf), dp(sp)))
        else:
            assert np.shares_memory(sp, space), "async copies need the caller's own contiguous float64 buffer"
            check(lib().pbso_batch_set_impulses_async(self._h, len(obj), ip(obj), ip(buf), dp(sp)))

    def render_mix(self, buf_size, n_buffers, precision=capi.PREC_F32_TILED, n_chunks=0):
        mix = np.empty(buf_size * n_buffers)
        check(lib().pbso_batch_render_mix(self._h, buf_size, n_buffers, precision, n_chunks, dp(mix)))
        return mix

    def render_mix_device(self, buf_size, n_buffers, d_mix_ptr, precision=capi.PREC_F32_TILED, n_chunks=0):
        check(lib().pbso_batch_render_mix_device(self._h, buf_size, n_buffers, precision, n_chunks, C.c_void_p(d_mix_ptr)))

    def render_stems(self, buf_size, n_buffers, precision=capi.PREC_F32_TILED):
        st = np.empty((self.n_obj, buf_size * n_buffers), dtype=np.float32)
        check(lib().pbso_batch_render_stems(self._h, buf_size, n_buffers, precision, st.ctypes.data_as(capi.c_fp)))
        return st

    def set_state(self, q_km1=None, q_km2=None):
        """Stateful range renders: the next renders start from (q[k-1], q[k-2]) per (object, mode), the pair
        ModalIntegrator keeps (modal_integrator.h:106-113); None = the zero state of a fresh solver."""
        if q_km1 is None:
            check(lib().pbso_batch_set_state(self._h, None, None)); return
        q1 = f64(q_km1).reshape(self.n_obj, self.n_modes); q2 = f64(q_km2).reshape(self.n_obj, self.n_modes)
        check(lib().pbso_batch_set_state(self._h, dp(q1), dp(q2)))

    def end_state(self, buf_size, n_buffers):
        """(q[k-1], q[k-2]) after n_buffers x buf_size samples of the current state + impulse script (FP64, closed form)."""
        q1 = np.empty((self.n_obj, self.n_modes)); q2 = np.empty((self.n_obj, self.n_modes))
        check(lib().pbso_batch_get_end_state(self._h, buf_size, n_buffers, dp(q1), dp(q2)))
        return q1, q2

    def sync(self):
        check(lib().pbso_batch_sync(self._h))

    def set_stream(self, cuda_stream_ptr):
        check(lib().pbso_batch_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def last_kernel_ms(self):
        ms = C.c_float(); n = C.c_int()
        check(lib().pbso_batch_last_kernel_ms(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def close(self):
        if getattr(self, "_h", None):
            try:
                lib().pbso_batch_destroy(self._h)
            except Exception:          # interpreter shutdown
                pass
            self._h = None

    __del__ = close
