"""ctypes binding of include/pbso_b200.h (the C ABI of libpbso_b200.so).

Python here is plumbing for tests and bench.py; the product is the shared library.  There is no
fallback: a missing library or a missing CUDA device raises.
"""
import ctypes as C
import os
import re
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libpbso_b200.so")
HEADER = os.path.join(HERE, "..", "include", "pbso_b200.h")

c_dp = C.POINTER(C.c_double)
c_fp = C.POINTER(C.c_float)
c_ip = C.POINTER(C.c_int)
c_vpp = C.POINTER(C.c_void_p)

OK, ERR_INVALID, ERR_CUDA, ERR_IO, ERR_FORMAT, ERR_RANGE, ERR_NO_DEVICE, ERR_UNSUPPORTED = range(8)
PREC_F64, PREC_F32_TILED, PREC_TF32X3, PREC_TC3X = 0, 1, 2, 3


class PbsoError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("pbso error %d: %s" % (code, msg))
        self.code = code


_lib = None


def header_symbols():
    """Every function name declared in include/pbso_b200.h."""
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pbso_[a-z0-9_]+)\s*\(", txt)))


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libpbso_b200.so is not built (run `python -m openpbso_b200.build`); "
                              "there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        L.pbso_last_error.restype = C.c_char_p
        vp = C.c_void_p
        sig = {
            "pbso_device_count": [c_ip],
            "pbso_set_device": [C.c_int],
            "pbso_device_info": [c_ip, c_ip, c_ip, c_dp],
            "pbso_integrator_build": [C.c_double, c_dp, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, c_vpp],
            "pbso_integrator_create": [C.c_int, C.c_double, c_dp, c_dp, c_vpp],
            "pbso_integrator_destroy": [vp],
            "pbso_integrator_size": [vp, c_ip],
            "pbso_integrator_listeners": [vp, c_ip],
            "pbso_integrator_stream": [vp, c_vpp],
            "pbso_modes_storm_buffer": [vp, vp, C.c_int, C.c_int, c_ip, c_dp, C.c_int, c_dp, c_dp, C.c_int],
            "pbso_integrator_get_coeffs": [vp, c_dp, c_dp, c_dp],
            "pbso_integrator_step": [vp, c_dp, c_dp],
            "pbso_integrator_get_state": [vp, c_dp, c_dp],
            "pbso_integrator_set_state": [vp, c_dp, c_dp],
            "pbso_integrator_set_transfer": [vp, c_dp, C.c_int, C.c_int],
            "pbso_render_buffer": [vp, c_dp, c_dp, C.c_int, c_dp, c_dp],
            "pbso_integrator_set_transfer_ffat": [vp, vp, C.c_int, c_dp, C.c_int],
            "pbso_render_buffer_device": [vp, vp, vp, C.c_int, vp, vp],
            "pbso_integrator_sync": [vp],
            "pbso_ffat_load_dir": [C.c_char_p, c_vpp],
            "pbso_ffat_load_file": [C.c_char_p, c_vpp],
            "pbso_ffat_create": [C.c_int, c_ip, c_dp, c_ip, c_dp, C.c_int, C.POINTER(C.c_ubyte), c_vpp],
            "pbso_ffat_destroy": [vp],
            "pbso_ffat_num_maps": [vp, c_ip],
            "pbso_ffat_mode_ids": [vp, c_ip],
            "pbso_ffat_get_map": [vp, C.c_int, c_dp, c_ip, c_ip, c_ip, c_ip, c_dp],
            "pbso_ffat_save_file": [vp, C.c_int, C.c_char_p],
            "pbso_ffat_save_legacy_file": [vp, C.c_int, C.c_char_p],
            "pbso_ffat_eval": [vp, C.c_int, c_dp, C.c_int, C.c_int, c_dp],
            "pbso_ffat_eval_device": [vp, C.c_int, vp, C.c_int, vp, vp],
            "pbso_ffat_eval_device_view": [vp, C.c_int, vp, C.c_int, C.c_int, vp, vp],
            "pbso_ffat_quantise": [vp, C.c_int, C.POINTER(C.c_ubyte), c_dp, c_dp],
            "pbso_ffat_set_compressed_u8": [vp, C.c_int, C.POINTER(C.c_ubyte), C.c_int, c_dp],
            "pbso_ffat_compress": [vp, C.c_int, c_dp],
            "pbso_ffat_get_compressed": [vp, C.c_int, C.POINTER(C.c_ubyte), c_dp, c_dp],
            "pbso_ffat_fitter_create": [C.c_double, c_dp, C.c_int, c_ip, C.c_int, c_vpp],
            "pbso_ffat_fitter_destroy": [vp],
            "pbso_ffat_fitter_info": [vp, c_ip, c_ip, c_ip, c_ip],
            "pbso_ffat_fitter_shell": [vp, C.c_int, c_dp, c_ip],
            "pbso_ffat_fitter_solve": [vp, C.c_int, c_dp, c_dp, C.c_int, c_dp, c_dp],
            "pbso_ffat_fitter_solve_device": [vp, C.c_int, vp, vp, C.c_int, vp, vp, vp],
            "pbso_ffat_fitter_last_kernel_ms": [vp, c_fp],
            "pbso_modes_upload": [c_dp, C.c_int, C.c_int, c_vpp],
            "pbso_modes_read_file": [C.c_char_p, c_vpp, c_ip, c_ip],
            "pbso_modes_omega_squared": [vp, c_dp],
            "pbso_modes_destroy": [vp],
            "pbso_modes_project_vertex": [vp, C.c_int, C.c_int, c_dp, c_dp],
            "pbso_modes_project_face": [vp, C.c_int, c_ip, c_dp, c_dp, c_dp],
            "pbso_modes_project_vertices": [vp, C.c_int, C.c_int, c_ip, c_dp, c_dp],
            "pbso_modes_project_dense": [vp, C.c_int, c_dp, C.c_int, c_dp, C.c_int],
            "pbso_modes_last_kernel_ms": [vp, c_fp],
            "pbso_modes_project_dense_device": [vp, C.c_int, vp, C.c_int, vp, vp],
            "pbso_batch_create": [C.c_int, C.c_int, C.c_double, c_dp, c_dp, c_vpp],
            "pbso_batch_destroy": [vp],
            "pbso_batch_set_transfer": [vp, c_dp],
            "pbso_batch_set_impulses": [vp, C.c_int, c_ip, c_ip, c_dp],
            "pbso_batch_set_transfer_async": [vp, c_dp],
            "pbso_batch_set_impulses_async": [vp, C.c_int, c_ip, c_ip, c_dp],
            "pbso_batch_render_mix": [vp, C.c_int, C.c_int, C.c_int, C.c_int, c_dp],
            "pbso_batch_render_mix_device": [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp],
            "pbso_batch_render_stems": [vp, C.c_int, C.c_int, C.c_int, c_fp],
            "pbso_batch_sync": [vp],
            "pbso_batch_set_state": [vp, c_dp, c_dp],
            "pbso_batch_get_end_state": [vp, C.c_int, C.c_int, c_dp, c_dp],
            "pbso_batch_set_stream": [vp, vp],
            "pbso_batch_last_kernel_ms": [vp, c_fp, c_ip],
            "pbso_comm_unique_id": [C.POINTER(C.c_ubyte)],
            "pbso_comm_init": [C.c_int, C.c_int, C.POINTER(C.c_ubyte), c_vpp],
            "pbso_comm_destroy": [vp],
            "pbso_comm_info": [vp, c_ip, c_ip, c_ip],
            "pbso_comm_shard": [vp, C.c_longlong, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)],
            "pbso_comm_reduce_audio": [vp, vp, C.c_size_t, C.c_int, vp],
            "pbso_comm_reduce_audio_host": [vp, c_dp, C.c_size_t, C.c_int],
            "pbso_measure_fma_peak": [C.c_int, c_dp, c_dp],
            "pbso_measure_tc_peak": [C.c_int, C.c_int, C.c_int, C.c_int, c_dp, c_dp, c_dp],
            "pbso_measure_tc_peak_sustained": [C.c_int, C.c_int, C.c_int, C.c_double, c_dp],
            "pbso_tc_selftest": [C.c_int, c_dp],
            "pbso_tc_gain": [c_dp],
            "pbso_measure_copy_bw": [C.c_size_t, c_dp],
            "pbso_flush_l2": [C.c_size_t],
        }
        for name, args in sig.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = C.c_int
        _lib = L
    return _lib


def check(rc):
    if rc != OK:
        raise PbsoError(rc, lib().pbso_last_error().decode(errors="replace"))


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def dp(a):
    return None if a is None else a.ctypes.data_as(c_dp)


def ip(a):
    return None if a is None else a.ctypes.data_as(c_ip)
