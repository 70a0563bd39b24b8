"""openpbso_b200 -- B200-native modal sound synthesis path behind openpbso's API.

The product is `libpbso_b200.so` (hand-written sm_100a CUDA behind the C ABI of
include/pbso_b200.h) plus the C++ header mirror of the reference API in include/openpbso/.
This package is the thin Python view of the C ABI used by tests/ and bench.py; class and method
names follow the reference (ModalIntegrator, ModalSolver::step hot loop, FFAT maps, ModeData).
"""
from ._capi import (PbsoError, PREC_F64, PREC_F32_TILED, PREC_TF32X3, PREC_TC3X, OK, ERR_INVALID, ERR_CUDA, ERR_IO, ERR_FORMAT,
                    ERR_RANGE, ERR_NO_DEVICE, ERR_UNSUPPORTED, lib, header_symbols, LIB_PATH)
from .api import (Comm, ModalIntegrator, FFATMaps, FFATFitter, ModeShapes, BatchRenderer, device_count, set_device,
                  device_info, measure_fma_peak, measure_tc_peak, measure_tc_peak_sustained, tc_selftest, tc_gain, measure_copy_bw, flush_l2)

__all__ = ["PbsoError", "Comm", "ModalIntegrator", "FFATMaps", "FFATFitter", "ModeShapes", "BatchRenderer", "device_count",
           "set_device", "device_info", "measure_fma_peak", "measure_tc_peak", "measure_tc_peak_sustained", "tc_selftest", "tc_gain", "measure_copy_bw", "flush_l2", "lib",
           "header_symbols", "PREC_F64", "PREC_F32_TILED", "PREC_TF32X3", "PREC_TC3X"]
