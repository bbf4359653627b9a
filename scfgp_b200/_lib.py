"""ctypes binding of ``include/scfgp_b200.h``.

There is no CPU fallback: if the CUDA library has not been built, importing any compute
entry point raises (``python -m scfgp_b200.build`` or ``__graft_entry__.build()`` builds it).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libscfgp_b200.so")

OK, ERR_ARG, ERR_CUDA, ERR_NOT_PD, ERR_NONFINITE = 0, 1, 2, 3, 4
FF_PHASE, FF_NOPHASE = 0, 1
RED1, RED2, RED3 = 1, 2, 3
PREC_FP64, PREC_EMU_I8 = 0, 1
FAM_NONE, FAM_FEATURES, FAM_TGEMM, FAM_BACKWARD, FAM_SYRK = 0, 1, 2, 3, 4

_dp = C.POINTER(C.c_double)
_h = C.c_void_p
_ll = C.c_longlong

# name -> (restype, argtypes); mirrors include/scfgp_b200.h declaration by declaration
PROTOTYPES = {
    "scfgp_abi_version": (C.c_int, []),
    "scfgp_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_h)]),
    "scfgp_destroy": (C.c_int, [_h]),
    "scfgp_last_error": (C.c_char_p, [_h]),
    "scfgp_last_pivot": (_ll, [_h]),
    "scfgp_num_params": (_ll, [_h]),
    "scfgp_num_features": (_ll, [_h]),
    "scfgp_set_stream": (C.c_int, [_h, C.c_void_p]),
    "scfgp_set_chunk_rows": (C.c_int, [_h, C.c_int]),
    "scfgp_set_streams": (C.c_int, [_h, C.c_int]),
    "scfgp_set_precision": (C.c_int, [_h, C.c_int, C.c_int]),
    "scfgp_get_precision": (C.c_int, [_h, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "scfgp_set_precision2": (C.c_int, [_h, C.c_int, C.c_int, C.c_int]),
    "scfgp_get_precision2": (C.c_int, [_h, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "scfgp_set_data_host": (C.c_int, [_h, _dp, _dp, _ll, _ll]),
    "scfgp_set_data_host_async": (C.c_int, [_h, _dp, _dp, _ll, _ll]),
    "scfgp_set_data_device": (C.c_int, [_h, C.c_void_p, C.c_void_p, _ll, _ll]),
    "scfgp_select_rows": (C.c_int, [_h, C.POINTER(_ll), _ll, _ll]),
    "scfgp_eval": (C.c_int, [_h, _dp, _dp, _dp, _dp, _dp]),
    "scfgp_predict": (C.c_int, [_h, _dp, _dp, _ll, _dp, _dp, _dp, _dp]),
    "scfgp_set_factors": (C.c_int, [_h, _dp, _dp]),
    "scfgp_predict_device": (C.c_int, [_h, _dp, C.c_void_p, _ll, C.c_void_p, C.c_void_p]),
    "scfgp_eval_begin": (C.c_int, [_h, _dp, C.c_int]),
    "scfgp_eval_pass1": (C.c_int, [_h]),
    "scfgp_eval_factor": (C.c_int, [_h]),
    "scfgp_eval_pass2": (C.c_int, [_h]),
    "scfgp_eval_backward": (C.c_int, [_h]),
    "scfgp_eval_pass3": (C.c_int, [_h]),
    "scfgp_eval_finish": (C.c_int, [_h, _dp, _dp, _dp, _dp]),
    "scfgp_reduce_buffer": (C.c_int, [_h, C.c_int, C.POINTER(C.c_void_p), C.POINTER(_ll)]),
    "scfgp_stream_handoff": (C.c_int, [_h, C.c_void_p, C.c_int]),
    "scfgp_eval_prefetch_factors": (C.c_int, [_h, _dp, _dp]),
    "scfgp_set_phi_cache": (C.c_int, [_h, C.c_int, C.c_double]),
    "scfgp_phi_cache_bytes": (_ll, [_h]),
    "scfgp_set_guard": (C.c_int, [_h, C.c_int, C.c_double, C.c_double]),
    "scfgp_guard_last": (C.c_int, [_h, _dp, _dp, _dp, C.POINTER(_ll), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "scfgp_guard_escalate": (C.c_int, [_h]),
    "scfgp_device_factors": (C.c_int, [_h, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(_ll)]),
    "scfgp_launch_count": (_ll, [_h]),
    "scfgp_last_pass_ms": (C.c_int, [_h, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float),
                                     C.POINTER(C.c_float)]),
    "scfgp_set_tracing": (C.c_int, [_h, C.c_int]),
    "scfgp_trace_dump": (C.c_int, [_h, C.c_char_p]),
    "scfgp_set_profiling": (C.c_int, [_h, C.c_int]),
    "scfgp_get_profile": (C.c_int, [_h, C.POINTER(C.c_double), C.POINTER(_ll)]),
    "scfgp_scaler_colstats": (C.c_int, [C.c_int, C.c_void_p, _ll, C.c_int, _ll, _dp, _dp, C.POINTER(C.c_int)]),
    "scfgp_scaler_gather_column": (C.c_int, [C.c_int, C.c_void_p, _ll, _ll, C.c_int, C.c_double, C.c_double,
                                             C.c_void_p]),
    "scfgp_scaler_column_moments": (C.c_int, [C.c_int, C.c_void_p, _ll, C.c_double, _dp, _dp, _dp]),
    "scfgp_scaler_column_sums": (C.c_int, [C.c_int, C.c_void_p, _ll, C.c_double, C.c_int, C.c_double, _dp, _dp]),
    "scfgp_scaler_column_distinct": (C.c_int, [C.c_int, C.c_void_p, _ll, C.c_int, C.POINTER(C.c_int)]),
    "scfgp_scaler_column_distinct_values": (C.c_int, [C.c_int, C.c_void_p, _ll, C.c_int, C.POINTER(C.c_int), _dp]),
    "scfgp_scaler_forward": (C.c_int, [C.c_int, C.c_int, C.c_void_p, _ll, _ll, C.POINTER(C.c_int), C.c_int,
                                       _dp, _dp, _dp, _dp, _dp, C.c_void_p, _ll]),
    "scfgp_dbg_gemm": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp]),
    "scfgp_dbg_emu_gemm": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp]),
    "scfgp_dbg_tile_timing": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_ulonglong), C.c_int]),
    "scfgp_dbg_mma_stream_peak": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp]),
    "scfgp_dbg_factor": (C.c_int, [C.c_int, C.c_int, _dp, _dp, _dp, C.POINTER(_ll)]),
    "scfgp_dbg_features": (C.c_int, [_h, _dp, _dp, _ll, _dp]),
}

_lib = None


def load():
    """Load the shared library (once) and attach prototypes.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"scfgp_b200: CUDA library not built ({LIB_PATH}); run `python -m scfgp_b200.build`. "
            "There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def dptr(a: np.ndarray):
    """float64 C-contiguous ndarray -> double*"""
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(_dp)


def as_f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)
