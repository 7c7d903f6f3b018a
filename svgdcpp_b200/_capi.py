"""ctypes binding of include/svgd_b200.h (plain pointers and sizes, no torch types)."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libsvgd_b200.so")

OK, ERR_INVALID, ERR_DIMENSION, ERR_UNSET, ERR_CUDA, ERR_NCCL, ERR_NOMEM, ERR_NUMERIC = 0, -1, -2, -3, -4, -5, -6, -7
PRECISION_F64, PRECISION_TC32 = 0, 1
SCALE_MEDIAN, SCALE_HESSIAN, SCALE_FIXED = 0, 1, 2
OPT_ADAGRAD, OPT_ADAM, OPT_RMSPROP = 0, 1, 2
TC32_AUTO, TC32_FAST, TC32_PRECISE = 0, 1, 2

_dp = C.POINTER(C.c_double)
_ctx = C.c_void_p

GRAD_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p)


class Stats(C.Structure):
    _fields_ = [
        ("iterations", C.c_uint64), ("kernel_launches", C.c_uint64), ("median_passes", C.c_uint64),
        ("median_bracket_hits", C.c_uint64), ("last_scale", C.c_double),
        ("ms_median", C.c_double), ("ms_grad", C.c_double), ("ms_phi", C.c_double), ("ms_comm", C.c_double),
        ("phi_launches", C.c_uint64), ("ms_phi_kernel", C.c_double), ("ms_grad_kernel", C.c_double),
    ]


# name -> (restype, argtypes); every symbol include/svgd_b200.h declares
SIGNATURES = {
    "svgdb_create": (C.c_int, [C.POINTER(_ctx), C.c_int, C.c_int64, C.c_int32, C.c_int]),
    "svgdb_destroy": (None, [_ctx]),
    "svgdb_last_error": (C.c_char_p, [_ctx]),
    "svgdb_version": (C.c_char_p, []),
    "svgdb_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "svgdb_set_stream": (C.c_int, [_ctx, C.c_void_p]),
    "svgdb_nccl_unique_id": (C.c_int, [C.c_void_p, C.c_size_t]),
    "svgdb_comm_init": (C.c_int, [_ctx, C.c_int, C.c_int, C.c_void_p, C.c_size_t]),
    "svgdb_set_particles": (C.c_int, [_ctx, _dp]),
    "svgdb_get_particles": (C.c_int, [_ctx, _dp]),
    "svgdb_set_model_mvn": (C.c_int, [_ctx, _dp, _dp]),
    "svgdb_set_model_mvn_sum": (C.c_int, [_ctx, C.c_int32, _dp, _dp]),
    "svgdb_set_model_device_hook": (C.c_int, [_ctx, C.c_void_p, C.c_void_p]),
    "svgdb_set_kernel_rbf": (C.c_int, [_ctx, C.c_int, C.c_double]),
    "svgdb_set_optimizer": (C.c_int, [_ctx, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double]),
    "svgdb_set_bounds": (C.c_int, [_ctx, _dp, _dp, C.c_int32]),
    "svgdb_initialize": (C.c_int, [_ctx]),
    "svgdb_step": (C.c_int, [_ctx, C.c_int64]),
    "svgdb_compute_phi": (C.c_int, [_ctx, _dp, _dp]),
    "svgdb_compute_scale": (C.c_int, [_ctx, _dp]),
    "svgdb_compute_log_model_grad": (C.c_int, [_ctx, _dp]),
    "svgdb_get_opt_state": (C.c_int, [_ctx, _dp, _dp, C.POINTER(C.c_uint64)]),
    "svgdb_set_opt_state": (C.c_int, [_ctx, _dp, _dp, C.c_uint64]),
    "svgdb_sync": (C.c_int, [_ctx]),
    "svgdb_set_profiling": (C.c_int, [_ctx, C.c_int]),
    "svgdb_get_stats": (C.c_int, [_ctx, C.POINTER(Stats)]),
    "svgdb_reset_stats": (C.c_int, [_ctx]),
    "svgdb_time_steps": (C.c_int, [_ctx, C.c_int64, C.POINTER(C.c_float)]),
    "svgdb_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "svgdb_host_free": (C.c_int, [C.c_void_p]),
    "svgdb_probe_peak": (C.c_int, [C.c_int, C.c_int, _dp]),
    "svgdb_get_scale_matrix": (C.c_int, [_ctx, _dp]),
    "svgdb_step_host": (C.c_int, [_ctx, _dp, _dp, C.c_int64]),
    "svgdb_compute_kernel_matrices": (C.c_int, [_ctx, _dp, _dp, _dp]),
    "svgdb_compute_log_model": (C.c_int, [_ctx, _dp]),
    "svgdb_local_rows": (C.c_int, [_ctx, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "svgdb_set_particles_rows": (C.c_int, [_ctx, _dp]),
    "svgdb_get_particles_rows": (C.c_int, [_ctx, _dp]),
    "svgdb_set_tc32_variant": (C.c_int, [_ctx, C.c_int]),
    "svgdb_time_kernel": (C.c_int, [_ctx, C.c_int, C.c_int, C.c_int, C.c_double, C.POINTER(C.c_float)]),
}

_lib = None


def load():
    """Loads libsvgd_b200.so.  Fails loudly if the CUDA extension has not been built: there is
    no Python/CPU stand-in for it."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build the CUDA extension first "
                "(python -m svgdcpp_b200.build, or __graft_entry__.build()). There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the library does not export it
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
