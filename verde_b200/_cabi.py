"""
ctypes binding of ``libverde_b200.so`` (the C ABI in ``include/verde_b200.h``).

This module is deliberately thin: numpy arrays (or raw device pointers) in,
status codes out.  There is NO CPU fallback -- if the shared library is missing
or no CUDA device is present every compute call raises.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libverde_b200.so")

c_double_p = ctypes.POINTER(ctypes.c_double)
i64 = ctypes.c_int64
dbl = ctypes.c_double

E_BADARG, E_CUDA, E_NOMEM, E_NOCONV = -1, -2, -3, -4


class FitInfo(ctypes.Structure):
    """Mirror of ``vb200_fit_info`` (include/verde_b200.h)."""

    _fields_ = [
        ("info", ctypes.c_int32),
        ("block_rows", ctypes.c_int32),
        ("rows", ctypes.c_int64),
        ("cols", ctypes.c_int64),
        ("gram_ms", ctypes.c_double),
        ("factor_ms", ctypes.c_double),
        ("solve_ms", ctypes.c_double),
        ("gemm_ms", ctypes.c_double),
        ("gemm_flops", ctypes.c_double),
        ("gemm_launches", ctypes.c_int64),
        ("kernel_launches", ctypes.c_int64),
        ("diag_min", ctypes.c_double),
        ("diag_max", ctypes.c_double),
        ("lambda_max", ctypes.c_double),
        ("lambda_min_est", ctypes.c_double),
        ("cond_est", ctypes.c_double),
        ("cond_upper", ctypes.c_double),
        ("backward_err", ctypes.c_double),
        ("diag_ms", ctypes.c_double),
        ("svd_rank", ctypes.c_int64),
        ("svd_sweeps", ctypes.c_int64),
        ("sigma_max", ctypes.c_double),
        ("sigma_min_kept", ctypes.c_double),
        ("power_its", ctypes.c_int32),
        ("inverse_its", ctypes.c_int32),
        ("shared_gemm_ms", ctypes.c_double),
        ("shared_gemm_flops", ctypes.c_double),
        ("shared_gemm_launches", ctypes.c_int64),
    ]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


# name -> (restype, argtypes); the single source of truth for the symbol check in
# tests/test_cabi_symbols.py (must match include/verde_b200.h)
SIGNATURES = {
    "vb200_version": (ctypes.c_int, []),
    "vb200_last_error": (ctypes.c_char_p, []),
    "vb200_device_count": (ctypes.c_int, []),
    "vb200_set_device": (ctypes.c_int, [ctypes.c_int]),
    "vb200_set_option": (ctypes.c_int, [ctypes.c_char_p, dbl]),
    "vb200_get_option": (ctypes.c_int, [ctypes.c_char_p, ctypes.POINTER(dbl)]),
    "vb200_release_workspace": (ctypes.c_int, []),
    "vb200_launch_count": (i64, []),
    "vb200_measure_fp64_peak_tflops": (dbl, [ctypes.c_int]),
    "vb200_jacobian": (ctypes.c_int, [ctypes.c_void_p] * 2 + [i64] + [ctypes.c_void_p] * 2 + [i64, dbl, ctypes.c_void_p]),
    "vb200_jacobian_2d": (ctypes.c_int, [ctypes.c_void_p] * 2 + [i64] + [ctypes.c_void_p] * 2 + [i64, dbl, dbl, ctypes.c_void_p]),
    "vb200_predict": (ctypes.c_int, [ctypes.c_void_p] * 2 + [i64] + [ctypes.c_void_p] * 3 + [i64, dbl, ctypes.c_void_p]),
    "vb200_predict_multi": (ctypes.c_int, [ctypes.c_void_p] * 2 + [i64] + [ctypes.c_void_p] * 3 + [i64, ctypes.c_int32, dbl, ctypes.c_void_p]),
    "vb200_predict_2d": (ctypes.c_int, [ctypes.c_void_p] * 2 + [i64] + [ctypes.c_void_p] * 3 + [i64, dbl, dbl] + [ctypes.c_void_p] * 2),
    "vb200_spline_fit": (ctypes.c_int, [ctypes.c_void_p] * 4 + [i64] + [ctypes.c_void_p] * 2 + [i64, dbl, dbl, ctypes.c_void_p, ctypes.POINTER(FitInfo)]),
    "vb200_spline_fit_2d": (ctypes.c_int, [ctypes.c_void_p] * 4 + [i64] + [ctypes.c_void_p] * 2 + [i64, dbl, dbl, dbl, ctypes.c_void_p, ctypes.POINTER(FitInfo)]),
    "vb200_gram_doubles": (i64, [i64]),
    "vb200_stats_doubles": (i64, [i64]),
    "vb200_gram_accumulate": (ctypes.c_int, [ctypes.c_int] + [ctypes.c_void_p] * 4 + [i64] + [ctypes.c_void_p] * 2 + [i64, dbl, dbl] + [ctypes.c_void_p] * 2 + [ctypes.c_int, ctypes.POINTER(FitInfo)]),
    "vb200_gram_solve": (ctypes.c_int, [ctypes.c_void_p] * 2 + [i64, i64, dbl, ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(FitInfo)]),
    "vb200_gram_solve_multi": (ctypes.c_int, [ctypes.c_void_p] * 2 + [i64, i64, ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(FitInfo)]),
    "vb200_chol_begin": (ctypes.c_int, [ctypes.c_void_p] * 2 + [i64, i64, dbl]),
    "vb200_chol_panel_range": (ctypes.c_int, [i64, ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(i64), ctypes.POINTER(i64)]),
    "vb200_chol_panel": (ctypes.c_int, [ctypes.c_void_p, i64, ctypes.c_int32, ctypes.c_int32]),
    "vb200_chol_update": (ctypes.c_int, [ctypes.c_void_p, i64] + [ctypes.c_int32] * 4),
    "vb200_chol_update_ranges": (ctypes.c_int, [ctypes.c_void_p, i64] + [ctypes.c_int32] * 3 + [ctypes.c_void_p] * 2),
    "vb200_shared_alloc": (ctypes.c_int, [i64, ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p]),
    "vb200_shared_free": (ctypes.c_int, [ctypes.c_void_p]),
    "vb200_shared_open": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)]),
    "vb200_shared_close": (ctypes.c_int, [ctypes.c_void_p]),
    "vb200_chol_wait_panel": (ctypes.c_int, [ctypes.c_void_p, i64]),
    "vb200_stream_wait_flag": (ctypes.c_int, [ctypes.c_void_p, i64, ctypes.c_void_p]),
    "vb200_peer_copy_async": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, i64, ctypes.c_void_p]),
    "vb200_chol_finish": (ctypes.c_int, [ctypes.c_void_p, i64, ctypes.c_void_p, ctypes.POINTER(FitInfo)]),
    "vb200_block_split": (ctypes.c_int, [ctypes.c_void_p] * 2 + [i64, ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.POINTER(i64)]),
    "vb200_block_ids": (ctypes.c_int, [ctypes.c_void_p]),
    "vb200_block_reduce_stream": (ctypes.c_int, [ctypes.c_void_p] * 2 + [i64, ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int32,
                                                 ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int32, i64, ctypes.c_void_p,
                                                 ctypes.c_void_p, ctypes.POINTER(i64)]),
    "vb200_block_aggregate": (ctypes.c_int, [ctypes.c_void_p] * 2 + [i64, ctypes.c_int32, ctypes.c_void_p]),
    "vb200_distance_mask": (ctypes.c_int, [ctypes.c_void_p] * 2 + [i64, dbl] + [ctypes.c_void_p] * 2 + [i64, ctypes.c_void_p]),
    "vb200_padded_cols": (i64, [i64]),
    "vb200_gram_factor": (ctypes.c_int, [ctypes.c_void_p] * 2 + [i64, i64, dbl, ctypes.c_void_p, ctypes.POINTER(FitInfo)]),
    "vb200_chol_solve": (ctypes.c_int, [ctypes.c_void_p] * 2 + [i64] + [ctypes.c_void_p] * 2 + [ctypes.POINTER(FitInfo)]),
    "vb200_jacobian_transpose_apply": (ctypes.c_int, [ctypes.c_void_p] * 2 + [i64] + [ctypes.c_void_p] * 2 + [i64, dbl] + [ctypes.c_void_p] * 2),
    "vb200_spline_fit_lstsq": (ctypes.c_int, [ctypes.c_void_p] * 4 + [i64] + [ctypes.c_void_p] * 2 + [i64, dbl, dbl, ctypes.c_void_p, ctypes.POINTER(FitInfo)]),
    "vb200_spline_fit_2d_lstsq": (ctypes.c_int, [ctypes.c_void_p] * 4 + [i64] + [ctypes.c_void_p] * 2 + [i64, dbl, dbl, dbl, ctypes.c_void_p, ctypes.POINTER(FitInfo)]),
    "vb200_least_squares_lstsq": (ctypes.c_int, [ctypes.c_void_p, i64, i64] + [ctypes.c_void_p] * 2 + [dbl, ctypes.c_void_p, ctypes.POINTER(FitInfo)]),
    "vb200_normal_residual": (ctypes.c_int, [ctypes.c_int] + [ctypes.c_void_p] * 4 + [i64] + [ctypes.c_void_p] * 2 + [i64, dbl, dbl] + [ctypes.c_void_p] * 2),
    "vb200_backward_error": (ctypes.c_int, [ctypes.c_void_p] * 2 + [i64, i64, dbl, ctypes.c_void_p, dbl, ctypes.POINTER(dbl)]),
    "vb200_least_squares": (ctypes.c_int, [ctypes.c_void_p, i64, i64] + [ctypes.c_void_p] * 2 + [dbl, ctypes.c_void_p, ctypes.POINTER(FitInfo)]),
}

_lib = None


class EngineError(RuntimeError):
    """The CUDA engine is missing or reported a CUDA/argument error."""


def load():
    """Load the shared library (raises EngineError when it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EngineError(
                "verde_b200 CUDA library not built: {} is missing. Build it with "
                "`python -c 'import __graft_entry__ as g; g.build()'` (there is no CPU "
                "fallback).".format(LIB_PATH)
            )
        lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


def require_device():
    lib = load()
    if lib.vb200_device_count() < 1:
        raise EngineError("verde_b200 needs a CUDA device (B200, sm_100a); none is visible and there is no CPU fallback")
    return lib


def get_option(key):
    "Current value of a library tunable (``vb200_get_option``)."
    value = dbl(0.0)
    check(load().vb200_get_option(key.encode() if isinstance(key, str) else key, ctypes.byref(value)), "get_option")
    return value.value


def ptr(a):
    """void* of a numpy array, a torch CUDA tensor, an int device address, or None."""
    if a is None:
        return None
    if isinstance(a, int):
        return ctypes.c_void_p(a)
    if isinstance(a, np.ndarray):
        return ctypes.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):
        return ctypes.c_void_p(a.data_ptr())
    raise TypeError("cannot take the address of {!r}".format(type(a)))


def f64(a):
    """Contiguous 1-D float64 view/copy of an array-like (host)."""
    return np.ascontiguousarray(np.ravel(np.atleast_1d(np.asarray(a))), dtype=np.float64)


def raise_status(rc, what, msg):
    """Raise the exception a non-zero status code stands for: <0 -> EngineError family, >0 -> LinAlgError."""
    if rc > 0:
        raise np.linalg.LinAlgError(
            "{}: the damped normal matrix is not positive definite (pivot {}): {}".format(what, rc, msg)
        )
    if rc == E_NOMEM:
        raise MemoryError("{}: out of device memory: {}".format(what, msg))
    if rc == E_BADARG:
        raise ValueError("{}: {}".format(what, msg))
    if rc == E_NOCONV:
        raise np.linalg.LinAlgError("{}: {}".format(what, msg))
    raise EngineError("{}: {}".format(what, msg))


def check(rc, what):
    """Translate a status code: <0 -> EngineError, >0 -> LinAlgError, 0 -> ok."""
    if rc == 0:
        return
    raise_status(rc, what, load().vb200_last_error().decode("utf-8", "replace"))
