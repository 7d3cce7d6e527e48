"""ctypes binding of libkcrn.so (the C ABI declared in include/kcrn.h).

There is deliberately no fallback: if the shared object is missing or a call fails, an exception
is raised.  The oracle under ``oracle/`` is test infrastructure and is never imported from here.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_int32, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libkcrn.so")

KCRN_OK, KCRN_E_ARG, KCRN_E_CUDA, KCRN_E_STATE, KCRN_E_NOTPD, KCRN_E_NCCL = 0, -1, -2, -3, -4, -5

# name -> (restype, argtypes); kept in one table so tests can check it against include/kcrn.h
SIGNATURES = {
    "kcrn_abi_version": (c_int, []),
    "kcrn_last_error": (c_char_p, []),
    "kcrn_create": (c_int, [POINTER(c_void_p), c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p,
                            c_void_p, c_double, c_int64, c_int, c_void_p]),
    "kcrn_destroy": (c_int, [c_void_p]),
    "kcrn_set_stream": (c_int, [c_void_p, c_void_p]),
    "kcrn_sync": (c_int, [c_void_p]),
    "kcrn_nccl_unique_id": (c_int, [c_void_p]),
    "kcrn_comm_init": (c_int, [c_void_p, c_void_p, c_int, c_int]),
    "kcrn_p2p_export": (c_int, [c_void_p, c_void_p]),
    "kcrn_p2p_attach": (c_int, [c_void_p, c_void_p]),
    "kcrn_p2p_enable": (c_int, [c_void_p, c_int]),
    "kcrn_csc_view": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "kcrn_set_x": (c_int, [c_void_p, c_void_p]),
    "kcrn_set_x_only": (c_int, [c_void_p, c_void_p]),
    "kcrn_value": (c_int, [c_void_p, c_void_p]),
    "kcrn_gradient": (c_int, [c_void_p, c_void_p]),
    "kcrn_hvp": (c_int, [c_void_p, c_void_p, c_void_p]),
    "kcrn_value_grad_host": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "kcrn_hvp_host": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "kcrn_cached_Ax": (c_int, [c_void_p, POINTER(c_void_p)]),
    "kcrn_set_Ax": (c_int, [c_void_p, c_void_p]),
    "kcrn_lanczos": (c_int, [c_void_p, c_void_p, c_int, c_int]),
    "kcrn_krylov_run": (c_int, [c_void_p, c_int, c_int, c_int, c_double, c_double, c_double, c_double, c_double, c_int,
                                c_double, c_int, c_int, c_double, c_void_p, c_void_p, c_void_p]),
    "kcrn_run_view": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "kcrn_current_x": (c_int, [c_void_p, c_void_p]),
    "kcrn_values_batch": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "kcrn_svmlight_count": (c_int, [c_char_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "kcrn_svmlight_parse": (c_int, [c_char_p, c_int64, c_int, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "kcrn_lanczos_view": (c_int, [c_void_p, POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p),
                                  POINTER(c_void_p)]),
    "kcrn_lanczos_begin": (c_int, [c_void_p, c_void_p, c_int, c_int]),
    "kcrn_lanczos_current": (c_int, [c_void_p, POINTER(c_void_p)]),
    "kcrn_lanczos_feed": (c_int, [c_void_p, c_void_p, POINTER(c_int)]),
    "kcrn_cubic_tridiag": (c_int, [c_void_p, c_void_p, c_int, c_double, c_double, c_double, c_double, c_int,
                                   c_void_p, c_void_p]),
    "kcrn_cubic_dense": (c_int, [c_void_p, c_void_p, c_int, c_double, c_double, c_double, c_int, c_void_p, c_void_p]),
    "kcrn_cubic_from_lanczos": (c_int, [c_void_p, c_double, c_double, c_double, c_int]),
    "kcrn_axpy_basis": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p]),
    "kcrn_trial_point": (c_int, [c_void_p, c_void_p, c_void_p]),
    "kcrn_step_record": (c_int, [c_void_p, c_int, POINTER(c_double)]),
    "kcrn_krylov_trial": (c_int, [c_void_p, c_int, c_int, c_int, c_double, c_double, c_double, c_int, c_void_p,
                                  POINTER(c_double)]),
    "kcrn_hessian_dense": (c_int, [c_void_p, c_void_p]),
    "kcrn_partial_gradient": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "kcrn_partial_hessian": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "kcrn_update_Ax": (c_int, [c_void_p, c_void_p, c_void_p, c_int]),
    "kcrn_counters": (c_int, [c_void_p, POINTER(c_int64), c_int]),
    "kcrn_time_sweep": (c_int, [c_void_p, c_int, c_int, POINTER(c_double)]),
    "kcrn_flush_l2": (c_int, [c_void_p, ctypes.c_size_t]),
    "kcrn_gather_variants": (c_int, [c_void_p, POINTER(c_int)]),
    "kcrn_plan_persist": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int64, POINTER(c_int64), c_void_p, c_int64,
                                  POINTER(c_int64), POINTER(c_int)]),
    "kcrn_plan_tiles": (c_int, [c_void_p, c_int64, c_void_p, c_int64, POINTER(c_int64), c_void_p, c_int64,
                                POINTER(c_int64), POINTER(c_int)]),
}


class KcrnError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libkcrn error {code}: {message}")
        self.code = code


class NotPositiveDefinite(KcrnError, ArithmeticError):
    """T + r I lost positive definiteness -- where the reference's dposv raises LinAlgError."""


_lib = None


def load():
    """Load libkcrn.so (once).  Raises if it has not been built: there is no CPU path."""
    global _lib
    if _lib is None:
        global LIB_PATH
        LIB_PATH = os.environ.get("KCRN_LIBRARY", LIB_PATH)      # development: a kernel variant built elsewhere
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing -- run `python -m kcrn_b200.build` "
                              "(__graft_entry__.build()); there is no fallback implementation")
        lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        if lib.kcrn_abi_version() != 1:
            raise ImportError("libkcrn.so ABI version mismatch; rebuild")
        _lib = lib
    return _lib


def check(code: int):
    if code == KCRN_OK:
        return
    msg = (load().kcrn_last_error() or b"").decode("utf-8", "replace")
    if code == KCRN_E_NOTPD:
        raise NotPositiveDefinite(code, msg)
    if code == KCRN_E_ARG:
        raise ValueError(f"libkcrn: {msg}")
    raise KcrnError(code, msg)
