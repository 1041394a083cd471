"""ctypes binding of include/tnqs_b200.h.  There is no CPU fallback: if the CUDA library is
missing or no GPU is visible, the first compute call raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtnqs_b200.so")

TNQS_C64, TNQS_C128 = 0, 1
ERRORS = {1: "EINVAL", 2: "ENOTADJ", 3: "ENSITES", 4: "ECUDA", 5: "EDOMAIN", 6: "ECAPACITY", 7: "ENOGPU"}


class ApplyOpts(C.Structure):
    _fields_ = [("maxdim", C.c_int32), ("mindim", C.c_int32), ("cutoff", C.c_double),
                ("normalize_tensors", C.c_int32), ("sqrt_cutoff", C.c_double),
                ("use_absolute_cutoff", C.c_int32), ("use_relative_cutoff", C.c_int32), ("svd_alg", C.c_int32),
                ("reserved", C.c_int32)]


class BpOpts(C.Structure):
    _fields_ = [("maxiter", C.c_int32), ("tolerance", C.c_double), ("use_tolerance", C.c_int32),
                ("edge_sequence", C.POINTER(C.c_int32)), ("n_seq", C.c_int32)]


class BpReport(C.Structure):
    _fields_ = [("niter", C.c_int32), ("converged", C.c_int32), ("diff", C.c_double)]


class Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_int64), ("bp_ms", C.c_double), ("su_ms", C.c_double),
                ("bp_messages", C.c_int64), ("two_site_gates", C.c_int64), ("bp_sweeps", C.c_int64),
                ("mode_ms", C.c_double), ("gram_ms", C.c_double), ("small_ms", C.c_double),
                ("mode_flops", C.c_double), ("gram_flops", C.c_double),
                ("mode_launches", C.c_int64), ("gram_launches", C.c_int64), ("tc_launches", C.c_int64),
                ("mode_bytes", C.c_double), ("gram_bytes", C.c_double), ("wall_ms", C.c_double), ("sync_ms", C.c_double),
                ("tma_launches", C.c_int64)]


class TnqsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"tnqs_b200 [{ERRORS.get(code, code)}]: {msg}")
        self.code = code


# every symbol include/tnqs_b200.h declares (tests check the library exports all of them)
SYMBOLS = ["tnqs_create", "tnqs_clone", "tnqs_destroy", "tnqs_set_site", "tnqs_site_shape",
           "tnqs_get_site", "tnqs_set_message", "tnqs_get_message", "tnqs_delete_messages",
           "tnqs_get_bond_dims", "tnqs_set_edge_sequence", "tnqs_apply_gates", "tnqs_bp_update",
           "tnqs_expect_local", "tnqs_expect_two_site", "tnqs_vertex_scalars", "tnqs_scale_sites", "tnqs_randomize_sites", "tnqs_site_contract", "tnqs_apply_leg_matrices", "tnqs_comm_unique_id", "tnqs_comm_init",
           "tnqs_get_stats", "tnqs_set_profiling", "tnqs_last_error", "tnqs_version"]

_lib = None


def load():
    """Load libtnqs_b200.so (built by `__graft_entry__.build()` / csrc/Makefile).  Raises
    ImportError if it has not been built — the product path never falls back to the CPU."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; "
                          "g.build()'` (nvcc, sm_100a). tnqs_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i32p, i64p, dp = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_double)
    ip = C.POINTER(C.c_int)
    sig = {
        "tnqs_create": [C.c_int, C.c_int, C.c_int, i32p, i32p, i32p, C.c_int, C.POINTER(vp)],
        "tnqs_clone": [vp, C.POINTER(vp)],
        "tnqs_set_site": [vp, C.c_int, vp, C.c_int, i64p],
        "tnqs_site_shape": [vp, C.c_int, ip, i64p],
        "tnqs_get_site": [vp, C.c_int, vp, C.c_int64],
        "tnqs_set_message": [vp, C.c_int, C.c_int, vp, C.c_int],
        "tnqs_get_message": [vp, C.c_int, C.c_int, vp, C.c_int64, ip, ip],
        "tnqs_delete_messages": [vp],
        "tnqs_get_bond_dims": [vp, i32p],
        "tnqs_set_edge_sequence": [vp, i32p, C.c_int],
        "tnqs_apply_gates": [vp, C.c_int, i32p, i32p, dp, C.POINTER(ApplyOpts), C.POINTER(BpOpts), C.c_int,
                             dp, C.POINTER(BpReport), C.c_int, ip],
        "tnqs_bp_update": [vp, C.POINTER(BpOpts), C.POINTER(BpReport)],
        "tnqs_expect_local": [vp, C.c_int, i32p, dp, dp],
        "tnqs_expect_two_site": [vp, C.c_int, i32p, dp, dp],
        "tnqs_vertex_scalars": [vp, C.c_int, i32p, dp],
        "tnqs_scale_sites": [vp, C.c_int, i32p, dp],
        "tnqs_randomize_sites": [vp, C.c_uint64, C.c_int],
        "tnqs_site_contract": [vp, C.c_int, C.c_int, i32p, dp, C.c_int, C.c_int, dp, dp, C.c_int64, ip],
        "tnqs_apply_leg_matrices": [vp, C.c_int, i32p, i32p, dp],
        "tnqs_comm_unique_id": [vp],
        "tnqs_comm_init": [vp, C.c_int, C.c_int, C.c_char_p, i32p],
        "tnqs_get_stats": [vp, C.POINTER(Stats), C.c_int],
        "tnqs_set_profiling": [vp, C.c_int],
    }
    for name, args in sig.items():
        f = getattr(lib, name)
        f.argtypes = args
        f.restype = C.c_int
    lib.tnqs_destroy.argtypes = [vp]
    lib.tnqs_destroy.restype = None
    lib.tnqs_last_error.restype = C.c_char_p
    lib.tnqs_version.restype = C.c_char_p
    _lib = lib
    return lib


def check(code):
    if code != 0:
        raise TnqsError(code, load().tnqs_last_error().decode("utf-8", "replace"))
