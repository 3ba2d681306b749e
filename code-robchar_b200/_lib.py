"""ctypes binding of librobchar_b200.so (include/robchar_b200.h).

There is deliberately NO fallback: if the CUDA library has not been built, or no CUDA device is
present when a compute entry point is called, this raises.  Build with
``python -c "import __graft_entry__ as g; g.build()"`` (or ``code-robchar_b200/csrc/build.sh``).
"""
from __future__ import annotations

import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RC_LIB_PATH") or os.path.join(_HERE, "librobchar_b200.so")  # override: tuning builds
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "robchar_b200.h")

RC_OK = 0
RC_ERR_BAD_ARG, RC_ERR_NULL, RC_ERR_CUDA, RC_ERR_NONCONV, RC_ERR_ILLEGAL_FIDS, RC_ERR_WORKSPACE = 1, 2, 3, 4, 5, 6
MODEL_COMPLEX3, MODEL_REAL2 = 0, 1
NUM_STATS = 15
MAX_NSPIN = 32


class RobcharLibraryError(RuntimeError):
    """The native library is missing or failed; there is no CPU path to fall back to."""


class EigensolverNonConvergence(RobcharLibraryError):
    pass


_i64, _i32, _u64, _f64, _vp, _sz = C.c_int64, C.c_int, C.c_uint64, C.c_double, C.c_void_p, C.c_size_t

class ObjectiveFrame(C.Structure):
    """rc_objective_frame (include/robchar_b200.h)."""
    _fields_ = [("x_host", C.c_void_p), ("rows_host", C.c_void_p), ("fids_host", C.c_void_p), ("stats_host", C.c_void_p),
                ("amps_host", C.c_void_p), ("stream", C.c_void_p), ("m", C.c_int64), ("dkw_eps", C.c_double),
                ("nspin", C.c_int32), ("inspin", C.c_int32), ("outspin", C.c_int32), ("model", C.c_int32),
                ("zz", C.c_int32), ("reserved", C.c_int32)]


_SIGNATURES = {
    "rc_version": (C.c_int, []),
    "rc_last_error": (C.c_char_p, []),
    "rc_launch_count": (C.c_ulonglong, []),
    "rc_device_info": (C.c_int, [_vp, _vp, _vp]),
    "rc_fidelity_mc": (C.c_int, [_vp, _i64, _i32, _i32, _i32, _vp, _i32, _i64, _i32, _i32, _u64, _i64, _i64, _vp, _vp,
                                 _vp, _vp]),
    "rc_fidelity_mc_stats": (C.c_int, [_vp, _i64, _i32, _i32, _i32, _vp, _i32, _i64, _i32, _i32, _u64, _i64, _i64, _vp, _f64,
                                       _vp, _vp, _vp, _vp, _vp]),
    "rc_stats_unsorted": (C.c_int, [_vp, _i64, _i64, _f64, _vp, _vp, _vp]),
    "rc_evolution_kernel_name": (C.c_int, [_i32, _i32, _i32, C.c_char_p, _sz]),
    "rc_rim_p": (C.c_int, [_vp, _i64, _i64, _f64, _vp, _vp, _vp]),
    "rc_spectral_fallbacks": (C.c_int, [_vp, _i32, _vp]),
    "rc_philox_normals": (C.c_int, [_i64, _i32, _i32, _i64, _i32, _u64, _i64, _i64, _vp, _vp]),
    "rc_stats_workspace_bytes": (_sz, [_i64, _i64]),
    "rc_stats": (C.c_int, [_vp, _i64, _i64, _f64, _vp, _vp, _vp, _vp, _sz, _vp]),
    "rc_fidelity_stats_workspace_bytes": (_sz, [_i64, _i64]),
    "rc_fidelity_stats": (C.c_int, [_vp, _i64, _i32, _i32, _i32, _vp, _i32, _i64, _i32, _i32, _u64, _i64, _i64, _vp, _f64,
                                    _vp, _vp, _vp, _sz, _vp]),
    "rc_draw_shard_range": (C.c_int, [_i64, _i32, _i32, _vp, _vp, _vp, _vp]),
    "rc_fidelity_stats_blocks_workspace_bytes": (_sz, [_i64, _i64, _i32]),
    "rc_fidelity_stats_blocks": (C.c_int, [_vp, _i64, _i32, _i32, _i32, _vp, _i32, _i64, _i32, _i32, _u64, _i64, _i64, _i32, _i32,
                                           _f64, _vp, _vp, _vp, _sz, _vp]),
    "rc_stats_from_blocks": (C.c_int, [_vp, _i64, _i64, _f64, _vp, _vp]),
    "rc_peer_alloc": (C.c_int, [_sz, _vp, _vp]),
    "rc_peer_open": (C.c_int, [_vp, _vp]),
    "rc_peer_close": (C.c_int, [_vp]),
    "rc_peer_free": (C.c_int, [_vp]),
    "rc_peer_push_columns": (C.c_int, [_vp, _i32, _vp, _i64, _i64, _i64, _i64, _vp]),
    "rc_peer_signal": (C.c_int, [_vp, _i32, _i32, _u64, _vp]),
    "rc_peer_wait": (C.c_int, [_vp, _i32, _u64, _f64, _vp, _vp]),
    "rc_ranks_workspace_bytes": (_sz, [_i64, _i64]),
    "rc_ranks": (C.c_int, [_vp, _i64, _i64, _vp, _vp, _sz, _vp]),
    "rc_clustered_ranks": (C.c_int, [_vp, _i64, _i64, _f64, _f64, _vp, _vp, _sz, _vp]),
    "rc_kendall_tau_b": (C.c_int, [_vp, _i64, _vp, _i64, _i64, _vp, _vp, _vp]),
    "rc_kendall_tau_b_batched": (C.c_int, [_vp, _vp, _i64, _i64, _i64, _i64, _vp, _vp, _vp]),
    "rc_kendall_large_workspace_bytes": (_sz, [_i64, _i64, _i64, _i64]),
    "rc_kendall_tau_b_large": (C.c_int, [_vp, _vp, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _sz, _vp]),
    "rc_rank_consistency_workspace_bytes": (_sz, [_i64, _i64, _i64, _i64]),
    "rc_rank_consistency": (C.c_int, [_vp, _i64, _i64, _i64, _i64, _f64, _vp, _vp, _vp, _vp, _sz, _vp]),
    "rc_robustness_sweep_host": (C.c_int, [_vp, _i64, _i32, _i32, _i32, _vp, _i32, _i64, _i32, _i32, _u64, _i64, _i64,
                                           _f64, _i32, _i64, _i64, _f64, _vp, _vp, _vp, _i32, _vp, _vp, _vp]),
    "rc_robustness_sweep_host_keep": (C.c_int, [_vp, _i64, _i32, _i32, _i32, _vp, _i32, _i64, _i32, _i32, _u64, _i64, _i64,
                                                _f64, _i32, _i64, _i64, _f64, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp]),
    "rc_robustness_sweep_workspace_bytes": (_sz, [_i64, _i32, _i64, _i32, _i64, _i64]),
    "rc_robustness_sweep": (C.c_int, [_vp, _i64, _i32, _i32, _i32, _vp, _i32, _i64, _i32, _i32, _u64, _i64, _i64, _f64, _i32,
                                      _i64, _i64, _f64, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _sz, _vp, _vp, _vp]),
    "rc_arim_bootstrap": (C.c_int, [_vp, _i64, _i64, _i32, _u64, _vp, _vp, _vp]),
    "rc_mc_sweep_host": (C.c_int, [_vp, _i64, _i32, _i32, _i32, _vp, _i32, _i64, _i32, _i32, _u64, _i64, _i64, _vp, _f64,
                                   _i32, _vp, _vp, _vp]),
    "rc_objective_host": (C.c_int, [_vp, _i32, _i32, _i32, _vp, _i64, _i32, _i32, _f64, _vp, _vp, _vp, _vp]),
    "rc_objective_release": (C.c_int, []),
    "rc_objective_call": (C.c_int, [_vp]),
    "rc_fidelity_grad": (C.c_int, [_vp, _i64, _i32, _i32, _i32, _vp, _i32, _vp, _vp, _vp, _vp]),
    "rc_fidelity_grad_host": (C.c_int, [_vp, _i64, _i32, _i32, _i32, _vp, _i32, _vp, _vp, _vp]),
    "rc_dense_fidelity_mc_workspace_bytes": (_sz, [_i32, _i64]),
    "rc_dense_fidelity_mc": (C.c_int, [_vp, _i64, _i32, _i32, _i32, _vp, _i32, _i64, _i32, _i32, _i32, _u64, _i64, _i64, _vp, _vp,
                                       _vp, _sz, _vp]),
    "rc_directional_fidelity_mc": (C.c_int, [_vp, _i64, _i32, _i32, _i32, _vp, _i32, _i64, _i32, _i32, _u64, _i64, _i64, _vp, _vp,
                                             _vp, _vp, _sz, _vp]),
    "rc_expm_batch": (C.c_int, [_vp, _i64, _i32, _vp, _vp]),
    "rc_fp64_peak_tflops": (C.c_int, [_vp, _vp]),
}

_lib = None


def declared_symbols() -> list[str]:
    """Every function declared in include/robchar_b200.h."""
    text = open(HEADER_PATH).read()
    return sorted(set(re.findall(r"\b(rc_[a-z0-9_]+)\s*\(", text)))


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RobcharLibraryError(
                f"{LIB_PATH} not found: build it with __graft_entry__.build() — robchar_b200 has no CPU fallback")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def last_error() -> str:
    return lib().rc_last_error().decode(errors="replace")


def check(code: int) -> None:
    """Map C status codes onto the exception types the reference raises."""
    if code == RC_OK:
        return
    msg = last_error()
    if code == RC_ERR_ILLEGAL_FIDS:
        raise AssertionError("illegal fids values - must be in [0,1]")  # wd_sortof_fast_implementation.py:25
    if code == RC_ERR_NONCONV:
        raise EigensolverNonConvergence(msg)
    if code in (RC_ERR_BAD_ARG, RC_ERR_NULL, RC_ERR_WORKSPACE):
        raise ValueError(msg)
    raise RobcharLibraryError(msg)
