"""ctypes binding of ``libb200nufft.so`` (the C ABI declared in ``include/b200nufft.h``).

The engine is CUDA-only by design: there is no CPU implementation and no silent
fallback.  If the shared library is missing, or a tensor lives on the CPU, the
calls below raise -- they never route around the kernels.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_int, c_int32, c_int64, c_size_t, c_void_p
from typing import Optional

MAX_DIMS = 3
MAX_NUMPOINTS = 16
C64, C128 = 0, 1
COIL_MAJOR, CHANNEL_LAST = 0, 1
ADJ_ATOMIC, ADJ_SORTED = 0, 1
ABI_VERSION = 4
OPT_TILED_KERNELS = 0
OPT_ADJ_ROW_OWNERSHIP = 1
OPT_FWD_COIL_CHUNK = 2
OPT_ADJ_COIL_CHUNK = 3
OPT_FAST_FFT = 4
OPT_PDL = 5
OPT_FFT_PREFETCH = 6
OPT_ADJ_OWNED = 7
OPT_OWN_CAP = 8
OPT_FFT_STREAM = 9
OPT_PEER_FORM = 10

_CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
# B2N_LIB_PATH: an alternative build of the same sources (A/B of compile-time kernel configurations, profiles/scripts)
LIB_PATH = os.environ.get("B2N_LIB_PATH") or os.path.join(_CSRC, "libb200nufft.so")


class Geom(Structure):
    """struct b2n_geom"""

    _fields_ = [
        ("ndim", c_int32),
        ("dtype", c_int32),
        ("grid_size", c_int64 * MAX_DIMS),
        ("numpoints", c_int32 * MAX_DIMS),
        ("table_oversamp", c_int32 * MAX_DIMS),
        ("table_len", c_int64 * MAX_DIMS),
        ("table_dev", c_void_p * MAX_DIMS),
        ("n_shift", c_double * MAX_DIMS),
        ("rtable_dev", c_void_p * MAX_DIMS),
        ("table_phase", c_double * MAX_DIMS),
    ]


class Points(Structure):
    """struct b2n_points"""

    _fields_ = [
        ("n_points", c_int64),
        ("n_traj", c_int64),
        ("ndim", c_int32),
        ("dtype", c_int32),
        ("coef_stride", c_int32),
        ("sub_cap", c_int32),
        ("tile", c_int32 * MAX_DIMS),
        ("n_tiles", c_int32 * MAX_DIMS),
        ("n_cells", c_int64),
        ("n_sub_max", c_int64),
        ("perm", c_void_p),
        ("inv_perm", c_void_p),
        ("base", c_void_p),
        ("coef", c_void_p),
        ("phase", c_void_p),
        ("cell_start", c_void_p),
        ("keys", c_void_p),
        ("sub_tile", c_void_p),
        ("sub_start", c_void_p),
        ("sub_count", c_void_p),
        ("n_sub", c_void_p),
        ("sub_slot", c_void_p),
        ("tile_sub_start", c_void_p),
        ("own_tile", c_int32),
        ("own_cap", c_int32),
        ("n_own_tiles", c_int32 * 3),
        ("own_pad_", c_int32),
        ("n_own_items_max", c_int64),
        ("own_visits", c_void_p),
        ("own_items", c_void_p),
        ("own_tiles", c_void_p),
        ("own_counts", c_void_p),
        ("own_hw", c_void_p),
        ("own_fac", c_void_p),
        ("own_q", c_void_p),
        ("own_exc", c_void_p),
        ("n_own_exc_max", c_int64),
        ("own_xt", c_void_p),
        ("own_xv", c_void_p),
        ("n_own_xv_max", c_int64),
    ]


PEER_MAX_RANKS = 16
PEER_HANDLE_BYTES = 64


class PeerComm(Structure):
    """struct b2n_peer_comm"""

    _fields_ = [
        ("rank", c_int32),
        ("world", c_int32),
        ("max_floats", c_int64),
        ("window", c_void_p * PEER_MAX_RANKS),
    ]


class EngineError(RuntimeError):
    """A libb200nufft call returned a non-zero status."""


# every symbol include/b200nufft.h declares: (restype, argtypes)
_I64P = POINTER(c_int64)
SIGNATURES = {
    "b2n_abi_version": (c_int, []),
    "b2n_struct_sizes": (c_int, [POINTER(c_size_t), POINTER(c_size_t)]),
    "b2n_launch_count": (ctypes.c_longlong, []),
    "b2n_last_error": (c_char_p, []),
    "b2n_device_count": (c_int, []),
    "b2n_set_option": (c_int, [c_int, c_int]),
    "b2n_get_option": (c_int, [c_int]),
    "b2n_set_trace_buffer": (c_int, [c_void_p, c_int64]),
    "b2n_points_workspace_bytes": (c_int, [POINTER(Geom), c_int64, c_int64, POINTER(c_size_t)]),
    "b2n_points_build": (c_int, [POINTER(Geom), c_void_p, c_int64, c_int64, c_void_p, c_size_t, POINTER(Points),
                                 c_void_p]),
    "b2n_export_indices": (c_int, [POINTER(Geom), c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "b2n_interp_forward": (c_int, [POINTER(Geom), POINTER(Points), c_void_p, c_int64, c_int64, c_int, c_void_p,
                                   c_void_p]),
    "b2n_interp_adjoint": (c_int, [POINTER(Geom), POINTER(Points), c_void_p, c_int64, c_int64, c_int, c_int,
                                   c_void_p, c_void_p]),
    "b2n_interp_adjoint_ordered_bytes": (c_int, [POINTER(Geom), POINTER(Points), c_int64, c_int64, c_int,
                                                 POINTER(c_size_t)]),
    "b2n_interp_adjoint_ordered_layout": (c_int, [POINTER(Geom), POINTER(Points), c_int64, c_int64, c_int, c_int64,
                                                  POINTER(c_size_t), POINTER(c_size_t)]),
    "b2n_interp_adjoint_ordered": (c_int, [POINTER(Geom), POINTER(Points), c_void_p, c_int64, c_int64, c_int, c_void_p,
                                           c_size_t, c_void_p, c_void_p]),
    "b2n_apod_pad": (c_int, [c_int, c_int, _I64P, _I64P, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_int64,
                             c_void_p, c_double, c_int, c_void_p, c_void_p]),
    "b2n_crop_apod_coilsum": (c_int, [c_int, c_int, _I64P, _I64P, c_int64, c_int64, c_void_p, c_int, c_void_p,
                                      c_int64, c_void_p, c_double, c_void_p, c_void_p]),
    "b2n_spectrum_mul": (c_int, [c_int, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int, c_double,
                                 c_void_p]),
    "b2n_fft_supported": (c_int, [c_int64]),
    "b2n_fft_work_bytes": (c_int, [c_int, _I64P, _I64P, c_int64, c_int64, POINTER(c_size_t)]),
    "b2n_fft_twiddles": (c_int, [c_int64, c_void_p, c_void_p]),
    "b2n_fft_forward_fused": (c_int, [c_int, _I64P, _I64P, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_int64,
                                      c_void_p, c_double, POINTER(c_void_p), c_void_p, c_void_p, c_void_p]),
    "b2n_fft_adjoint_fused": (c_int, [c_int, _I64P, _I64P, c_int64, c_int64, c_void_p, c_void_p, c_int64, c_void_p,
                                      c_int64, c_void_p, c_double, POINTER(c_void_p), c_void_p, c_void_p, c_void_p]),
    "b2n_fft_toeplitz_fused": (c_int, [c_int, _I64P, _I64P, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_int64,
                                       c_void_p, c_int64, c_double, POINTER(c_void_p), c_void_p, c_void_p, c_void_p]),
    "b2n_peer_window_bytes": (c_int, [c_int, c_int64, POINTER(c_size_t)]),
    "b2n_peer_window_create": (c_int, [c_size_t, POINTER(c_void_p), c_void_p]),
    "b2n_peer_window_open": (c_int, [c_void_p, POINTER(c_void_p)]),
    "b2n_peer_window_close": (c_int, [c_void_p]),
    "b2n_peer_window_destroy": (c_int, [c_void_p]),
    "b2n_peer_allreduce_sum": (c_int, [POINTER(PeerComm), c_void_p, c_void_p, c_int64, c_void_p]),
    "b2n_fft_adjoint_fused_allreduce": (c_int, [c_int, _I64P, _I64P, c_int64, c_int64, c_void_p, c_void_p, c_int64,
                                                c_void_p, c_int64, c_void_p, c_double, POINTER(c_void_p), c_void_p,
                                                c_void_p, POINTER(PeerComm), c_void_p]),
}

_lib: Optional[ctypes.CDLL] = None


def load() -> ctypes.CDLL:
    """Load the engine; raise if it has not been built (no fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineError(
            f"{LIB_PATH} not found. The B200 NUFFT engine has no CPU/PyTorch fallback: build it with "
            "`python -m torchkbnufft_b200._build` (needs nvcc) before calling any operator."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library lacks a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.b2n_abi_version() != ABI_VERSION:
        raise EngineError(f"libb200nufft ABI {lib.b2n_abi_version()} != expected {ABI_VERSION}; rebuild")
    gsz, psz = c_size_t(0), c_size_t(0)
    lib.b2n_struct_sizes(ctypes.byref(gsz), ctypes.byref(psz))
    if gsz.value != ctypes.sizeof(Geom) or psz.value != ctypes.sizeof(Points):
        raise EngineError(f"struct mirrors out of date: b2n_geom {gsz.value} vs {ctypes.sizeof(Geom)} bytes, "
                          f"b2n_points {psz.value} vs {ctypes.sizeof(Points)} bytes")
    _lib = lib
    return lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = load().b2n_last_error()
        raise EngineError(f"{what} failed with status {status}: {msg.decode() if msg else '?'}")


def i64_array(values) -> ctypes.Array:
    vals = [int(v) for v in values]
    return (c_int64 * len(vals))(*vals)
