"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.

numpy / ctypes front end of ``kbnufft_oracle.c``: a CPU restatement of the
reference's table-interpolation NUFFT path (mmuckley/torchkbnufft).  It is the
checker for the CUDA engine in ``torchkbnufft_b200`` and the CPU baseline that
``bench.py`` times.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py`` (``cpu_baseline`` / ``--impl reference``) may import this module;
the product package never does (``tests/test_boundary.py`` greps for it).

Parity status: PINNED against the reference's golden vectors and against
outputs of the reference itself (``tests/test_oracle.py``, fixtures made by
``oracle/make_golden.py``).

Each function cites the reference file:line it restates (paths relative to the
reference checkout, ``torchkbnufft/...``).  Array conventions follow the
reference: grids ``(B, C, *K)``, k-space data ``(B, C, M)``, trajectories
``(d, M)`` or ``(B, d, M)`` in radians/voxel, numpy complex64 / complex128.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import List, Optional, Sequence, Tuple

import numpy as np

try:  # scipy's pocketfft runs single precision natively and takes `workers`
    import scipy.fft as _fft

    _HAVE_SCIPY_FFT = True
except Exception:  # pragma: no cover - scipy is in the image
    import numpy.fft as _fft

    _HAVE_SCIPY_FFT = False

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libkbnufft_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile ``kbnufft_oracle.c`` with the committed Makefile."""
    if force or not os.path.exists(_LIB_PATH):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True)
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        assert _lib.orc_abi_version() == 1
    return _lib


def _i64(seq) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(seq, dtype=np.int64).reshape(-1))


def _suffix(real_dtype) -> str:
    return "f32" if np.dtype(real_dtype) == np.float32 else "f64"


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def _prep_tables(tables: Sequence[np.ndarray], cdtype):
    tabs = [np.ascontiguousarray(np.asarray(t).astype(cdtype, copy=False)) for t in tables]
    arr = (ctypes.c_void_p * len(tabs))(*[t.ctypes.data for t in tabs])
    lens = _i64([t.shape[0] for t in tabs])
    return tabs, arr, lens


def _prep_omega(omega: np.ndarray, real_dtype, nbatch: int) -> Tuple[np.ndarray, int]:
    """omega validation of table_interp / table_interp_adjoint
    (_nufft/interp.py:345-356 and :621-633)."""
    omega = np.asarray(omega)
    if omega.ndim not in (2, 3):
        raise ValueError("omega must have 2 or 3 dimensions.")
    if omega.ndim == 3 and omega.shape[0] == 1:
        omega = omega[0]
    if omega.ndim == 3 and omega.shape[0] != nbatch:
        raise ValueError("If omega has batch dim, omega batch dimension must match.")
    n_traj = omega.shape[0] if omega.ndim == 3 else 1
    return np.ascontiguousarray(omega.astype(real_dtype, copy=False)), n_traj


def calc_coef_and_indices(omega, grid_size, numpoints, table_oversamp):
    """Integer indices of calc_coef_and_indices (_nufft/interp.py:89-150) for
    every neighbour offset (row-major offsets, _nufft/utils.py:329).

    Returns ``arr_ind[W, M]`` (flat wrapped grid index) and ``tab_idx[W, d, M]``
    (table index incl. centre), both int64.
    """
    omega = np.ascontiguousarray(omega)
    assert omega.ndim == 2
    d, M = omega.shape
    K, J, L = _i64(grid_size), _i64(numpoints), _i64(table_oversamp)
    W = int(np.prod(J))
    arr_ind = np.empty((W, M), dtype=np.int64)
    tab_idx = np.empty((W, d, M), dtype=np.int64)
    fn = getattr(_load(), "orc_indices_" + _suffix(omega.dtype))
    fn(ctypes.c_int(d), ctypes.c_int64(M), _ptr(omega), _ptr(K), _ptr(J), _ptr(L), _ptr(arr_ind), _ptr(tab_idx))
    return arr_ind, tab_idx


def table_interp(image, omega, tables, n_shift, numpoints, table_oversamp, nthreads: int = 1):
    """Forward table interpolation; reference table_interp
    (_nufft/interp.py:315-403 -> table_interp_one_batch :154-203)."""
    image = np.asarray(image)
    cdtype = image.dtype
    rdtype = np.float32 if cdtype == np.complex64 else np.float64
    B, C = image.shape[:2]
    K = _i64(image.shape[2:])
    d = K.size
    omega, n_traj = _prep_omega(omega, rdtype, B)
    M = omega.shape[-1]
    J, L = _i64(numpoints), _i64(table_oversamp)
    tabs, tab_arr, tab_len = _prep_tables(tables, cdtype)
    ns = np.ascontiguousarray(np.asarray(n_shift, dtype=rdtype))
    img = np.ascontiguousarray(image).reshape(B, C, -1)
    out = np.empty((B, C, M), dtype=cdtype)
    fn = getattr(_load(), "orc_interp_forward_" + _suffix(rdtype))
    fn(ctypes.c_int(d), _ptr(K), _ptr(J), _ptr(L), _ptr(ns), tab_arr, _ptr(tab_len), _ptr(omega),
       ctypes.c_int64(n_traj), ctypes.c_int64(M), _ptr(img), ctypes.c_int64(B), ctypes.c_int64(C),
       _ptr(out), ctypes.c_int(nthreads))
    return out


def table_interp_adjoint(data, omega, tables, n_shift, numpoints, table_oversamp, grid_size, nthreads: int = 1):
    """Adjoint table interpolation; reference table_interp_adjoint
    (_nufft/interp.py:587-726), CPU summation order."""
    data = np.asarray(data)
    cdtype = data.dtype
    rdtype = np.float32 if cdtype == np.complex64 else np.float64
    B, C, M = data.shape
    K = _i64(grid_size)
    d = K.size
    omega, n_traj = _prep_omega(omega, rdtype, B)
    assert omega.shape[-1] == M
    J, L = _i64(numpoints), _i64(table_oversamp)
    tabs, tab_arr, tab_len = _prep_tables(tables, cdtype)
    ns = np.ascontiguousarray(np.asarray(n_shift, dtype=rdtype))
    dat = np.ascontiguousarray(data)
    out = np.empty((B, C) + tuple(int(k) for k in K), dtype=cdtype)
    fn = getattr(_load(), "orc_interp_adjoint_" + _suffix(rdtype))
    fn(ctypes.c_int(d), _ptr(K), _ptr(J), _ptr(L), _ptr(ns), tab_arr, _ptr(tab_len), _ptr(omega),
       ctypes.c_int64(n_traj), ctypes.c_int64(M), _ptr(dat), ctypes.c_int64(B), ctypes.c_int64(C),
       _ptr(out), ctypes.c_int(nthreads))
    return out


# ---------------------------------------------------------------------------
# FFT / apodisation / SENSE / Toeplitz steps (numpy + pocketfft)
# ---------------------------------------------------------------------------
def _fftn(x, axes, inverse: bool, norm: Optional[str], workers: int):
    """fft_fn / ifft_fn (_nufft/fft.py:9-22): forward is unscaled or 'ortho';
    the inverse uses norm='forward' (i.e. UNSCALED) unless 'ortho'."""
    if norm not in (None, "ortho"):
        raise ValueError("Only option for norm is 'ortho'.")
    kw = {"workers": workers} if _HAVE_SCIPY_FFT else {}
    if inverse:
        return _fft.ifftn(x, axes=axes, norm="ortho" if norm == "ortho" else "forward", **kw)
    return _fft.fftn(x, axes=axes, norm="ortho" if norm == "ortho" else "backward", **kw)


def fft_and_scale(image, scaling_coef, im_size, grid_size, norm=None, workers: int = 1):
    """x * scaling_coef -> zero-pad at the END of each dim -> FFT
    (_nufft/fft.py:36-76)."""
    d = len(grid_size)
    x = image * scaling_coef
    pad = [(0, 0), (0, 0)] + [(0, int(g) - int(n)) for g, n in zip(grid_size, im_size)]
    x = np.pad(x, pad)
    return _fftn(x, tuple(range(-d, 0)), False, norm, workers).astype(image.dtype, copy=False)


def ifft_and_scale(grid, scaling_coef, im_size, grid_size, norm=None, workers: int = 1):
    """IFFT (unscaled or ortho) -> keep the first N_d entries of each dim ->
    * conj(scaling_coef)   (_nufft/fft.py:80-118, crop_dims :25-32)."""
    d = len(grid_size)
    x = _fftn(grid, tuple(range(-d, 0)), True, norm, workers)
    sl = (slice(None), slice(None)) + tuple(slice(0, int(n)) for n in im_size)
    return (x[sl] * np.conj(scaling_coef)).astype(grid.dtype, copy=False)


def fft_filter(image, kernel, norm: Optional[str] = "ortho", workers: int = 1):
    """crop(IFFT(kernel * FFT(pad(image))))   (_nufft/fft.py:121-173)."""
    d = image.ndim - 2
    im_size = image.shape[2:]
    grid_size = kernel.shape[-d:]
    pad = [(0, 0), (0, 0)] + [(0, int(g) - int(n)) for g, n in zip(grid_size, im_size)]
    x = _fftn(np.pad(image, pad), tuple(range(-d, 0)), False, norm, workers)
    x = _fftn(x * kernel, tuple(range(-d, 0)), True, norm, workers)
    sl = (slice(None), slice(None)) + tuple(slice(0, int(n)) for n in im_size)
    return x[sl].astype(image.dtype, copy=False)


def nufft_forward(image, omega, tables, n_shift, numpoints, table_oversamp, scaling_coef, im_size,
                  grid_size, smaps=None, norm=None, nthreads: int = 1):
    """KbNufft.forward (modules/kbnufft.py:125-228): image*smaps ->
    kb_table_nufft (functional/nufft.py:126-191)."""
    x = image if smaps is None else image * smaps
    g = fft_and_scale(x, scaling_coef, im_size, grid_size, norm, workers=nthreads)
    return table_interp(g, omega, tables, n_shift, numpoints, table_oversamp, nthreads)


def nufft_adjoint(data, omega, tables, n_shift, numpoints, table_oversamp, scaling_coef, im_size,
                  grid_size, smaps=None, norm=None, nthreads: int = 1):
    """KbNufftAdjoint.forward (modules/kbnufft.py:307-410):
    kb_table_nufft_adjoint (functional/nufft.py:194-260) -> sum(x*conj(smaps), dim=1)."""
    g = table_interp_adjoint(data, omega, tables, n_shift, numpoints, table_oversamp, grid_size, nthreads)
    x = ifft_and_scale(g, scaling_coef, im_size, grid_size, norm, workers=nthreads)
    if smaps is not None:
        x = np.sum(x * np.conj(smaps), axis=1, keepdims=True).astype(data.dtype, copy=False)
    return x


def toep_nufft(image, kernel, smaps=None, norm=None, workers: int = 1):
    """ToepNufft.forward (modules/kbnufft.py:486-547) incl. toep_batch_loop
    (:441-484); kernel is ``(*2N)`` or ``(B, *2N)``."""
    d = image.ndim - 2
    if kernel.ndim > d and kernel.shape[0] == 1:
        kernel = kernel[0]
    if kernel.ndim > d and kernel.shape[0] != image.shape[0]:
        raise ValueError("If using batch dimension, kernel must have same batch size as image")
    if smaps is None:
        if kernel.ndim > d:  # (B, *2N) broadcasts against (B, C, *2N) only via a coil axis
            kernel = kernel[:, None]
        return fft_filter(image, kernel, norm, workers)
    out = []
    for b in range(image.shape[0]):
        s = smaps[b if smaps.shape[0] > 1 else 0][None]
        k = kernel[b] if kernel.ndim > d else kernel
        y = fft_filter(image[b][None] * s, k, norm, workers)
        out.append(np.sum(y * np.conj(s), axis=1, keepdims=True)[0])
    return np.stack(out).astype(image.dtype, copy=False)
