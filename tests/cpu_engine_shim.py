"""TEST-ONLY stand-in for the CUDA engine, used by the ``-m "not gpu"`` suite.

The product package is CUDA-only (no CPU path, by design).  To exercise its HOST
logic without a GPU -- argument validation, real/complex view handling, autograd
wiring, Toeplitz kernel assembly, density compensation, sharding -- this module
monkey-patches the five engine entry points with the CPU oracle
(``oracle/kbnufft_oracle.py``) for the duration of a ``with`` block.  It lives in
``tests/``; nothing in ``torchkbnufft_b200`` imports it, and the GPU tests never
use it.
"""
from __future__ import annotations

import contextlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import kbnufft_oracle as orc  # noqa: E402

import torchkbnufft_b200._autograd.interp as ag_interp  # noqa: E402
import torchkbnufft_b200._nufft.fft as eng_fft  # noqa: E402
import torchkbnufft_b200._nufft.interp as eng_interp  # noqa: E402
from torchkbnufft_b200 import _lib  # noqa: E402
from torchkbnufft_b200._nufft.plan import host_ints, normalize_omega  # noqa: E402


def _np(t):
    return t.detach().cpu().numpy()


def _table_interp(image, omega, tables, n_shift, numpoints, table_oversamp, offsets=None, min_kspace_per_fork=1024,
                  layout=_lib.COIL_MAJOR, n_coils=None):
    assert layout == _lib.COIL_MAJOR
    omega = normalize_omega(omega, image.shape[0], "image")
    out = orc.table_interp(_np(image), _np(omega), [_np(t) for t in tables], _np(n_shift), _np(numpoints),
                           _np(table_oversamp))
    return torch.from_numpy(out)


def _table_interp_adjoint(data, omega, tables, n_shift, numpoints, table_oversamp, offsets, grid_size,
                          layout=_lib.COIL_MAJOR, mode=None):
    assert layout == _lib.COIL_MAJOR
    omega = normalize_omega(omega, data.shape[0], "data")
    out = orc.table_interp_adjoint(_np(data), _np(omega), [_np(t) for t in tables], _np(n_shift), _np(numpoints),
                                   _np(table_oversamp), host_ints(grid_size))
    return torch.from_numpy(out)


def _apod_pad(image, grid_size, smaps=None, scaling_coef=None, scale=1.0, n_coils=None, layout=_lib.COIL_MAJOR):
    x = image
    if smaps is not None:
        x = x * smaps
    if scaling_coef is not None:
        x = x * scaling_coef
    x = x * scale
    pad = []
    for g, n in zip(reversed(host_ints(grid_size)), reversed(image.shape[2:])):
        pad += [0, g - n]
    return torch.nn.functional.pad(x, pad).contiguous()


def _crop_apod_coilsum(grid, im_size, smaps=None, scaling_coef=None, scale=1.0, layout=_lib.COIL_MAJOR):
    sl = (slice(None), slice(None)) + tuple(slice(0, n) for n in host_ints(im_size))
    x = grid[sl]
    if scaling_coef is not None:
        x = x * scaling_coef.conj()
    if smaps is not None:
        x = torch.sum(x * smaps.conj(), dim=1, keepdim=True)
    return (x * scale).contiguous()


def _spectrum_mul_(spectrum, kernel, scale=1.0, layout=_lib.COIL_MAJOR):
    ndim = spectrum.ndim - 2
    k = kernel if kernel.ndim == ndim else kernel.unsqueeze(1)
    spectrum.mul_(k * scale)
    return spectrum


@contextlib.contextmanager
def oracle_engine():
    """Patch the engine entry points with oracle-backed CPU stand-ins."""
    saved = [
        (eng_interp, "table_interp", eng_interp.table_interp),
        (eng_interp, "table_interp_adjoint", eng_interp.table_interp_adjoint),
        (ag_interp, "table_interp", ag_interp.table_interp),
        (ag_interp, "table_interp_adjoint", ag_interp.table_interp_adjoint),
        (eng_fft, "apod_pad", eng_fft.apod_pad),
        (eng_fft, "crop_apod_coilsum", eng_fft.crop_apod_coilsum),
        (eng_fft, "spectrum_mul_", eng_fft.spectrum_mul_),
        (eng_fft, "fused_fft_available", eng_fft.fused_fft_available),
    ]
    try:
        for mod in (eng_interp, ag_interp):
            mod.table_interp = _table_interp
            mod.table_interp_adjoint = _table_interp_adjoint
        eng_fft.apod_pad = _apod_pad
        eng_fft.crop_apod_coilsum = _crop_apod_coilsum
        eng_fft.spectrum_mul_ = _spectrum_mul_
        eng_fft.fused_fft_available = lambda dtype, grid_size, n_rows=1: False  # CPU stand-in uses torch.fft
        yield
    finally:
        for mod, name, fn in saved:
            setattr(mod, name, fn)
