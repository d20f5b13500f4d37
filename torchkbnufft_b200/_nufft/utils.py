"""Host-side one-time precompute for the B200 NUFFT engine (numpy/scipy, float64).

Rebuilds, from the published formulas, the buffers the reference derives in
``torchkbnufft/_nufft/utils.py``: the per-dimension Kaiser-Bessel lookup tables
(``build_table`` :122-179, via the 1-D column trick over ``build_numpy_spmatrix``
:16-119), the image-domain apodisation coefficients (``compute_scaling_coefs``
:212-254 / ``kaiser_bessel_ft`` :182-209), argument normalisation
(``validate_args`` :361-427) and tensor packaging (``init_fn`` :257-358).

This is setup, not the hot path: it runs once per module on the host and its
numbers are pinned against the reference's own buffers
(``tests/golden/ref_buffers.npz``, ``tests/test_precompute.py``).
"""
from __future__ import annotations

import functools
import itertools
from typing import List, NamedTuple, Optional, Sequence, Tuple, Union

import numpy as np
import torch
from scipy import special
from torch import Tensor

# complex <-> real dtype pairing (reference: DTYPE_MAP, _nufft/utils.py:10-13)
DTYPE_MAP = [
    (torch.complex128, torch.float64),
    (torch.complex64, torch.float32),
]


def paired_dtypes(dtype: torch.dtype) -> Tuple[torch.dtype, torch.dtype]:
    """Return ``(complex_dtype, real_dtype)`` for a real or complex dtype."""
    for cplx, real in DTYPE_MAP:
        if dtype in (cplx, real):
            return cplx, real
    raise TypeError("Unrecognized dtype.")


class Geometry(NamedTuple):
    im_size: Tuple[int, ...]
    grid_size: Tuple[int, ...]
    numpoints: Tuple[int, ...]
    n_shift: Tuple[int, ...]
    table_oversamp: Tuple[int, ...]
    order: Tuple[float, ...]
    alpha: Tuple[float, ...]
    dtype: torch.dtype
    device: torch.device


def _per_dim(value, ndim: int, scalar_type) -> tuple:
    if isinstance(value, scalar_type):
        return tuple(value for _ in range(ndim))
    return tuple(value)


def validate_args(
    im_size: Sequence[int],
    grid_size: Optional[Sequence[int]] = None,
    numpoints: Union[int, Sequence[int]] = 6,
    n_shift: Optional[Sequence[int]] = None,
    table_oversamp: Union[int, Sequence[int]] = 2**10,
    kbwidth: float = 2.34,
    order: Union[float, Sequence[float]] = 0.0,
    dtype: Optional[torch.dtype] = None,
    device: Optional[torch.device] = None,
) -> Geometry:
    """Fill in defaults (grid = 2N, n_shift = N//2, alpha = kbwidth*J) and check
    that every per-dimension argument has one entry per image dimension."""
    im_size = tuple(int(n) for n in im_size)
    ndim = len(im_size)
    grid_size = tuple(2 * n for n in im_size) if grid_size is None else tuple(int(k) for k in grid_size)
    numpoints = _per_dim(numpoints, len(grid_size), int)
    n_shift = tuple(n // 2 for n in im_size) if n_shift is None else tuple(n_shift)
    table_oversamp = _per_dim(table_oversamp, len(grid_size), int)
    alpha = tuple(kbwidth * j for j in numpoints)
    order = _per_dim(order, len(grid_size), float)
    if dtype is None:
        dtype = torch.get_default_dtype()
    if device is None:
        device = torch.device("cpu")
    for name, seq in (("grid_size", grid_size), ("n_shift", n_shift), ("numpoints", numpoints),
                      ("alpha", alpha), ("order", order), ("table_oversamp", table_oversamp)):
        assert len(seq) == ndim, f"{name} must have one entry per image dimension"
    return Geometry(im_size, grid_size, numpoints, n_shift, table_oversamp, order, alpha, dtype, torch.device(device))


@functools.lru_cache(maxsize=64)
def _kaiser_bessel_table_1d_cached(im_size: int, grid_size: int, numpoints: int, table_oversamp: int, order: float,
                                   alpha: float) -> np.ndarray:
    table = _kaiser_bessel_table_1d(im_size, grid_size, numpoints, table_oversamp, order, alpha)
    table.setflags(write=False)  # shared by every caller with these parameters
    return table


def kaiser_bessel_table_1d(im_size: int, grid_size: int, numpoints: int, table_oversamp: int, order: float,
                           alpha: float) -> np.ndarray:
    """One dimension's interpolation table (see ``_kaiser_bessel_table_1d``), memoised: the Bessel
    evaluations take milliseconds, and setup-side callers (density compensation, Toeplitz kernels,
    every module constructor) ask for the same few tables again and again.  Read-only array."""
    return _kaiser_bessel_table_1d_cached(int(im_size), int(grid_size), int(numpoints), int(table_oversamp),
                                          float(order), float(alpha))


def _kaiser_bessel_table_1d(im_size: int, grid_size: int, numpoints: int, table_oversamp: int, order: float,
                            alpha: float) -> np.ndarray:
    """One dimension's interpolation table, complex128, length ``J*L + 1``.

    Entry ``i`` (``i < J*L``) samples the kernel at ``x = i/L - J/2``:
    ``kb(x) * exp(-i*pi*(N-1)*x/K)`` with
    ``kb(x) = I_order(alpha*sqrt(1-(2x/J)^2)) / I_order(alpha)`` for ``|x| < J/2``
    and 0 otherwise; the final entry is 0.  The evaluation mirrors the floating
    point path of the reference's column trick (table row ``l`` of column ``c``
    is the coefficient of neighbour ``c+1`` at fractional position ``l/L``) so
    the numbers agree to the last bits.
    """
    J, L, K, N = int(numpoints), int(table_oversamp), int(grid_size), int(im_size)
    frac = J / 2 - 1 + np.arange(L) / L  # sample positions in grid units
    gam = 2 * np.pi / K
    q = (frac * 2 * np.pi / K) / gam
    koff = np.floor(q - J / 2)
    dist = q - koff
    table = np.zeros(J * L + 1, dtype=np.complex128)
    denom = special.iv(order, alpha)
    phase_scale = 1j * gam * (N - 1) / 2
    for col in range(J):
        neighbour = col - koff  # the neighbour number (1..J) that lands in this column
        x = -neighbour + dist
        inside = (np.abs(x) < J / 2) & (neighbour >= 1) & (neighbour <= J)
        kb = np.zeros(L)
        kb[inside] = np.real(special.iv(order, alpha * np.sqrt(1 - (x[inside] / (J / 2)) ** 2)) / denom)
        coef = np.conj(np.exp(phase_scale * x) * kb)
        coef[~inside] = 0
        start = (J - 1 - col) * L
        table[start:start + L] = coef
    return table


def build_table(im_size, grid_size, numpoints, table_oversamp, order, alpha) -> List[Tensor]:
    """Tables for every dimension as complex128 tensors (reference name kept)."""
    return [
        torch.tensor(kaiser_bessel_table_1d(n, k, j, l, o, a))  # a copy: the memoised array is read-only
        for n, k, j, l, o, a in zip(im_size, grid_size, numpoints, table_oversamp, order, alpha)
    ]


def kaiser_bessel_ft(omega: np.ndarray, numpoints: int, alpha: float, order: float, d: int) -> np.ndarray:
    """Fourier transform of the Kaiser-Bessel kernel (d-dimensional form),
    evaluated at normalised frequencies ``omega``."""
    z = np.sqrt((2 * np.pi * (numpoints / 2) * omega) ** 2 - alpha**2 + 0j)
    nu = d / 2 + order
    ft = ((2 * np.pi) ** (d / 2) * ((numpoints / 2) ** d) * (alpha**order) / special.iv(order, alpha)
          * special.jv(nu, z) / (z**nu))
    return np.real(ft)


def compute_scaling_coefs(im_size, grid_size, numpoints, alpha, order) -> Tensor:
    """Separable apodisation ``prod_d 1/FT{kb}((n_d-(N_d-1)/2)/K_d)`` as a
    float64 tensor of shape ``im_size`` (all ones along a dim with J=1)."""
    coef = None
    for n, k, j, a, o in zip(im_size, grid_size, numpoints, alpha, order):
        line = _scaling_line(int(n), int(k), int(j), float(a), float(o))
        coef = line if coef is None else coef[..., np.newaxis] * line
    return torch.from_numpy(np.array(coef))


@functools.lru_cache(maxsize=64)
def _scaling_line(n: int, k: int, j: int, alpha: float, order: float) -> np.ndarray:
    pos = np.arange(n) - (n - 1) / 2
    line = np.ones(n) if j == 1 else 1 / kaiser_bessel_ft(pos / k, j, alpha, order, 1)
    line.setflags(write=False)
    return line


class Precomputed(NamedTuple):
    tables: List[Tensor]
    im_size: Tensor
    grid_size: Tensor
    n_shift: Tensor
    numpoints: Tensor
    offsets: Tensor
    table_oversamp: Tensor
    order: Tensor
    alpha: Tensor


def init_fn(
    im_size: Sequence[int],
    grid_size: Optional[Sequence[int]] = None,
    numpoints: Union[int, Sequence[int]] = 6,
    n_shift: Optional[Sequence[int]] = None,
    table_oversamp: Union[int, Sequence[int]] = 2**10,
    kbwidth: float = 2.34,
    order: Union[float, Sequence[float]] = 0.0,
    dtype: Optional[torch.dtype] = None,
    device: Optional[torch.device] = None,
) -> Precomputed:
    """All per-module tensors, in the reference's order: tables, im_size,
    grid_size, n_shift, numpoints, offsets, table_oversamp, order, alpha.
    Sizes are int64, n_shift/order/alpha the real dtype, tables complex."""
    geo = validate_args(im_size, grid_size, numpoints, n_shift, table_oversamp, kbwidth, order, dtype, device)
    if not (geo.dtype.is_floating_point or geo.dtype.is_complex):
        raise TypeError("Unrecognized dtype.")
    complex_dtype, real_dtype = paired_dtypes(geo.dtype)
    tables = build_table(geo.im_size, geo.grid_size, geo.numpoints, geo.table_oversamp, geo.order, geo.alpha)
    assert len(tables) == len(geo.im_size)
    offsets = list(itertools.product(*[range(j) for j in geo.numpoints]))  # row-major neighbour offsets
    dev = geo.device

    def ints(v):
        return torch.tensor(v, dtype=torch.long, device=dev)

    def reals(v):
        return torch.tensor(v, dtype=real_dtype, device=dev)

    return Precomputed(
        [t.to(dtype=complex_dtype, device=dev) for t in tables],
        ints(geo.im_size), ints(geo.grid_size), reals(geo.n_shift), ints(geo.numpoints), ints(offsets),
        ints(geo.table_oversamp), reals(geo.order), reals(geo.alpha),
    )
