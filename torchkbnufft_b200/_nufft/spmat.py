"""Sparse-matrix interpolation mode: API-completeness shim.

The reference offers precomputed sparse interpolation matrices as an alternative to table
interpolation (``torchkbnufft/_nufft/spmat.py:10-105``, builder ``_nufft/utils.py:16-120``, apply
``_nufft/interp.py:13-85``) and labels that mode slow / not recommended (``README.md:33-37``).  It
is NOT part of the accelerated path: the matrices are built on the host with numpy/scipy exactly as
the reference does (direct Kaiser-Bessel evaluation, no table), and applied with ``torch.sparse``
(cuSPARSE) -- no kernel of ``libb200nufft.so`` is involved.  It exists so that code passing
``interp_mats=`` keeps working; use the default table path for speed.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple, Union

import numpy as np
import torch
from scipy import special
from torch import Tensor

from .utils import validate_args


def _dimension_terms(om: np.ndarray, im_size: int, grid_size: int, numpoints: int, alpha: float,
                     order: float) -> Tuple[np.ndarray, np.ndarray]:
    """Neighbour grid columns ``(M, J)`` (wrapped into ``[0, K)``) and their complex coefficients
    ``kb(x) * exp(i * gam * (N-1)/2 * x)`` for one dimension, ``x`` = signed distance from the sample to
    neighbour ``j = 1..J`` in grid units (reference ``interp_coeff`` / ``kd``, utils.py:45-85)."""
    J, K = int(numpoints), int(grid_size)
    gam = 2 * np.pi / K
    u = om / gam
    first = np.floor(u - J / 2)
    j = np.arange(1, J + 1)
    x = (u - first)[:, None] - j[None, :]
    inside = np.abs(x) < J / 2
    kb = np.zeros(x.shape, dtype=np.float64)
    kb[inside] = np.real(special.iv(order, alpha * np.sqrt(1 - (x[inside] / (J / 2)) ** 2)) / special.iv(order, alpha))
    coef = kb * np.exp(1j * gam * (im_size - 1) / 2 * x)
    cols = np.mod(j[None, :] + first[:, None], K).astype(np.int64)
    return cols, coef


def build_interp_coo(omega: np.ndarray, im_size, grid_size, numpoints, n_shift, order, alpha):
    """``(rows, cols, values, shape)`` of the ``M x prod(grid_size)`` interpolation matrix:
    entry (m, k) = conj(prod_d coef_d) * exp(i * omega_m . n_shift) for every neighbour k of sample m."""
    ndim, M = omega.shape
    cols = np.zeros((M, 1), dtype=np.int64)
    vals = np.ones((M, 1), dtype=np.complex128)
    for d in range(ndim):
        c, v = _dimension_terms(omega[d].astype(np.float64), im_size[d], grid_size[d], numpoints[d], alpha[d], order[d])
        stride = int(np.prod(grid_size[d + 1:]))
        cols = (cols[:, :, None] + (c * stride)[:, None, :]).reshape(M, -1)   # row-major neighbourhood
        vals = (vals[:, :, None] * v[:, None, :]).reshape(M, -1)
    shift_phase = np.exp(1j * (omega.astype(np.float64).T @ np.asarray(n_shift, dtype=np.float64)))
    vals = np.conj(vals) * shift_phase[:, None]
    rows = np.repeat(np.arange(M, dtype=np.int64), cols.shape[1])
    return rows, cols.reshape(-1), vals.reshape(-1), (M, int(np.prod(grid_size)))


def calc_tensor_spmatrix(
    omega: Tensor,
    im_size: Sequence[int],
    grid_size: Optional[Sequence[int]] = None,
    numpoints: Union[int, Sequence[int]] = 6,
    n_shift: Optional[Sequence[int]] = None,
    table_oversamp: Union[int, Sequence[int]] = 2**10,
    kbwidth: float = 2.34,
    order: Union[float, Sequence[float]] = 0.0,
) -> Tuple[Tensor, Tensor]:
    """(real, imaginary) sparse COO interpolation matrices for ``omega (d, M)``, on omega's device
    and in omega's dtype -- same signature and result as the reference's ``calc_tensor_spmatrix``."""
    if not omega.ndim == 2:
        raise ValueError("Sparse matrix calculation not implemented for batched omega.")
    geo = validate_args(im_size, grid_size, numpoints, n_shift, table_oversamp, kbwidth, order, omega.dtype,
                        omega.device)
    rows, cols, vals, shape = build_interp_coo(omega.detach().cpu().numpy(), geo.im_size, geo.grid_size,
                                               geo.numpoints, geo.n_shift, geo.order, geo.alpha)
    index = torch.from_numpy(np.stack((rows, cols))).to(geo.device)
    real = torch.from_numpy(np.ascontiguousarray(vals.real)).to(device=geo.device, dtype=geo.dtype)
    imag = torch.from_numpy(np.ascontiguousarray(vals.imag)).to(device=geo.device, dtype=geo.dtype)
    return (torch.sparse_coo_tensor(index, real, torch.Size(shape), check_invariants=False),
            torch.sparse_coo_tensor(index, imag, torch.Size(shape), check_invariants=False))


def _check_mats(interp_mats) -> Tuple[Tensor, Tensor]:
    if not isinstance(interp_mats, tuple):
        raise TypeError("interp_mats must be 2-tuple of (real_mat, imag_mat.")
    return interp_mats


def spmat_interp(image: Tensor, interp_mats: Tuple[Tensor, Tensor]) -> Tensor:
    """``(B, C, *K)`` complex grid -> ``(B, C, M)``: ``y = (A_r + i A_i) x`` with real sparse products
    (``torch.mm(sparse, dense)``; differentiable with respect to ``image``)."""
    a_r, a_i = _check_mats(interp_mats)
    B, C = image.shape[:2]
    x = torch.view_as_real(image.reshape(B * C, -1))
    x_r, x_i = x[..., 0].t().contiguous(), x[..., 1].t().contiguous()
    y_r = torch.mm(a_r, x_r) - torch.mm(a_i, x_i)
    y_i = torch.mm(a_r, x_i) + torch.mm(a_i, x_r)
    return torch.complex(y_r.t(), y_i.t()).reshape(B, C, -1)


def spmat_interp_adjoint(data: Tensor, interp_mats: Tuple[Tensor, Tensor], grid_size) -> Tensor:
    """``(B, C, M)`` samples -> ``(B, C, *grid_size)``: ``x = (A_r + i A_i)^H y``."""
    a_r, a_i = _check_mats(interp_mats)
    B, C = data.shape[:2]
    sizes = [int(k) for k in (grid_size.tolist() if isinstance(grid_size, Tensor) else grid_size)]
    y = torch.view_as_real(data.reshape(B * C, -1))
    y_r, y_i = y[..., 0].t().contiguous(), y[..., 1].t().contiguous()
    at_r, at_i = a_r.t(), a_i.t()
    x_r = torch.mm(at_r, y_r) + torch.mm(at_i, y_i)
    x_i = torch.mm(at_r, y_i) - torch.mm(at_i, y_r)
    return torch.complex(x_r.t(), x_i.t()).reshape([B, C] + sizes)
