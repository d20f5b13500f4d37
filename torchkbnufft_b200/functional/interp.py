"""Functional interpolation API (stateless), mirroring
``torchkbnufft/functional/interp.py``: complex tensors, or real tensors whose last
dimension of size 2 holds (re, im) (``:109-124``, ``:150-172``)."""
from __future__ import annotations

from typing import Callable, List, Tuple

import torch
from torch import Tensor

from .._autograd.interp import KbTableInterpAdjoint, KbTableInterpForward
from .._nufft import spmat as _spmat
from .._nufft.plan import require_cuda


def with_complex_view(fn: Callable[[Tensor], Tensor], x: Tensor) -> Tensor:
    """Run ``fn`` on a complex view of ``x`` and give the result back in the caller's
    convention (complex stays complex, real ``(..., 2)`` stays real)."""
    if x.is_complex():
        return fn(x)
    if not x.shape[-1] == 2:
        raise ValueError("For real inputs, last dimension must be size 2.")
    return torch.view_as_real(fn(torch.view_as_complex(x)))


def kb_table_interp(image: Tensor, omega: Tensor, tables: List[Tensor], n_shift: Tensor, numpoints: Tensor,
                    table_oversamp: Tensor, offsets: Tensor) -> Tensor:
    """Kaiser-Bessel table interpolation: gridded ``image (B, C, *K)`` -> samples
    ``(B, C, M)`` at ``omega`` (radians/voxel)."""
    return with_complex_view(
        lambda x: KbTableInterpForward.apply(x, omega, tables, n_shift, numpoints, table_oversamp, offsets), image
    )


def kb_table_interp_adjoint(data: Tensor, omega: Tensor, tables: List[Tensor], n_shift: Tensor, numpoints: Tensor,
                            table_oversamp: Tensor, offsets: Tensor, grid_size: Tensor) -> Tensor:
    """Adjoint Kaiser-Bessel table interpolation: samples ``(B, C, M)`` -> grid
    ``(B, C, *grid_size)``."""
    return with_complex_view(
        lambda x: KbTableInterpAdjoint.apply(x, omega, tables, n_shift, numpoints, table_oversamp, offsets, grid_size),
        data,
    )


def kb_spmat_interp(image: Tensor, interp_mats: Tuple[Tensor, Tensor]) -> Tensor:
    """Sparse-matrix interpolation (reference ``functional/interp.py:14-44``): API-completeness shim
    on ``torch.sparse``, not part of the accelerated table path (see ``_nufft/spmat.py``)."""
    require_cuda(image, "image")
    return with_complex_view(lambda x: _spmat.spmat_interp(x, interp_mats), image)


def kb_spmat_interp_adjoint(data: Tensor, interp_mats: Tuple[Tensor, Tensor], grid_size: Tensor) -> Tensor:
    """Adjoint sparse-matrix interpolation (reference ``functional/interp.py:47-79``)."""
    require_cuda(data, "data")
    return with_complex_view(lambda x: _spmat.spmat_interp_adjoint(x, interp_mats, grid_size), data)
