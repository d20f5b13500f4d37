"""Functional interpolation API (stateless), mirroring
``torchkbnufft/functional/interp.py``: complex tensors, or real tensors whose last
dimension of size 2 holds (re, im) (``:109-124``, ``:150-172``)."""
from __future__ import annotations

from typing import Callable, List, Tuple

import torch
from torch import Tensor

from .._autograd.interp import KbTableInterpAdjoint, KbTableInterpForward

_SPMAT_MSG = (
    "sparse-matrix interpolation is outside the B200 engine's scope: only the table-interpolation path "
    "is accelerated (the reference itself labels the sparse mode 'not recommended', README.md:33-37). "
    "Pass interp_mats=None."
)


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
    raise NotImplementedError(_SPMAT_MSG)


def kb_spmat_interp_adjoint(data: Tensor, interp_mats: Tuple[Tensor, Tensor], grid_size: Tensor) -> Tensor:
    raise NotImplementedError(_SPMAT_MSG)
