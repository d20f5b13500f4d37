"""Density compensation by the iterative method of Pipe & Menon (1999), with the
contract of the reference's ``calc_density_compensation_function``
(``torchkbnufft/_nufft/dcomp.py:10-119``): repeatedly push the current weights
through interpolation-adjoint then interpolation and divide by the magnitude."""
from __future__ import annotations

from typing import Optional, Sequence, Union

import torch
from torch import Tensor

from . import interp as _interp
from .utils import init_fn


def calc_density_compensation_function(
    ktraj: Tensor,
    im_size: Sequence[int],
    num_iterations: int = 10,
    grid_size: Optional[Sequence[int]] = None,
    numpoints: Union[int, Sequence[int]] = 6,
    n_shift: Optional[Sequence[int]] = None,
    table_oversamp: Union[int, Sequence[int]] = 2**10,
    kbwidth: float = 2.34,
    order: Union[float, Sequence[float]] = 0.0,
) -> Tensor:
    """Density compensation weights ``(B, 1, M)`` (complex, zero imaginary part) for
    ``ktraj`` of shape ``(d, M)`` or ``(B, d, M)``."""
    if ktraj.ndim not in (2, 3):
        raise ValueError("ktraj must have 2 or 3 dimensions")
    batch_size = 1
    if ktraj.ndim == 3:
        if ktraj.shape[0] == 1:
            ktraj = ktraj[0]
        else:
            batch_size = ktraj.shape[0]
    pre = init_fn(im_size=im_size, grid_size=grid_size, numpoints=numpoints, n_shift=n_shift,
                  table_oversamp=table_oversamp, kbwidth=kbwidth, order=order, dtype=ktraj.dtype,
                  device=ktraj.device)
    weights = torch.ones([batch_size, 1, ktraj.shape[-1]], dtype=pre.tables[0].dtype, device=ktraj.device)
    for _ in range(num_iterations):
        # nothing here is differentiated (the reference's weights do not require grad either): call the
        # engine directly rather than through the autograd Functions
        gridded = _interp.table_interp_adjoint(weights, ktraj, pre.tables, pre.n_shift, pre.numpoints, pre.table_oversamp,
                                               pre.offsets, pre.grid_size)
        resampled = _interp.table_interp(gridded, ktraj, pre.tables, pre.n_shift, pre.numpoints, pre.table_oversamp,
                                         pre.offsets)
        weights = weights / torch.abs(resampled)
    return weights
