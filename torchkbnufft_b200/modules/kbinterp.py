"""Interpolation-only modules (no FFT, no apodisation):
``KbInterp`` (grid -> samples) and ``KbInterpAdjoint`` (samples -> grid), mirroring
``torchkbnufft/modules/kbinterp.py:41-294``."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

from .. import functional as tkbnF
from ._kbmodule import KbModule


class KbInterp(KbModule):
    """Non-uniform Kaiser-Bessel interpolation layer.

    ``forward(image, omega)``: ``image`` is ``(B, C, *grid_size)`` complex (or real
    with a trailing dim of 2), ``omega`` is ``(d, M)`` or ``(B, d, M)`` in
    radians/voxel; returns ``(B, C, M)``.  Construct with ``grid_size=im_size`` to
    interpolate an un-oversampled array (as the reference's tests do)."""

    def forward(self, image: Tensor, omega: Tensor, interp_mats: Optional[Tuple[Tensor, Tensor]] = None) -> Tensor:
        if interp_mats is not None:
            return tkbnF.kb_spmat_interp(image=image, interp_mats=interp_mats)
        return tkbnF.kb_table_interp(image=image, omega=omega, tables=self.tables, n_shift=self.n_shift,
                                     numpoints=self.numpoints, table_oversamp=self.table_oversamp,
                                     offsets=self.offsets)


class KbInterpAdjoint(KbModule):
    """Adjoint interpolation layer: ``forward(data, omega, grid_size=None)`` spreads
    ``data (B, C, M)`` onto a ``(B, C, *grid_size)`` grid (default: the module's)."""

    def forward(self, data: Tensor, omega: Tensor, interp_mats: Optional[Tuple[Tensor, Tensor]] = None,
                grid_size: Optional[Tensor] = None) -> Tensor:
        if grid_size is None:
            grid_size = self.grid_size
        if interp_mats is not None:
            return tkbnF.kb_spmat_interp_adjoint(data=data, interp_mats=interp_mats, grid_size=grid_size)
        return tkbnF.kb_table_interp_adjoint(data=data, omega=omega, tables=self.tables, n_shift=self.n_shift,
                                             numpoints=self.numpoints, table_oversamp=self.table_oversamp,
                                             offsets=self.offsets, grid_size=grid_size)
