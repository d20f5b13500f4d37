"""``torch.autograd.Function`` wrappers of the table-interpolation kernels.

This is the drop-in boundary named by the reference's ``_autograd/interp.py``
(``KbTableInterpForward`` :78-130, ``KbTableInterpAdjoint`` :133-178): identical
``apply`` argument lists and gradient structure (the gather's backward is the
spread and vice versa; only the first argument receives a gradient).
"""
from __future__ import annotations

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .._nufft.interp import table_interp, table_interp_adjoint


class KbTableInterpForward(Function):
    @staticmethod
    def forward(ctx, image, omega, tables, n_shift, numpoints, table_oversamp, offsets):
        """grid ``(B, C, *K)`` -> k-space samples ``(B, C, M)``."""
        output = table_interp(image, omega, tables, n_shift, numpoints, table_oversamp, offsets)
        ctx.grid_size = tuple(int(k) for k in image.shape[2:])
        ctx.save_for_backward(omega, n_shift, numpoints, table_oversamp, offsets, *tables)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, data):
        omega, n_shift, numpoints, table_oversamp, offsets = ctx.saved_tensors[:5]
        tables = list(ctx.saved_tensors[5:])
        image = table_interp_adjoint(data, omega, tables, n_shift, numpoints, table_oversamp, offsets, ctx.grid_size)
        return image, None, None, None, None, None, None


class KbTableInterpAdjoint(Function):
    @staticmethod
    def forward(ctx, data, omega, tables, n_shift, numpoints, table_oversamp, offsets, grid_size):
        """k-space samples ``(B, C, M)`` -> grid ``(B, C, *grid_size)``."""
        image = table_interp_adjoint(data, omega, tables, n_shift, numpoints, table_oversamp, offsets, grid_size)
        ctx.save_for_backward(omega, n_shift, numpoints, table_oversamp, offsets, *tables)
        return image

    @staticmethod
    @once_differentiable
    def backward(ctx, image):
        omega, n_shift, numpoints, table_oversamp, offsets = ctx.saved_tensors[:5]
        tables = list(ctx.saved_tensors[5:])
        data = table_interp(image, omega, tables, n_shift, numpoints, table_oversamp, offsets)
        return data, None, None, None, None, None, None, None
