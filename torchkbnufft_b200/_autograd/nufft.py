"""Autograd wrappers of the fused steps around the FFT and of the Toeplitz filter.

In the reference these steps are plain differentiable ATen ops
(``_nufft/fft.py:36-118``, ``modules/kbnufft.py:182-183``, ``:404-405``,
``:441-484``); here each is one kernel, so each gets an explicit backward.  The
two fused steps are exact adjoints of one another (the reference's inverse FFT is
unnormalised, ``_nufft/fft.py:19``), which makes every backward a single call of
the opposite kernel.  Sensitivity maps and the scaling coefficients are treated as
constants here; callers route through plain torch ops when those need gradients.

Every backward calls a raw kernel, so it is marked ``once_differentiable``: asking for a
second derivative through these Functions (``create_graph=True``) raises instead of
silently returning gradients cut off from the graph.  (The operators are linear in their
differentiable argument, so Hessian-vector products of a least-squares loss never need it:
apply the operator to the vector instead.)
"""
from __future__ import annotations

from typing import Optional

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .._nufft import fft as _fft


class ApodPad(Function):
    """``image (B, 1|C, *N) -> zero_pad(image * smaps * scaling_coef) * scale``."""

    @staticmethod
    def forward(ctx, image, smaps, scaling_coef, grid_size, scale):
        ctx.save_for_backward(smaps, scaling_coef)
        ctx.im_size = tuple(image.shape[2:])
        ctx.scale = scale
        return _fft.apod_pad(image, grid_size, smaps, scaling_coef, scale)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_grid):
        smaps, scaling_coef = ctx.saved_tensors
        grad = _fft.crop_apod_coilsum(grad_grid, ctx.im_size, smaps, scaling_coef, ctx.scale)
        return grad, None, None, None, None


class CropApodCoilsum(Function):
    """``grid (B, C, *K) -> sum_c crop(grid) * conj(scaling_coef) * conj(smaps) * scale``."""

    @staticmethod
    def forward(ctx, grid, smaps, scaling_coef, im_size, scale):
        ctx.save_for_backward(smaps, scaling_coef)
        ctx.grid_size = tuple(grid.shape[2:])
        ctx.n_coils = grid.shape[1]
        ctx.scale = scale
        return _fft.crop_apod_coilsum(grid, im_size, smaps, scaling_coef, scale)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_image):
        smaps, scaling_coef = ctx.saved_tensors
        grad = _fft.apod_pad(grad_image, ctx.grid_size, smaps, scaling_coef, ctx.scale, n_coils=ctx.n_coils)
        return grad, None, None, None, None


class FusedFftForward(Function):
    """``image -> fftn(zero_pad(image * smaps * scaling_coef)) * scale`` with the engine's own
    pruned FFT passes; the backward is :class:`FusedFftAdjoint`'s forward (exact adjoint)."""

    @staticmethod
    def forward(ctx, image, smaps, scaling_coef, grid_size, scale):
        ctx.save_for_backward(smaps, scaling_coef)
        ctx.im_size = tuple(image.shape[2:])
        ctx.scale = scale
        return _fft.fused_fft_forward(image, grid_size, smaps, scaling_coef, scale)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_grid):
        smaps, scaling_coef = ctx.saved_tensors
        return _fft.fused_fft_adjoint(grad_grid, ctx.im_size, smaps, scaling_coef, ctx.scale), None, None, None, None


class FusedFftAdjoint(Function):
    """``grid -> sum_c crop(ifftn_unnormalised(grid)) * conj(scaling_coef) * conj(smaps) * scale``."""

    @staticmethod
    def forward(ctx, grid, smaps, scaling_coef, im_size, scale):
        ctx.save_for_backward(smaps, scaling_coef)
        ctx.grid_size = tuple(grid.shape[2:])
        ctx.n_coils = grid.shape[1]
        ctx.scale = scale
        return _fft.fused_fft_adjoint(grid, im_size, smaps, scaling_coef, scale)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_image):
        smaps, scaling_coef = ctx.saved_tensors
        grad = _fft.fused_fft_forward(grad_image, ctx.grid_size, smaps, scaling_coef, ctx.scale, n_coils=ctx.n_coils)
        return grad, None, None, None, None


fuse_toeplitz_columns = True  # False: forward passes + inverse passes through the full spectrum (A/B, tests)


def toeplitz_apply(image, kernel, smaps, normalized: bool):
    """``sum_c conj(S_c) crop(IFFT(kernel * FFT(pad(S_c * image))))`` for the whole
    batch at once (the reference loops over the batch in Python,
    ``modules/kbnufft.py:441-484``).  kernel is ``(*K2)`` or ``(B, *K2)``."""
    ndim = image.ndim - 2
    grid_size = tuple(kernel.shape[-ndim:])
    n_grid = 1
    for k in grid_size:
        n_grid *= k
    n_rows = image.shape[0] * (smaps.shape[1] if smaps is not None else image.shape[1])
    if (fuse_toeplitz_columns and _fft.fused_fft_available(image.dtype, grid_size, n_rows)
            and _fft.toeplitz_fused_available(image.dtype, image.shape[2:], grid_size)):
        # three passes: the column pass transforms, filters and transforms back inside the CTA
        return _fft.fused_toeplitz(image, kernel, smaps, (1.0 / n_grid) if normalized else 1.0)
    if ndim > 1 and _fft.fused_fft_available(image.dtype, grid_size, n_rows):
        # pruned passes both ways, kernel multiply fused into the first inverse pass
        grid = _fft.fused_fft_forward(image, grid_size, smaps, None, 1.0)
        return _fft.fused_fft_adjoint(grid, image.shape[2:], smaps, None, (1.0 / n_grid) if normalized else 1.0,
                                      kernel=kernel)
    grid = _fft.apod_pad(image, grid_size, smaps, None, 1.0)
    grid = _fft.fft_grid(grid, ndim, inverse=False)
    # 'ortho' scales both transforms by 1/sqrt(prod K2): fold 1/prod(K2) into the filter pass
    _fft.spectrum_mul_(grid, kernel, (1.0 / n_grid) if normalized else 1.0)
    grid = _fft.fft_grid(grid, ndim, inverse=True)
    return _fft.crop_apod_coilsum(grid, image.shape[2:], smaps, None, 1.0)


class ToeplitzFilter(Function):
    """Differentiable (w.r.t. ``image``) fused Toeplitz normal operator."""

    @staticmethod
    def forward(ctx, image, kernel, smaps, normalized):
        ctx.save_for_backward(kernel, smaps)
        ctx.normalized = normalized
        return toeplitz_apply(image, kernel, smaps, normalized)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad):
        kernel, smaps = ctx.saved_tensors
        # the operator is S^H P^H F^H D F P S; its adjoint swaps D for conj(D)
        return toeplitz_apply(grad, kernel.conj(), smaps, ctx.normalized), None, None, None
