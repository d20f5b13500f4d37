"""Functional NUFFT API (stateless), mirroring ``torchkbnufft/functional/nufft.py``
(``kb_table_nufft`` :126-191, ``kb_table_nufft_adjoint`` :194-260) and
``fft_filter`` (``_nufft/fft.py:121-173``)."""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
from torch import Tensor

from .._autograd.interp import KbTableInterpAdjoint, KbTableInterpForward
from .._autograd.nufft import ApodPad, CropApodCoilsum, FusedFftAdjoint, FusedFftForward, ToeplitzFilter
from .._nufft import fft as _fft
from .._nufft import graphs as _graphs
from .._nufft import interp as _interp
from .._nufft.plan import host_ints as _ints
from .._nufft import spmat as _spmat
from .interp import with_complex_view


def _needs_grad(t: Tensor) -> bool:
    """Whether autograd has to record this call; when it does not, the kernels are called directly --
    the ``Function.apply`` round trip costs more host time than some of the kernels run."""
    return torch.is_grad_enabled() and t.requires_grad


def sense_nufft_forward(image: Tensor, smaps: Optional[Tensor], scaling_coef: Tensor, grid_size, omega: Tensor,
                        tables: List[Tensor], n_shift: Tensor, numpoints: Tensor, table_oversamp: Tensor,
                        offsets: Tensor, norm: Optional[str] = None) -> Tensor:
    """Complex-only core of the forward NUFFT with the optional SENSE multiply fused
    into the apodisation/zero-pad kernel."""
    normalized = _fft.check_norm(norm)
    if image.shape[0] == 0:
        raise ValueError("image has an empty batch dimension")
    if smaps is not None and (image.shape[1] != 1 or smaps.requires_grad):
        image, smaps = image * smaps, None  # general broadcast / d(smaps): plain torch multiply
    grid_size = _ints(grid_size)
    scale = _fft.ortho_scale(grid_size, normalized)
    record = _needs_grad(image)
    if not record and _graphs.get_graph_mode() and not _graphs.in_replay_scope():
        with _graphs.replay_scope():
            return _graphs.replay_or_run(
                "nufft_forward", (image, smaps), (scaling_coef, n_shift, numpoints, table_oversamp) + tuple(tables), omega,
                (grid_size, scale),
                lambda: sense_nufft_forward(image, smaps, scaling_coef, grid_size, omega, tables, n_shift, numpoints,
                                            table_oversamp, offsets, norm),
                lambda: _interp.lookup_plan(omega, image.shape[0], tables, n_shift, numpoints, table_oversamp, grid_size,
                                            image.device))
    n_rows = image.shape[0] * (smaps.shape[1] if smaps is not None else image.shape[1])
    if _fft.fused_fft_available(image.dtype, grid_size, n_rows):
        grid = (FusedFftForward.apply(image, smaps, scaling_coef, grid_size, scale) if record else
                _fft.fused_fft_forward(image, grid_size, smaps, scaling_coef, scale))
    else:
        grid = (ApodPad.apply(image, smaps, scaling_coef, grid_size, scale) if record else
                _fft.apod_pad(image, grid_size, smaps, scaling_coef, scale))
        grid = _fft.fft_grid(grid, len(grid_size), inverse=False)
    if not record:
        return _interp.table_interp(grid, omega, tables, n_shift, numpoints, table_oversamp, offsets)
    return KbTableInterpForward.apply(grid, omega, tables, n_shift, numpoints, table_oversamp, offsets)


def sense_nufft_adjoint(data: Tensor, smaps: Optional[Tensor], scaling_coef: Tensor, im_size, grid_size,
                        omega: Tensor, tables: List[Tensor], n_shift: Tensor, numpoints: Tensor,
                        table_oversamp: Tensor, offsets: Tensor, norm: Optional[str] = None) -> Tensor:
    """Complex-only core of the adjoint NUFFT with crop, conjugate apodisation and the
    SENSE coil combination fused into one kernel."""
    normalized = _fft.check_norm(norm)
    grid_sizes = _ints(grid_size)
    scale = _fft.ortho_scale(grid_sizes, normalized)
    fused = _fft.fused_fft_available(data.dtype, grid_sizes, data.shape[0] * data.shape[1])
    epilogue = _graphs.current_epilogue()  # e.g. the all-reduce of a coil-sharded adjoint (parallel.py)
    if (_graphs.get_graph_mode() and not _graphs.in_replay_scope() and not _needs_grad(data)
            and not (smaps is not None and smaps.requires_grad)):
        with _graphs.replay_scope():
            out = _graphs.replay_or_run(
                "nufft_adjoint", (data, smaps), (scaling_coef, n_shift, numpoints, table_oversamp) + tuple(tables), omega,
                (grid_sizes, _ints(im_size), scale, _interp.get_adjoint_mode(), epilogue.key if epilogue else None),
                lambda: sense_nufft_adjoint(data, smaps, scaling_coef, im_size, grid_size, omega, tables, n_shift,
                                            numpoints, table_oversamp, offsets, norm),
                lambda: _interp.lookup_plan(omega, data.shape[0], tables, n_shift, numpoints, table_oversamp, grid_sizes,
                                            data.device))
        if epilogue is not None:
            epilogue.applied = True  # by the eager body or by the replayed graph
        return out
    if not _needs_grad(data) and not (smaps is not None and smaps.requires_grad):
        grid = _interp.table_interp_adjoint(data, omega, tables, n_shift, numpoints, table_oversamp, offsets, grid_size)
        if fused and epilogue is not None and epilogue.peer_comm is not None:
            # the all-reduce is part of the last inverse pass (or one kernel behind it: the C side decides)
            out = _fft.fused_fft_adjoint(grid, _ints(im_size), smaps, scaling_coef, scale, peer_comm=epilogue.peer_comm)
            epilogue.applied = True
            return out
        if fused:
            out = _fft.fused_fft_adjoint(grid, _ints(im_size), smaps, scaling_coef, scale)
        else:
            grid = _fft.fft_grid(grid, len(grid_sizes), inverse=True)
            out = _fft.crop_apod_coilsum(grid, _ints(im_size), smaps, scaling_coef, scale)
        if epilogue is not None:
            out = epilogue.fn(out)
            epilogue.applied = True
        return out
    grid = KbTableInterpAdjoint.apply(data, omega, tables, n_shift, numpoints, table_oversamp, offsets, grid_size)
    finish = FusedFftAdjoint if fused else CropApodCoilsum
    if not fused:
        grid = _fft.fft_grid(grid, len(grid_sizes), inverse=True)
    if smaps is not None and smaps.requires_grad:
        image = finish.apply(grid, None, scaling_coef, _ints(im_size), scale)
        return torch.sum(image * smaps.conj(), dim=1, keepdim=True)
    return finish.apply(grid, smaps, scaling_coef, _ints(im_size), scale)


def kb_table_nufft(image: Tensor, scaling_coef: Tensor, im_size: Tensor, grid_size: Tensor, omega: Tensor,
                   tables: List[Tensor], n_shift: Tensor, numpoints: Tensor, table_oversamp: Tensor, offsets: Tensor,
                   norm: Optional[str] = None) -> Tensor:
    """Forward NUFFT with table interpolation: ``image (B, C, *N)`` -> ``(B, C, M)``."""
    return with_complex_view(
        lambda x: sense_nufft_forward(x, None, scaling_coef, grid_size, omega, tables, n_shift, numpoints,
                                      table_oversamp, offsets, norm),
        image,
    )


def kb_table_nufft_adjoint(data: Tensor, scaling_coef: Tensor, im_size: Tensor, grid_size: Tensor, omega: Tensor,
                           tables: List[Tensor], n_shift: Tensor, numpoints: Tensor, table_oversamp: Tensor,
                           offsets: Tensor, norm: Optional[str] = None) -> Tensor:
    """Adjoint NUFFT with table interpolation: ``data (B, C, M)`` -> ``(B, C, *N)``."""
    return with_complex_view(
        lambda x: sense_nufft_adjoint(x, None, scaling_coef, im_size, grid_size, omega, tables, n_shift, numpoints,
                                      table_oversamp, offsets, norm),
        data,
    )


def toeplitz_filter(image: Tensor, kernel: Tensor, smaps: Optional[Tensor], norm: Optional[str]) -> Tensor:
    """Complex-only batched Toeplitz normal operator (fused kernels + cuFFT)."""
    normalized = _fft.check_norm(norm)
    if kernel.requires_grad or (smaps is not None and smaps.requires_grad):
        # differentiable-in-everything composition out of the same fused pieces
        ndim = image.ndim - 2
        grid_size = tuple(kernel.shape[-ndim:])
        n_grid = 1
        for k in grid_size:
            n_grid *= k
        x = image if smaps is None else image * smaps
        grid = _fft.fft_grid(ApodPad.apply(x, None, None, grid_size, 1.0), ndim, inverse=False)
        kern = kernel if kernel.ndim == ndim else kernel.unsqueeze(1)
        grid = _fft.fft_grid(grid * kern * ((1.0 / n_grid) if normalized else 1.0), ndim, inverse=True)
        out = CropApodCoilsum.apply(grid, None, None, tuple(image.shape[2:]), 1.0)
        return out if smaps is None else torch.sum(out * smaps.conj(), dim=1, keepdim=True)
    if smaps is not None and image.shape[1] != 1:
        raise ValueError("with smaps, image must have a single coil dimension (B, 1, *N)")
    return ToeplitzFilter.apply(image, kernel, smaps, normalized)


def fft_filter(image: Tensor, kernel: Tensor, norm: Optional[str] = "ortho") -> Tensor:
    """``crop(IFFT(kernel * FFT(zero_pad(image))))`` on the grid of ``kernel``
    (reference: ``_nufft/fft.py:121-173``).

    ``kernel`` is ``(*K)`` (shared by every batch element and coil) or ``(B, *K)`` / ``(1, *K)``: one kernel
    PER BATCH ELEMENT, applied to all coils of that element -- which is how ``ToepNufft`` with a batched
    trajectory uses it (``modules/kbnufft.py:441-484``).  Note the difference from calling the reference's
    ``fft_filter`` directly with an ``(ndim+1)``-D kernel: plain broadcasting there lines the extra axis up with
    the COIL axis.  Pass ``kernel.unsqueeze(0)`` permuted to your intent, or loop, if you relied on that."""
    return toeplitz_filter(image, kernel, None, norm)


def kb_spmat_nufft(image: Tensor, scaling_coef: Tensor, im_size: Tensor, grid_size: Tensor,
                   interp_mats: Tuple[Tensor, Tensor], norm: Optional[str] = None) -> Tensor:
    """Forward NUFFT with sparse-matrix interpolation (reference ``functional/nufft.py:11-64``): the
    engine's apodise/pad/FFT step, then ``torch.sparse`` for the interpolation (API-completeness shim)."""
    def run(x: Tensor) -> Tensor:
        normalized = _fft.check_norm(norm)
        sizes = _ints(grid_size)
        scale = _fft.ortho_scale(sizes, normalized)
        if _fft.fused_fft_available(x.dtype, sizes, x.shape[0] * x.shape[1]):
            grid = FusedFftForward.apply(x, None, scaling_coef, sizes, scale)
        else:
            grid = _fft.fft_grid(ApodPad.apply(x, None, scaling_coef, sizes, scale), len(sizes), inverse=False)
        return _spmat.spmat_interp(grid, interp_mats)

    return with_complex_view(run, image)


def kb_spmat_nufft_adjoint(data: Tensor, scaling_coef: Tensor, im_size: Tensor, grid_size: Tensor,
                           interp_mats: Tuple[Tensor, Tensor], norm: Optional[str] = None) -> Tensor:
    """Adjoint NUFFT with sparse-matrix interpolation (reference ``functional/nufft.py:67-123``)."""
    def run(y: Tensor) -> Tensor:
        normalized = _fft.check_norm(norm)
        sizes = _ints(grid_size)
        scale = _fft.ortho_scale(sizes, normalized)
        grid = _spmat.spmat_interp_adjoint(y, interp_mats, sizes)
        if _fft.fused_fft_available(y.dtype, sizes, y.shape[0] * y.shape[1]):
            return FusedFftAdjoint.apply(grid, None, scaling_coef, _ints(im_size), scale)
        grid = _fft.fft_grid(grid, len(sizes), inverse=True)
        return CropApodCoilsum.apply(grid, None, scaling_coef, _ints(im_size), scale)

    return with_complex_view(run, data)
