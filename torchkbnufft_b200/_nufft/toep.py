"""Toeplitz embedding kernel for the NUFFT normal operator ``A^H diag(w) A``.

Same contract as the reference's ``calc_toeplitz_kernel``
(``torchkbnufft/_nufft/toep.py:11-120``): the returned FFT-domain kernel of shape
``2*im_size`` (``(B, *2*im_size)`` for batched trajectories) is consumed by
``ToepNufft``.  The point-spread function ``psf[n] = sum_m w_m exp(i omega_m . n)``
is needed for ``n`` in ``(-N, N)`` per dimension.  An adjoint NUFFT with
``n_shift = 0`` yields the all-non-negative orthant; negating trajectory axes
``1..d-1`` yields the orthants with negative indices along those axes, and the
conjugate symmetry ``psf[-n] = conj(psf[n])`` supplies negative indices along
axis 0.  The reference runs the ``2^(d-1)`` adjoints one after another through a
recursion (:185-237); here they are ONE batched-trajectory adjoint NUFFT.
"""
from __future__ import annotations

import itertools
from typing import Optional, Sequence, Union

import torch
from torch import Tensor


def _reverse_mod(x: Tensor, dim: int) -> Tensor:
    """``y[i] = x[(-i) mod n]`` along ``dim`` (index 0 stays in place)."""
    return torch.roll(torch.flip(x, (dim,)), 1, dims=dim)


def _embed_axis(pos: Tensor, neg: Tensor, dim: int) -> Tensor:
    """Lay out indices ``0..N-1`` (from ``pos``), a zero at ``N`` and ``-(N-1)..-1``
    (from ``neg``, which holds ``psf`` at negated index) in FFT wrap-around order."""
    n = pos.shape[dim]
    shape = list(pos.shape)
    shape[dim] = 1
    zero = torch.zeros(shape, dtype=pos.dtype, device=pos.device)
    return torch.cat((pos, zero, torch.flip(neg.narrow(dim, 1, n - 1), (dim,))), dim)


def calc_toeplitz_kernel(
    omega: Tensor,
    im_size: Sequence[int],
    weights: Optional[Tensor] = None,
    norm: Optional[str] = None,
    grid_size: Optional[Sequence[int]] = None,
    numpoints: Union[int, Sequence[int]] = 6,
    table_oversamp: Union[int, Sequence[int]] = 2**10,
    kbwidth: float = 2.34,
    order: Union[float, Sequence[float]] = 0.0,
) -> Tensor:
    """Compute the Toeplitz kernel for ``omega`` (``(d, M)`` or ``(B, d, M)``).

    ``weights`` (e.g. density compensation) is ``(1, M)`` -- or ``(B, 1, M)`` /
    ``(B, M)`` for batched trajectories; ``norm`` is ``None`` or ``"ortho"``."""
    from ..modules import KbNufftAdjoint  # local import: modules import this package's functional layer

    if omega.ndim not in (2, 3):
        raise ValueError("Unrecognized k-space shape.")
    if weights is not None:
        if weights.ndim not in (2, 3):
            raise ValueError("Unrecognized weights dimension.")
        if omega.ndim == 3 and weights.ndim == 2:
            if weights.shape[0] == 1:
                weights = weights.repeat(omega.shape[0], 1)
            if not weights.shape[0] == omega.shape[0]:
                raise ValueError("weights and omega do not have same batch size")
    batched = omega.ndim == 3
    trajs = omega if batched else omega.unsqueeze(0)  # (T, d, M)
    n_traj, ndim, n_points = trajs.shape
    normalized = norm == "ortho"

    adj_ob = KbNufftAdjoint(im_size=im_size, grid_size=grid_size, numpoints=numpoints, n_shift=[0] * ndim,
                            table_oversamp=table_oversamp, kbwidth=kbwidth, order=order, dtype=omega.dtype,
                            device=omega.device)
    cdtype = adj_ob.table_0.dtype
    if weights is None:
        w = torch.ones((n_traj, 1, n_points), dtype=cdtype, device=omega.device)
    else:
        w = weights.to(dtype=cdtype, device=omega.device).reshape(n_traj, -1, n_points)
        if w.shape[1] != 1:
            raise ValueError("weights must have one row per trajectory")

    # sign patterns over axes 1..d-1 (axis 0 is handled by conjugate symmetry)
    patterns = list(itertools.product((1.0, -1.0), repeat=ndim - 1))
    signs = torch.tensor([(1.0,) + p for p in patterns], dtype=omega.dtype, device=omega.device)  # (P, d)
    flipped = (trajs[:, None] * signs[None, :, :, None]).reshape(n_traj * len(patterns), ndim, n_points)
    w_rep = w[:, None].expand(n_traj, len(patterns), 1, n_points).reshape(-1, 1, n_points)
    orthants = adj_ob(w_rep.contiguous(), flipped.contiguous(), norm=norm)  # (T*P, 1, *N)
    orthants = orthants.reshape(n_traj, len(patterns), *orthants.shape[2:])

    # stitch orthants together, last axis first
    pieces = {p: orthants[:, i] for i, p in enumerate(patterns)}  # each (T, *N)
    for axis in range(ndim - 1, 0, -1):
        merged = {}
        for p in {q[: axis - 1] for q in pieces}:
            merged[p] = _embed_axis(pieces[p + (1.0,)], pieces[p + (-1.0,)], dim=axis + 1)
        pieces = merged
    half = pieces[()]  # (T, N0, 2N1, ..., 2N_{d-1})

    # negative indices along axis 0 by conjugate symmetry, then enforce exact Hermitian symmetry
    mirrored = half.conj()
    for dim in range(1, ndim + 1):
        mirrored = _reverse_mod(mirrored, dim)
    shape = list(half.shape)
    shape[1] = 1
    zero = torch.zeros(shape, dtype=half.dtype, device=half.device)
    kernel = torch.cat((half, zero, mirrored.narrow(1, 1, half.shape[1] - 1)), 1)
    reflected = kernel
    for dim in range(1, ndim + 1):
        reflected = _reverse_mod(reflected, dim)
    kernel = (kernel + reflected.conj()) / 2

    n_embed = 1
    for n in im_size:
        n_embed *= 2 * int(n)
    n_grid = 1
    for k in adj_ob.grid_size.tolist():
        n_grid *= int(k)
    scale = (n_embed / n_grid) ** 0.5 if normalized else 1.0 / n_embed
    kernel = torch.fft.fftn(kernel, dim=list(range(-ndim, 0)), norm="ortho" if normalized else None) * scale
    return kernel if batched else kernel[0]
