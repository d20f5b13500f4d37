"""NUFFT modules: ``KbNufft`` (image -> k-space), ``KbNufftAdjoint`` (k-space ->
image) with optional SENSE maps, and ``ToepNufft`` (Toeplitz normal operator),
mirroring ``torchkbnufft/modules/kbnufft.py`` (:11-52, :125-228, :307-410,
:413-547).  The SENSE multiply / coil combination, the apodisation and the
zero-pad / crop are fused into one kernel on each side of the FFT."""
from __future__ import annotations

from typing import Optional, Sequence, Tuple, Union

import torch
from torch import Tensor

from .. import functional as tkbnF
from .._nufft.utils import compute_scaling_coefs
from ..functional.nufft import sense_nufft_adjoint, sense_nufft_forward, toeplitz_filter
from ._kbmodule import KbModule


class KbNufftModule(KbModule):
    """Adds the image-domain apodisation (``scaling_coef``) to :class:`KbModule`."""

    def __init__(
        self,
        im_size: Sequence[int],
        grid_size: Optional[Sequence[int]] = None,
        numpoints: Union[int, Sequence[int]] = 6,
        n_shift: Optional[Sequence[int]] = None,
        table_oversamp: Union[int, Sequence[int]] = 2**10,
        kbwidth: float = 2.34,
        order: Union[float, Sequence[float]] = 0.0,
        dtype: Optional[torch.dtype] = None,
        device: Optional[torch.device] = None,
    ):
        super().__init__(im_size=im_size, grid_size=grid_size, numpoints=numpoints, n_shift=n_shift,
                         table_oversamp=table_oversamp, kbwidth=kbwidth, order=order, dtype=dtype, device=device)
        # as the reference does (:41-47), alpha/order are read back from the (possibly
        # float32-rounded) buffers so that complex64 modules get bit-identical coefficients
        scaling_coef = compute_scaling_coefs(
            im_size=self.im_size.tolist(), grid_size=self.grid_size.tolist(), numpoints=self.numpoints.tolist(),
            alpha=self.alpha.tolist(), order=self.order.tolist(),
        )
        self.register_buffer("scaling_coef", scaling_coef.to(dtype=self.table_0.dtype, device=device))


def _complex_views(x: Tensor, smaps: Optional[Tensor], x_name: str):
    """dtype / real-view checks shared by both NUFFT directions
    (reference: modules/kbnufft.py:165-180 and :347-362)."""
    if smaps is not None and not smaps.dtype == x.dtype:
        raise TypeError(f"{x_name} dtype does not match smaps dtype.")
    if x.is_complex():
        return x, smaps, True
    if not x.shape[-1] == 2:
        raise ValueError("For real inputs, last dimension must be size 2.")
    if smaps is not None:
        if not smaps.shape[-1] == 2:
            raise ValueError("For real inputs, last dimension must be size 2.")
        smaps = torch.view_as_complex(smaps)
    return torch.view_as_complex(x), smaps, False


class KbNufft(KbNufftModule):
    """Forward NUFFT: ``forward(image, omega, interp_mats=None, smaps=None, norm=None)``.

    ``image`` is ``(B, C, *im_size)`` (``(B, 1, *im_size)`` with ``smaps (Bs, C,
    *im_size)`` for SENSE); returns k-space data ``(B, C, M)``.  ``norm`` is ``None``
    or ``"ortho"``."""

    def forward(self, image: Tensor, omega: Tensor, interp_mats: Optional[Tuple[Tensor, Tensor]] = None,
                smaps: Optional[Tensor] = None, norm: Optional[str] = None) -> Tensor:
        image, smaps, is_complex = _complex_views(image, smaps, "image")
        if interp_mats is not None:
            if smaps is not None:
                image = image * smaps
            output = tkbnF.kb_spmat_nufft(image=image, scaling_coef=self.scaling_coef, im_size=self.im_size,
                                          grid_size=self.grid_size, interp_mats=interp_mats, norm=norm)
        else:
            output = sense_nufft_forward(image, smaps, self.scaling_coef, self.grid_size, omega, self.tables,
                                         self.n_shift, self.numpoints, self.table_oversamp, self.offsets, norm)
        return output if is_complex else torch.view_as_real(output)


class KbNufftAdjoint(KbNufftModule):
    """Adjoint NUFFT: ``forward(data, omega, interp_mats=None, smaps=None, norm=None)``.

    ``data`` is ``(B, C, M)``; returns ``(B, C, *im_size)``, or the coil-combined
    ``(B, 1, *im_size)`` when ``smaps`` is given."""

    def forward(self, data: Tensor, omega: Tensor, interp_mats: Optional[Tuple[Tensor, Tensor]] = None,
                smaps: Optional[Tensor] = None, norm: Optional[str] = None) -> Tensor:
        data, smaps, is_complex = _complex_views(data, smaps, "data")
        if interp_mats is not None:
            output = tkbnF.kb_spmat_nufft_adjoint(data=data, scaling_coef=self.scaling_coef, im_size=self.im_size,
                                                  grid_size=self.grid_size, interp_mats=interp_mats, norm=norm)
            if smaps is not None:
                output = torch.sum(output * smaps.conj(), dim=1, keepdim=True)
        else:
            output = sense_nufft_adjoint(data, smaps, self.scaling_coef, self.im_size, self.grid_size, omega,
                                         self.tables, self.n_shift, self.numpoints, self.table_oversamp,
                                         self.offsets, norm)
        return output if is_complex else torch.view_as_real(output)


class ToepNufft(torch.nn.Module):
    """Forward/adjoint NUFFT pair as one Toeplitz-embedded FFT filter.

    ``forward(image, kernel, smaps=None, norm=None)`` with ``kernel`` from
    :func:`calc_toeplitz_kernel` (shape ``2*im_size`` or ``(B, *2*im_size)``).
    The whole batch is filtered in one pass; a single set of ``smaps`` or a single
    kernel broadcasts over the batch."""

    def __init__(self):
        super().__init__()

    def forward(self, image: Tensor, kernel: Tensor, smaps: Optional[Tensor] = None,
                norm: Optional[str] = None) -> Tensor:
        if not kernel.dtype == image.dtype:
            raise TypeError("kernel and image must have same dtype.")
        if smaps is not None and not smaps.dtype == image.dtype:
            raise TypeError("image dtype does not match smaps dtype.")
        is_complex = image.is_complex()
        if not is_complex:
            for t in (image, kernel) + ((smaps,) if smaps is not None else ()):
                if not t.shape[-1] == 2:
                    raise ValueError("For real inputs, last dimension must be size 2.")
            image, kernel = torch.view_as_complex(image), torch.view_as_complex(kernel)
            smaps = torch.view_as_complex(smaps) if smaps is not None else None
        ndim = image.ndim - 2
        if kernel.ndim > ndim:
            if kernel.shape[0] == 1:
                kernel = kernel[0]
            elif not kernel.shape[0] == image.shape[0]:
                raise ValueError("If using batch dimension, kernel must have same batch size as image")
        if smaps is not None and smaps.shape[0] not in (1, image.shape[0]):
            raise ValueError("smaps batch dimension must be 1 or match image")
        output = toeplitz_filter(image, kernel, smaps, norm)
        return output if is_complex else torch.view_as_real(output)
