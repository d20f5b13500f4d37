"""Base class of the NUFFT modules: precompute + buffer registration.

State-dict compatible with the reference's ``KbModule``
(``torchkbnufft/modules/_kbmodule.py:9-116``): the same buffer names, dtypes and
shapes (``table_{i}``, ``im_size``, ``grid_size``, ``n_shift``, ``numpoints``,
``offsets``, ``table_oversamp``, ``order``, ``alpha``), so checkpoints interchange.
Trajectory plans are runtime caches and never enter the state dict.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Union

import torch
import torch.nn as nn
from torch import Tensor

from .._nufft.utils import init_fn, paired_dtypes


class KbModule(nn.Module):
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
        super().__init__()
        pre = init_fn(im_size=im_size, grid_size=grid_size, numpoints=numpoints, n_shift=n_shift,
                      table_oversamp=table_oversamp, kbwidth=kbwidth, order=order, dtype=dtype, device=device)
        for i, table in enumerate(pre.tables):
            self.register_buffer(f"table_{i}", table)
        for name in ("im_size", "grid_size", "n_shift", "numpoints", "offsets", "table_oversamp", "order", "alpha"):
            self.register_buffer(name, getattr(pre, name))

    @property
    def tables(self) -> List[Tensor]:
        return [getattr(self, f"table_{i}") for i in range(len(self.im_size))]

    def to(self, *args, **kwargs):
        """``nn.Module.to`` with real/complex pairing: asking for ``float32`` (or
        ``complex64``) moves real buffers to float32 and complex ones to complex64,
        likewise for double precision; integer buffers keep their dtype."""
        device, dtype, non_blocking, convert_to_format = torch._C._nn._parse_to(*args, **kwargs)
        complex_dtype = real_dtype = None
        if dtype is not None:
            if not (dtype.is_floating_point or dtype.is_complex):
                raise TypeError(
                    "KbModule.to only accepts floating point or complex "
                    "dtypes, but got desired dtype={}".format(dtype)
                )
            complex_dtype, real_dtype = paired_dtypes(dtype)

        def convert(t: Tensor) -> Tensor:
            target = None
            if dtype is not None:
                if t.is_floating_point():
                    target = real_dtype
                elif t.is_complex():
                    target = complex_dtype
            if convert_to_format is not None and t.dim() == 4:
                return t.to(device, target, non_blocking, memory_format=convert_to_format)
            return t.to(device, target, non_blocking)

        return self._apply(convert)

    def __repr__(self):
        lines = ["", self.__class__.__name__, "-" * 40, "buffers"]
        for name, buf in self._buffers.items():
            lines.append(f"\ttensor: {name}, shape: {tuple(buf.shape)}")
        return "\n".join(lines) + "\n"
