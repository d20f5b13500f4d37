"""Small complex-arithmetic helpers for tensors in either convention (complex
dtype, or real with a size-2 axis holding re/im), with the names and semantics of
``torchkbnufft/_math.py``.  Plain torch ops; not part of the accelerated path."""
from __future__ import annotations

import torch
from torch import Tensor


def _require_pair(*vals: Tensor, dim: int) -> None:
    if not all(v.shape[dim] == 2 for v in vals):
        raise ValueError("Real input does not have dimension size 2 at dim.")


def _same_dtype(val1: Tensor, val2: Tensor) -> None:
    if not val1.dtype == val2.dtype:
        raise ValueError("val1 has different dtype than val2.")


def absolute(val: Tensor, dim: int = -1) -> Tensor:
    """Complex magnitude; for real-view input the re/im axis is kept with size 1."""
    if torch.is_complex(val):
        return torch.abs(val)
    _require_pair(val, dim=dim)
    re, im = val.select(dim, 0), val.select(dim, 1)
    return torch.sqrt(re**2 + im**2).unsqueeze(dim)


def complex_mult(val1: Tensor, val2: Tensor, dim: int = -1) -> Tensor:
    """``val1 * val2``."""
    _same_dtype(val1, val2)
    if torch.is_complex(val1):
        return val1 * val2
    _require_pair(val1, val2, dim=dim)
    a, b = val1.select(dim, 0), val1.select(dim, 1)
    c, d = val2.select(dim, 0), val2.select(dim, 1)
    return torch.stack((a * c - b * d, b * c + a * d), dim)


def conj_complex_mult(val1: Tensor, val2: Tensor, dim: int = -1) -> Tensor:
    """``val1 * conj(val2)``."""
    _same_dtype(val1, val2)
    if torch.is_complex(val1):
        return val1 * val2.conj()
    _require_pair(val1, val2, dim=dim)
    a, b = val1.select(dim, 0), val1.select(dim, 1)
    c, d = val2.select(dim, 0), val2.select(dim, 1)
    return torch.stack((a * c + b * d, b * c - a * d), dim)


def imag_exp(val: Tensor, dim: int = -1, return_complex: bool = True) -> Tensor:
    """``exp(i * val)`` for real ``val`` (``cos + i sin``); real-view output stacks on
    a new last axis."""
    out = torch.stack((torch.cos(val), torch.sin(val)), -1)
    return torch.view_as_complex(out) if return_complex else out


def complex_sign(val: Tensor, dim: int = -1) -> Tensor:
    """Unit-magnitude phase factor ``exp(i * angle(val))``."""
    is_complex = torch.is_complex(val)
    if is_complex:
        val, dim = torch.view_as_real(val), -1
    else:
        _require_pair(val, dim=dim)
    angle = torch.atan2(val.select(dim, 1), val.select(dim, 0))
    return imag_exp(angle, dim=dim, return_complex=is_complex)


def inner_product(val1: Tensor, val2: Tensor, dim: int = -1) -> Tensor:
    """``sum(conj(val1) * val2)``; real-view input returns a 2-vector (re, im)."""
    _same_dtype(val1, val2)
    if torch.is_complex(val1):
        return torch.sum(val2 * val1.conj())
    _require_pair(val1, val2, dim=dim)
    prod = conj_complex_mult(val2, val1, dim=dim)
    return torch.stack((prod.select(dim, 0).sum(), prod.select(dim, 1).sum()))
