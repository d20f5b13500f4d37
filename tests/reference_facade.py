"""``torchkbnufft``-shaped facade over ``torchkbnufft_b200`` for running the REFERENCE'S OWN test files unchanged.

The upstream tests (``/root/reference/tests/test_{interp,nufft,sense_nufft,toep,dcomp}.py``) build CPU tensors and CPU
modules; the engine is CUDA-only by design.  This facade is the thinnest possible adapter: same names as
``torchkbnufft/__init__.py:30-47``; every module lives on ``cuda:0``; tensor arguments are moved there and results are
moved back to the device of the first tensor argument.  No arithmetic happens here -- the numbers the upstream
assertions see come from the CUDA kernels.  Test infrastructure only (``tests/test_reference_suite.py``).
"""
from __future__ import annotations

import sys
import types

import torch

import torchkbnufft_b200 as _eng

_DEV = torch.device("cuda:0")


def _to_dev(x):
    if isinstance(x, torch.Tensor):
        return x.to(_DEV)
    if isinstance(x, (tuple, list)):
        return type(x)(_to_dev(v) for v in x)
    return x


def _first_device(args, kwargs):
    for v in list(args) + list(kwargs.values()):
        if isinstance(v, torch.Tensor):
            return v.device
        if isinstance(v, (tuple, list)):
            for w in v:
                if isinstance(w, torch.Tensor):
                    return w.device
    return torch.device("cpu")


def _back(out, device):
    if isinstance(out, torch.Tensor):
        return out.to(device)
    if isinstance(out, (tuple, list)):
        return type(out)(_back(v, device) for v in out)
    return out


def _on_gpu(fn):
    def call(*args, **kwargs):
        device = _first_device(args, kwargs)
        return _back(fn(*_to_dev(args), **{k: _to_dev(v) for k, v in kwargs.items()}), device)

    call.__name__ = getattr(fn, "__name__", "call")
    call.__doc__ = fn.__doc__
    return call


def _wrap_module(cls):
    class OnGpu(cls):
        __doc__ = cls.__doc__

        def __init__(self, *args, **kwargs):
            kwargs.pop("device", None)
            super().__init__(*args, **kwargs)
            torch.nn.Module.to(self, _DEV)

        def to(self, *args, **kwargs):  # dtype moves are honoured, device moves are not (the engine has no CPU path)
            for a in list(args) + list(kwargs.values()):
                if isinstance(a, torch.dtype):
                    super().to(a)
                elif isinstance(a, torch.Tensor):
                    super().to(a.dtype)
            return self

        def forward(self, *args, **kwargs):
            device = _first_device(args, kwargs)
            out = super().forward(*_to_dev(args), **{k: _to_dev(v) for k, v in kwargs.items()})
            return _back(out, device)

    OnGpu.__name__ = cls.__name__
    OnGpu.__qualname__ = cls.__name__
    return OnGpu


def install() -> types.ModuleType:
    """Register the facade as ``torchkbnufft`` in ``sys.modules`` and return it."""
    mod = types.ModuleType("torchkbnufft")
    mod.__doc__ = __doc__
    for name in ("KbInterp", "KbInterpAdjoint", "KbNufft", "KbNufftAdjoint", "ToepNufft"):
        setattr(mod, name, _wrap_module(getattr(_eng, name)))
    for name in ("calc_density_compensation_function", "calc_tensor_spmatrix", "calc_toeplitz_kernel"):
        setattr(mod, name, _on_gpu(getattr(_eng, name)))
    for name in ("absolute", "complex_mult", "complex_sign", "conj_complex_mult", "imag_exp", "inner_product"):
        setattr(mod, name, getattr(_eng, name))  # pure torch helpers: device agnostic
    mod.functional = _eng.functional
    mod.__version__ = "b200-facade"
    sys.modules["torchkbnufft"] = mod
    return mod
