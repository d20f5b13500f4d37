"""torchkbnufft_b200 -- a B200-native (sm_100a) engine behind torchkbnufft's
table-interpolation NUFFT API.

Drop-in surface (names and call signatures of ``torchkbnufft/__init__.py:30-47``):
``KbInterp``, ``KbInterpAdjoint``, ``KbNufft``, ``KbNufftAdjoint``, ``ToepNufft``,
``calc_toeplitz_kernel``, ``calc_density_compensation_function``, the
``functional`` module and the complex-math helpers.  The compute path is
hand-written CUDA behind a C ABI (``include/b200nufft.h``); it is CUDA-only and
fails loudly on CPU tensors or when ``libb200nufft.so`` has not been built.
"""
from . import functional, modules
from ._math import absolute, complex_mult, complex_sign, conj_complex_mult, imag_exp, inner_product
from ._nufft import utils as nufft_utils
from ._nufft.dcomp import calc_density_compensation_function
from ._nufft.graphs import get_graph_mode, set_graph_mode
from ._nufft.interp import get_adjoint_mode, get_tiled_kernels, set_adjoint_mode, set_tiled_kernels
from ._nufft.plan import clear_caches, get_plan_cache_mode, invalidate_plans, set_plan_cache_mode
from ._nufft.spmat import calc_tensor_spmatrix
from ._nufft.toep import calc_toeplitz_kernel
from .modules import KbInterp, KbInterpAdjoint, KbNufft, KbNufftAdjoint, ToepNufft

__version__ = "0.1.0"

__all__ = [
    "KbInterp",
    "KbInterpAdjoint",
    "KbNufft",
    "KbNufftAdjoint",
    "ToepNufft",
    "absolute",
    "calc_density_compensation_function",
    "calc_tensor_spmatrix",
    "calc_toeplitz_kernel",
    "clear_caches",
    "complex_mult",
    "complex_sign",
    "conj_complex_mult",
    "functional",
    "get_adjoint_mode",
    "get_graph_mode",
    "get_plan_cache_mode",
    "get_tiled_kernels",
    "imag_exp",
    "inner_product",
    "invalidate_plans",
    "modules",
    "set_adjoint_mode",
    "set_graph_mode",
    "set_plan_cache_mode",
    "set_tiled_kernels",
]
