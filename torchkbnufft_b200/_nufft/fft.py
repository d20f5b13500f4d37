"""The FFT stage of the NUFFT with its element-wise neighbours fused.

Reference (``torchkbnufft/_nufft/fft.py``): ``fft_and_scale`` :36-76 is
mul -> F.pad -> fftn, ``ifft_and_scale`` :80-118 is ifftn -> index_select per dim ->
mul, ``fft_filter`` :121-173 is pad -> fftn -> mul -> ifftn -> crop; the SENSE
coil multiply / coil sum sit in the modules (``modules/kbnufft.py:182-183``,
``:404-405``).  Here the scaling, SENSE multiply, zero-pad (resp. crop, conjugate
multiplies and coil sum) are ONE kernel on each side of a cuFFT transform, and the
'ortho' factor is folded into that kernel so cuFFT never runs a scaling pass.

The raw (non-differentiable) host calls live here; autograd wrappers are in
``_autograd/nufft.py``.
"""
from __future__ import annotations

import ctypes
import math
from typing import Optional, Sequence

import torch
from torch import Tensor

from .. import _lib
from .plan import current_stream_ptr, dense, device_guard, engine_dtype, host_ints, require_cuda


def check_norm(norm: Optional[str]) -> bool:
    if norm is not None and norm != "ortho":
        raise ValueError("Only option for norm is 'ortho'.")
    return norm == "ortho"


def _sizes(v) -> list:
    return list(host_ints(v))


def _check_operands(what: str, x: Tensor, spatial: Sequence[int], smaps: Optional[Tensor],
                    scaling_coef: Optional[Tensor], layout: int = _lib.COIL_MAJOR) -> None:
    """The kernels take raw pointers and trust ``spatial`` for every operand: refuse anything the reference would
    reject with a broadcast error (``_nufft/fft.py:72``, ``modules/kbnufft.py:183``, ``:405``) before it can
    become an out-of-bounds read."""
    spatial = tuple(int(n) for n in spatial)
    if not x.is_complex():
        raise TypeError(f"{what} must be complex")
    if scaling_coef is not None:
        if tuple(scaling_coef.shape) != spatial:
            raise ValueError(f"scaling_coef has shape {tuple(scaling_coef.shape)} but the image size is {spatial}")
        if scaling_coef.dtype != x.dtype:
            raise TypeError(f"scaling_coef dtype {scaling_coef.dtype} does not match {what} dtype {x.dtype}")
        if scaling_coef.device != x.device:
            raise ValueError(f"scaling_coef is on {scaling_coef.device} but {what} is on {x.device}")
    if smaps is not None:
        if smaps.ndim != len(spatial) + 2:
            raise ValueError(f"smaps must have {len(spatial) + 2} dimensions, got {smaps.ndim}")
        sm_spatial = tuple(smaps.shape[1:-1]) if layout == _lib.CHANNEL_LAST else tuple(smaps.shape[2:])
        if sm_spatial != spatial:
            raise ValueError(f"smaps spatial size {sm_spatial} does not match the image size {spatial}")
        if smaps.dtype != x.dtype:
            raise TypeError(f"smaps dtype {smaps.dtype} does not match {what} dtype {x.dtype}")
        if smaps.device != x.device:
            raise ValueError(f"smaps is on {smaps.device} but {what} is on {x.device}")
        if smaps.shape[0] not in (1, x.shape[0]):
            raise ValueError(f"smaps batch size {smaps.shape[0]} must be 1 or match the batch size {x.shape[0]}")


def apod_pad(image: Tensor, grid_size: Sequence[int], smaps: Optional[Tensor] = None,
             scaling_coef: Optional[Tensor] = None, scale: float = 1.0, n_coils: Optional[int] = None,
             layout: int = _lib.COIL_MAJOR) -> Tensor:
    """``zero_pad_end(image * smaps * scaling_coef) * scale`` in one pass.

    image ``(B, Ci, *N)`` with ``Ci in {1, C}``; smaps ``(Bs, C, *N)`` (channel-last
    ``(Bs, *N, C)`` when ``layout`` is channel-last); returns the grid ``(B, C, *K)``
    or ``(B, *K, C)``."""
    require_cuda(image, "image")
    lib = _lib.load()
    image = dense(image)
    B, Ci = image.shape[:2]
    im_size = list(image.shape[2:])
    grid_size = _sizes(grid_size)
    _check_operands("image", image, im_size, smaps, scaling_coef, layout)
    if smaps is not None:
        smaps = dense(smaps)
        C = smaps.shape[-1] if layout == _lib.CHANNEL_LAST else smaps.shape[1]
        Bs = smaps.shape[0]
    else:
        C = Ci if n_coils is None else n_coils
        Bs = 1
    shape = [B] + grid_size + [C] if layout == _lib.CHANNEL_LAST else [B, C] + grid_size
    out = torch.empty(shape, dtype=image.dtype, device=image.device)
    if out.numel() == 0:
        return out
    if scaling_coef is not None:
        scaling_coef = dense(scaling_coef)
    with device_guard(image.device):
        _lib.check(
            lib.b2n_apod_pad(len(im_size), engine_dtype(image.dtype), _lib.i64_array(im_size), _lib.i64_array(grid_size),
                             B, C, image.data_ptr(), Ci, smaps.data_ptr() if smaps is not None else None, Bs,
                             scaling_coef.data_ptr() if scaling_coef is not None else None, float(scale), layout,
                             out.data_ptr(), current_stream_ptr(image.device)),
            "b2n_apod_pad",
        )
    return out


def crop_apod_coilsum(grid: Tensor, im_size: Sequence[int], smaps: Optional[Tensor] = None,
                      scaling_coef: Optional[Tensor] = None, scale: float = 1.0,
                      layout: int = _lib.COIL_MAJOR) -> Tensor:
    """``sum_c crop(grid) * conj(scaling_coef) * conj(smaps) * scale`` in one pass;
    without smaps the coil axis is kept.  Returns ``(B, 1 or C, *N)``."""
    require_cuda(grid, "grid")
    lib = _lib.load()
    grid = dense(grid)
    im_size = _sizes(im_size)
    B = grid.shape[0]
    if layout == _lib.CHANNEL_LAST:
        C, grid_size = grid.shape[-1], list(grid.shape[1:-1])
    else:
        C, grid_size = grid.shape[1], list(grid.shape[2:])
    Bs = 1
    _check_operands("grid", grid, im_size, smaps, scaling_coef, layout)
    if smaps is not None:
        smaps = dense(smaps)
        Bs = smaps.shape[0]
        if (smaps.shape[-1] if layout == _lib.CHANNEL_LAST else smaps.shape[1]) != C:
            raise ValueError("smaps and grid disagree on the number of coils")
    out = torch.empty([B, 1 if smaps is not None else C] + im_size, dtype=grid.dtype, device=grid.device)
    if out.numel() == 0:
        return out
    if scaling_coef is not None:
        scaling_coef = dense(scaling_coef)
    with device_guard(grid.device):
        _lib.check(
            lib.b2n_crop_apod_coilsum(len(im_size), engine_dtype(grid.dtype), _lib.i64_array(im_size),
                                      _lib.i64_array(grid_size), B, C, grid.data_ptr(), layout,
                                      smaps.data_ptr() if smaps is not None else None, Bs,
                                      scaling_coef.data_ptr() if scaling_coef is not None else None, float(scale),
                                      out.data_ptr(), current_stream_ptr(grid.device)),
            "b2n_crop_apod_coilsum",
        )
    return out


def _check_kernel(kernel: Tensor, x: Tensor, grid_size: Sequence[int]) -> None:
    """Toeplitz kernel ``(*K)`` or per-batch ``(B, *K)`` of the data's dtype on the data's device."""
    grid_size = tuple(int(k) for k in grid_size)
    nd = len(grid_size)
    if kernel.ndim not in (nd, nd + 1) or tuple(kernel.shape[-nd:]) != grid_size:
        raise ValueError(f"kernel shape {tuple(kernel.shape)} does not match the grid {grid_size}")
    if kernel.ndim == nd + 1 and kernel.shape[0] not in (1, x.shape[0]):
        raise ValueError(f"kernel batch size {kernel.shape[0]} must be 1 or match the batch size {x.shape[0]}")
    if kernel.dtype != x.dtype:
        raise TypeError(f"kernel dtype {kernel.dtype} does not match data dtype {x.dtype}")
    if kernel.device != x.device:
        raise ValueError(f"kernel is on {kernel.device} but the data is on {x.device}")


def spectrum_mul_(spectrum: Tensor, kernel: Tensor, scale: float = 1.0, layout: int = _lib.COIL_MAJOR) -> Tensor:
    """In-place ``spectrum[b, c] *= kernel[b or 0] * scale`` (Toeplitz filter)."""
    require_cuda(spectrum, "spectrum")
    assert spectrum.is_contiguous() and not spectrum.is_conj()
    if kernel.dtype != spectrum.dtype or kernel.device != spectrum.device:
        raise TypeError("kernel must have the spectrum's dtype and device")
    kernel = dense(kernel)
    B = spectrum.shape[0]
    C = spectrum.shape[-1] if layout == _lib.CHANNEL_LAST else spectrum.shape[1]
    n_grid = spectrum.numel() // max(B * C, 1)
    kb = kernel.numel() // max(n_grid, 1)
    if spectrum.numel() == 0:
        return spectrum
    with device_guard(spectrum.device):
        _lib.check(
            _lib.load().b2n_spectrum_mul(engine_dtype(spectrum.dtype), spectrum.data_ptr(), kernel.data_ptr(), B, C,
                                         n_grid, kb, layout, float(scale), current_stream_ptr(spectrum.device)),
            "b2n_spectrum_mul",
        )
    return spectrum


def fft_grid(grid: Tensor, ndim: int, inverse: bool) -> Tensor:
    """Unnormalised cuFFT transform over the last ``ndim`` axes of a coil-major grid.
    (``norm='backward'`` on the forward and ``norm='forward'`` on the inverse both
    mean "no scaling pass"; reference: fft_fn / ifft_fn, ``_nufft/fft.py:9-22``.)"""
    dims = list(range(-ndim, 0))
    if inverse:
        return torch.fft.ifftn(grid, dim=dims, norm="forward")
    return torch.fft.fftn(grid, dim=dims, norm="backward")


def ortho_scale(grid_size: Sequence[int], normalized: bool) -> float:
    if not normalized:
        return 1.0
    n = 1
    for k in grid_size:
        n *= int(k)
    return 1.0 / math.sqrt(n)


# ---------------------------------------------------------------------------------------
# pruned, fused FFT passes (own Stockham kernels, complex64): see csrc/b2n_fft.cu
# ---------------------------------------------------------------------------------------
_FFT_SUPPORTED: dict = {}
# "auto" (default): own passes when every grid length has a compile-time plan (b2n_fft_fast.cuh; measured faster
# than cuFFT + element-wise kernels, profiles/r01_g); True: whenever the lengths factor into primes <= 13 (the
# run-time Stockham passes are slower than cuFFT, kept for A/B); False: always cuFFT.
use_fused_fft = "auto"


def fused_fft_available(dtype: torch.dtype, grid_size: Sequence[int], n_rows: int = 1) -> bool:
    """True when the engine's own FFT passes are to be used for this transform (complex64 only;
    ``n_rows`` = batch x coils, the passes index the whole grid stack with 32 bits); otherwise callers
    use cuFFT around the fused element-wise kernels."""
    if not use_fused_fft or dtype != torch.complex64:
        return False
    if n_rows * math.prod(int(k) for k in grid_size) >= 2 ** 31:
        return False
    lib = _lib.load()
    need = 2 if use_fused_fft == "auto" else 1
    for n in grid_size:
        n = int(n)
        if n not in _FFT_SUPPORTED:
            _FFT_SUPPORTED[n] = int(lib.b2n_fft_supported(n))  # 0 unsupported, 1 run-time passes, 2 compile-time plan
        if _FFT_SUPPORTED[n] < need:
            return False
    return True


_TWIDDLES: dict = {}


_TWIDDLE_SETS: dict = {}


def _twiddles(grid_size, device):
    """Per-(device, length) twiddle tables (2n entries: plain table + the staged tables of the
    compile-time plan, see ``b2n_fft_twiddles``), built once on the device; returns the tables
    and the ``void*[ndim]`` argument (cached per size tuple)."""
    key = (device.index, tuple(grid_size))
    hit = _TWIDDLE_SETS.get(key)
    if hit is not None:
        return hit
    lib = _lib.load()
    tabs = []
    for n in grid_size:
        tkey = (device.index, int(n))
        t = _TWIDDLES.get(tkey)
        if t is None:
            if torch.cuda.is_current_stream_capturing():
                # the fill kernel would be recorded instead of run, and later eager calls would read zeros
                raise RuntimeError(
                    f"FFT twiddle tables for length {int(n)} must be built before CUDA-graph capture: run the "
                    "operator once eagerly (warm-up) on this device before capturing it")
            t = torch.zeros(2 * int(n), dtype=torch.complex64, device=device)
            with device_guard(device):
                _lib.check(lib.b2n_fft_twiddles(int(n), t.data_ptr(), current_stream_ptr(device)), "b2n_fft_twiddles")
            _TWIDDLES[tkey] = t
        tabs.append(t)
    ptrs = (ctypes.c_void_p * len(tabs))(*[t.data_ptr() for t in tabs])
    _TWIDDLE_SETS[key] = (tabs, ptrs)
    return tabs, ptrs


_WORK_BYTES: dict = {}


def _fft_work(im_size, grid_size, B, C, device) -> Optional[Tensor]:
    key = (tuple(im_size), tuple(grid_size), B, C)
    n = _WORK_BYTES.get(key)
    if n is None:
        nbytes = ctypes.c_size_t(0)
        _lib.check(_lib.load().b2n_fft_work_bytes(len(im_size), _lib.i64_array(im_size), _lib.i64_array(grid_size), B, C,
                                                  ctypes.byref(nbytes)), "b2n_fft_work_bytes")
        n = _WORK_BYTES[key] = int(nbytes.value)
    return torch.empty(n, dtype=torch.uint8, device=device) if n else None


_FFT_CTX: dict = {}


def _fft_ctx(im_size, grid_size, B, C, device):
    """Per (sizes, batch, coils, device): the ctypes size arrays, twiddle pointer array and scratch size --
    rebuilt arguments cost more host time than some of these kernels run."""
    key = (tuple(im_size), tuple(grid_size), B, C, device.index)
    ctx = _FFT_CTX.get(key)
    if ctx is None:
        n = _WORK_BYTES.get((tuple(im_size), tuple(grid_size), B, C))
        if n is None:
            nbytes = ctypes.c_size_t(0)
            _lib.check(_lib.load().b2n_fft_work_bytes(len(im_size), _lib.i64_array(im_size), _lib.i64_array(grid_size),
                                                      B, C, ctypes.byref(nbytes)), "b2n_fft_work_bytes")
            n = _WORK_BYTES[(tuple(im_size), tuple(grid_size), B, C)] = int(nbytes.value)
        tabs, tw = _twiddles(grid_size, device)
        if len(_FFT_CTX) > 256:
            _FFT_CTX.clear()
        ctx = _FFT_CTX[key] = (_lib.i64_array(im_size), _lib.i64_array(grid_size), tw, n, tabs)
    return ctx


def fused_fft_forward(image: Tensor, grid_size: Sequence[int], smaps: Optional[Tensor] = None,
                      scaling_coef: Optional[Tensor] = None, scale: float = 1.0,
                      n_coils: Optional[int] = None) -> Tensor:
    """``fftn(zero_pad(image * smaps * scaling_coef)) * scale`` (unnormalised transform) in
    ``ndim`` pruned passes; returns the coil-major grid ``(B, C, *K)``."""
    require_cuda(image, "image")
    image = dense(image)
    B, Ci = image.shape[:2]
    im_size, grid_size = list(image.shape[2:]), _sizes(grid_size)
    _check_operands("image", image, im_size, smaps, scaling_coef)
    Bs = 1
    if smaps is not None:
        smaps = dense(smaps)
        C, Bs = smaps.shape[1], smaps.shape[0]
    else:
        C = Ci if n_coils is None else n_coils
    out = torch.empty([B, C] + grid_size, dtype=image.dtype, device=image.device)
    if out.numel() == 0:
        return out
    if scaling_coef is not None:
        scaling_coef = dense(scaling_coef)
    n_arr, k_arr, tw, nwork, _tabs = _fft_ctx(im_size, grid_size, B, C, image.device)
    work = torch.empty(nwork, dtype=torch.uint8, device=image.device) if nwork else None
    with device_guard(image.device):
        _lib.check(
            _lib.load().b2n_fft_forward_fused(
                len(im_size), n_arr, k_arr, B, C, image.data_ptr(), Ci,
                smaps.data_ptr() if smaps is not None else None, Bs,
                scaling_coef.data_ptr() if scaling_coef is not None else None, float(scale), tw, out.data_ptr(),
                work.data_ptr() if work is not None else None, current_stream_ptr(image.device)),
            "b2n_fft_forward_fused",
        )
    return out


def fused_fft_adjoint(grid: Tensor, im_size: Sequence[int], smaps: Optional[Tensor] = None,
                      scaling_coef: Optional[Tensor] = None, scale: float = 1.0,
                      kernel: Optional[Tensor] = None, peer_comm=None) -> Tensor:
    """``sum_c crop(ifftn_unnormalised(grid * kernel)) * conj(scaling_coef) * conj(smaps) * scale``
    in ``ndim`` pruned passes; ``kernel`` (Toeplitz, ``(*K)`` or ``(B, *K)``) is optional.
    Returns ``(B, 1, *N)`` with smaps, ``(B, C, *N)`` without.  ``peer_comm`` (a ``_lib.PeerComm``, see
    ``parallel.PeerAllReduce``): the result is additionally summed over the ranks of the communicator, inside the
    last pass where that pass can carry the exchange (``b2n_fft_adjoint_fused_allreduce``)."""
    require_cuda(grid, "grid")
    grid = dense(grid)
    im_size = _sizes(im_size)
    B, C = grid.shape[:2]
    grid_size = list(grid.shape[2:])
    Bs = 1
    _check_operands("grid", grid, im_size, smaps, scaling_coef)
    if smaps is not None:
        smaps = dense(smaps)
        Bs = smaps.shape[0]
        if smaps.shape[1] != C:
            raise ValueError("smaps and grid disagree on the number of coils")
    out = torch.empty([B, 1 if smaps is not None else C] + im_size, dtype=grid.dtype, device=grid.device)
    if out.numel() == 0:
        return out
    if scaling_coef is not None:
        scaling_coef = dense(scaling_coef)
    kb = 1
    if kernel is not None:
        _check_kernel(kernel, grid, grid_size)
        kernel = dense(kernel)
        kb = kernel.shape[0] if kernel.ndim > len(grid_size) else 1
    n_arr, k_arr, tw, nwork, _tabs = _fft_ctx(im_size, grid_size, B, C, grid.device)
    work = torch.empty(nwork, dtype=torch.uint8, device=grid.device) if nwork else None
    with device_guard(grid.device):
        args = (len(im_size), n_arr, k_arr, B, C, grid.data_ptr(),
                kernel.data_ptr() if kernel is not None else None, kb,
                smaps.data_ptr() if smaps is not None else None, Bs,
                scaling_coef.data_ptr() if scaling_coef is not None else None, float(scale), tw, out.data_ptr(),
                work.data_ptr() if work is not None else None)
        if peer_comm is None:
            _lib.check(_lib.load().b2n_fft_adjoint_fused(*args, current_stream_ptr(grid.device)), "b2n_fft_adjoint_fused")
        else:
            import ctypes

            _lib.check(_lib.load().b2n_fft_adjoint_fused_allreduce(*args, ctypes.byref(peer_comm),
                                                                   current_stream_ptr(grid.device)),
                       "b2n_fft_adjoint_fused_allreduce")
    return out


def toeplitz_fused_available(dtype: torch.dtype, im_size: Sequence[int], grid_size: Sequence[int]) -> bool:
    """True when ``b2n_fft_toeplitz_fused`` takes the shape: 2-D complex64, compile-time planned column length,
    grid at least twice the image along the slow dimension (what ``calc_toeplitz_kernel`` produces)."""
    if use_fused_fft is False or dtype != torch.complex64 or len(grid_size) != 2:
        return False
    k0, k1 = int(grid_size[0]), int(grid_size[1])
    if 2 * int(im_size[0]) > k0 or k1 % 2 or 4 * 16 * (2 * k0 + k0 // 8 + 1) > 227 * 1024:
        return False
    lib = _lib.load()
    return lib.b2n_fft_supported(k0) == 2 and lib.b2n_fft_supported(k1) > 0


def fused_toeplitz(image: Tensor, kernel: Tensor, smaps: Optional[Tensor] = None, scale: float = 1.0) -> Tensor:
    """``scale * sum_c conj(S_c) crop(ifft2(kernel * fft2(zero_pad(S_c * image))))`` (unnormalised transforms) in
    three passes; the column pass transforms, filters and transforms back without writing the spectrum.
    Returns ``(B, 1, *N)`` with smaps, ``(B, C, *N)`` without."""
    require_cuda(image, "image")
    image = dense(image)
    kernel = dense(kernel)
    B, Ci = image.shape[:2]
    im_size = list(image.shape[2:])
    grid_size = list(kernel.shape[-2:])
    kb = kernel.shape[0] if kernel.ndim > 2 else 1
    _check_operands("image", image, im_size, smaps, None)
    _check_kernel(kernel, image, grid_size)
    Bs, C = 1, Ci
    if smaps is not None:
        smaps = dense(smaps)
        C, Bs = smaps.shape[1], smaps.shape[0]
    out = torch.empty([B, 1 if smaps is not None else C] + im_size, dtype=image.dtype, device=image.device)
    if out.numel() == 0:
        return out
    n_arr, k_arr, tw, nwork, _tabs = _fft_ctx(im_size, grid_size, B, C, image.device)
    work = torch.empty(nwork, dtype=torch.uint8, device=image.device)
    with device_guard(image.device):
        _lib.check(
            _lib.load().b2n_fft_toeplitz_fused(
                2, n_arr, k_arr, B, C, image.data_ptr(), Ci, smaps.data_ptr() if smaps is not None else None, Bs,
                kernel.data_ptr(), kb, float(scale), tw, out.data_ptr(), work.data_ptr(),
                current_stream_ptr(image.device)),
            "b2n_fft_toeplitz_fused",
        )
    return out
