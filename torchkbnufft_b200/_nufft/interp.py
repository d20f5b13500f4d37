"""Table interpolation backends: host entry points into the sm_100a kernels.

Same call signatures and error behaviour as the reference's
``torchkbnufft/_nufft/interp.py`` (``table_interp`` :315-403,
``table_interp_adjoint`` :587-726) so the autograd Functions and everything above
them are drop-in; the work itself is one kernel launch per direction on a cached
trajectory plan instead of ``prod(J)`` passes of ATen index/mul/add ops.

There is deliberately no CPU implementation here (see ``plan.require_cuda``).
"""
from __future__ import annotations

import ctypes
import os
from typing import List, Optional

import torch
from torch import Tensor

from .. import _lib
from .plan import (current_stream_ptr, dense, device_guard, get_geometry, get_plan, normalize_omega, omega_for,
                   require_cuda)

ADJOINT_MODES = {"atomic": _lib.ADJ_ATOMIC, "sorted": _lib.ADJ_SORTED}
_default_adjoint_mode = os.environ.get("B200NUFFT_ADJOINT_MODE", "atomic")


def set_adjoint_mode(mode: str) -> None:
    """Select the adjoint accumulation mode: ``"atomic"`` (fastest; like
    ``index_add_`` on CUDA the summation order is not reproducible) or ``"sorted"``
    (deterministic: every grid cell gathers its samples in a fixed order)."""
    global _default_adjoint_mode
    if mode not in ADJOINT_MODES:
        raise ValueError(f"adjoint mode must be one of {sorted(ADJOINT_MODES)}")
    _default_adjoint_mode = mode


def get_adjoint_mode() -> str:
    return _default_adjoint_mode


def set_tiled_kernels(enabled) -> None:
    """``True`` (default): shared-memory tiled kernels where they apply and pay off (a single 2-D
    (batch, coil) row goes to the per-point kernels, which are faster there); ``"force"``: tiled
    wherever they apply; ``False``: per-point kernels only.  For A/B measurements and tests."""
    value = 2 if enabled == "force" else (1 if enabled else 0)
    _lib.check(_lib.load().b2n_set_option(_lib.OPT_TILED_KERNELS, value), "b2n_set_option")


def get_tiled_kernels() -> bool:
    return bool(_lib.load().b2n_get_option(_lib.OPT_TILED_KERNELS))


class KernelTimer:
    """Optional CUDA-event bracket around the two interpolation launches, for
    benchmarks (``bench.py`` installs one during its timed region)."""

    def __init__(self):
        self.events = {}

    def bracket(self, name: str, device):
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.events.setdefault(name, []).append((start, stop))
        start.record(torch.cuda.current_stream(device))
        return stop

    def mean_ms(self, name: str) -> float:
        pairs = self.events.get(name, [])
        return sum(a.elapsed_time(b) for a, b in pairs) / max(len(pairs), 1)


kernel_timer: Optional[KernelTimer] = None

# owner-tile (output-stationary, register-accumulator) spread: on by default; rows = batch x coils
owned_spread = True
owned_spread_min_rows = 1


def _check_offsets(offsets: Optional[Tensor], n_offsets: int, ndim: int) -> None:
    # The engine always visits the full row-major neighbourhood (the only thing the
    # reference's modules ever pass, _nufft/utils.py:329); a custom subset is rejected.
    if offsets is not None and tuple(offsets.shape) != (n_offsets, ndim):
        raise ValueError(
            f"offsets must list all {n_offsets} neighbour offsets (shape {(n_offsets, ndim)}), got {tuple(offsets.shape)}"
        )


def lookup_plan(omega: Tensor, n_batch: int, tables: List[Tensor], n_shift: Tensor, numpoints: Tensor,
                table_oversamp: Tensor, grid_size, device) -> TrajectoryPlan:
    """The cached trajectory plan an interpolation call with these operator buffers uses (built if missing)."""
    omega = normalize_omega(omega, n_batch, "image")
    geo = get_geometry(tables, n_shift, numpoints, table_oversamp, tuple(grid_size))
    return get_plan(geo, omega_for(omega, geo.cdtype, device, "image"))


def table_interp(
    image: Tensor,
    omega: Tensor,
    tables: List[Tensor],
    n_shift: Tensor,
    numpoints: Tensor,
    table_oversamp: Tensor,
    offsets: Optional[Tensor] = None,
    min_kspace_per_fork: int = 1024,
    layout: int = _lib.COIL_MAJOR,
    n_coils: Optional[int] = None,
) -> Tensor:
    """Interpolate gridded data ``image (B, C, *K)`` to the off-grid locations
    ``omega`` (``(d, M)`` or ``(B, d, M)``, radians/voxel); returns ``(B, C, M)``.

    ``min_kspace_per_fork`` is accepted for signature parity and ignored (it tunes
    the reference's CPU thread forking).  ``layout``/``n_coils`` are engine
    extensions used by the fused NUFFT path for channel-last grids ``(B, *K, C)``.
    """
    require_cuda(image, "image")
    require_cuda(omega, "omega")
    if not image.is_complex():
        raise TypeError("image must be complex (use the functional API for real views).")
    omega = normalize_omega(omega, image.shape[0], "image")
    if layout == _lib.CHANNEL_LAST:
        grid_size = image.shape[1:-1]
        C = image.shape[-1]
    else:
        grid_size = image.shape[2:]
        C = image.shape[1]
    geo = get_geometry(tables, n_shift, numpoints, table_oversamp, tuple(grid_size))
    if geo.cdtype != image.dtype:
        raise TypeError(f"image dtype {image.dtype} does not match table dtype {geo.cdtype}")
    if omega.shape[-2] != geo.ndim:
        raise ValueError(f"omega has {omega.shape[-2]} coordinate rows for a {geo.ndim}-D grid")
    _check_offsets(offsets, geo.n_offsets, geo.ndim)
    plan = get_plan(geo, omega_for(omega, geo.cdtype, image.device, "image"))
    B = image.shape[0]
    image = dense(image)
    out = torch.empty((B, C, plan.n_points), dtype=image.dtype, device=image.device)
    if out.numel() == 0:
        return out
    with device_guard(image.device):
        stop = kernel_timer.bracket("interp_fwd", image.device) if kernel_timer is not None else None
        _lib.check(
            _lib.load().b2n_interp_forward(ctypes.byref(geo.struct), ctypes.byref(plan.struct), image.data_ptr(), B, C,
                                           layout, out.data_ptr(), current_stream_ptr(image.device)),
            "b2n_interp_forward",
        )
        if stop is not None:
            stop.record(torch.cuda.current_stream(image.device))
    return out


def table_interp_adjoint(
    data: Tensor,
    omega: Tensor,
    tables: List[Tensor],
    n_shift: Tensor,
    numpoints: Tensor,
    table_oversamp: Tensor,
    offsets: Optional[Tensor],
    grid_size: Tensor,
    layout: int = _lib.COIL_MAJOR,
    mode: Optional[str] = None,
) -> Tensor:
    """Spread off-grid ``data (B, C, M)`` at ``omega`` onto the grid; returns
    ``(B, C, *grid_size)`` (or ``(B, *grid_size, C)`` for the channel-last layout)."""
    require_cuda(data, "data")
    require_cuda(omega, "omega")
    if not data.is_complex():
        raise TypeError("data must be complex (use the functional API for real views).")
    omega = normalize_omega(omega, data.shape[0], "data")
    geo = get_geometry(tables, n_shift, numpoints, table_oversamp, grid_size)
    if geo.cdtype != data.dtype:
        raise TypeError(f"data dtype {data.dtype} does not match table dtype {geo.cdtype}")
    if omega.shape[-2] != geo.ndim:
        raise ValueError(f"omega has {omega.shape[-2]} coordinate rows for a {geo.ndim}-D grid")
    if omega.shape[-1] != data.shape[-1]:
        raise ValueError("omega and data disagree on the number of k-space samples")
    _check_offsets(offsets, geo.n_offsets, geo.ndim)
    plan = get_plan(geo, omega_for(omega, geo.cdtype, data.device, "data"))
    B, C = data.shape[:2]
    data = dense(data)
    if layout == _lib.CHANNEL_LAST:
        shape = [B] + geo.grid_size + [C]
    else:
        shape = [B, C] + geo.grid_size
    if data.numel() == 0 or plan.n_points == 0:
        return torch.zeros(shape, dtype=data.dtype, device=data.device)  # nothing to spread
    out = torch.empty(shape, dtype=data.dtype, device=data.device)
    if out.numel() == 0:
        return out
    mode_id = ADJOINT_MODES[_default_adjoint_mode if mode is None else mode]
    lib = _lib.load()
    with device_guard(data.device):
        stop = kernel_timer.bracket("interp_adj", data.device) if kernel_timer is not None else None
        done = False
        # The output-stationary owner-tile spread (2-D complex64 J = 6 plans) is deterministic AND the fastest kernel
        # from `owned_spread_min_rows` (batch x coil) rows on, so it serves both modes; it keeps a small persistent
        # scratch (arrival counters + partial tiles of the dense tiles) with the plan.
        if (plan.struct.own_tile and layout == _lib.COIL_MAJOR and owned_spread and
                (mode_id == _lib.ADJ_SORTED or B * C >= owned_spread_min_rows)):
            scratch, nbytes = plan.own_scratch(lib, geo, B, C, layout)
            if scratch is not None:
                _lib.check(
                    lib.b2n_interp_adjoint_ordered(ctypes.byref(geo.struct), ctypes.byref(plan.struct), data.data_ptr(),
                                                   B, C, layout, scratch.data_ptr(), nbytes, out.data_ptr(),
                                                   current_stream_ptr(data.device)),
                    "b2n_interp_adjoint_ordered",
                )
                done = True
        scratch_bytes = ctypes.c_size_t(0)
        if not done and mode_id == _lib.ADJ_SORTED:
            # deterministic mode elsewhere: the tiled kernels with per-sub-problem scratch tiles and a fixed-order
            # merge where they apply (3-D complex64, J = 6), else the per-cell gather
            zero_bytes = ctypes.c_size_t(0)
            _lib.check(lib.b2n_interp_adjoint_ordered_layout(ctypes.byref(geo.struct), ctypes.byref(plan.struct), B, C,
                                                             layout, 0, ctypes.byref(scratch_bytes),
                                                             ctypes.byref(zero_bytes)),
                       "b2n_interp_adjoint_ordered_layout")
            if zero_bytes.value:  # owner-tile path switched off in Python but on in the library: needs zeroed counters
                scratch_bytes = ctypes.c_size_t(0)
        if done:
            pass
        elif scratch_bytes.value:
            scratch = torch.empty(scratch_bytes.value, dtype=torch.uint8, device=data.device)
            _lib.check(
                lib.b2n_interp_adjoint_ordered(ctypes.byref(geo.struct), ctypes.byref(plan.struct), data.data_ptr(), B,
                                               C, layout, scratch.data_ptr(), scratch_bytes.value, out.data_ptr(),
                                               current_stream_ptr(data.device)),
                "b2n_interp_adjoint_ordered",
            )
        else:
            _lib.check(
                lib.b2n_interp_adjoint(ctypes.byref(geo.struct), ctypes.byref(plan.struct), data.data_ptr(), B, C,
                                       layout, mode_id, out.data_ptr(), current_stream_ptr(data.device)),
                "b2n_interp_adjoint",
            )
        if stop is not None:
            stop.record(torch.cuda.current_stream(data.device))
    return out


def export_indices(omega: Tensor, tables, n_shift, numpoints, table_oversamp, grid_size):
    """Parity/debug export of the integer indices the reference's
    ``calc_coef_and_indices`` (``_nufft/interp.py:89-150``) produces for every
    neighbour offset: ``arr_ind (W, M)`` int64 and ``tab_idx (W, d, M)`` int32."""
    require_cuda(omega, "omega")
    if omega.ndim != 2:
        raise ValueError("export_indices takes a single (d, M) trajectory")
    geo = get_geometry(tables, n_shift, numpoints, table_oversamp, grid_size)
    if omega.shape[0] != geo.ndim:
        raise ValueError(f"omega has {omega.shape[0]} coordinate rows for a {geo.ndim}-D grid")
    om = dense(omega_for(omega, geo.cdtype, geo.device, "the tables"))
    M = om.shape[1]
    arr_ind = torch.empty((geo.n_offsets, M), dtype=torch.int64, device=om.device)
    tab_idx = torch.empty((geo.n_offsets, geo.ndim, M), dtype=torch.int32, device=om.device)
    with device_guard(om.device):
        _lib.check(
            _lib.load().b2n_export_indices(ctypes.byref(geo.struct), om.data_ptr(), M, arr_ind.data_ptr(),
                                           tab_idx.data_ptr(), current_stream_ptr(om.device)),
            "b2n_export_indices",
        )
    return arr_ind, tab_idx
