"""Operator geometry and trajectory plans (host side of ``b2n_geom`` / ``b2n_points``).

The reference recomputes the normalised coordinates, table lookups, the point sort
and the fftshift phase on every call (``torchkbnufft/_nufft/interp.py:171-177``,
``:129-148``, ``:552-584``, ``:663-686``).  Iterative reconstructions call the
operator many times on one trajectory, so the engine builds that state once per
``(geometry, omega)`` pair and keeps it in a small LRU cache.  Cache entries hold
strong references to the tensors they were built from, so a ``data_ptr`` cannot be
recycled under a live entry; in-place edits are seen through the tensor version.
"""
from __future__ import annotations

import ctypes
import threading
from collections import OrderedDict
from typing import List, Optional, Sequence, Tuple

import torch
from torch import Tensor

from .. import _lib

_GEOM_CACHE: "OrderedDict[tuple, Geometry]" = OrderedDict()
_PLAN_CACHE: "OrderedDict[tuple, TrajectoryPlan]" = OrderedDict()
_GRID_SIZE_CACHE: dict = {}
GEOM_CACHE_SIZE = 32
PLAN_CACHE_SIZE = 8
PLAN_CACHE_BYTES = 16 << 30  # device memory the cached plans may hold together (the newest plan always stays)
# every mutation of the process-global caches below happens under this lock (threaded / DataParallel callers)
_LOCK = threading.RLock()
# "version" (default): plans are keyed on (storage pointer, tensor version, shape, stride): in-place edits through
#   torch are seen, writes that bypass the version counter (DLPack consumers, custom kernels, ``omega.data``) are not;
# "content": additionally compare the trajectory with the copy the plan was built from (one device comparison and a
#   host synchronisation per call -- for callers that rewrite trajectories behind torch's back);
# "off": rebuild the plan on every call.
_plan_cache_mode = "version"


def set_plan_cache_mode(mode: str) -> None:
    """``"version"`` (default), ``"content"`` or ``"off"`` -- see the comment above ``_plan_cache_mode``."""
    global _plan_cache_mode
    if mode not in ("version", "content", "off"):
        raise ValueError("plan cache mode must be 'version', 'content' or 'off'")
    _plan_cache_mode = mode


def get_plan_cache_mode() -> str:
    return _plan_cache_mode


def dense(t: Tensor) -> Tensor:
    """The tensor as plain contiguous memory, ready for ``data_ptr()``: lazy conjugate / negative bits are
    materialised (``.contiguous()`` alone keeps them, and a raw pointer ignores them)."""
    if t.is_conj():
        t = t.resolve_conj()
    if t.is_neg():
        t = t.resolve_neg()
    return t if t.is_contiguous() else t.contiguous()


REAL_OF = {torch.complex64: torch.float32, torch.complex128: torch.float64}
_OMEGA_CAST: "OrderedDict[tuple, tuple]" = OrderedDict()


def omega_for(omega: Tensor, cdtype: torch.dtype, device: torch.device, what: str) -> Tensor:
    """The trajectory in the real dtype the kernels read (float32 with complex64 tables, float64 with complex128):
    the C side reinterprets the buffer by the TABLE dtype, so a mismatched omega must be converted, never passed
    through.  Conversions are cached per (tensor, version) so that the trajectory plan keyed on the result is
    reused across calls.  Raises when omega lives on another device than ``what``."""
    if omega.device != device:
        raise ValueError(f"omega is on {omega.device} but {what} is on {device}")
    want = REAL_OF[cdtype]
    if omega.dtype == want:
        return omega
    if not omega.dtype.is_floating_point:
        raise TypeError(f"omega must be a real floating-point tensor, got {omega.dtype}")
    key = (id(omega), omega._version, want)
    with _LOCK:
        hit = _OMEGA_CAST.get(key)
        if hit is not None and hit[0] is omega:
            _OMEGA_CAST.move_to_end(key)
            return hit[1]
        cast = omega.to(want)
        _OMEGA_CAST[key] = (omega, cast)
        while len(_OMEGA_CAST) > PLAN_CACHE_SIZE:
            _OMEGA_CAST.popitem(last=False)
    return cast


def require_cuda(t: Tensor, what: str) -> None:
    """The engine has no CPU path; refuse CPU tensors loudly."""
    if not t.is_cuda:
        raise RuntimeError(
            f"torchkbnufft_b200 is a CUDA-only (sm_100a) engine with no CPU fallback: `{what}` is on "
            f"{t.device}. Move the module and its inputs to a CUDA device."
        )


def engine_dtype(cdtype: torch.dtype) -> int:
    if cdtype == torch.complex64:
        return _lib.C64
    if cdtype == torch.complex128:
        return _lib.C128
    raise TypeError(f"unsupported complex dtype {cdtype}")


def current_stream_ptr(device: torch.device) -> ctypes.c_void_p:
    index = device.index if device.index is not None else torch.cuda.current_device()
    return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(index))


class _NoGuard:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NO_GUARD = _NoGuard()


def device_guard(device: torch.device):
    """``torch.cuda.device(device)`` only when ``device`` is not already current (the guard costs
    several microseconds of host time per call, which matters next to 20-100 us kernels)."""
    if device.index is None or device.index == torch.cuda.current_device():
        return _NO_GUARD
    return torch.cuda.device(device)


def _tkey(t: Tensor) -> tuple:
    return (t.data_ptr(), t._version, tuple(t.shape), tuple(t.stride()), t.dtype, t.device.index)


class Geometry:
    """Host copy of the operator buffers + the ``b2n_geom`` struct."""

    def __init__(self, tables: Sequence[Tensor], n_shift: Tensor, numpoints: Tensor, table_oversamp: Tensor,
                 grid_size: Sequence[int]):
        self.ndim = len(tables)
        if not 1 <= self.ndim <= _lib.MAX_DIMS:
            raise ValueError(f"only 1-3 dimensions are supported, got {self.ndim}")
        self.cdtype = tables[0].dtype
        self.device = tables[0].device
        # one host read of the small buffers per geometry (cached afterwards)
        self.numpoints = [int(v) for v in numpoints.tolist()]
        self.table_oversamp = [int(v) for v in table_oversamp.tolist()]
        self.n_shift = [float(v) for v in n_shift.tolist()]
        self.grid_size = [int(v) for v in grid_size]
        if not (len(self.numpoints) == len(self.table_oversamp) == len(self.n_shift) == len(self.grid_size)
                == self.ndim):
            raise ValueError("tables, n_shift, numpoints, table_oversamp and grid_size must agree in length")
        # private copies (a few KB each): the geometry is shared by every operator with the same contents (see
        # get_geometry), so nobody else may hold a handle that can rewrite its tables
        self.tables = [dense(t).detach().clone() for t in tables]
        self.content_key: tuple = ()
        self.n_grid = 1
        for k in self.grid_size:
            self.n_grid *= k
        self.n_offsets = 1
        for j in self.numpoints:
            self.n_offsets *= j
        g = _lib.Geom()
        g.ndim = self.ndim
        g.dtype = engine_dtype(self.cdtype)
        for d in range(self.ndim):
            g.grid_size[d] = self.grid_size[d]
            g.numpoints[d] = self.numpoints[d]
            g.table_oversamp[d] = self.table_oversamp[d]
            g.table_len[d] = self.tables[d].numel()
            g.table_dev[d] = self.tables[d].data_ptr()
            g.n_shift[d] = self.n_shift[d]
        # real kernel + linear phase form of the tables (b2n_geom.rtable_dev): checked once per geometry on the host
        self.rtables = real_table_form(self.tables, self.numpoints, self.table_oversamp, self.grid_size)
        if self.rtables is not None:
            for d, (rt, slope) in enumerate(self.rtables):
                g.rtable_dev[d] = rt.data_ptr()
                g.table_phase[d] = slope
        self.struct = g
        self.key: tuple = ()


OWN_SCRATCH_SYNC_BYTES = 1 << 30  # upper-bound scratch size from which the plan's counts are read synchronously
REAL_TABLE_TOL = 2e-6  # |table - r exp(-i p x)| / max|table| allowed when the tables are declared to be of that form


def real_table_form(tables: Sequence[Tensor], numpoints, table_oversamp, grid_size):
    """``[(real table on the device, phase slope p_d)]`` if every table is ``r_d[i] exp(-1j p_d x_i)`` with real
    ``r_d``, ``x_i = i / L_d - J_d / 2`` and ``p_d = pi (N_d - 1) / K_d`` for an integer ``N_d`` -- the form the
    reference builds (``torchkbnufft/_nufft/utils.py:160-204``) -- else ``None`` (modified tables keep the complex
    kernels).  One host read of the tables per geometry."""
    import numpy as np
    out = []
    for t, J, L, K in zip(tables, numpoints, table_oversamp, grid_size):
        tab = t.detach().cpu().numpy().astype(np.complex128)
        if tab.shape[0] != J * L + 1 or not np.all(np.isfinite(tab)):
            return None
        x = np.arange(J * L + 1) / L - J / 2
        mag = np.abs(tab)
        big = mag > 0.25 * mag.max()
        if big.sum() < 4:
            return None
        # slope from a least-squares line through the unwrapped phase of the significant entries
        ang = np.unwrap(np.angle(tab[big]))
        slope = -np.polyfit(x[big], ang, 1)[0]
        n_minus_1 = round(slope * K / np.pi)
        slope = np.pi * n_minus_1 / K
        real = tab * np.exp(1j * slope * x)
        scale = mag.max()
        if scale == 0 or np.abs(real.imag).max() > REAL_TABLE_TOL * scale:
            return None
        rdtype = torch.float32 if t.dtype == torch.complex64 else torch.float64
        out.append((torch.from_numpy(np.ascontiguousarray(real.real)).to(device=t.device, dtype=rdtype), float(slope)))
    return out


_HOST_INTS_FAST: dict = {}  # id(tensor) -> (_version, ints, tensor): identity front for _GRID_SIZE_CACHE


def host_ints(sizes) -> Tuple[int, ...]:
    """Integer tuple of a size argument.  Device tensors (module buffers) are read
    once and then looked up by identity so the hot path never synchronises."""
    if not isinstance(sizes, Tensor):
        return tuple(int(k) for k in sizes)
    fast = _HOST_INTS_FAST.get(id(sizes))
    if fast is not None and fast[0] == sizes._version and fast[2] is sizes:
        return fast[1]
    tkey = _tkey(sizes)
    hit = _GRID_SIZE_CACHE.get(tkey)
    if hit is None:
        hit = (tuple(int(k) for k in sizes.tolist()), sizes)  # keep the tensor alive with its key
        if len(_GRID_SIZE_CACHE) > 8 * GEOM_CACHE_SIZE:
            _GRID_SIZE_CACHE.clear()
        _GRID_SIZE_CACHE[tkey] = hit
    if len(_HOST_INTS_FAST) > 16 * GEOM_CACHE_SIZE:
        _HOST_INTS_FAST.clear()
    _HOST_INTS_FAST[id(sizes)] = (sizes._version, hit[0], sizes)
    return hit[0]


_GEOM_BY_CONTENT: "OrderedDict[tuple, Geometry]" = OrderedDict()


def _content_key(tables, n_shift, numpoints, table_oversamp, grid_ints) -> tuple:
    import hashlib
    parts = []
    for t in tables:
        raw = torch.view_as_real(dense(t).detach()).cpu().numpy().tobytes()
        parts.append((str(t.dtype), t.numel(), hashlib.blake2b(raw, digest_size=16).digest()))
    return (tables[0].device.index, tuple(parts), tuple(float(v) for v in n_shift.tolist()),
            tuple(int(v) for v in numpoints.tolist()), tuple(int(v) for v in table_oversamp.tolist()), tuple(grid_ints))


_GEOM_FAST: dict = {}  # (id, _version) of every buffer -> (Geometry, the buffers): a cheap front for _GEOM_CACHE


def get_geometry(tables: Sequence[Tensor], n_shift: Tensor, numpoints: Tensor, table_oversamp: Tensor,
                 grid_size) -> Geometry:
    """Cached :class:`Geometry` for a set of operator buffers.  ``grid_size`` may be a
    tensor (read once) or a sequence of ints."""
    # Hot path: the same tensor OBJECTS come back call after call (module buffers); identity + version is enough and
    # an order of magnitude cheaper than data_ptr/shape/stride keys.  The entry keeps the tensors alive, so an id
    # cannot be recycled while it is cached.
    gfast = (id(grid_size), grid_size._version) if isinstance(grid_size, Tensor) else tuple(grid_size)
    fkey = (tuple((id(t), t._version) for t in tables), id(n_shift), n_shift._version, id(numpoints),
            numpoints._version, id(table_oversamp), table_oversamp._version, gfast)
    hit = _GEOM_FAST.get(fkey)
    if hit is not None:
        return hit[0]
    # sizes as ints: the forward (sizes from image.shape) and the adjoint (grid_size buffer)
    # then share one geometry and one trajectory plan
    gkey = host_ints(grid_size)
    key = (tuple(_tkey(t) for t in tables), _tkey(n_shift), _tkey(numpoints), _tkey(table_oversamp), gkey)
    with _LOCK:
        geo = _GEOM_CACHE.get(key)
        if geo is not None:
            _GEOM_CACHE.move_to_end(key)
        else:
            for t in tables:
                require_cuda(t, "tables")
            # Operators with the same CONTENTS share one geometry, hence one trajectory plan: KbNufft and
            # KbNufftAdjoint built with the same arguments own different buffer tensors, and a trajectory that changes
            # every call would otherwise be planned once per operator (measured: two plan builds per forward+adjoint
            # pair, profiles/r02_spread_notes.txt).  One host read of the small buffers per new buffer identity.
            ckey = _content_key(tables, n_shift, numpoints, table_oversamp, gkey)
            geo = _GEOM_BY_CONTENT.get(ckey)
            if geo is None:
                geo = Geometry(tables, n_shift, numpoints, table_oversamp, gkey)
                geo.content_key = ckey
                _GEOM_BY_CONTENT[ckey] = geo
                while len(_GEOM_BY_CONTENT) > GEOM_CACHE_SIZE:
                    _GEOM_BY_CONTENT.popitem(last=False)
            else:
                _GEOM_BY_CONTENT.move_to_end(ckey)
            geo.key = key
            geo._refs = (n_shift, numpoints, table_oversamp)  # keep key pointers alive
            _GEOM_CACHE[key] = geo
            while len(_GEOM_CACHE) > GEOM_CACHE_SIZE:
                _GEOM_CACHE.popitem(last=False)
        if len(_GEOM_FAST) > 8 * GEOM_CACHE_SIZE:
            _GEOM_FAST.clear()
        _GEOM_FAST[fkey] = (geo, (list(tables), n_shift, numpoints, table_oversamp, grid_size))
    return geo


_RING = None          # pinned int32 ring for the asynchronous read-back of the sub-problem counts
_RING_SLOTS = 256
_RING_NEXT = 0


class TrajectoryPlan:
    """Device-resident ``b2n_points`` plan for one ``(geometry, omega)`` pair."""

    def __init__(self, geo: Geometry, omega: Tensor):
        lib = _lib.load()
        self.geo = geo
        self.omega = omega  # strong ref (see module docstring)
        om = omega
        if om.ndim == 2:
            self.n_traj = 1
            self.n_points = om.shape[1]
        else:
            self.n_traj = om.shape[0]
            self.n_points = om.shape[2]
        om = dense(om)
        nbytes = ctypes.c_size_t(0)
        _lib.check(lib.b2n_points_workspace_bytes(ctypes.byref(geo.struct), self.n_points, self.n_traj,
                                                  ctypes.byref(nbytes)), "b2n_points_workspace_bytes")
        self.workspace = torch.empty(max(int(nbytes.value), 256), dtype=torch.uint8, device=omega.device)
        self.struct = _lib.Points()
        with device_guard(omega.device):
            _lib.check(
                lib.b2n_points_build(ctypes.byref(geo.struct), om.data_ptr(), self.n_points, self.n_traj,
                                     self.workspace.data_ptr(), self.workspace.numel(), ctypes.byref(self.struct),
                                     current_stream_ptr(omega.device)),
                "b2n_points_build",
            )
        # the build reads `om` asynchronously; keep a possible contiguous copy alive with the plan
        self._om_contig = om
        stream = torch.cuda.current_stream(omega.device)
        self.workspace.record_stream(stream)
        # other streams that use the plan wait for this event first (use_on_current_stream)
        self._build_stream = stream
        self._streams_seen = {stream.cuda_stream}
        self._build_event = None
        if not torch.cuda.is_current_stream_capturing():
            self._build_event = torch.cuda.Event()
            self._build_event.record(stream)
        self._content = om.clone() if _plan_cache_mode == "content" else None
        self.own_slots = 0      # partial-sum slots of the owner-tile spread: 0 = not read back yet (upper bound)
        self._own_scratch = {}
        # The number of sub-problems is only known on the device; kernels are launched over the upper bound
        # n_sub_max and the surplus CTAs exit at once (~3 us of tail per launch at BASELINE config 2).  When a
        # plan is REUSED, the exact count is fetched asynchronously -- no synchronisation -- and the bound is
        # tightened once it has arrived.  (Not at build time: a device-to-host copy in the compute stream queues
        # behind bulk downloads on the copy engine and would stall pipelines that change trajectory every call.)
        self._n_sub_host = None
        self._n_sub_event = None
        self._count_requested = not (self.struct.n_sub and self.struct.n_sub_max > 0)

    def use_on_current_stream(self) -> None:
        """Make the plan safe to read from the current stream: a stream other than the one that built it waits
        for the build and is recorded as a user of the workspace (so the caching allocator does not hand the
        memory out while that stream still reads it)."""
        stream = torch.cuda.current_stream(self.omega.device)
        if stream.cuda_stream in self._streams_seen:
            return
        if torch.cuda.is_current_stream_capturing():
            # A capturing stream must not wait on work recorded outside its capture (cudaErrorStreamCaptureIsolation).
            # The build was enqueued before the capture began, and torch.cuda.graph synchronises the device on entry,
            # so the plan is complete; the stream is not remembered, a later eager use of it takes the normal path.
            return
        if self._build_event is not None:
            stream.wait_event(self._build_event)
        self.workspace.record_stream(stream)
        with _LOCK:
            self._streams_seen.add(stream.cuda_stream)

    def content_matches(self, omega: Tensor) -> bool:
        return self._content is None or bool(torch.equal(self._content, dense(omega)))

    def request_count(self) -> None:
        """Start the asynchronous read-back of the device-side counts (first reuse of the plan): sub-problems,
        and the work items / partial-sum slots of the owner-tile spread."""
        self._count_requested = True
        # a slot of a pinned ring allocated once (pinning memory per plan would cost a blocking cudaHostAlloc)
        global _RING, _RING_NEXT
        with _LOCK:
            if _RING is None:
                _RING = torch.empty((_RING_SLOTS, 4), dtype=torch.int32).pin_memory()
            self._n_sub_token = _RING_NEXT
            _RING_NEXT += 1
        self._n_sub_host = _RING[self._n_sub_token % _RING_SLOTS]
        base = self.workspace.data_ptr()
        off = int(self.struct.n_sub) - base
        self._n_sub_host[0:1].copy_(self.workspace[off:off + 4].view(torch.int32), non_blocking=True)
        if self.struct.own_tile:
            off = int(self.struct.own_counts) - base
            self._n_sub_host[1:4].copy_(self.workspace[off:off + 12].view(torch.int32), non_blocking=True)
        self._n_sub_event = torch.cuda.Event()
        self._n_sub_event.record(torch.cuda.current_stream(self.omega.device))

    def tighten(self) -> None:
        """Replace the launch bounds (n_sub_max, n_own_items_max) by the exact counts once the asynchronous read-back
        of the device counters has completed (cheap no-op before that and after it has been applied)."""
        ev = self._n_sub_event
        if ev is None:
            return
        if _RING_NEXT - self._n_sub_token >= _RING_SLOTS:  # the ring slot has been handed to a newer plan
            self._n_sub_event = None
            return
        if torch.cuda.is_current_stream_capturing() or not ev.query():
            return  # (event queries are not allowed while a CUDA graph is being captured)
        self._n_sub_event = None
        n = int(self._n_sub_host[0])
        if 0 < n < self.struct.n_sub_max:
            self.struct.n_sub_max = n
        if self.struct.own_tile:
            n_items, n_slots, n_exc = (int(v) for v in self._n_sub_host[1:4])
            if 0 < n_items <= self.struct.n_own_items_max:
                self.struct.n_own_items_max = n_items
                self.own_slots = max(n_slots, 1)
                self._own_scratch.clear()  # sized by the upper bound until now
                if 0 <= n_exc <= self.struct.n_own_exc_max:
                    self.struct.n_own_exc_max = n_exc  # 0 (the usual case): no fix-up launch from now on

    def own_scratch(self, lib, geo, B: int, C: int, layout: int):
        """Persistent scratch of the owner-tile spread for this (batch, coils, stream): arrival counters (zeroed once,
        the kernel leaves them zero) followed by the partial-sum slots.  ``None`` when that path does not apply."""
        if not self.struct.own_tile:
            return None
        stream = torch.cuda.current_stream(self.omega.device)
        key = (B, C, layout, stream.cuda_stream)
        hit = self._own_scratch.get(key)
        if hit is None:
            nbytes, nzero = ctypes.c_size_t(0), ctypes.c_size_t(0)
            if self.own_slots == 0 and not torch.cuda.is_current_stream_capturing():
                # The partial-sum slots are sized by the plan's upper bound until the device counts have been read
                # back; for large 3-D plans that bound is gigabytes, so read the counts now (one synchronisation per
                # plan) rather than allocate it.
                _lib.check(lib.b2n_interp_adjoint_ordered_layout(ctypes.byref(geo.struct), ctypes.byref(self.struct), B,
                                                                 C, layout, 0, ctypes.byref(nbytes), ctypes.byref(nzero)),
                           "b2n_interp_adjoint_ordered_layout")
                if nbytes.value > OWN_SCRATCH_SYNC_BYTES:
                    base = self.workspace.data_ptr()
                    off = int(self.struct.own_counts) - base
                    n_items, n_slots, n_exc = (int(v) for v in self.workspace[off:off + 12].view(torch.int32).cpu())
                    if 0 < n_items <= self.struct.n_own_items_max:
                        self.struct.n_own_items_max = n_items
                        self.own_slots = max(n_slots, 1)
                        if 0 <= n_exc <= self.struct.n_own_exc_max:
                            self.struct.n_own_exc_max = n_exc
            _lib.check(lib.b2n_interp_adjoint_ordered_layout(ctypes.byref(geo.struct), ctypes.byref(self.struct), B, C,
                                                             layout, self.own_slots, ctypes.byref(nbytes),
                                                             ctypes.byref(nzero)), "b2n_interp_adjoint_ordered_layout")
            if not nzero.value:
                hit = (None, 0)  # another deterministic path (or none) handles this case
            else:
                buf = torch.empty(nbytes.value, dtype=torch.uint8, device=self.omega.device)
                buf[:nzero.value].zero_()
                hit = (buf, nbytes.value)
            if len(self._own_scratch) >= 4:
                self._own_scratch.clear()
            self._own_scratch[key] = hit
        return hit


_PLAN_FAST: dict = {}  # (id(geo), id(omega), omega._version) -> plan (the plan keeps geo and omega alive)


def get_plan(geo: Geometry, omega: Tensor) -> TrajectoryPlan:
    if _plan_cache_mode == "off":
        return TrajectoryPlan(geo, omega)
    fkey = (id(geo), id(omega), omega._version)
    plan = _PLAN_FAST.get(fkey)
    if plan is not None and plan._content is not None and not plan.content_matches(omega):
        invalidate_plans(omega)
        plan = None
    if plan is not None:
        if not plan._count_requested:
            if not torch.cuda.is_current_stream_capturing():
                plan.request_count()
        elif plan._n_sub_event is not None:
            plan.tighten()
        plan.use_on_current_stream()
        return plan
    key = (geo.key, _tkey(omega))
    with _LOCK:
        plan = _PLAN_CACHE.get(key)
        if plan is not None and plan._content is not None and not plan.content_matches(omega):
            del _PLAN_CACHE[key]
            plan = None
        if plan is not None:
            _PLAN_CACHE.move_to_end(key)
        else:
            plan = TrajectoryPlan(geo, omega)
            _PLAN_CACHE[key] = plan
            # bounded by count AND by bytes: a 3-D plan with owner-tile visit lists is gigabytes (5 GB at BASELINE
            # config 4), a 2-D one tens of megabytes
            while len(_PLAN_CACHE) > 1 and (len(_PLAN_CACHE) > PLAN_CACHE_SIZE or
                                            sum(p.workspace.numel() for p in _PLAN_CACHE.values()) > PLAN_CACHE_BYTES):
                evicted = _PLAN_CACHE.popitem(last=False)[1]
                for k in [k for k, v in _PLAN_FAST.items() if v is evicted]:
                    del _PLAN_FAST[k]  # an evicted plan must free its device memory
        if len(_PLAN_FAST) > 8 * PLAN_CACHE_SIZE:
            _PLAN_FAST.clear()
        if plan.omega is omega:  # only identity-key the tensor the plan itself keeps alive
            _PLAN_FAST[fkey] = plan
    plan.use_on_current_stream()
    return plan


def invalidate_plans(omega: Optional[Tensor] = None) -> None:
    """Forget the cached trajectory plans built from ``omega`` (all plans when ``None``).  Call this after writing
    to a trajectory tensor through anything that does not bump torch's version counter (``omega.data``, DLPack,
    a custom kernel): the cache key is (storage pointer, version), so such a write is otherwise invisible."""
    from . import graphs as _graphs
    _graphs.clear_graphs()  # captured graphs hold pointers into the plans
    with _LOCK:
        if omega is None:
            _PLAN_CACHE.clear()
            _PLAN_FAST.clear()
            _OMEGA_CAST.clear()
            return
        ptr = omega.data_ptr()
        for k in [k for k, v in _PLAN_CACHE.items() if v.omega is omega or v.omega.data_ptr() == ptr]:
            del _PLAN_CACHE[k]
        for k in [k for k, v in _PLAN_FAST.items() if v.omega is omega or v.omega.data_ptr() == ptr]:
            del _PLAN_FAST[k]
        for k in [k for k, v in _OMEGA_CAST.items() if v[0] is omega or v[0].data_ptr() == ptr]:
            plans = [pk for pk, pv in _PLAN_CACHE.items() if pv.omega is _OMEGA_CAST[k][1]]
            for pk in plans:
                del _PLAN_CACHE[pk]
            for fk in [fk for fk, fv in _PLAN_FAST.items() if fv.omega is _OMEGA_CAST[k][1]]:
                del _PLAN_FAST[fk]
            del _OMEGA_CAST[k]


def clear_caches() -> None:
    """Drop every cached geometry and trajectory plan (frees their device memory)."""
    from . import graphs as _graphs
    _graphs.clear_graphs()
    _PLAN_CACHE.clear()
    _PLAN_FAST.clear()
    _OMEGA_CAST.clear()
    _GEOM_CACHE.clear()
    _GEOM_BY_CONTENT.clear()
    _GEOM_FAST.clear()
    _GRID_SIZE_CACHE.clear()
    _HOST_INTS_FAST.clear()


def normalize_omega(omega: Tensor, n_batch: int, what: str) -> Tensor:
    """Trajectory validation shared by both directions
    (reference: ``_nufft/interp.py:345-356`` and ``:621-633``)."""
    if omega.ndim not in (2, 3):
        raise ValueError("omega must have 2 or 3 dimensions.")
    if omega.ndim == 3 and omega.shape[0] == 1:
        omega = omega[0]  # a single trajectory is broadcast over the batch
    if omega.ndim == 3 and not omega.shape[0] == n_batch:
        raise ValueError(f"If omega has batch dim, omega batch dimension must match {what}.")
    return omega
