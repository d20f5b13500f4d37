"""Multi-GPU partitioning of the SENSE NUFFT: one process per GPU, ``torch.distributed``.

Every ``(batch, coil)`` NUFFT is independent in both directions (the reference
broadcasts over B and C, ``_nufft/interp.py:197`` and ``:412-417``); only the SENSE
adjoint's coil sum (``modules/kbnufft.py:404-405``) couples coils.  So:

* **batch sharding** (primary; BASELINE config 5): each rank owns a contiguous block
  of slices with all coils, trajectory / tables / smaps replicated -- no communication;
* **coil sharding** (when B < number of GPUs): each rank owns a block of coils; the
  forward needs no communication (its output stays coil-sharded), the adjoint ends
  with ONE sum all-reduce of the coil-combined image ``(B, 1, *N)``, a small
  (<= tens of MB) message issued right after the fused crop/apodise/coil-sum kernel
  on the same stream (NCCL over NVLink on GPUs; gloo in the CPU unit tests) -- or, with a
  :class:`PeerAllReduce`, the engine's own exchange over NVLink peer memory
  (``csrc/b2n_peer.cu``): the adjoint's last FFT pass pushes every finished image row into
  the peers' windows and adds what arrives from them in rank order, compute and collective in
  one kernel (a stand-alone kernel where that pass cannot carry it); bit-identical sums on
  every rank; config 2 over 2 / 4 / 8 GPUs: 116 / 103 / 94 us per pair against 127 / 119 / 111 us
  through NCCL.

The reference has no distributed code at all; this module is new surface.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
import torch.distributed as dist
from torch import Tensor


def shard_bounds(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous balanced partition: the first ``n_items % world_size`` ranks get one
    extra item.  Returns ``(start, stop)``; empty when there are more ranks than items."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError(f"bad rank {rank} for world size {world_size}")
    base, extra = divmod(n_items, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def _rank_world(group) -> Tuple[int, int]:
    if not (dist.is_available() and dist.is_initialized()):
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def local_batch(x: Tensor, group=None) -> Tensor:
    """This rank's block of the batch axis (dim 0)."""
    rank, world = _rank_world(group)
    lo, hi = shard_bounds(x.shape[0], rank, world)
    return x[lo:hi]


def local_coils(x: Tensor, group=None) -> Tensor:
    """This rank's block of the coil axis (dim 1)."""
    rank, world = _rank_world(group)
    lo, hi = shard_bounds(x.shape[1], rank, world)
    return x[:, lo:hi]


def all_reduce_complex_(x: Tensor, group=None) -> Tensor:
    """In-place sum all-reduce of a complex tensor (sent as interleaved reals)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        if not x.is_contiguous():
            raise ValueError("all_reduce_complex_ needs a contiguous tensor")
        dist.all_reduce(torch.view_as_real(x) if x.is_complex() else x, op=dist.ReduceOp.SUM, group=group)
    return x


# coil_sharded_adjoint with a PeerAllReduce: let the last inverse FFT pass push its rows into the peers' windows itself
# (b2n_fft_adjoint_fused_allreduce) instead of launching the all-reduce kernel behind it; off for A/B measurements
fuse_allreduce = True


class PeerAllReduce:
    """In-place sum all-reduce of a float32 / complex64 CUDA tensor over NVLink peer memory
    (``b2n_peer_allreduce_sum``): one process per GPU, all on one node.

    Construction is collective over ``group``: every rank creates its window (device memory
    owned by the engine), the CUDA IPC handles travel through ``all_gather_object``, every rank
    maps the others' windows.  Calls are collective too (same order, same sizes on all ranks, one
    stream at a time), enqueue one kernel on the current stream and never touch the host, so they
    can be captured into CUDA graphs.  ``max_values`` is the largest tensor, counted in elements
    of the tensor's own dtype (complex counts double internally).
    """

    _serial = 0

    def __init__(self, max_values: int, dtype: torch.dtype = torch.complex64, group=None,
                 device: Optional[torch.device] = None):
        import ctypes

        from . import _lib

        self._lib, self._ctypes = _lib, ctypes
        PeerAllReduce._serial += 1
        self.key = ("peer_allreduce", PeerAllReduce._serial)  # identifies this reducer in the CUDA-graph cache
        self.rank, self.world = _rank_world(group)
        if self.world > _lib.PEER_MAX_RANKS:
            raise ValueError(f"PeerAllReduce serves up to {_lib.PEER_MAX_RANKS} ranks, got {self.world}")
        if dtype not in (torch.complex64, torch.float32):
            raise TypeError("PeerAllReduce sums float32 / complex64 tensors")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.max_floats = int(max_values) * (2 if dtype.is_complex else 1)
        lib = _lib.load()
        self._group = group
        self._own, self._peers = None, []
        self._comm = _lib.PeerComm()
        self._comm.rank, self._comm.world, self._comm.max_floats = self.rank, self.world, self.max_floats
        # every step below is followed by an exchange of its outcome, so that a failure on ONE rank (no peer access,
        # IPC refused by the container) raises on ALL of them instead of leaving the others in a collective
        error, handle = None, ctypes.create_string_buffer(_lib.PEER_HANDLE_BYTES)
        try:
            nbytes = ctypes.c_size_t(0)
            _lib.check(lib.b2n_peer_window_bytes(self.world, self.max_floats, ctypes.byref(nbytes)), "b2n_peer_window_bytes")
            own = ctypes.c_void_p(None)
            with torch.cuda.device(self.device):
                _lib.check(lib.b2n_peer_window_create(nbytes.value, ctypes.byref(own), handle), "b2n_peer_window_create")
            self._own = own
            self._comm.window[self.rank] = own.value
        except Exception as exc:  # noqa: BLE001
            error = f"rank {self.rank}: {exc}"
        handles = self._exchange((error, bytes(handle.raw)))
        self._raise_if_any([h[0] for h in handles])
        try:
            for r in range(self.world):
                if r == self.rank:
                    continue
                mapped = ctypes.c_void_p(None)
                with torch.cuda.device(self.device):
                    _lib.check(lib.b2n_peer_window_open(ctypes.create_string_buffer(handles[r][1], _lib.PEER_HANDLE_BYTES),
                                                        ctypes.byref(mapped)), f"b2n_peer_window_open(rank {r})")
                self._comm.window[r] = mapped.value
                self._peers.append(mapped)
        except Exception as exc:  # noqa: BLE001
            error = f"rank {self.rank}: {exc}"
        # also the barrier that keeps anyone from pushing into a window that is not mapped and filled yet
        self._raise_if_any(self._exchange(error))

    def _exchange(self, item):
        if self.world == 1:
            return [item]
        out = [None] * self.world
        dist.all_gather_object(out, item, group=self._group)
        return out

    def _raise_if_any(self, errors) -> None:
        errors = [e for e in errors if e]
        if errors:
            self._release()
            raise RuntimeError("PeerAllReduce could not be set up: " + "; ".join(errors))

    def _release(self) -> None:
        lib = self._lib.load()
        for mapped in self._peers:
            lib.b2n_peer_window_close(mapped)
        self._peers = []
        if self._own is not None:
            lib.b2n_peer_window_destroy(self._own)
            self._own = None

    # Past this size a pipelined ring (NCCL) is the better tool; up to it the engine's kernel picks the one-shot exchange
    # (latency-bound messages) or the two-shot form (reduce-scatter + all-gather through the same windows) by itself.
    MAX_BYTES = 64 << 20

    def takes(self, n_values: int, dtype: torch.dtype) -> bool:
        """Whether a tensor of ``n_values`` elements of ``dtype`` should go through this reducer: float32 / complex64,
        inside the window, at most ``MAX_BYTES``."""
        if self._own is None or dtype not in (torch.complex64, torch.float32):
            return False
        n = int(n_values) * (2 if dtype.is_complex else 1)
        return 0 < n <= self.max_floats and 4 * n <= self.MAX_BYTES

    @property
    def comm(self):
        """The ``_lib.PeerComm`` (``struct b2n_peer_comm``) of this rank, for C entries that take a communicator."""
        if self._own is None:
            raise RuntimeError("PeerAllReduce is closed")
        return self._comm

    def __call__(self, x: Tensor) -> Tensor:
        if x.device != self.device or x.dtype not in (torch.complex64, torch.float32) or not x.is_contiguous():
            raise ValueError("PeerAllReduce needs a contiguous float32 / complex64 tensor on its own device")
        if self._own is None:
            raise RuntimeError("PeerAllReduce is closed")
        n = x.numel() * (2 if x.is_complex() else 1)
        if n == 0:
            return x
        if x.is_complex() and (x.is_conj() or x.is_neg()):
            raise ValueError("PeerAllReduce works in place: materialise lazy conj / neg views first (resolve_conj())")
        stream = torch.cuda.current_stream(self.device).cuda_stream
        ptr = x.data_ptr()
        self._lib.check(self._lib.load().b2n_peer_allreduce_sum(self._ctypes.byref(self._comm), ptr, ptr, n, stream),
                        "b2n_peer_allreduce_sum")
        return x

    def close(self) -> None:
        """Collective: unmap the peers' windows, then free this rank's own (after everyone has unmapped)."""
        if self._own is None:
            return
        from ._nufft import graphs as _graphs

        _graphs.clear_graphs()  # captured adjoints may hold this reducer's kernel and window pointers
        lib = self._lib.load()
        torch.cuda.synchronize(self.device)
        for mapped in self._peers:
            lib.b2n_peer_window_close(mapped)
        self._peers = []
        if self.world > 1 and dist.is_initialized():
            dist.barrier(group=self._group)
        self._release()


def coil_sharded_forward(nufft_ob, image: Tensor, omega: Tensor, smaps_local: Tensor,
                         norm: Optional[str] = None) -> Tensor:
    """Forward SENSE NUFFT on this rank's coils: ``image (B, 1, *N)`` replicated,
    ``smaps_local (Bs, C_local, *N)``; returns ``(B, C_local, M)`` (coil-sharded)."""
    return nufft_ob(image, omega, smaps=smaps_local, norm=norm)


def coil_sharded_adjoint(adj_ob, data_local: Tensor, omega: Tensor, smaps_local: Tensor,
                         norm: Optional[str] = None, group=None, reducer: Optional[PeerAllReduce] = None) -> Tensor:
    """Adjoint SENSE NUFFT with coils sharded over ranks: local adjoint + local coil
    combination, then one sum all-reduce of ``(B, 1, *N)`` -- through ``reducer`` (the
    engine's peer-memory kernel) when one is given, else ``dist.all_reduce``.  Every rank
    returns the full coil-combined image.  Not differentiable across ranks (inference/recon path)."""
    if data_local.shape[1] == 0:  # more ranks than coils: contribute zeros
        shape = (data_local.shape[0], 1) + tuple(smaps_local.shape[2:])
        partial = torch.zeros(shape, dtype=data_local.dtype, device=data_local.device)
    elif reducer is not None and reducer.takes(data_local.shape[0] * math.prod(smaps_local.shape[2:]), data_local.dtype):
        # the all-reduce kernel is the last launch OF the adjoint: in graph mode it is replayed with it
        from ._nufft import graphs as _graphs

        with _graphs.adjoint_epilogue(reducer, reducer.key + (fuse_allreduce,),
                                      peer_comm=reducer.comm if fuse_allreduce else None) as epilogue:
            partial = adj_ob(data_local, omega, smaps=smaps_local, norm=norm)
        if epilogue.applied:
            return partial
    else:
        partial = adj_ob(data_local, omega, smaps=smaps_local, norm=norm)
    partial = partial.contiguous()
    if reducer is not None and reducer.takes(partial.numel(), partial.dtype):
        return reducer(partial)
    return all_reduce_complex_(partial, group)


def coil_sharded_toeplitz(toep_ob, image: Tensor, kernel: Tensor, smaps_local: Tensor, norm: Optional[str] = None,
                          group=None, reducer: Optional[PeerAllReduce] = None) -> Tensor:
    """Toeplitz normal operator ``sum_c conj(S_c) filter(S_c x)`` with the coils sharded over ranks: the local
    ``ToepNufft`` apply on this rank's coils (image and kernel replicated), then the same sum all-reduce of
    ``(B, 1, *N)`` as the adjoint (reference coupling point ``modules/kbnufft.py:484``)."""
    if smaps_local.shape[1] == 0:
        partial = torch.zeros_like(image)
    else:
        partial = toep_ob(image, kernel, smaps=smaps_local, norm=norm)
    partial = partial.contiguous()
    if reducer is not None and reducer.takes(partial.numel(), partial.dtype):
        return reducer(partial)
    return all_reduce_complex_(partial, group)


def batch_sharded_pair(nufft_ob, adj_ob, image_local: Tensor, omega: Tensor, smaps: Tensor,
                       norm: Optional[str] = None) -> Tuple[Tensor, Tensor]:
    """Forward then adjoint SENSE NUFFT on this rank's slices (no communication)."""
    kdata = nufft_ob(image_local, omega, smaps=smaps, norm=norm)
    return kdata, adj_ob(kdata, omega, smaps=smaps, norm=norm)
