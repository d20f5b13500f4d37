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
  on the same stream (NCCL over NVLink on GPUs; gloo in the CPU unit tests).

The reference has no distributed code at all; this module is new surface.
"""
from __future__ import annotations

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


def coil_sharded_forward(nufft_ob, image: Tensor, omega: Tensor, smaps_local: Tensor,
                         norm: Optional[str] = None) -> Tensor:
    """Forward SENSE NUFFT on this rank's coils: ``image (B, 1, *N)`` replicated,
    ``smaps_local (Bs, C_local, *N)``; returns ``(B, C_local, M)`` (coil-sharded)."""
    return nufft_ob(image, omega, smaps=smaps_local, norm=norm)


def coil_sharded_adjoint(adj_ob, data_local: Tensor, omega: Tensor, smaps_local: Tensor,
                         norm: Optional[str] = None, group=None) -> Tensor:
    """Adjoint SENSE NUFFT with coils sharded over ranks: local adjoint + local coil
    combination, then one sum all-reduce of ``(B, 1, *N)``.  Every rank returns the
    full coil-combined image.  Not differentiable across ranks (inference/recon path)."""
    if data_local.shape[1] == 0:  # more ranks than coils: contribute zeros
        shape = (data_local.shape[0], 1) + tuple(smaps_local.shape[2:])
        partial = torch.zeros(shape, dtype=data_local.dtype, device=data_local.device)
    else:
        partial = adj_ob(data_local, omega, smaps=smaps_local, norm=norm)
    return all_reduce_complex_(partial.contiguous(), group)


def batch_sharded_pair(nufft_ob, adj_ob, image_local: Tensor, omega: Tensor, smaps: Tensor,
                       norm: Optional[str] = None) -> Tuple[Tensor, Tensor]:
    """Forward then adjoint SENSE NUFFT on this rank's slices (no communication)."""
    kdata = nufft_ob(image_local, omega, smaps=smaps, norm=norm)
    return kdata, adj_ob(kdata, omega, smaps=smaps, norm=norm)
