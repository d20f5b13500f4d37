"""Library-owned CUDA-graph replay of the forward / adjoint NUFFT (opt-in: ``set_graph_mode(True)``).

One forward + adjoint SENSE NUFFT of BASELINE config 2 is 7 kernels and ~170 us of GPU time, but ~145 us of Python /
ctypes work per pair on the host (``profiles/r01_g_host_overhead.log``): any host hiccup opens gaps on the device, and
small problems (config 1: 45 us of GPU work) are host-bound outright.  With graph mode on, the second call of an
operator with the SAME argument buffers (same storage pointers, shapes and trajectory version -- contents may change)
captures its launches into a ``torch.cuda.CUDAGraph`` and every later call replays it: one ``cudaGraphLaunch`` instead
of the whole host path.

Semantics to know about (the reason this is opt-in):
  * the result of a replayed call is a library-owned static tensor that the next call with the same arguments
    overwrites -- clone it if it must outlive that;
  * only calls that autograd does not record are replayed (inference / iterative reconstruction loops);
  * a trajectory edited in place bumps its version and gets a new plan and a new graph; edits that bypass the version
    counter need ``invalidate_plans`` (as for the plan cache), which also drops the graphs;
  * the capture call synchronises the device once (``torch.cuda.graph``).

The reference has no counterpart (it launches ATen ops eagerly, ``torchkbnufft/_nufft/interp.py:185-197``).
"""
from __future__ import annotations

import threading
from collections import OrderedDict
from typing import Callable, Optional, Sequence

import torch
from torch import Tensor

GRAPH_CACHE_SIZE = 16
_ENABLED = False
_CACHE: "OrderedDict[tuple, _Entry]" = OrderedDict()


_REPLAYED_KERNELS = 0  # kernels of this library launched through graph replays (b2n_launch_count sees eager ones only)


def replayed_kernel_count() -> int:
    return _REPLAYED_KERNELS


class _Entry:
    __slots__ = ("calls", "graph", "out", "refs", "n_kernels", "failed")

    def __init__(self):
        self.calls = 0
        self.n_kernels = 0
        self.failed = False
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.out: Optional[Tensor] = None
        self.refs: tuple = ()


def set_graph_mode(enabled: bool) -> None:
    """Switch the library-owned CUDA-graph replay on or off (off by default; see the module docstring)."""
    global _ENABLED
    _ENABLED = bool(enabled)
    if not _ENABLED:
        clear_graphs()


def get_graph_mode() -> bool:
    return _ENABLED


def clear_graphs() -> None:
    _CACHE.clear()


_SCOPE = threading.local()


class replay_scope:
    """Marks the eager body of a graphed operator: the functional entry points call themselves through
    ``replay_or_run`` and must not start a nested lookup."""

    def __enter__(self):
        _SCOPE.depth = getattr(_SCOPE, "depth", 0) + 1

    def __exit__(self, *exc):
        _SCOPE.depth -= 1


def in_replay_scope() -> bool:
    return getattr(_SCOPE, "depth", 0) > 0


class adjoint_epilogue:
    """Context manager: ``fn(image) -> image`` runs on the coil-combined result of every SENSE adjoint that autograd
    does not record inside the block, as the last step OF the operator -- so that it is captured into the operator's
    graph and replayed with it.  ``parallel.coil_sharded_adjoint`` puts the peer-memory all-reduce of the partial
    coil sums here: one ``cudaGraphLaunch`` per adjoint, collective included.  ``key`` distinguishes epilogues in the
    graph cache; ``applied`` tells the caller whether the operator ran it (it does not on the autograd path)."""

    def __init__(self, fn: Callable[[Tensor], Tensor], key, peer_comm=None):
        # peer_comm: the epilogue is the engine's own all-reduce over this communicator, which the last FFT pass can
        # perform itself (b2n_fft_adjoint_fused_allreduce) -- fn is then only the fallback for the unfused FFT route
        self.fn, self.key, self.applied, self.peer_comm = fn, key, False, peer_comm

    def __enter__(self):
        self._prev = getattr(_SCOPE, "epilogue", None)
        _SCOPE.epilogue = self
        return self

    def __exit__(self, *exc):
        _SCOPE.epilogue = self._prev


def current_epilogue() -> Optional[adjoint_epilogue]:
    return getattr(_SCOPE, "epilogue", None)


def _tensor_key(t: Optional[Tensor]) -> tuple:
    if t is None:
        return ()
    return (t.data_ptr(), tuple(t.shape), tuple(t.stride()), t.dtype, t.is_conj(), t.is_neg())


def replay_or_run(kind: str, operands: Sequence[Optional[Tensor]], statics: Sequence[Tensor], omega: Tensor,
                  extra: tuple, fn: Callable[[], Tensor], pin: Callable[[], object]) -> Tensor:
    """Run ``fn`` (which launches the operator on the current stream and returns its result), through a captured graph
    from the fourth call with the same key on.  ``operands``: every tensor whose STORAGE the launches read (a different
    buffer is a different graph); ``statics``: operator buffers, keyed by identity and version (cheaper); ``omega``: the trajectory (its version is part of the key: new contents, new plan);
    ``pin``: returns the objects the captured pointers live in besides ``operands`` (the trajectory plan), kept alive
    with the graph."""
    if not _ENABLED or torch.cuda.is_current_stream_capturing():
        return fn()
    dev = omega.device
    key = (kind, extra, tuple(_tensor_key(t) for t in operands), tuple((id(t), t._version) for t in statics),
           _tensor_key(omega), omega._version,
           torch.cuda.current_stream(dev).cuda_stream)
    ent = _CACHE.get(key)
    if ent is None:
        ent = _Entry()
        _CACHE[key] = ent
        while len(_CACHE) > GRAPH_CACHE_SIZE:
            _CACHE.popitem(last=False)
    else:
        _CACHE.move_to_end(key)
    ent.calls += 1
    global _REPLAYED_KERNELS
    if ent.graph is not None:
        ent.graph.replay()
        _REPLAYED_KERNELS += ent.n_kernels
        return ent.out
    if ent.calls < 3 or ent.failed:
        return fn()  # first sights: run eagerly (builds the plan, twiddles, scratch -- nothing of that may be captured)
    plan = pin()
    if getattr(plan, "_n_sub_event", None) is not None or not getattr(plan, "_count_requested", True):
        return fn()  # the plan's launch bounds are still upper bounds (device counts not read back yet): capture later
    from .. import _lib
    lib = _lib.load()
    graph = torch.cuda.CUDAGraph()
    before = lib.b2n_launch_count()
    try:
        with torch.cuda.graph(graph):
            out = fn()
    except Exception:  # something in this call cannot be captured (e.g. a lazy allocation): stay eager for this key
        ent.failed = True
        torch.cuda.synchronize(dev)
        return fn()
    ent.n_kernels = int(lib.b2n_launch_count() - before)  # launches recorded into the graph, not executed
    ent.graph, ent.out = graph, out
    ent.refs = (tuple(operands), tuple(statics), omega, plan)
    graph.replay()
    _REPLAYED_KERNELS += ent.n_kernels
    return out
