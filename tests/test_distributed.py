"""Multi-rank sharding logic (torchkbnufft_b200/parallel.py) with world_size 2 on the
gloo backend; compute is the oracle-backed CPU stand-in (tests/cpu_engine_shim.py),
so this covers partitioning and the coil-sum all-reduce, not the kernels."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from torchkbnufft_b200.parallel import shard_bounds


def test_shard_bounds_cover_without_overlap():
    for n in (0, 1, 7, 16, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "tests"), os.path.join(root, "oracle")):
        sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import torchkbnufft_b200 as tkbn
    from cpu_engine_shim import oracle_engine
    from golden_cases import CASES, case_inputs
    from torchkbnufft_b200 import parallel

    case = CASES["d2_radial"]  # B=1, C=4
    inp = case_inputs(case, np.complex128)
    T = lambda k: torch.from_numpy(inp[k])
    kw = dict(im_size=case["im_size"], dtype=torch.complex128)
    nu, na = tkbn.KbNufft(**kw), tkbn.KbNufftAdjoint(**kw)
    with oracle_engine():
        full_f = nu(T("image"), T("omega"), smaps=T("smaps"), norm="ortho")
        full_a = na(T("kdata"), T("omega"), smaps=T("smaps"), norm="ortho")
        # coil sharding: forward stays sharded, adjoint all-reduces the coil-combined image
        smaps_l = parallel.local_coils(T("smaps"))
        kdata_l = parallel.local_coils(T("kdata"))
        lo, hi = parallel.shard_bounds(T("smaps").shape[1], rank, world)
        f_local = parallel.coil_sharded_forward(nu, T("image"), T("omega"), smaps_l, norm="ortho")
        ok_f = torch.allclose(f_local, full_f[:, lo:hi])
        a_all = parallel.coil_sharded_adjoint(na, kdata_l, T("omega"), smaps_l, norm="ortho")
        ok_a = torch.allclose(a_all, full_a)
        # a reducer object (the role parallel.PeerAllReduce plays on GPUs) runs as the adjoint's epilogue, exactly once
        class GlooReducer:
            key, comm, calls = ("gloo_reducer", 1), None, 0

            def takes(self, n_values, dtype):
                return True

            def __call__(self, x):
                GlooReducer.calls += 1
                return parallel.all_reduce_complex_(x)

        parallel.fuse_allreduce = False  # .comm is a C communicator on GPUs; here the callable is the whole reducer
        try:
            a_red = parallel.coil_sharded_adjoint(na, kdata_l, T("omega"), smaps_l, norm="ortho", reducer=GlooReducer())
        finally:
            parallel.fuse_allreduce = True
        ok_a = ok_a and torch.allclose(a_red, full_a) and GlooReducer.calls == 1
        # Toeplitz normal operator with sharded coils: same all-reduce after the local coil sum
        toep = tkbn.ToepNufft()
        kern = tkbn.calc_toeplitz_kernel(T("omega"), case["im_size"], norm="ortho")
        full_t = toep(T("image"), kern, smaps=T("smaps"), norm="ortho")
        t_all = parallel.coil_sharded_toeplitz(toep, T("image"), kern, smaps_l, norm="ortho")
        ok_a = ok_a and torch.allclose(t_all, full_t)
        # batch sharding: no communication
        case_b = CASES["d2"]  # B=2
        inp_b = case_inputs(case_b, np.complex128)
        Tb = lambda k: torch.from_numpy(inp_b[k])
        kwb = dict(im_size=case_b["im_size"], dtype=torch.complex128)
        nub, nab = tkbn.KbNufft(**kwb), tkbn.KbNufftAdjoint(**kwb)
        img_l = parallel.local_batch(Tb("image"))
        k_l, a_l = parallel.batch_sharded_pair(nub, nab, img_l, Tb("omega"), Tb("smaps"))
        full_k = nub(Tb("image"), Tb("omega"), smaps=Tb("smaps"))
        blo, bhi = parallel.shard_bounds(2, rank, world)
        ok_b = torch.allclose(k_l, full_k[blo:bhi]) and a_l.shape == img_l.shape
    with open(os.path.join(out_dir, f"rank{rank}.txt"), "w") as f:
        f.write(f"{int(ok_f)} {int(ok_a)} {int(ok_b)}")
    dist.barrier()
    dist.destroy_process_group()


def test_coil_and_batch_sharding_world_size_2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for rank in range(world):
        assert open(tmp_path / f"rank{rank}.txt").read() == "1 1 1"
