"""Worker of tests/test_gpu_multi.py: one process per GPU under torchrun, NCCL.  Checks the two north-star partitions on
real devices: coil sharding (C / N coils per rank, one NCCL sum all-reduce of the coil-combined image -- the coupling
point of the reference, torchkbnufft/modules/kbnufft.py:404-405) against the unsharded result of the same rank, and
batch sharding (no collective) against the oracle-free property that the gathered shards equal the full-batch result."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torchkbnufft_b200 as tkbn  # noqa: E402
from torchkbnufft_b200 import _lib, parallel, workloads  # noqa: E402


def rel(a, b):
    return float(torch.linalg.vector_norm(a - b) / torch.linalg.vector_norm(b))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    wl = workloads.WORKLOADS["cfg2"].scaled(0.25)
    image, smaps, kdata, omega = workloads.make_inputs(wl, seed=3, n_batch=4)  # same inputs on every rank
    x, s, y, om = (torch.from_numpy(a).to(dev) for a in (image, smaps, kdata, omega))
    nu = tkbn.KbNufft(im_size=wl.im_size, dtype=torch.complex64).to(dev)
    na = tkbn.KbNufftAdjoint(im_size=wl.im_size, dtype=torch.complex64).to(dev)
    want_k = nu(x, om, smaps=s)
    want_im = na(y, om, smaps=s)

    # coil sharding: forward stays sharded, adjoint ends with the all-reduce
    lo, hi = parallel.shard_bounds(wl.n_coils, rank, world)
    s_loc = s[:, lo:hi].contiguous()
    k_loc = parallel.coil_sharded_forward(nu, x, om, s_loc)
    assert rel(k_loc, want_k[:, lo:hi]) < 1e-6, "coil-sharded forward differs from the unsharded coils"
    im = parallel.coil_sharded_adjoint(na, y[:, lo:hi].contiguous(), om, s_loc)
    err = float(torch.linalg.vector_norm(im - want_im) / torch.linalg.vector_norm(want_im))
    assert err < 1e-5, f"coil-sharded adjoint: rel-L2 {err}"
    # every rank holds the same reduced image, bit for bit
    gathered = [torch.empty_like(im) for _ in range(world)]
    dist.all_gather(gathered, im)
    assert all(torch.equal(g, gathered[0]) for g in gathered), "ranks disagree after the all-reduce"

    # the same partition with the engine's own all-reduce over NVLink peer memory (b2n_peer_allreduce_sum): equal to
    # the NCCL result up to summation order, bit-identical on every rank, and stable over repeated calls of different
    # sizes (double-buffered slots, odd lengths take the scalar path)
    peer = parallel.PeerAllReduce(max_values=4 * want_im.numel(), dtype=torch.complex64)  # = 8 x numel floats
    for rep in range(5):
        im_p = parallel.coil_sharded_adjoint(na, y[:, lo:hi].contiguous(), om, s_loc, reducer=peer)
        err_p = rel(im_p, want_im)
        assert err_p < 1e-5, f"coil-sharded adjoint through the peer all-reduce: rel-L2 {err_p} (call {rep})"
        dist.all_gather(gathered, im_p)
        assert all(torch.equal(g, gathered[0]) for g in gathered), "ranks disagree after the peer all-reduce"
    # library-owned graph replay: the all-reduce kernel is the last launch of the adjoint's captured graph
    tkbn.set_graph_mode(True)
    y_loc = y[:, lo:hi].contiguous()
    launches = _lib.load().b2n_launch_count
    for rep in range(8):
        y_loc.copy_(y[:, lo:hi] * (1.0 + rep))
        before = launches()
        im_g = parallel.coil_sharded_adjoint(na, y_loc, om, s_loc, reducer=peer).clone()
        eager_launches = launches() - before
        assert rel(im_g, want_im * (1.0 + rep)) < 1e-5, f"graph-mode coil-sharded adjoint, call {rep}"
    assert eager_launches == 0, f"{eager_launches} eager launches in a replayed coil-sharded adjoint"
    tkbn.set_graph_mode(False)
    # Toeplitz normal operator with the coils sharded: NCCL and peer all-reduce against the unsharded apply
    toep = tkbn.ToepNufft()
    kern = tkbn.calc_toeplitz_kernel(om, wl.im_size)
    want_t = toep(x, kern, smaps=s)
    for red in (None, peer):
        got_t = parallel.coil_sharded_toeplitz(toep, x, kern, s_loc, reducer=red)
        assert rel(got_t, want_t) < 1e-5, f"coil-sharded Toeplitz ({'peer' if red else 'NCCL'}): rel-L2 {rel(got_t, want_t)}"
    gen = torch.Generator(device="cpu").manual_seed(100 + rank)
    for n in (1, 5, 4096, 4097, 12345, 3 * want_im.numel() + 1):
        mine = torch.randn(n, generator=gen).to(dev)
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        want_sum = parts[0].clone()
        for p in parts[1:]:
            want_sum += p  # rank order, as the kernel adds
        assert torch.equal(peer(mine.clone()), want_sum), f"peer all-reduce of {n} floats is not the rank-ordered sum"
    # the two-shot form (reduce-scatter + all-gather through the same windows), forced: the same rank-ordered sums, mixed
    # freely with one-shot calls (the scalar path of an odd length is always one-shot)
    lib = _lib.load()
    for form in (2, 1, 2):
        lib.b2n_set_option(_lib.OPT_PEER_FORM, form)
        for n in (4, 4096, 8192 + 4, 12345, 4 * 30001, 8 * want_im.numel()):
            mine = torch.randn(n, generator=gen).to(dev)
            parts = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(parts, mine)
            want_sum = parts[0].clone()
            for p in parts[1:]:
                want_sum += p
            assert torch.equal(peer(mine.clone()), want_sum), f"peer all-reduce (form {form}) of {n} floats"
    lib.b2n_set_option(_lib.OPT_PEER_FORM, 0)
    # captured into a CUDA graph: the call counter lives in the window, replays stay in step across ranks
    buf = torch.full((8192,), float(rank + 1), device=dev)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        peer(buf.clone())  # warm-up on the capture stream
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = peer(buf.clone())
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, torch.full_like(out, world * (world + 1) / 2)), "graph-replayed peer all-reduce is wrong"
    peer.close()

    # batch sharding: no collective on the data path; the shards put together are the full-batch result
    blo, bhi = parallel.shard_bounds(x.shape[0], rank, world)
    k_b, im_b = parallel.batch_sharded_pair(nu, na, x[blo:bhi].contiguous(), om, s)
    want_pair = na(want_k, om, smaps=s)
    assert rel(k_b, want_k[blo:bhi]) < 1e-6, "batch shard of the forward differs"
    assert rel(im_b, want_pair[blo:bhi]) < 1e-6, "batch shard of the adjoint differs"
    dist.barrier()
    if rank == 0:
        print(f"MGPU_OK world={world} coil_sharded_rel_l2={err:.3e} peer_allreduce_rel_l2={err_p:.3e}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
