"""NCCL all-reduce vs the engine's peer-memory all-reduce (b2n_peer_allreduce_sum) on the coil-combined image of
BASELINE config 2 (320 x 320 complex64 = 0.8 MB) and a few other sizes; device time (CUDA events), max over ranks,
with and without an L2 flush before every call.  torchrun --nproc-per-node N profiles/scripts/peer_ab.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.distributed as dist
from torchkbnufft_b200 import _lib, parallel

lib = _lib.load()

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timed(fn, n=40, do_flush=True):
    for _ in range(8):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    e0 = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
    e1 = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
    for i in range(n):
        if do_flush:
            flush.fill_(i & 0xFF)
        e0[i].record()
        fn()
        e1[i].record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in zip(e0, e1))
    t = torch.tensor([sum(ms) / n, ms[n // 2]], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return 1e3 * float(t[0]), 1e3 * float(t[1])


for values in (320 * 320, 64 * 1024, 256 * 256 * 8, 1 << 22):
    buf = torch.randn(values, dtype=torch.complex64, device=dev)
    peer = parallel.PeerAllReduce(max_values=values, dtype=torch.complex64)
    lib.b2n_set_option(_lib.OPT_PEER_FORM, 1)
    for do_flush in (True, False):
        t_nccl = timed(lambda: parallel.all_reduce_complex_(buf), do_flush=do_flush)
        t_peer = timed(lambda: peer(buf), do_flush=do_flush)
        lib.b2n_set_option(_lib.OPT_PEER_FORM, 2)
        t_two = timed(lambda: peer(buf), do_flush=do_flush)
        t_two10 = timed(lambda: [peer(buf) for _ in range(10)], do_flush=do_flush)
        lib.b2n_set_option(_lib.OPT_PEER_FORM, 1)
        # back to back without events in between: 10 calls per timed region
        t_peer10 = timed(lambda: [peer(buf) for _ in range(10)], do_flush=do_flush)
        t_nccl10 = timed(lambda: [parallel.all_reduce_complex_(buf) for _ in range(10)], do_flush=do_flush)
        if rank == 0:
            print(f"world {world} {values * 8 / 1e6:7.2f} MB flush={int(do_flush)}: NCCL mean {t_nccl[0]:6.1f} median {t_nccl[1]:6.1f} us | "
                  f"peer one-shot mean {t_peer[0]:6.1f} median {t_peer[1]:6.1f} us | two-shot mean {t_two[0]:6.1f} median {t_two[1]:6.1f} us | "
                  f"10 back to back: NCCL {t_nccl10[0] / 10:6.1f} one-shot {t_peer10[0] / 10:6.1f} two-shot {t_two10[0] / 10:6.1f} us per call",
                  flush=True)
    lib.b2n_set_option(_lib.OPT_PEER_FORM, 0)
    peer.close()
dist.destroy_process_group()
