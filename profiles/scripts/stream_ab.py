"""A/B inside ONE process of B2N_OPT_FFT_STREAM (streamed persistent FFT passes) on the forward and the adjoint SENSE
NUFFT, with the results compared bit for bit against the classic passes.  python profiles/scripts/stream_ab.py cfg2 cfg3 -- 0 1 2 3"""
import sys, os, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torchkbnufft_b200 as tkbn
from torchkbnufft_b200 import _lib, workloads

dev = torch.device("cuda:0")
lib = _lib.load()
flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)

def timed(fn, n=60):
    for _ in range(5):
        fn()
    st = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
    en = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
    torch.cuda.synchronize()
    for i in range(n):
        flush.fill_(i & 0xFF)
        st[i].record()
        fn()
        en[i].record()
    torch.cuda.synchronize()
    return 1e3 * statistics.median(a.elapsed_time(b) for a, b in zip(st, en))

args = sys.argv[1:]
names = [a for a in args if a.startswith("cfg")] or ["cfg2"]
masks = [int(a) for a in args if a.isdigit()] or [0, 1, 2, 3]
for name in names:
    wl = workloads.WORKLOADS[name]
    image, smaps, kdata, omega = workloads.make_inputs(wl, seed=0, n_batch=min(wl.n_batch, 8))
    nu = tkbn.KbNufft(im_size=wl.im_size, dtype=torch.complex64).to(dev)
    na = tkbn.KbNufftAdjoint(im_size=wl.im_size, dtype=torch.complex64).to(dev)
    x, s, y, om = (torch.from_numpy(a).to(dev) for a in (image, smaps, kdata, omega))
    kw = dict(smaps=s) if wl.n_coils > 1 else {}
    lib.b2n_set_option(_lib.OPT_FFT_STREAM, 0)
    ref_f, ref_a = nu(x, om, **kw).clone(), na(y, om, **kw).clone()
    for rep in range(2):
        for m in masks:
            lib.b2n_set_option(_lib.OPT_FFT_STREAM, m)
            df = float((nu(x, om, **kw) - ref_f).abs().max())
            da = float((na(y, om, **kw) - ref_a).abs().max())
            tf = timed(lambda: nu(x, om, **kw))
            ta = timed(lambda: na(y, om, **kw))
            print(f"{name} rep{rep} stream={m:3d}: fwd {tf:6.1f} us  adj {ta:6.1f} us  pair {tf + ta:6.1f} us   max|diff| fwd {df:.1e} adj {da:.1e}", flush=True)
    lib.b2n_set_option(_lib.OPT_FFT_STREAM, 0)
