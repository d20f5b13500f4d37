"""A/B inside ONE process of B2N_OPT_PDL x B2N_OPT_FFT_PREFETCH on the forward and the adjoint NUFFT separately."""
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

for name in sys.argv[1:] or ["cfg3"]:
    wl = workloads.WORKLOADS[name]
    image, smaps, kdata, omega = workloads.make_inputs(wl, seed=0)
    nu = tkbn.KbNufft(im_size=wl.im_size, dtype=torch.complex64).to(dev)
    na = tkbn.KbNufftAdjoint(im_size=wl.im_size, dtype=torch.complex64).to(dev)
    x, s, y, om = (torch.from_numpy(a).to(dev) for a in (image, smaps, kdata, omega))
    kern = tkbn.calc_toeplitz_kernel(om, wl.im_size, norm="ortho")
    toep = tkbn.ToepNufft()
    for rep in range(2):
        for pdl in (3, 1):
            for pf in (19,):
                lib.b2n_set_option(_lib.OPT_PDL, pdl)
                lib.b2n_set_option(_lib.OPT_FFT_PREFETCH, pf)
                tf = timed(lambda: nu(x, om, smaps=s))
                ta = timed(lambda: na(y, om, smaps=s))
                tt = timed(lambda: toep(x, kern, smaps=s, norm="ortho"))
                print(f"{name} rep{rep} pdl={pdl} prefetch={pf}: fwd {tf:.1f} us  adj {ta:.1f} us  toeplitz {tt:.1f} us")
    lib.b2n_set_option(_lib.OPT_PDL, 1)
    lib.b2n_set_option(_lib.OPT_FFT_PREFETCH, 19)
