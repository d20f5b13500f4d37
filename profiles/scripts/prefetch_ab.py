"""A/B inside ONE process: forward+adjoint pair with and without programmatic dependent launch (B2N_OPT_FFT_PREFETCH).
Also checks that the results are identical (the forward is deterministic; the adjoint to rounding)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import torchkbnufft_b200 as tkbn
from torchkbnufft_b200 import _lib, workloads

dev = torch.device("cuda:0")
lib = _lib.load()
flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)
for name in sys.argv[1:] or ["cfg2", "cfg1"]:
    wl = workloads.WORKLOADS[name]
    image, smaps, kdata, omega = workloads.make_inputs(wl, seed=0)
    nu = tkbn.KbNufft(im_size=wl.im_size, dtype=torch.complex64).to(dev)
    na = tkbn.KbNufftAdjoint(im_size=wl.im_size, dtype=torch.complex64).to(dev)
    x, s, y, om = (torch.from_numpy(a).to(dev) for a in (image, smaps, kdata, omega))

    def step():
        k = nu(x, om, smaps=s)
        return k, na(k, om, smaps=s)

    def timed(n=100):
        for _ in range(5):
            step()
        st = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
        en = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
        torch.cuda.synchronize()
        for i in range(n):
            flush.fill_(i & 0xFF)
            st[i].record()
            step()
            en[i].record()
        torch.cuda.synchronize()
        t = sorted(a.elapsed_time(b) for a, b in zip(st, en))
        return 1e3 * sum(t) / n, 1e3 * t[n // 2]

    lib.b2n_set_option(_lib.OPT_FFT_PREFETCH, 0)
    k0, a0 = step()
    lib.b2n_set_option(_lib.OPT_FFT_PREFETCH, 19)
    k1, a1 = step()
    torch.cuda.synchronize()
    print(name, "fwd identical:", bool(torch.equal(k0, k1)), " adj rel diff:",
          float((a0 - a1).norm() / a0.norm()))
    for rep in range(3):
        for pdl in (0, 3, 15):
            lib.b2n_set_option(_lib.OPT_FFT_PREFETCH, pdl)
            mean, med = timed()
            print(f"{name} rep{rep} prefetch={pdl}: mean {mean:.1f} us  median {med:.1f} us")
    lib.b2n_set_option(_lib.OPT_FFT_PREFETCH, 0)
    lib.b2n_set_option(_lib.OPT_FFT_PREFETCH, 0)
# ToepNufft apply (three-pass route) at the same settings
from torchkbnufft_b200._autograd import nufft as auto_nufft
for name in sys.argv[1:] or ["cfg2", "cfg1"]:
    wl = workloads.WORKLOADS[name]
    image, smaps, kdata, omega = workloads.make_inputs(wl, seed=0)
    x, s, om = (torch.from_numpy(a).to(dev) for a in (image, smaps, omega))
    kern = tkbn.calc_toeplitz_kernel(om, wl.im_size, norm="ortho")
    toep = tkbn.ToepNufft()
    for rep in range(2):
        for pf in (0, 3, 16, 19, 31):
            lib.b2n_set_option(_lib.OPT_FFT_PREFETCH, pf)
            for _ in range(5):
                toep(x, kern, smaps=s, norm="ortho")
            n = 100
            st = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
            en = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
            torch.cuda.synchronize()
            for i in range(n):
                flush.fill_(i & 0xFF)
                st[i].record()
                toep(x, kern, smaps=s, norm="ortho")
                en[i].record()
            torch.cuda.synchronize()
            t = sorted(a.elapsed_time(b) for a, b in zip(st, en))
            print(f"{name} toeplitz apply rep{rep} prefetch={pf}: median {1e3 * t[n // 2]:.1f} us")
    lib.b2n_set_option(_lib.OPT_FFT_PREFETCH, 19)
