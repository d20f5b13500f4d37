"""Adjoint SENSE NUFFT with 2 coils (what a rank of an 8-way coil split of config 2 runs): the coil sum fused into the
row pass (mostly idle CTAs) against the unfused route (row pass + crop/coil-sum kernel).  A/B in one process."""
import sys, os, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torchkbnufft_b200 as tkbn
from torchkbnufft_b200 import _lib, workloads
dev = torch.device("cuda:0")
lib = _lib.load()
flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)
def timed(fn, n=100):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = 0.0
    for i in range(n):
        flush.fill_(i & 0xFF)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return 1e3 * tot / n
wl = workloads.WORKLOADS["cfg2"]
image, smaps, kdata, omega = workloads.make_inputs(wl, seed=0)
na = tkbn.KbNufftAdjoint(im_size=wl.im_size, dtype=torch.complex64).to(dev)
for C in (2, 4):
    s, y, om = (torch.from_numpy(a).to(dev) for a in (smaps[:, :C].copy(), kdata[:, :C].copy(), omega))
    ref = None
    for rep in range(2):
        for mask in (1, 129):
            lib.b2n_set_option(_lib.OPT_FFT_STREAM, mask)
            out = na(y, om, smaps=s)
            ref = out.clone() if ref is None else ref
            err = float((out - ref).abs().max() / ref.abs().max())
            print(f"C={C} rep{rep} option {mask:3d}: adjoint {timed(lambda: na(y, om, smaps=s)):6.1f} us   max rel diff vs first {err:.1e}", flush=True)
lib.b2n_set_option(_lib.OPT_FFT_STREAM, 1)
