"""One forward + one adjoint SENSE NUFFT of a BASELINE config (for ncu captures).  python profiles/run_cfg.py cfg4 [reps]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torchkbnufft_b200 as tkbn
from torchkbnufft_b200 import workloads
dev = torch.device("cuda:0")
wl = workloads.WORKLOADS[sys.argv[1]]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
image, smaps, kdata, omega = workloads.make_inputs(wl, seed=0, n_batch=min(wl.n_batch, 8))
x, s, y, om = (torch.from_numpy(a).to(dev) for a in (image, smaps, kdata, omega))
nu = tkbn.KbNufft(im_size=wl.im_size, dtype=torch.complex64).to(dev)
na = tkbn.KbNufftAdjoint(im_size=wl.im_size, dtype=torch.complex64).to(dev)
kw = dict(smaps=s) if wl.n_coils > 1 else {}
for _ in range(reps):
    k = nu(x, om, **kw); im = na(y, om, **kw)
torch.cuda.synchronize()
print("done", float(k.abs().sum()), float(im.abs().sum()))
