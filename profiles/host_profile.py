"""cProfile of the host path of one forward+adjoint pair (cfg2).  python profiles/host_profile.py"""
import cProfile, os, pstats, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torchkbnufft_b200 as tkbn
from torchkbnufft_b200 import workloads
dev = torch.device("cuda:0")
wl = workloads.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
image, smaps, kdata, omega = workloads.make_inputs(wl, seed=0)
x, s, y, om = (torch.from_numpy(a).to(dev) for a in (image, smaps, kdata, omega))
nu = tkbn.KbNufft(im_size=wl.im_size, dtype=torch.complex64).to(dev)
na = tkbn.KbNufftAdjoint(im_size=wl.im_size, dtype=torch.complex64).to(dev)
kw = dict(smaps=s) if wl.n_coils > 1 else {}
for _ in range(5):
    k = nu(x, om, **kw); na(k, om, **kw)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(300):
    k = nu(x, om, **kw); na(k, om, **kw)
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr); st.sort_stats("tottime"); st.print_stats(28)
