"""Host-side cost of the end-to-end step when the trajectory changes every step (plan rebuilt): wall time of the
host calls with nothing to wait for, then a cProfile of the same loop.  python profiles/scripts/e2e_host.py [cfg2]"""
import cProfile, os, pstats, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torchkbnufft_b200 as tkbn
from torchkbnufft_b200 import workloads
dev = torch.device("cuda:0")
wl = workloads.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
image, smaps, kdata, omega = workloads.make_inputs(wl, seed=0)
x, s, y, om = (torch.from_numpy(a).to(dev) for a in (image, smaps, kdata, omega))
hom = torch.from_numpy(omega).pin_memory()
nu = tkbn.KbNufft(im_size=wl.im_size, dtype=torch.complex64).to(dev)
na = tkbn.KbNufftAdjoint(im_size=wl.im_size, dtype=torch.complex64).to(dev)
kw = dict(smaps=s) if wl.n_coils > 1 else {}

def step():
    om.copy_(hom, non_blocking=True)  # new trajectory contents -> plan is rebuilt
    k = nu(x, om, **kw)
    return na(y, om, **kw)

for _ in range(5):
    step()
torch.cuda.synchronize()
N = 50
t0 = time.perf_counter()
for _ in range(N):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"{wl.name if hasattr(wl, 'name') else ''} host time per step (calls only) {(t1 - t0) / N * 1e6:8.1f} us; with the final drain {(t2 - t0) / N * 1e6:8.1f} us",
      flush=True)
pr = cProfile.Profile()
pr.enable()
for _ in range(N):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr); st.sort_stats("tottime"); st.print_stats(22)
