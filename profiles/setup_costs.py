"""Steady-state cost of the setup-side callers (density compensation, Toeplitz kernel).  python profiles/setup_costs.py"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torchkbnufft_b200 as tkbn
from torchkbnufft_b200 import workloads
dev = torch.device("cuda:0")

def wall(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return min(ts) * 1e3

for name in ("cfg1", "cfg3", "cfg4"):
    wl = workloads.WORKLOADS[name]
    om = torch.from_numpy(wl.trajectory(np.float32)).to(dev)
    if name != "cfg4":
        t = wall(lambda: tkbn.calc_density_compensation_function(om, wl.im_size, num_iterations=10))
        print(f"{name}: calc_density_compensation_function(10 it)  {t:8.2f} ms", flush=True)
    t = wall(lambda: tkbn.calc_toeplitz_kernel(om, wl.im_size, norm="ortho"), reps=3)
    print(f"{name}: calc_toeplitz_kernel                      {t:8.2f} ms", flush=True)
    om2 = om.clone()
    t = wall(lambda: tkbn.calc_toeplitz_kernel(om2.add_(0), wl.im_size, norm="ortho"), reps=3)
    print(f"{name}: calc_toeplitz_kernel, new trajectory        {t:8.2f} ms", flush=True)
