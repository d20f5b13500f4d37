"""Cost of building one trajectory plan (the per-step setup of the end-to-end path when the trajectory changes):
python profiles/scripts/plan_build.py [cfg2 cfg5 ...]   (run under `ncu --metrics gpu__time_duration.sum` for the
per-kernel list)"""
import os, statistics, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torchkbnufft_b200 as tkbn
from torchkbnufft_b200 import workloads
from torchkbnufft_b200._nufft import plan as P

dev = torch.device("cuda:0")
for name in sys.argv[1:] or ["cfg2"]:
    wl = workloads.WORKLOADS[name]
    om = torch.from_numpy(wl.trajectory(np.float32)).to(dev)
    ob = tkbn.KbInterp(im_size=wl.im_size, dtype=torch.complex64).to(dev)
    geo = P.get_geometry(ob.tables, ob.n_shift, ob.numpoints, ob.table_oversamp, ob.grid_size)
    ts = []
    for r in range(8):
        tkbn.clear_caches()
        geo = P.get_geometry(ob.tables, ob.n_shift, ob.numpoints, ob.table_oversamp, ob.grid_size)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); pl = P.TrajectoryPlan(geo, om); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    print(f"{name}: plan build {statistics.median(ts[2:]):8.1f} us (median of 6), workspace {pl.workspace.numel() / 1e6:.1f} MB",
          flush=True)
