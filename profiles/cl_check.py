"""Channel-last tiled spread vs the coil-major one (cfg2 geometry): agreement and time.  python profiles/cl_check.py"""
import os, statistics, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torchkbnufft_b200 as tkbn
from torchkbnufft_b200 import workloads, _lib
from torchkbnufft_b200._nufft import interp as ei
dev = torch.device("cuda:0")
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, reps=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for r in range(reps):
        flush.fill_(r & 255)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
    return statistics.median(ts)
for name, C, B in (("cfg2", 16, 1), ("cfg3", 32, 1), ("cfg5", 16, 4)):
    wl = workloads.WORKLOADS[name]
    om = torch.from_numpy(wl.trajectory(np.float32)).to(dev)
    ob = tkbn.KbInterp(im_size=wl.im_size, dtype=torch.complex64).to(dev)
    args = (ob.tables, ob.n_shift, ob.numpoints, ob.table_oversamp)
    y = torch.randn((B, C, om.shape[-1]), dtype=torch.complex64, device=dev)
    cm = ei.table_interp_adjoint(y, om, *args, None, ob.grid_size, mode="atomic")
    cl = ei.table_interp_adjoint(y, om, *args, None, ob.grid_size, mode="atomic", layout=_lib.CHANNEL_LAST)
    err = float(torch.linalg.norm(cl.movedim(-1, 1) - cm) / torch.linalg.norm(cm))
    t_cm = timeit(lambda: ei.table_interp_adjoint(y, om, *args, None, ob.grid_size, mode="atomic"))
    t_cl = timeit(lambda: ei.table_interp_adjoint(y, om, *args, None, ob.grid_size, mode="atomic", layout=_lib.CHANNEL_LAST))
    print(f"{name} B={B} C={C}: channel-last vs coil-major rel err {err:.2e}; spread coil-major {t_cm:.1f} us, channel-last {t_cl:.1f} us", flush=True)
