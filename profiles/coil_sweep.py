"""Interpolation-only times vs coil count, tiled vs generic kernels (cfg2 geometry: 640^2 grid, M = 128 000).
Decides the dispatch threshold in b2n_interp.cu.   python profiles/coil_sweep.py"""
import os, statistics, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torchkbnufft_b200 as tkbn
from torchkbnufft_b200 import workloads
from torchkbnufft_b200._nufft import interp as eng_interp

dev = torch.device("cuda:0")
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

def timeit(fn, reps=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for r in range(reps):
        flush.fill_(r & 255)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return statistics.median(ts)

args_in = [a for a in sys.argv[1:] if not a.startswith("--opt1=")]
for a in sys.argv[1:]:
    if a.startswith("--opt1="):  # adjoint variant (B2N_OPT_ADJ_ROW_OWNERSHIP) for A/B runs
        from torchkbnufft_b200 import _lib
        _lib.load().b2n_set_option(_lib.OPT_ADJ_ROW_OWNERSHIP, int(a.split("=")[1]))
for name in args_in or ["cfg2", "cfg1"]:
    wl = workloads.WORKLOADS[name]
    om = torch.from_numpy(wl.trajectory(np.float32)).to(dev)
    ob = tkbn.KbInterp(im_size=wl.im_size, dtype=torch.complex64).to(dev)
    args = (ob.tables, ob.n_shift, ob.numpoints, ob.table_oversamp)
    K = tuple(int(k) for k in ob.grid_size)
    for C in (1, 2, 3, 4, 6, 8, 12, 16, 32):
        g = torch.randn((1, C) + K, dtype=torch.complex64, device=dev)
        y = torch.randn((1, C, om.shape[-1]), dtype=torch.complex64, device=dev)
        row = f"{name} C={C:2d}"
        for tiled in ("force", False):
            tkbn.set_tiled_kernels(tiled)
            tf = timeit(lambda: eng_interp.table_interp(g, om, *args))
            ta = timeit(lambda: eng_interp.table_interp_adjoint(y, om, *args, None, ob.grid_size, mode="atomic"))
            row += f"   {'tiled  ' if tiled else 'generic'} fwd {tf:7.1f} us adj {ta:7.1f} us"
        tkbn.set_tiled_kernels(True)
        print(row, flush=True)
