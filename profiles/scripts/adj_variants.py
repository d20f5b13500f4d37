"""Adjoint-spread variants (B2N_OPT_ADJ_ROW_OWNERSHIP) timed alone at 16 coils, in one process.
python profiles/scripts/adj_variants.py [cfg2 cfg5s ...] [--variants=0,3,4]"""
import os, statistics, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torchkbnufft_b200 as tkbn
from torchkbnufft_b200 import workloads, _lib
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

variants = [0, 3, 4, 5, 2]
names = []
for a in sys.argv[1:]:
    if a.startswith("--variants="):
        variants = [int(v) for v in a.split("=")[1].split(",")]
    else:
        names.append(a)
lib = _lib.load()
for name in names or ["cfg2"]:
    wl = workloads.WORKLOADS[name]
    om = torch.from_numpy(wl.trajectory(np.float32)).to(dev)
    ob = tkbn.KbInterp(im_size=wl.im_size, dtype=torch.complex64).to(dev)
    args = (ob.tables, ob.n_shift, ob.numpoints, ob.table_oversamp)
    B = min(wl.n_batch, 8)
    y = torch.randn((B, 16, om.shape[-1]), dtype=torch.complex64, device=dev)
    ref = None
    for v in variants:
        lib.b2n_set_option(_lib.OPT_ADJ_ROW_OWNERSHIP, v)
        fn = lambda: eng_interp.table_interp_adjoint(y, om, *args, None, ob.grid_size, mode="atomic")
        out = fn()
        if ref is None:
            ref = out
        err = float((out - ref).norm() / ref.norm())
        t = timeit(fn)
        print(f"{name} B={B} C=16 variant {v}: adjoint interp {t:8.1f} us   rel diff vs variant {variants[0]}: {err:.2e}", flush=True)
    lib.b2n_set_option(_lib.OPT_ADJ_ROW_OWNERSHIP, 0)
