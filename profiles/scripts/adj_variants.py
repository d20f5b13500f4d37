"""Adjoint-spread variants timed alone at 16 coils, in one process: the shared-memory tiled kernels
(B2N_OPT_ADJ_ROW_OWNERSHIP values) and the owner-tile register spread with several work-item caps.
python profiles/scripts/adj_variants.py [cfg2 cfg5 ...] [--variants=0,3,4] [--caps=32,64,128] [--rows=4,8] [--owned=1,5] [--coils=16]"""
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

variants, caps, names, coils, owned, rows_list = [0], [32, 64, 128], [], [16], [1], [4]
for a in sys.argv[1:]:
    if a.startswith("--variants="):
        variants = [int(v) for v in a.split("=")[1].split(",") if v]
    elif a.startswith("--caps="):
        caps = [int(v) for v in a.split("=")[1].split(",") if v]
    elif a.startswith("--owned="):
        owned = [int(v) for v in a.split("=")[1].split(",") if v]
    elif a.startswith("--rows="):
        rows_list = [int(v) for v in a.split("=")[1].split(",") if v]
    elif a.startswith("--coils="):
        coils = [int(v) for v in a.split("=")[1].split(",") if v]
    else:
        names.append(a)
lib = _lib.load()
for name in names or ["cfg2"]:
    wl = workloads.WORKLOADS[name]
    om = torch.from_numpy(wl.trajectory(np.float32)).to(dev)
    ob = tkbn.KbInterp(im_size=wl.im_size, dtype=torch.complex64).to(dev)
    args = (ob.tables, ob.n_shift, ob.numpoints, ob.table_oversamp)
    B = min(wl.n_batch, 8)
    for C in coils:
        y = torch.randn((B, C, om.shape[-1]), dtype=torch.complex64, device=dev)
        fn = lambda: eng_interp.table_interp_adjoint(y, om, *args, None, ob.grid_size, mode="atomic")
        ref = None
        eng_interp.owned_spread = False
        for v in variants:
            lib.b2n_set_option(_lib.OPT_ADJ_ROW_OWNERSHIP, v)
            out = fn()
            ref = out if ref is None else ref
            err = float((out - ref).norm() / ref.norm())
            print(f"{name} B={B} C={C} tiled variant {v}: adjoint interp {timeit(fn):8.1f} us   rel diff {err:.2e}", flush=True)
        lib.b2n_set_option(_lib.OPT_ADJ_ROW_OWNERSHIP, 0)
        eng_interp.owned_spread = True
        for rows, cap, ow in [(r, c, o) for r in rows_list for o in owned for c in caps]:
            lib.b2n_set_option(_lib.OPT_ADJ_OWNED, ow)
            lib.b2n_set_option(_lib.OPT_OWN_CAP, cap)
            tkbn.clear_caches()
            out = fn(); out2 = fn()
            torch.cuda.synchronize()
            err = float((out - ref).norm() / ref.norm()) if ref is not None else float("nan")
            same = bool(torch.equal(out, out2))
            from torchkbnufft_b200._nufft import plan as P
            pl = list(P._PLAN_CACHE.values())[-1]
            fn(); torch.cuda.synchronize(); fn()
            print(f"{name} B={B} C={C} owner tiles rows {rows} (opt {ow}) cap {cap:4d}: adjoint interp {timeit(fn):8.1f} us   rel diff {err:.2e} "
                  f"bit-reproducible {same} items {pl.struct.n_own_items_max} slots {pl.own_slots}", flush=True)
        lib.b2n_set_option(_lib.OPT_OWN_CAP, 0)
        lib.b2n_set_option(_lib.OPT_ADJ_OWNED, 1)
        tkbn.clear_caches()
