"""Per-CTA timeline of the tiled forward kernel (development aid): python profiles/trace_cta.py"""
import ctypes, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchkbnufft_b200 as tkbn
from torchkbnufft_b200 import _lib, workloads
from torchkbnufft_b200._nufft import interp as eng

wl = workloads.WORKLOADS["cfg2"]
image, smaps, kdata, omega = workloads.make_inputs(wl, seed=0)
dev = torch.device("cuda:0")
ob = tkbn.KbInterp(im_size=wl.im_size, dtype=torch.complex64).to(dev)
grid = torch.randn((1, wl.n_coils) + wl.grid_size, dtype=torch.complex64, device=dev)
om = torch.from_numpy(omega).to(dev)
args = (ob.tables, ob.n_shift, ob.numpoints, ob.table_oversamp)
for _ in range(3):
    eng.table_interp(grid, om, *args)
cap = 8192
buf = torch.zeros(cap * 6, dtype=torch.int64, device=dev)
lib = _lib.load()
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev); flush.fill_(1)
which = sys.argv[1] if len(sys.argv) > 1 else "fwd"
if len(sys.argv) > 2:
    lib.b2n_set_option(int(sys.argv[2].split("=")[0]), int(sys.argv[2].split("=")[1]))
kd = torch.randn((1, wl.n_coils, wl.n_points), dtype=torch.complex64, device=dev)
if which == "adj":
    for _ in range(2):
        eng.table_interp_adjoint(kd, om, *args, None, ob.grid_size, mode="atomic")
    flush.fill_(2)
lib.b2n_set_trace_buffer(ctypes.c_void_p(buf.data_ptr()), cap)
if which == "adj":
    eng.table_interp_adjoint(kd, om, *args, None, ob.grid_size, mode="atomic")
else:
    eng.table_interp(grid, om, *args)
torch.cuda.synchronize()
lib.b2n_set_trace_buffer(None, 0)
print("kernel:", which, sys.argv[2:] )
r = buf.cpu().numpy().reshape(cap, 6)
r = r[r[:, 2] > 0]
t0 = r[:, 2].min()
start, staged, done = (r[:, 2] - t0) / 1e3, (r[:, 3] - t0) / 1e3, (r[:, 4] - t0) / 1e3
print(f"CTAs traced {len(r)}  kernel span {done.max():.1f} us")
print(f"staging  mean {np.mean(staged - start):.2f} us  p50 {np.median(staged - start):.2f}  p95 {np.percentile(staged - start, 95):.2f}  max {np.max(staged - start):.2f}")
print(f"compute  mean {np.mean(done - staged):.2f} us  p50 {np.median(done - staged):.2f}  p95 {np.percentile(done - staged, 95):.2f}  max {np.max(done - staged):.2f}")
if which == "fwd":
    issued = ((r[:, 5] >> 8) & 0xFFFFF) / 1e3
    records = (r[:, 0] >> 32) / 1e3
    r[:, 0] &= 0xFFFFFFFF
    r[:, 5] &= 0xFF
    print(f"fwd staging detail: copies issued at +{np.median(issued):.2f} us, records landed at +{np.median(records):.2f} us, "
          f"tile landed at +{np.median(staged - start):.2f} us (medians)")
pts = r[:, 1]
print(f"points per CTA mean {pts.mean():.1f} max {pts.max()}  compute ns/point {1e3 * np.sum(done - staged) / pts.sum():.1f}")
for tma in (0, 1):
    m = r[:, 5] == tma
    if m.any():
        print(f"tma={tma}: n={m.sum()} staging mean {np.mean((staged - start)[m]):.2f} us compute mean {np.mean((done - staged)[m]):.2f} us points {pts[m].mean():.1f}")
# busy time per SM
for q in (0, 25, 50, 75, 100):
    print(f"start-time percentile {q}: {np.percentile(start, q):.1f} us   end-time percentile {q}: {np.percentile(done, q):.1f} us")
sm_busy = {}
for sm, a_, b_ in zip(r[:, 0], start, done):
    sm_busy.setdefault(sm, []).append((a_, b_))
ends = [max(b for _, b in v) for v in sm_busy.values()]
print(f"SMs used {len(sm_busy)}  per-SM last end: min {min(ends):.1f} median {np.median(ends):.1f} max {max(ends):.1f} us")
conc = np.mean([len(v) for v in sm_busy.values()])
print(f"CTAs per SM mean {conc:.1f}")
