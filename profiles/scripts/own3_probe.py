"""Config 4 (3-D kooshball) adjoint interpolation alone: visit statistics of the plan and the time of the launch.
python profiles/scripts/own3_probe.py [fraction of spokes]"""
import os, statistics, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torchkbnufft_b200 as tkbn
from torchkbnufft_b200 import workloads
from torchkbnufft_b200._nufft import interp as eng_interp, plan as P

dev = torch.device("cuda:0")
wl = workloads.WORKLOADS["cfg4"]
if len(sys.argv) > 1:
    wl = wl.scaled(float(sys.argv[1]))
FWD = len(sys.argv) > 2 and sys.argv[2] == "fwd"  # also run the forward gather once (for ncu captures)
om = torch.from_numpy(wl.trajectory(np.float32)).to(dev)
ob = tkbn.KbInterp(im_size=wl.im_size, dtype=torch.complex64).to(dev)
args = (ob.tables, ob.n_shift, ob.numpoints, ob.table_oversamp)
y = torch.randn((1, wl.n_coils, om.shape[-1]), dtype=torch.complex64, device=dev)
fn = lambda: eng_interp.table_interp_adjoint(y, om, *args, None, ob.grid_size, mode="atomic")
out = fn(); torch.cuda.synchronize()
if FWD:
    k = eng_interp.table_interp(out, om, *args, None); torch.cuda.synchronize()
pl = list(P._PLAN_CACHE.values())[-1]
st = pl.struct
base = pl.workspace.data_ptr()
cnt = pl.workspace[int(st.own_counts) - base:int(st.own_counts) - base + 12].view(torch.int32).cpu().tolist()
nt = st.n_own_tiles[0] * st.n_own_tiles[1] * st.n_own_tiles[2]
tiles = pl.workspace[int(st.own_tiles) - base:int(st.own_tiles) - base + 16 * nt].view(torch.int32).view(nt, 4).cpu()
visits = int(tiles[:, 0].sum())
print(f"points {om.shape[-1]}  tiles {nt}  visits {visits} ({visits / om.shape[-1]:.2f} per point)  items {cnt[0]}  "
      f"partial slots {cnt[1]}  exceptions {cnt[2]}  max visits per tile {int(tiles[:, 0].max())}  cap {st.own_cap}", flush=True)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
ts = []
for r in range(5):
    flush.fill_(r)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
print(f"adjoint interp {statistics.median(ts):.3f} ms", flush=True)
