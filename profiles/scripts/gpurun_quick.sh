for i in 1 2; do python profiles/setup_costs.py 2>&1 | grep -E "cfg[13]: calc_density"; done
python - <<'PY'
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import torchkbnufft_b200 as tkbn
from torchkbnufft_b200 import workloads, _lib
from torchkbnufft_b200._nufft import interp as ei
dev = torch.device('cuda:0')
wl = workloads.WORKLOADS['cfg1']
om = torch.from_numpy(wl.trajectory(np.float32)).to(dev)
ob = tkbn.KbInterp(im_size=wl.im_size, dtype=torch.complex64).to(dev)
args = (ob.tables, ob.n_shift, ob.numpoints, ob.table_oversamp)
w = torch.ones(1, 1, om.shape[1], dtype=torch.complex64, device=dev)
def it(n):
    global w
    for _ in range(n):
        g = ei.table_interp_adjoint(w, om, *args, None, ob.grid_size)
        r = ei.table_interp(g, om, *args)
it(5); torch.cuda.synchronize()
for opt in (0, 3):
    _lib.load().b2n_set_option(1, opt)
    it(5); torch.cuda.synchronize()
    t0 = time.perf_counter(); it(100); torch.cuda.synchronize(); t = (time.perf_counter() - t0) / 100
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); it(100); e1.record(); torch.cuda.synchronize()
    print(f"adj variant {opt}: wall {t*1e6:.1f} us/iter  gpu {e0.elapsed_time(e1)*10:.1f} us/iter")
PY
