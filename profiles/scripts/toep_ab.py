"""A/B inside ONE process: ToepNufft apply at BASELINE config 3 (and config 2's shape) with the three-pass route
(column pass transforms / filters / transforms back) against the four-pass route through the full spectrum."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torchkbnufft_b200 as tkbn
from torchkbnufft_b200 import workloads
from torchkbnufft_b200._autograd import nufft as auto_nufft

dev = torch.device("cuda:0")
flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)
for name in sys.argv[1:] or ["cfg3", "cfg2"]:
    wl = workloads.WORKLOADS[name]
    image, smaps, kdata, omega = workloads.make_inputs(wl, seed=0)
    x, s, om = (torch.from_numpy(a).to(dev) for a in (image, smaps, omega))
    kern = tkbn.calc_toeplitz_kernel(om, wl.im_size, norm="ortho")
    toep = tkbn.ToepNufft()

    def timed(n=100):
        for _ in range(5):
            toep(x, kern, smaps=s, norm="ortho")
        st = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
        en = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
        torch.cuda.synchronize()
        for i in range(n):
            flush.fill_(i & 0xFF)
            st[i].record()
            toep(x, kern, smaps=s, norm="ortho")
            en[i].record()
        torch.cuda.synchronize()
        t = sorted(a.elapsed_time(b) for a, b in zip(st, en))
        return 1e3 * sum(t) / n, 1e3 * t[n // 2]

    res = {}
    for fuse in (True, False):
        auto_nufft.fuse_toeplitz_columns = fuse
        res[fuse] = toep(x, kern, smaps=s, norm="ortho")
    print(name, "three-pass vs four-pass rel diff:", float((res[True] - res[False]).norm() / res[False].norm()))
    for rep in range(2):
        for fuse in (False, True):
            auto_nufft.fuse_toeplitz_columns = fuse
            mean, med = timed()
            print(f"{name} rep{rep} three_pass={fuse}: mean {mean:.1f} us  median {med:.1f} us")
    auto_nufft.fuse_toeplitz_columns = True
