"""Times the five BASELINE.json configs on one B200 (device-resident inputs, CUDA events, L2 flushed
between repetitions) and checks each against the oracle where that finishes in seconds.
  python profiles/bench_configs.py [cfg1 cfg2 ...]
"""
import os, statistics, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torchkbnufft_b200 as tkbn
from torchkbnufft_b200 import workloads

dev = torch.device("cuda:0")
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

def timeit(fn, reps=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for r in range(reps):
        flush.fill_(r & 255)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)

def D(a): return torch.from_numpy(np.ascontiguousarray(a)).to(dev)

which = sys.argv[1:] or ["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"]
for name in which:
    wl = workloads.WORKLOADS[name]
    B = wl.n_batch if name != "cfg5" else 8          # cfg5: the 8 slices one GPU owns under 8-way batch sharding
    t0 = time.time()
    image, smaps, kdata, omega = workloads.make_inputs(wl, seed=0, n_batch=B)
    nu = tkbn.KbNufft(im_size=wl.im_size, dtype=torch.complex64).to(dev)
    na = tkbn.KbNufftAdjoint(im_size=wl.im_size, dtype=torch.complex64).to(dev)
    x, s, y, om = D(image), D(smaps), D(kdata), D(omega)
    units = B * wl.n_coils * wl.n_points
    t_plan0 = time.time(); k = nu(x, om, smaps=s); torch.cuda.synchronize(); t_first = time.time() - t_plan0
    reps = 5 if name == "cfg4" else 10
    tf = timeit(lambda: nu(x, om, smaps=s), reps)
    ta = timeit(lambda: na(y, om, smaps=s), reps)
    line = f"{name}: B={B} C={wl.n_coils} M={wl.n_points} N={wl.im_size}  fwd {tf*1e3:9.1f} us  adj {ta*1e3:9.1f} us  pair {2*units/((tf+ta)*1e-3)/1e9:7.2f} G coil-pts/s  (first call incl. plan {t_first*1e3:.1f} ms)"
    tkbn.set_adjoint_mode("sorted")
    try:
        ts = timeit(lambda: na(y, om, smaps=s), 3, 1)
        line += f"  adj-sorted {ts*1e3:9.1f} us"
    finally:
        tkbn.set_adjoint_mode("atomic")
    print(line, flush=True)
    # roofline fractions of each direction: HBM (algorithmic bytes of SURVEY 8(d) over the measured copy bandwidth) and
    # FP32 (8 flops per complex multiply-add x J^d neighbours x coil-points over 148 SMs x 128 lanes x 2 x 1.965 GHz)
    d = len(wl.im_size)
    N, K = int(np.prod(wl.im_size)), int(np.prod(wl.grid_size))
    C, M = wl.n_coils, wl.n_points
    fwd_b = 8 * (B * N + C * N + N + 4 * B * C * K + B * C * M) + 4 * d * M
    adj_b = 8 * (B * C * M + 3 * B * C * K + B * C * N + C * N + B * N + N) + 4 * d * M
    flops = 8.0 * 6 ** d * units
    hbm, fp32 = 6559.4e9, 148 * 128 * 2 * 1.965e9
    print(f"{name}: roofline  fwd HBM {fwd_b / (tf * 1e-3) / hbm:.3f} FP32 {flops / (tf * 1e-3) / fp32:.3f}   "
          f"adj HBM {adj_b / (ta * 1e-3) / hbm:.3f} FP32 {flops / (ta * 1e-3) / fp32:.3f}   "
          f"(interpolation flops {flops / 1e9:.1f} G per direction)", flush=True)
    if name == "cfg3":
        t_k = time.time(); kern = tkbn.calc_toeplitz_kernel(om, wl.im_size, norm="ortho"); torch.cuda.synchronize(); t_k = time.time() - t_k
        toep = tkbn.ToepNufft()
        tt = timeit(lambda: toep(x, kern, smaps=s, norm="ortho"), reps)
        fbn = na(nu(x, om, smaps=s, norm="ortho"), om, smaps=s, norm="ortho"); fbt = toep(x, kern, smaps=s, norm="ortho")
        print(f"cfg3: ToepNufft apply {tt*1e3:9.1f} us; calc_toeplitz_kernel {t_k*1e3:.1f} ms (first call); |A^H A x - T x|/|A^H A x| = {float(torch.norm(fbn-fbt)/torch.norm(fbn)):.2e}", flush=True)
    if name == "cfg4":
        xg = x.clone().requires_grad_(True)
        def fb():
            xg.grad = None
            out = nu(xg, om, smaps=s); (out.abs() ** 2 / 2).sum().backward()
        tb = timeit(fb, 3, 1)
        print(f"cfg4: forward + autograd backward of 0.5*||Ax||^2: {tb*1e3:9.1f} us", flush=True)
    if name == "cfg1":
        t_d = time.time(); dcf = tkbn.calc_density_compensation_function(om, wl.im_size); torch.cuda.synchronize(); t_d = time.time() - t_d
        print(f"cfg1: calc_density_compensation_function (10 iterations) {t_d*1e3:.1f} ms (first call)", flush=True)
    del x, s, y, om, nu, na
    tkbn.clear_caches(); torch.cuda.empty_cache()
