"""Host-side cost of one call through the module API vs its GPU time (cfg2 and cfg1).
   python profiles/host_overhead.py"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torchkbnufft_b200 as tkbn
from torchkbnufft_b200 import workloads

dev = torch.device("cuda:0")
for name in ("cfg2", "cfg1"):
    wl = workloads.WORKLOADS[name]
    image, smaps, kdata, omega = workloads.make_inputs(wl, seed=0)
    x, s, y, om = (torch.from_numpy(a).to(dev) for a in (image, smaps, kdata, omega))
    nu = tkbn.KbNufft(im_size=wl.im_size, dtype=torch.complex64).to(dev)
    na = tkbn.KbNufftAdjoint(im_size=wl.im_size, dtype=torch.complex64).to(dev)
    kw = dict(smaps=s) if wl.n_coils > 1 else {}
    for _ in range(5):
        k = nu(x, om, **kw); na(k, om, **kw)
    torch.cuda.synchronize()
    n = 300
    # (a) host time to ENQUEUE n pairs while the GPU is kept busy by a long kernel (pure CPU cost)
    big = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
    for _ in range(20): big.fill_(1)
    t0 = time.perf_counter()
    for _ in range(n):
        k = nu(x, om, **kw); na(k, om, **kw)
    t_host = (time.perf_counter() - t0) / n
    torch.cuda.synchronize()
    # (b) back-to-back pairs, wall clock incl. GPU
    t0 = time.perf_counter()
    for _ in range(n):
        k = nu(x, om, **kw); na(k, om, **kw)
    torch.cuda.synchronize()
    t_wall = (time.perf_counter() - t0) / n
    # (c) GPU time of the same loop
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        k = nu(x, om, **kw); na(k, om, **kw)
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: host enqueue {t_host*1e6:7.1f} us/pair   wall {t_wall*1e6:7.1f} us/pair   gpu-timeline {e0.elapsed_time(e1)/n*1e3:7.1f} us/pair", flush=True)
    # (d) the same pair captured once in a CUDA graph and replayed: no per-call host work at all
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        k = nu(x, om, **kw); im = na(k, om, **kw)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        graph.replay()
    torch.cuda.synchronize()
    print(f"{name}: CUDA-graph replay wall {(time.perf_counter()-t0)/n*1e6:7.1f} us/pair", flush=True)
    with torch.no_grad():
        t0 = time.perf_counter()
        for _ in range(n):
            k = nu(x, om, **kw); na(k, om, **kw)
        torch.cuda.synchronize()
        print(f"{name}: no_grad wall {(time.perf_counter()-t0)/n*1e6:7.1f} us/pair", flush=True)
