import torch, time
dev = torch.device("cuda:0")
sizes = [819200, 13107200, 16384000, 1024000]
host = [torch.empty(s, dtype=torch.uint8).pin_memory() for s in sizes]
devb = [torch.empty(s, dtype=torch.uint8, device=dev) for s in sizes]
tot = sum(sizes)
hostall = torch.empty(tot, dtype=torch.uint8).pin_memory()
devall = torch.empty(tot, dtype=torch.uint8, device=dev)
d2h_src = torch.empty(17203200, dtype=torch.uint8, device=dev)
d2h_dst = torch.empty(17203200, dtype=torch.uint8).pin_memory()
s = torch.cuda.Stream(); s2 = torch.cuda.Stream()
def timeit(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s):
        a.record(s)
        for _ in range(n): fn()
        b.record(s)
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n
def four():
    with torch.cuda.stream(s):
        for h, d in zip(host, devb): d.copy_(h, non_blocking=True)
def one():
    with torch.cuda.stream(s):
        devall.copy_(hostall, non_blocking=True)
def one_duplex():
    with torch.cuda.stream(s):
        devall.copy_(hostall, non_blocking=True)
    with torch.cuda.stream(s2):
        d2h_dst.copy_(d2h_src, non_blocking=True)
for name, fn in (("4 copies", four), ("1 copy", one), ("1 copy + concurrent D2H 17 MB", one_duplex)):
    ms = timeit(fn)
    print(f"{name}: {ms:.3f} ms per {tot/1e6:.1f} MB -> {tot/ms/1e6:.1f} GB/s")
