import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from torchkbnufft_b200._nufft import fft as eng_fft
dev = torch.device("cuda:0")
flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)
N, K = (320, 320), (640, 640)
def timed(fn, n=200):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    tot = 0.0
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(n):
        flush.fill_(i & 0xFF)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return 1e3 * tot / n
scal = torch.randn(N, dtype=torch.complex64, device=dev)
for C in (4, 8, 16, 32):
    grid = torch.randn((1, C) + K, dtype=torch.complex64, device=dev)
    smaps = torch.randn((1, C) + N, dtype=torch.complex64, device=dev)
    print(f"C={C:2d}: fused adjoint FFT (columns + rows/coil sum) {timed(lambda: eng_fft.fused_fft_adjoint(grid, N, smaps, scal, 1.0)):6.1f} us", flush=True)
