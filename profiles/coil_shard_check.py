"""2-GPU check of coil sharding with the NCCL all-reduce (torchrun --nproc-per-node 2):
the coil-combined adjoint image from sharded coils must equal the unsharded one."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchkbnufft_b200 as tkbn
from torchkbnufft_b200 import parallel, workloads

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
wl = workloads.WORKLOADS["cfg2"]
image, smaps, kdata, omega = workloads.make_inputs(wl, seed=0)
nu = tkbn.KbNufft(im_size=wl.im_size, dtype=torch.complex64).to(dev)
na = tkbn.KbNufftAdjoint(im_size=wl.im_size, dtype=torch.complex64).to(dev)
x, s, y, om = (torch.from_numpy(a).to(dev) for a in (image, smaps, kdata, omega))
full = na(y, om, smaps=s)
part = parallel.coil_sharded_adjoint(na, parallel.local_coils(y).contiguous(), om, parallel.local_coils(s).contiguous())
err = float(torch.norm(part - full) / torch.norm(full))
fwd_local = parallel.coil_sharded_forward(nu, x, om, parallel.local_coils(s).contiguous())
lo, hi = parallel.shard_bounds(wl.n_coils, rank, world)
err_f = float(torch.norm(fwd_local - nu(x, om, smaps=s)[:, lo:hi]) / torch.norm(fwd_local))
# timing of the sharded adjoint incl. the all-reduce
for _ in range(5):
    parallel.coil_sharded_adjoint(na, parallel.local_coils(y).contiguous(), om, parallel.local_coils(s).contiguous())
torch.cuda.synchronize(); dist.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
yl, sl = parallel.local_coils(y).contiguous(), parallel.local_coils(s).contiguous()
a.record()
for _ in range(20):
    parallel.coil_sharded_adjoint(na, yl, om, sl)
b.record(); torch.cuda.synchronize()
print(f"rank {rank}/{world}: coil-sharded adjoint rel err {err:.2e}, forward shard rel err {err_f:.2e}, "
      f"sharded adjoint + all-reduce {a.elapsed_time(b) / 20 * 1e3:.1f} us per call")
dist.barrier(); dist.destroy_process_group()
