# round 2, call 4d (2 GPUs): NCCL vs peer-memory all-reduce latency
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 profiles/scripts/peer_ab.py > gpurun_out/r4d_peer_ab.log 2>&1
grep "^world" gpurun_out/r4d_peer_ab.log || tail -20 gpurun_out/r4d_peer_ab.log
