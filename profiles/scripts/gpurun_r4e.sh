# round 2, call 4e (2 GPUs): peer all-reduce (flag-free protocol): one-device protocol test, NCCL partitions test, latency A/B
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_robustness.py -m gpu -x -q -k "peer_allreduce" > gpurun_out/r4e_peer.log 2>&1
tail -3 gpurun_out/r4e_peer.log
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r4e_pytest.log 2>&1
tail -15 gpurun_out/r4e_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 profiles/scripts/peer_ab.py > gpurun_out/r4e_peer_ab.log 2>&1
grep "^world" gpurun_out/r4e_peer_ab.log || tail -20 gpurun_out/r4e_peer_ab.log
