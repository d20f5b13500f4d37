# round 2: two-shot form of the peer all-reduce: one-device protocol tests (both forms), multi-process test, latency A/B at 2 GPUs
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_robustness.py -m gpu -x -q -k "peer_allreduce or allreduce_inside" > gpurun_out/r8a_peer.log 2>&1
tail -3 gpurun_out/r8a_peer.log
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r8a_pytest.log 2>&1
tail -12 gpurun_out/r8a_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 profiles/scripts/peer_ab.py > gpurun_out/r8a_peer_ab_2gpu.log 2>&1
grep "^world" gpurun_out/r8a_peer_ab_2gpu.log | cut -c1-260 || tail -20 gpurun_out/r8a_peer_ab_2gpu.log
