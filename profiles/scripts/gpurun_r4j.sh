# round 2, call 4j (2 GPUs): fused adjoint + all-reduce: one-device protocol tests, multi-process test, bench at N = 2 (fused vs separate kernel)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_robustness.py -m gpu -x -q -k "peer_allreduce or allreduce_inside" > gpurun_out/r4j_peer.log 2>&1
tail -8 gpurun_out/r4j_peer.log
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r4j_pytest.log 2>&1
tail -15 gpurun_out/r4j_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-reference-cuda > gpurun_out/r4j_bench_2gpu.log 2>&1
tail -1 gpurun_out/r4j_bench_2gpu.log | python -c "import sys, json; d = json.loads(sys.stdin.readline()); print(d['value'], d['ms_per_step']); print(json.dumps(d['partitions']['coil_sharded'], indent=1))"
