# round 2, call 4g (2 GPUs): partitions test incl. graph-mode adjoint with the all-reduce inside, bench at N = 2
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r4g_pytest.log 2>&1
tail -15 gpurun_out/r4g_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-reference-cuda > gpurun_out/r4g_bench_2gpu.log 2>&1
tail -1 gpurun_out/r4g_bench_2gpu.log | python -c "import sys, json; d = json.loads(sys.stdin.readline()); print(d['value'], d['ms_per_step']); print(json.dumps(d['partitions']['coil_sharded'], indent=1))"
