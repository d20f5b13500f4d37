# round 2, 2-GPU call: NCCL partition test + bench.py under torchrun (cfg2 replicas, cfg5 strong scaling, coil sharding)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_distributed.py -m gpu -x -q > gpurun_out/r2mg_pytest.log 2>&1
tail -4 gpurun_out/r2mg_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2mg_bench_2gpu.log 2>&1
tail -1 gpurun_out/r2mg_bench_2gpu.log | cut -c1-300
