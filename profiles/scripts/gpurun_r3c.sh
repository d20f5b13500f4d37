# round 2, call 3c (2 GPUs): NCCL partition test + bench.py under torchrun with the default (graph) launch mode; single-GPU bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r3c_pytest.log 2>&1
tail -2 gpurun_out/r3c_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r3c_bench_1gpu.log 2>&1
tail -1 gpurun_out/r3c_bench_1gpu.log | cut -c1-200
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r3c_bench_2gpu.log 2>&1
tail -1 gpurun_out/r3c_bench_2gpu.log | cut -c1-200
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r3c_bench_reference.log 2>&1
tail -1 gpurun_out/r3c_bench_reference.log | cut -c1-300
