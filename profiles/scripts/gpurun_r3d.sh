# round 2, call 3d (8 GPUs): bench.py under torchrun at N = 8 and N = 4 (default launch mode), as the driver runs it
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r3d_bench_8gpu.log 2>&1
tail -1 gpurun_out/r3d_bench_8gpu.log | cut -c1-200
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 20 --warmup 5 --no-reference-cuda > gpurun_out/r3d_bench_4gpu.log 2>&1
tail -1 gpurun_out/r3d_bench_4gpu.log | cut -c1-200
