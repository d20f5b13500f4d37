# round 2, call 4f (8 GPUs): peer all-reduce latency A/B and the partitions of bench.py at N = 8
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 profiles/scripts/peer_ab.py > gpurun_out/r4f_peer_ab_8gpu.log 2>&1
grep "^world" gpurun_out/r4f_peer_ab_8gpu.log || tail -20 gpurun_out/r4f_peer_ab_8gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 --no-reference-cuda > gpurun_out/r4f_bench_8gpu.log 2>&1
tail -1 gpurun_out/r4f_bench_8gpu.log | python -c "import sys, json; d = json.loads(sys.stdin.readline()); print(d['value'], d['ms_per_step']); print(json.dumps(d['partitions'], indent=1))"
