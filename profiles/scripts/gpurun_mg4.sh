mkdir -p gpurun_out
nvidia-smi -L | wc -l
N=$1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/r1_bench_${N}gpu.log 2>&1; echo "exit $?"; tail -1 gpurun_out/r1_bench_${N}gpu.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','n_gpus','ms_per_step')}, d['e2e']['value'], d['clocks'])"
