# round 2, final validation of the last code (2 GPUs visible): full GPU suite, smoke, bench on one GPU, bench on two
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r9a_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r9a_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r9a_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/r9a_smoke.log
CUDA_VISIBLE_DEVICES=0 timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r9a_bench_cfg2.log 2>&1; echo "bench exit $?" >> gpurun_out/r9a_bench_cfg2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r9a_bench_2gpu.log 2>&1
tail -3 gpurun_out/r9a_pytest_gpu.log; tail -2 gpurun_out/r9a_smoke.log; tail -2 gpurun_out/r9a_bench_cfg2.log | cut -c1-250; tail -1 gpurun_out/r9a_bench_2gpu.log | cut -c1-250
