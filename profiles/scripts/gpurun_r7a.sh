# round 2, final validation (2 GPUs visible: the multi-GPU test runs inside the suite): full GPU suite, smoke, bench 1 GPU
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r7a_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r7a_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r7a_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/r7a_smoke.log
CUDA_VISIBLE_DEVICES=0 timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r7a_bench_cfg2.log 2>&1; echo "bench exit $?" >> gpurun_out/r7a_bench_cfg2.log
CUDA_VISIBLE_DEVICES=0 timeout 900 python profiles/bench_configs.py cfg4 > gpurun_out/r7a_cfg4.log 2>&1
tail -3 gpurun_out/r7a_pytest_gpu.log; tail -2 gpurun_out/r7a_smoke.log; tail -2 gpurun_out/r7a_bench_cfg2.log | cut -c1-300; cat gpurun_out/r7a_cfg4.log
