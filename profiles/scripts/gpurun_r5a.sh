# round 2, final evidence (1 GPU): full GPU suite, smoke, bench (both arms), all configs
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r5a_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r5a_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r5a_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/r5a_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r5a_bench_cfg2.log 2>&1; echo "bench exit $?" >> gpurun_out/r5a_bench_cfg2.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r5a_bench_reference.log 2>&1
timeout 1200 python profiles/bench_configs.py > gpurun_out/r5a_configs.log 2>&1; echo "configs exit $?" >> gpurun_out/r5a_configs.log
tail -3 gpurun_out/r5a_pytest_gpu.log; tail -2 gpurun_out/r5a_smoke.log; tail -2 gpurun_out/r5a_bench_cfg2.log | cut -c1-1500; tail -1 gpurun_out/r5a_bench_reference.log | cut -c1-600; cat gpurun_out/r5a_configs.log
