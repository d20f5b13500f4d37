set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r1_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r1_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r1_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/r1_smoke.log
timeout 600 python bench.py --breakdown > gpurun_out/r1_bench.log 2>&1; echo "bench exit $?" >> gpurun_out/r1_bench.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r1_bench_reference.log 2>&1
timeout 1200 python profiles/bench_configs.py > gpurun_out/r1_configs.log 2>&1; echo "configs exit $?" >> gpurun_out/r1_configs.log
tail -3 gpurun_out/r1_pytest_gpu.log; tail -2 gpurun_out/r1_smoke.log; tail -4 gpurun_out/r1_bench.log | cut -c1-2500; cat gpurun_out/r1_bench_reference.log | cut -c1-600; cat gpurun_out/r1_configs.log
