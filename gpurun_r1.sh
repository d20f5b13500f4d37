set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r1_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r1_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r1_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/r1_smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 --breakdown > gpurun_out/r1_bench.log 2>&1; echo "bench exit $?" >> gpurun_out/r1_bench.log
timeout 300 ./profiles/micro/micro_r1 > gpurun_out/r1_micro.log 2>&1; echo "micro exit $?" >> gpurun_out/r1_micro.log
tail -5 gpurun_out/r1_pytest_gpu.log; cat gpurun_out/r1_smoke.log | tail -3; tail -4 gpurun_out/r1_bench.log; cat gpurun_out/r1_micro.log
