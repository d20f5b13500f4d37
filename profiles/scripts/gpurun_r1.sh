set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r1_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r1_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r1_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/r1_smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 --breakdown > gpurun_out/r1_bench.log 2>&1; echo "bench exit $?" >> gpurun_out/r1_bench.log
timeout 600 python bench.py --steps 20 --warmup 3 --breakdown --workload cfg1 --no-cpu-baseline > gpurun_out/r1_bench_cfg1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r1_ncu_bench.log 2>&1
tail -15 gpurun_out/r1_pytest_gpu.log; tail -3 gpurun_out/r1_smoke.log; tail -4 gpurun_out/r1_bench.log; tail -3 gpurun_out/r1_bench_cfg1.log; tail -70 gpurun_out/r1_launches.csv | cut -c1-200
