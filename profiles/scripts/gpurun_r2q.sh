# round 2, call Q: faster plan build (cooperative visit-record writes, parallel exception list, one-pass tile scan)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_robustness.py -m gpu -x -q -k "not cfg4 and not cfg3" > gpurun_out/r2q_pytest.log 2>&1
tail -3 gpurun_out/r2q_pytest.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2q_plan_launches.csv python profiles/scripts/plan_build.py cfg2 > gpurun_out/r2q_setup.log 2>&1
tail -2 gpurun_out/r2q_setup.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2q_bench.log 2>&1
tail -1 gpurun_out/r2q_bench.log | cut -c1-300
timeout 600 python profiles/scripts/e2e_host.py cfg2 > gpurun_out/r2q_e2e_host.log 2>&1
head -12 gpurun_out/r2q_e2e_host.log | cut -c1-160
