# round 2, call Y: visit lists written per work item -- tests, plan-build launch lists (cfg2, cfg4), e2e
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_robustness.py -m gpu -x -q -k "not cfg3" > gpurun_out/r2y_pytest.log 2>&1
tail -3 gpurun_out/r2y_pytest.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2y_plan_launches.csv python profiles/scripts/plan_build.py cfg2 cfg4 > gpurun_out/r2y_setup.log 2>&1
tail -2 gpurun_out/r2y_setup.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-partitions > gpurun_out/r2y_bench.log 2>&1
tail -1 gpurun_out/r2y_bench.log | cut -c1-200
