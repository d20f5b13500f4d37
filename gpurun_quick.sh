mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ordered or cfg2 or cfg1 or layouts or cases" > gpurun_out/q_pytest.log 2>&1; tail -12 gpurun_out/q_pytest.log
timeout 900 python profiles/bench_configs.py cfg2 cfg1 cfg5 2>&1 | tail -4
