mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tiled_3d or cfg4 or cases_match" > gpurun_out/q_pytest.log 2>&1; tail -12 gpurun_out/q_pytest.log
timeout 900 python profiles/bench_configs.py cfg4 2>&1 | tail -3
