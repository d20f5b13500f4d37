mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ordered or cfg4 or tiled_3d" > gpurun_out/q_pytest.log 2>&1; tail -5 gpurun_out/q_pytest.log
timeout 900 python profiles/bench_configs.py cfg4 2>&1 | tail -2
