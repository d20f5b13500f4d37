mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/q_pytest_all.log 2>&1; tail -16 gpurun_out/q_pytest_all.log
timeout 600 python profiles/bench_configs.py cfg1 2>&1 | tail -2
