mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/q_pytest_all.log 2>&1; tail -3 gpurun_out/q_pytest_all.log
