mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_spmat.py -m gpu -x -q > gpurun_out/q_pytest.log 2>&1; tail -12 gpurun_out/q_pytest.log
