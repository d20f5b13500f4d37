# round 2, call 3b: strengthened parity tests (fallback kernels with the owner path off, partial owner tiles, modified tables)
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_gpu_parity.py tests/test_gpu_robustness.py -m gpu -x -q -k "not cfg3 and not cfg4" > gpurun_out/r3b_pytest.log 2>&1
tail -6 gpurun_out/r3b_pytest.log
