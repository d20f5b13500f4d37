mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_robustness.py tests/test_gpu_multi.py -m gpu -x -q -k "peer_allreduce or allreduce_inside or partitions" > gpurun_out/r7b_pytest.log 2>&1
tail -3 gpurun_out/r7b_pytest.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused_pruned or toeplitz or cases_match" > gpurun_out/r7b_pytest2.log 2>&1
tail -2 gpurun_out/r7b_pytest2.log
