timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_autograd.py -m gpu -x -q -k "fused or cfg1 or cfg2 or toep or sense" 2>&1 | tail -2
python profiles/host_overhead.py 2>&1 | tail -4
