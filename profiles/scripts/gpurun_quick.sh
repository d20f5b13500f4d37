timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused_pruned or fast_fft or cfg4" 2>&1 | tail -2
timeout 900 python profiles/bench_configs.py cfg4 cfg2 2>&1 | tail -3
