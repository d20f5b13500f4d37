timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused_pruned or fast_fft" 2>&1 | tail -2
timeout 600 python bench.py --steps 100 --warmup 5 --breakdown --no-cpu-baseline 2>&1 | grep -E "stage ms|step ms"
