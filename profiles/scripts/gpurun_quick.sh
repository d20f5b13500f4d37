mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tiled or cfg2 or edge or cases or ordered" > gpurun_out/q_pytest.log 2>&1; tail -3 gpurun_out/q_pytest.log
timeout 600 python bench.py --steps 100 --warmup 3 --breakdown --no-cpu-baseline 2>&1 | grep -E "stage ms|step ms"
