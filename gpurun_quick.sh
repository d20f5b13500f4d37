mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused_pruned" > gpurun_out/q_pytest.log 2>&1; tail -15 gpurun_out/q_pytest.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/q_pytest_all.log 2>&1; tail -5 gpurun_out/q_pytest_all.log
timeout 600 python bench.py --steps 20 --warmup 3 --breakdown --no-cpu-baseline 2>&1 | grep -E "stage ms|step ms"
timeout 600 python bench.py --steps 20 --warmup 3 --breakdown --no-cpu-baseline --cufft 2>&1 | grep -E "stage ms|step ms"
