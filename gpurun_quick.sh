mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tiled or cfg2 or edge or cases" > gpurun_out/q_pytest.log 2>&1; tail -3 gpurun_out/q_pytest.log
for o in "1=0" "1=4"; do timeout 600 python bench.py --steps 100 --warmup 3 --breakdown --no-cpu-baseline --opt $o 2>&1 | grep -E "stage ms|step ms"; done
