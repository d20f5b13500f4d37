mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tiled or cfg2 or cfg1 or cases_match or edge" > gpurun_out/q_pytest.log 2>&1; tail -5 gpurun_out/q_pytest.log
for o in "--opt 2=0" "--opt 2=16"; do echo "== $o"; timeout 600 python bench.py --steps 50 --warmup 3 --breakdown --no-cpu-baseline $o 2>&1 | grep -E "stage ms|step ms"; done
