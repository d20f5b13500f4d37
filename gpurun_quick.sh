mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/q_pytest_all.log 2>&1; tail -3 gpurun_out/q_pytest_all.log
timeout 600 python bench.py --steps 100 --warmup 3 --breakdown --no-cpu-baseline --fft cufft 2>&1 | grep -E "stage ms|step ms"
timeout 900 python profiles/bench_configs.py cfg1 cfg3 cfg5 2>&1 | tail -6
