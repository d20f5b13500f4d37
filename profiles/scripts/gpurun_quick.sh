mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/q_pytest_all.log 2>&1; tail -3 gpurun_out/q_pytest_all.log
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['cuda_graph']['ms_per_step'], d['gpu_launches'], d['stages_ms'])"
