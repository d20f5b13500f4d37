# round 2, call R: library-owned CUDA-graph replay (set_graph_mode) -- tests, bench cfg2 + cfg1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_robustness.py tests/test_gpu_parity.py -m gpu -x -q -k "not cfg4 and not cfg3" > gpurun_out/r2r_pytest.log 2>&1
tail -12 gpurun_out/r2r_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2r_bench.log 2>&1
tail -1 gpurun_out/r2r_bench.log | cut -c1-300
timeout 600 python bench.py --steps 50 --warmup 5 --workload cfg1 --no-cpu-baseline --no-reference-cuda > gpurun_out/r2r_bench_cfg1.log 2>&1
tail -1 gpurun_out/r2r_bench_cfg1.log | cut -c1-300
