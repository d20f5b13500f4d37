# round 2, call W: lane-per-cell spread for 1-4 coils; new FFT lengths (radix 9 / 15) on the GPU
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_robustness.py tests/test_gpu_autograd.py -m gpu -x -q -k "not cfg3 and not cfg4" > gpurun_out/r2w_pytest.log 2>&1
tail -5 gpurun_out/r2w_pytest.log
timeout 600 python profiles/scripts/adj_variants.py cfg1 cfg2 --variants=6 --caps=128 --owned=1 --coils=1,2,3,4 > gpurun_out/r2w_coils.log 2>&1
grep -v Warn gpurun_out/r2w_coils.log | tail -20
timeout 600 python bench.py --steps 50 --warmup 5 --workload cfg1 --no-cpu-baseline --no-reference-cuda > gpurun_out/r2w_bench_cfg1.log 2>&1
tail -1 gpurun_out/r2w_bench_cfg1.log | cut -c1-300
