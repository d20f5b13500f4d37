mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/q_pytest_all.log 2>&1; tail -5 gpurun_out/q_pytest_all.log
timeout 600 python bench.py --steps 100 --warmup 3 --breakdown --no-cpu-baseline 2>&1 | grep -E "stage ms|step ms"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fast_fft or (tiled_3d and grid_size0 and 3-False) or (fused_pruned and (N12 or N21))" > gpurun_out/r1_sanitizer_racecheck.log 2>&1; echo "racecheck exit $?"; tail -4 gpurun_out/r1_sanitizer_racecheck.log
