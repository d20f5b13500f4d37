mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_spmat.py -m gpu -x -q -k "not cfg4_full and not cfg3_full and not cfg5 and not cfg2_full and not cfg1_full" > gpurun_out/r1_sanitizer_memcheck.log 2>&1; echo "memcheck exit $?"; tail -4 gpurun_out/r1_sanitizer_memcheck.log
