# round 2, call B: first run of the owner-tile spread: parity tests that touch the adjoint, then timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_robustness.py -m gpu -x -q -k "not cfg4 and not cfg3" > gpurun_out/r2b_pytest.log 2>&1
tail -15 gpurun_out/r2b_pytest.log
timeout 600 python profiles/scripts/adj_variants.py cfg2 cfg5 --caps=32,64,128,256 > gpurun_out/r2b_variants.log 2>&1
grep -v Warn gpurun_out/r2b_variants.log | tail -14
