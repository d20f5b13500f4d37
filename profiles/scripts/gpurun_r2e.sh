# round 2, call E: run-based owner-tile spread: parity, variants (default / 8-coil warps / no unroll), ncu of default
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_robustness.py -m gpu -x -q -k "not cfg4 and not cfg3" > gpurun_out/r2e_pytest.log 2>&1
tail -3 gpurun_out/r2e_pytest.log
timeout 600 python profiles/scripts/adj_variants.py cfg2 cfg5 --caps=96,128,192 --owned=1,3,5 > gpurun_out/r2e_variants.log 2>&1
grep -v Warn gpurun_out/r2e_variants.log | tail -24
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_adj_own' -s 3 -c 1 -o gpurun_out/r2e_prof_own -f python profiles/scripts/adj_variants.py cfg2 --variants= --caps=128 > gpurun_out/r2e_prof.log 2>&1
tail -2 gpurun_out/r2e_prof.log | cut -c1-200
