# round 2, call H: seven window classes + cheaper staging
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_robustness.py -m gpu -x -q -k "not cfg4 and not cfg3" > gpurun_out/r2h_pytest.log 2>&1
tail -3 gpurun_out/r2h_pytest.log
timeout 600 python profiles/scripts/adj_variants.py cfg2 cfg5 --caps=128,192 --owned=1,4,5 > gpurun_out/r2h_variants.log 2>&1
grep -v Warn gpurun_out/r2h_variants.log | tail -24
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_adj_own' -s 3 -c 1 -o gpurun_out/r2h_prof_own -f python profiles/scripts/adj_variants.py cfg2 --variants= --caps=128 > gpurun_out/r2h_prof.log 2>&1
tail -2 gpurun_out/r2h_prof.log | cut -c1-200
