# round 2, call J: four-row owner tiles with pre-packed samples (k_own_pack) vs direct staging
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_robustness.py -m gpu -x -q -k "not cfg4 and not cfg3" > gpurun_out/r2j_pytest.log 2>&1
tail -3 gpurun_out/r2j_pytest.log
timeout 600 python profiles/scripts/adj_variants.py cfg2 cfg5 --caps=128 --rows=4 --owned=1,4,5,6,7 > gpurun_out/r2j_variants.log 2>&1
grep -v Warn gpurun_out/r2j_variants.log | tail -30
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_adj_own4|k_own_pack' -s 6 -c 2 -o gpurun_out/r2j_prof_own4 -f python profiles/scripts/adj_variants.py cfg2 --variants= --caps=128 --rows=4 > gpurun_out/r2j_prof.log 2>&1
tail -2 gpurun_out/r2j_prof.log | cut -c1-200
