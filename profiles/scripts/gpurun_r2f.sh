# round 2, call F: row-window owner-tile spread: parity, variants (unroll 2 / 8-coil warps / unroll 1), ncu of both
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_robustness.py -m gpu -x -q -k "not cfg4 and not cfg3" > gpurun_out/r2f_pytest.log 2>&1
tail -3 gpurun_out/r2f_pytest.log
timeout 600 python profiles/scripts/adj_variants.py cfg2 cfg5 --caps=96,128,192 --owned=1,3,5 > gpurun_out/r2f_variants.log 2>&1
grep -v Warn gpurun_out/r2f_variants.log | tail -24
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_adj_own' -s 3 -c 1 -o gpurun_out/r2f_prof_own -f python profiles/scripts/adj_variants.py cfg2 --variants= --caps=128 > gpurun_out/r2f_prof.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_adj_own' -s 3 -c 1 -o gpurun_out/r2f_prof_own_u1 -f python profiles/scripts/adj_variants.py cfg2 --variants= --caps=128 --owned=5 > gpurun_out/r2f_prof_u1.log 2>&1
tail -2 gpurun_out/r2f_prof_u1.log | cut -c1-200
