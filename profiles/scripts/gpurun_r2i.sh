# round 2, call I: four-row owner tiles (one branch-free loop) vs eight-row tiles; FFMA2 / LDS micro-benchmarks
mkdir -p gpurun_out
./profiles/micro/micro_r2 > gpurun_out/r2i_micro.log 2>&1
cat gpurun_out/r2i_micro.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_robustness.py -m gpu -x -q -k "not cfg4 and not cfg3" > gpurun_out/r2i_pytest.log 2>&1
tail -3 gpurun_out/r2i_pytest.log
timeout 600 python profiles/scripts/adj_variants.py cfg2 cfg5 --caps=64,128 --rows=4 --owned=1,4,5,6,7 > gpurun_out/r2i_variants.log 2>&1
timeout 600 python profiles/scripts/adj_variants.py cfg2 cfg5 --variants= --caps=128 --rows=8 --owned=5 >> gpurun_out/r2i_variants.log 2>&1
grep -v Warn gpurun_out/r2i_variants.log | tail -30
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_adj_own4' -s 3 -c 1 -o gpurun_out/r2i_prof_own4 -f python profiles/scripts/adj_variants.py cfg2 --variants= --caps=128 --rows=4 > gpurun_out/r2i_prof.log 2>&1
tail -2 gpurun_out/r2i_prof.log | cut -c1-200
