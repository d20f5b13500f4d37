# round 2, call K: real-weight owner-tile spread (64-byte visit records, phase factored into samples / cells / sign)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_robustness.py -m gpu -x -q -k "not cfg4 and not cfg3" > gpurun_out/r2k_pytest.log 2>&1
tail -15 gpurun_out/r2k_pytest.log
timeout 600 python profiles/scripts/adj_variants.py cfg2 cfg5 --caps=64,128,256 --owned=1,4,5,6 > gpurun_out/r2k_variants.log 2>&1
grep -v Warn gpurun_out/r2k_variants.log | tail -30
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_adj_own|k_own_pack' -s 6 -c 2 -o gpurun_out/r2k_prof_own -f python profiles/scripts/adj_variants.py cfg2 --variants= --caps=128 > gpurun_out/r2k_prof.log 2>&1
tail -2 gpurun_out/r2k_prof.log | cut -c1-200
