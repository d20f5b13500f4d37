# round 2, call C: ncu of the owner-tile spread at cap 128 + finer cap sweep
mkdir -p gpurun_out
timeout 600 python profiles/scripts/adj_variants.py cfg2 --variants= --caps=96,128,160,192 > gpurun_out/r2c_variants.log 2>&1
grep -v Warn gpurun_out/r2c_variants.log | tail -6
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_adj_own' -s 3 -c 1 -o gpurun_out/r2c_prof_own -f python profiles/scripts/adj_variants.py cfg2 --variants= --caps=128 > gpurun_out/r2c_prof.log 2>&1
tail -2 gpurun_out/r2c_prof.log | cut -c1-200
