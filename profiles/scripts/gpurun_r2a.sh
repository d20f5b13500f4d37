# round 2, call A: calibrate the adjoint-spread variants (times + ncu of rows-ownership vs warp-private tiles)
mkdir -p gpurun_out
timeout 600 python profiles/scripts/adj_variants.py cfg2 cfg5 > gpurun_out/r2a_variants.log 2>&1
tail -12 gpurun_out/r2a_variants.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_adj_' -s 3 -c 1 -o gpurun_out/r2a_prof_v3 -f python profiles/scripts/adj_variants.py cfg2 --variants=3 > gpurun_out/r2a_prof_v3.log 2>&1
tail -2 gpurun_out/r2a_prof_v3.log | cut -c1-200
