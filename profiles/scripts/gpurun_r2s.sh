# round 2, call S: steady-state host cost per pair, cProfile of the host path, single-coil spread variants
mkdir -p gpurun_out
timeout 600 python profiles/host_overhead.py > gpurun_out/r2s_host_overhead.log 2>&1
tail -8 gpurun_out/r2s_host_overhead.log
timeout 600 python profiles/host_profile.py cfg2 > gpurun_out/r2s_host_profile.log 2>&1
head -45 gpurun_out/r2s_host_profile.log | cut -c1-170
timeout 600 python profiles/scripts/adj_variants.py cfg1 cfg2 --variants=0 --caps=128 --owned=1 --coils=1,2,4,8 > gpurun_out/r2s_coils.log 2>&1
grep -v Warn gpurun_out/r2s_coils.log | tail -20
