# round 2 evidence pass: launch list of the bench command + --set full captures of the kernels of the step (cfg2), the
# Toeplitz column pass (cfg3), the 3-D kernels (cfg4, 1/4 of the spokes) and the batched spread (cfg5)
mkdir -p gpurun_out
ARGS="--no-cpu-baseline --no-reference-cuda --no-partitions"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 4 --warmup 3 $ARGS > gpurun_out/r02_ncu_bench.log 2>&1
grep -c "k_" gpurun_out/r02_launches.csv
# 6 set-up + 3 warm-up steps x 7 kernels of the library come first
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_adj_own|k_own_pack|k_fwd_tiled_2d|k_fft_' -s 70 -c 7 -o gpurun_out/r02_prof_step -f python bench.py --steps 3 --warmup 3 $ARGS > gpurun_out/r02_prof_step.log 2>&1
tail -1 gpurun_out/r02_prof_step.log | cut -c1-160
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_fft_cols_toep' -s 2 -c 1 -o gpurun_out/r02_prof_toep -f python profiles/scripts/toep_ab.py > gpurun_out/r02_prof_toep.log 2>&1
tail -1 gpurun_out/r02_prof_toep.log | cut -c1-160
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_adj_own_2d' -s 2 -c 1 -o gpurun_out/r02_prof_cfg5 -f python profiles/scripts/adj_variants.py cfg5 --variants= --caps=0 > gpurun_out/r02_prof_cfg5.log 2>&1
tail -1 gpurun_out/r02_prof_cfg5.log | cut -c1-160
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_fwd_tiled_3d' -c 1 -o gpurun_out/r02_prof_fwd3d -f python profiles/scripts/own3_probe.py 0.25 fwd > gpurun_out/r02_prof_fwd3d.log 2>&1
tail -1 gpurun_out/r02_prof_fwd3d.log | cut -c1-160
