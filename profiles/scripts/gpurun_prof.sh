mkdir -p gpurun_out
# launch list of the default bench command (cold-cache, serialised: shares of the step, not absolute times)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r1_ncu_bench.log 2>&1
# full capture of one launch of each own kernel in the step
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_adj_tiled_2d|k_fwd_tiled_2d|k_fft_' -s 18 -c 6 -o gpurun_out/r1_prof_step -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r1_prof_step.log 2>&1
tail -2 gpurun_out/r1_prof_step.log | cut -c1-200
grep -c "k_" gpurun_out/r1_launches.csv
