# round 2, call 4i (1 GPU): bench A/B of the streamed forward column pass (option 9) + ncu --set full of the streamed kernel
mkdir -p gpurun_out
for v in 0 1 3; do
  timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-reference-cuda --no-partitions --opt 9=$v > gpurun_out/r4i_bench_stream$v.log 2>&1
  tail -1 gpurun_out/r4i_bench_stream$v.log | python -c "import sys, json; d = json.loads(sys.stdin.readline()); print('stream=$v', d['ms_per_step'], d['other_launch_mode']['ms_per_step'], d['stages_ms'])"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_fft_cols_stream|k_fft_cols_fast' -s 12 -c 4 -o gpurun_out/r4i_prof_cols -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-reference-cuda --no-partitions --opt 9=3 > gpurun_out/r4i_prof.log 2>&1
tail -2 gpurun_out/r4i_prof.log | cut -c1-200
