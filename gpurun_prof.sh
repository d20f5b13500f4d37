mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_fft_' -s 8 -c 4 -o gpurun_out/r1_prof_fft -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --fused-fft > gpurun_out/r1_prof_fft.log 2>&1
tail -2 gpurun_out/r1_prof_fft.log | cut -c1-300
