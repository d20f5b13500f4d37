mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_fft_|k_crop' -s 10 -c 5 -o gpurun_out/r1_prof_fft -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r1_prof_fft.log 2>&1
tail -2 gpurun_out/r1_prof_fft.log | cut -c1-300
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/q_pytest_all.log 2>&1; tail -5 gpurun_out/q_pytest_all.log
