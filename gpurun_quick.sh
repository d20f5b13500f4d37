mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused_pruned or fast_fft" > gpurun_out/q_pytest.log 2>&1; tail -12 gpurun_out/q_pytest.log
timeout 600 python bench.py --steps 50 --warmup 3 --breakdown --no-cpu-baseline --fft own 2>&1 | grep -E "stage ms|step ms"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_fft_|k_crop' -s 8 -c 10 --csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --fft own 2>/dev/null | grep -E "k_fft|k_crop" | cut -d'"' -f10,30 | head -12
