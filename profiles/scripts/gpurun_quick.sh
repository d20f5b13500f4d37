mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ordered or cfg2" > gpurun_out/q_pytest.log 2>&1; tail -3 gpurun_out/q_pytest.log
timeout 900 python profiles/bench_configs.py cfg2 cfg5 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_adj_merge|k_adj_tiled' -c 6 --csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --adjoint-mode sorted 2>/dev/null | grep -E "k_adj" | cut -d'"' -f10,30 | cut -c1-100 | tail -4
