# round 2, call T: 3-D owner-tile spread (k_adj_own_3d) -- 3-D parity tests, new FFT lengths, config 4 / 3 timings
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "d3 or 3d or cfg4 or three or fused_pruned or cases_match or ordered or tiled" > gpurun_out/r2t_pytest.log 2>&1
tail -8 gpurun_out/r2t_pytest.log
timeout 900 python profiles/bench_configs.py cfg4 cfg3 cfg1 > gpurun_out/r2t_configs.log 2>&1
grep -v Warn gpurun_out/r2t_configs.log | tail -8
