# 3-D fix-up kernel with 4 tiles per CTA: parity tests with exception points, cfg4 timing, launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cases_match_reference or ordered_tiled or cfg4 or tiled_3d" > gpurun_out/r5c_pytest.log 2>&1
tail -2 gpurun_out/r5c_pytest.log
timeout 900 python profiles/bench_configs.py cfg4 > gpurun_out/r5c_cfg4.log 2>&1
cat gpurun_out/r5c_cfg4.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r5c_cfg4_launches.csv python profiles/run_cfg.py cfg4 2 > gpurun_out/r5c_cfg4_ncu.log 2>&1
grep "k_own_fix" gpurun_out/r5c_cfg4_launches.csv | grep duration | cut -d, -f 5,15- | head -3
