# round 2, call V: tile-parallel exception fix-up; 3-D spread timings and one full ncu capture
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_robustness.py -m gpu -x -q -k "not cfg3" > gpurun_out/r2v_pytest.log 2>&1
tail -5 gpurun_out/r2v_pytest.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2v_launches.csv python profiles/scripts/own3_probe.py > gpurun_out/r2v_probe.log 2>&1
grep -v Warn gpurun_out/r2v_probe.log | tail -3
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_adj_own_3d' -s 2 -c 1 -o gpurun_out/r2v_prof_own3 -f python profiles/scripts/own3_probe.py 0.25 > gpurun_out/r2v_prof.log 2>&1
tail -2 gpurun_out/r2v_prof.log | cut -c1-200
timeout 900 python profiles/bench_configs.py cfg4 cfg2 > gpurun_out/r2v_configs.log 2>&1
grep -v Warn gpurun_out/r2v_configs.log | tail -4
