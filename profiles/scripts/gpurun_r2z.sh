# round 2, call Z: full GPU suite (incl. cfg3 / cfg4 / upstream suite), all five configs, smoke
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2z_pytest.log 2>&1
tail -4 gpurun_out/r2z_pytest.log
timeout 1200 python profiles/bench_configs.py > gpurun_out/r2z_configs.log 2>&1
grep -v Warn gpurun_out/r2z_configs.log | tail -14
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2z_smoke.log 2>&1
tail -2 gpurun_out/r2z_smoke.log
