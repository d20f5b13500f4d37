# round 2, call M: full GPU suite, bench with graph capture, launch list of one e2e step (plan build included)
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2m_pytest.log 2>&1
tail -5 gpurun_out/r2m_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2m_bench.log 2>&1
tail -1 gpurun_out/r2m_bench.log | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2m_plan_launches.csv python profiles/scripts/plan_build.py cfg2 > gpurun_out/r2m_setup.log 2>&1
tail -5 gpurun_out/r2m_setup.log
