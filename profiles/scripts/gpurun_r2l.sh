# round 2, call L: real-weight owner-tile spread + exception fix-up; whole GPU suite except the two slowest configs
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "not cfg4 and not cfg3" > gpurun_out/r2l_pytest.log 2>&1
tail -15 gpurun_out/r2l_pytest.log
timeout 600 python profiles/scripts/adj_variants.py cfg2 cfg5 --caps=128,256 --owned=1 > gpurun_out/r2l_variants.log 2>&1
grep -v Warn gpurun_out/r2l_variants.log | tail -30
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2l_bench.log 2>&1
tail -3 gpurun_out/r2l_bench.log | cut -c1-1500
