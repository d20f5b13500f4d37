# final launch list of the bench command (cold caches, serialised: shares only)
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r9b_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-reference-cuda --no-partitions > gpurun_out/r9b_ncu_bench.log 2>&1
grep -c "k_" gpurun_out/r9b_launches.csv
