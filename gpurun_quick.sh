mkdir -p gpurun_out
for o in "1=0" "1=1" "1=2"; do echo "opt $o"; timeout 600 python bench.py --steps 100 --warmup 3 --breakdown --no-cpu-baseline --opt $o 2>&1 | grep -E "stage ms|step ms"; done
timeout 600 python bench.py --steps 100 --warmup 3 > gpurun_out/q_bench.log 2>&1; tail -1 gpurun_out/q_bench.log | cut -c1-2500
