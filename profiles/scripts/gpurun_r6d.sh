# EXPERIMENT (results invalid with 65): upper bound of what overlapping the forward column pass with the tail of the row pass could give
for v in 1 65; do
  timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-reference-cuda --no-partitions --launch eager --opt 9=$v 2>/dev/null | tail -1 | python -c "import sys, json; d = json.loads(sys.stdin.readline()); print('stream=$v', d['ms_per_step'], d['stages_ms'])"
done
