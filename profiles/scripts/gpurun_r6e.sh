# EXPERIMENT (results invalid for the variant): 2-D gather with no weight loads (what removing the per-point weight reads from shared memory would give at most)
V=$PWD/torchkbnufft_b200/csrc/variants/libb200nufft_freeweights.so
for lib in default $V; do
  if [ "$lib" = "default" ]; then unset B2N_LIB_PATH; else export B2N_LIB_PATH=$lib; fi
  timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-reference-cuda --no-partitions --launch eager 2>/dev/null | tail -1 | python -c "import sys, json; d = json.loads(sys.stdin.readline()); print('$lib', d['ms_per_step'], d['stages_ms'])"
  timeout 600 python profiles/bench_configs.py cfg5 2>&1 | grep "fwd " | cut -c1-120
done
