# round 2, call 6a: CTA shapes of the tiled forward gathers (variants built by profiles/scripts/gather_cfg_ab.sh)
mkdir -p gpurun_out
V=torchkbnufft_b200/csrc/variants
: > gpurun_out/r6a_gather_cfg_ab.log
for name in default f2w16b2 f2w12b3; do
  if [ "$name" = "default" ]; then unset B2N_LIB_PATH; else export B2N_LIB_PATH=$PWD/$V/libb200nufft_$name.so; fi
  timeout 600 python profiles/bench_configs.py cfg2 cfg5 cfg3 2>&1 | grep "fwd " | sed "s/^/$name /" | cut -c1-150 >> gpurun_out/r6a_gather_cfg_ab.log
done
for name in default f3w16; do
  if [ "$name" = "default" ]; then unset B2N_LIB_PATH; else export B2N_LIB_PATH=$PWD/$V/libb200nufft_$name.so; fi
  timeout 600 python profiles/bench_configs.py cfg4 2>&1 | grep "fwd " | sed "s/^/$name /" | cut -c1-150 >> gpurun_out/r6a_gather_cfg_ab.log
done
cat gpurun_out/r6a_gather_cfg_ab.log
