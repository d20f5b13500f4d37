# round 2, call X: A/B of the FFT passes' CTA shapes (compile-time variants of the library), cfg2 and cfg3
mkdir -p gpurun_out
for v in default row80 row320 col2 col8; do
  if [ $v = default ]; then unset B2N_LIB_PATH; else export B2N_LIB_PATH=$PWD/torchkbnufft_b200/csrc/variants/libb200nufft_$v.so; fi
  for wl in cfg2 cfg3; do
    timeout 300 python bench.py --steps 30 --warmup 5 --workload $wl --no-cpu-baseline --no-reference-cuda --no-partitions > gpurun_out/r2x_$v_$wl.log 2>&1
    tail -1 gpurun_out/r2x_$v_$wl.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$v', '$wl', 'step %.1f us' % (d['ms_per_step']*1e3), {k: round(v*1e3,1) for k,v in d['stages_ms'].items()})"
  done
done 2>&1 | tee gpurun_out/r2x_fft_cfg_ab.log
