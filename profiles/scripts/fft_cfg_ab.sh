#!/bin/sh
# A/B of the compile-time CTA shapes of the planned FFT passes: builds libb200nufft variants next to the default one
# (only the translation units that depend on the setting are rebuilt: plans_c holds length 640, plans_b 320).
#   sh profiles/scripts/fft_cfg_ab.sh          (on the build machine; the variants travel with the gpurun snapshot)
set -e
cd "$(dirname "$0")/../../torchkbnufft_b200/csrc"
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-ffp-contract=off,-fvisibility=hidden --expt-relaxed-constexpr"
mkdir -p variants
for v in "row80:-DB2N_FFT_ROW_TARGET=80" "row320:-DB2N_FFT_ROW_TARGET=320" "col2:-DB2N_FFT_COL_PAIRS=2" "col8:-DB2N_FFT_COL_PAIRS=8"; do
  name=${v%%:*}; def=${v#*:}
  objs=""
  for f in b2n_fft_plans_a b2n_fft_plans_b b2n_fft_plans_c b2n_fft_plans_d b2n_fft_plans_e b2n_fft_plans_f b2n_fft_plans_g b2n_fft_plans_h b2n_fft; do
    nvcc $FLAGS $def -c $f.cu -o variants/${f}_$name.o &
    objs="$objs variants/${f}_$name.o"
  done
  wait
  others=$(ls build/*.o | grep -v "b2n_fft_plans_\|b2n_fft.o")
  nvcc -shared -o variants/libb200nufft_$name.so $objs $others -gencode arch=compute_100a,code=sm_100a -lcudart
  echo built variants/libb200nufft_$name.so
done
