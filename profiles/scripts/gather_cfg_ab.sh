#!/bin/sh
# A/B of the compile-time CTA shapes of the tiled forward gathers: builds libb200nufft variants next to the default one
# (only b2n_interp_tiled.cu / b2n_interp_tiled3d.cu are rebuilt).   sh profiles/scripts/gather_cfg_ab.sh
set -e
cd "$(dirname "$0")/../../torchkbnufft_b200/csrc"
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-ffp-contract=off,-fvisibility=hidden --expt-relaxed-constexpr"
mkdir -p variants
for v in "f2w12b2:-DB2N_FWD2_WARPS=12 -DB2N_FWD2_MINB=2" "f2w16b2:-DB2N_FWD2_WARPS=16 -DB2N_FWD2_MINB=2" "f2w10b3:-DB2N_FWD2_WARPS=10 -DB2N_FWD2_MINB=3" "f2w12b3:-DB2N_FWD2_WARPS=12 -DB2N_FWD2_MINB=3" "f3w12:-DB2N_FWD3_WARPS=12" "f3w16:-DB2N_FWD3_WARPS=16"; do
  name=${v%%:*}; def=${v#*:}
  nvcc $FLAGS $def -Xptxas=-v -c b2n_interp_tiled.cu -o variants/b2n_interp_tiled_$name.o 2> variants/ptxas_tiled_$name.log &
  nvcc $FLAGS $def -Xptxas=-v -c b2n_interp_tiled3d.cu -o variants/b2n_interp_tiled3d_$name.o 2> variants/ptxas_tiled3d_$name.log &
  wait
  others=$(ls build/*.o | grep -v "b2n_interp_tiled.o\|b2n_interp_tiled3d.o")
  nvcc -shared -o variants/libb200nufft_$name.so variants/b2n_interp_tiled_$name.o variants/b2n_interp_tiled3d_$name.o $others -gencode arch=compute_100a,code=sm_100a -lcudart
  echo built variants/libb200nufft_$name.so
done
