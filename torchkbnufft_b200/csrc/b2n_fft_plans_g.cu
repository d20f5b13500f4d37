// b2n_fft_plans_g.cu -- instantiates the compile-time planned FFT passes for the lengths with a radix-9 or radix-15
// stage: 72, 120, 144, 360, 600, 720, 1200, 1440 (see b2n_fft_fast_kernels.cuh).
#include "b2n_fft_fast_kernels.cuh"

namespace b2n {

B2N_DEFINE_PLAN(72)
B2N_DEFINE_PLAN(120)
B2N_DEFINE_PLAN(144)
B2N_DEFINE_PLAN(360)

}  // namespace b2n
