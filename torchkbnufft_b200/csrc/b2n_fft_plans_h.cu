// b2n_fft_plans_h.cu -- instantiates the compile-time planned FFT passes for 600, 720, 1200, 1440
// (radix-9 / radix-15 stages; see b2n_fft_fast_kernels.cuh).
#include "b2n_fft_fast_kernels.cuh"

namespace b2n {

B2N_DEFINE_PLAN(600)
B2N_DEFINE_PLAN(720)
B2N_DEFINE_PLAN(1200)
B2N_DEFINE_PLAN(1440)

}  // namespace b2n
