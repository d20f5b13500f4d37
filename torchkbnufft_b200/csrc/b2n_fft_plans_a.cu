// b2n_fft_plans_a.cu -- instantiates the compile-time planned FFT passes for lengths 64, 96, 128, 192, 224
// (see b2n_fft_fast_kernels.cuh; the plans are spread over several translation units so that they compile in parallel).
#include "b2n_fft_fast_kernels.cuh"

namespace b2n {

B2N_DEFINE_PLAN(64)
B2N_DEFINE_PLAN(96)
B2N_DEFINE_PLAN(128)
B2N_DEFINE_PLAN(192)
B2N_DEFINE_PLAN(224)

}  // namespace b2n
