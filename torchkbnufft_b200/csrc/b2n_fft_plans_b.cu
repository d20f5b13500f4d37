// b2n_fft_plans_b.cu -- instantiates the compile-time planned FFT passes for lengths 256, 288, 320, 384, 448
// (see b2n_fft_fast_kernels.cuh; the plans are spread over several translation units so that they compile in parallel).
#include "b2n_fft_fast_kernels.cuh"

namespace b2n {

B2N_DEFINE_PLAN(256)
B2N_DEFINE_PLAN(288)
B2N_DEFINE_PLAN(320)
B2N_DEFINE_PLAN(384)
B2N_DEFINE_PLAN(448)

}  // namespace b2n
