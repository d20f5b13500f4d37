// b2n_fft_plans_d.cu -- instantiates the compile-time planned FFT passes for lengths 896, 960, 1024, 1280, 2048
// (see b2n_fft_fast_kernels.cuh; the plans are spread over several translation units so that they compile in parallel).
#include "b2n_fft_fast_kernels.cuh"

namespace b2n {

B2N_DEFINE_PLAN(896)
B2N_DEFINE_PLAN(960)
B2N_DEFINE_PLAN(1024)
B2N_DEFINE_PLAN(1280)
B2N_DEFINE_PLAN(2048)

}  // namespace b2n
