// b2n_fft_plans_f.cu -- instantiates the compile-time planned FFT passes for lengths 1152, 1536, 1600, 1920
// (see b2n_fft_fast_kernels.cuh; the plans are spread over several translation units so that they compile in parallel).
#include "b2n_fft_fast_kernels.cuh"

namespace b2n {

B2N_DEFINE_PLAN(1152)
B2N_DEFINE_PLAN(1536)
B2N_DEFINE_PLAN(1600)
B2N_DEFINE_PLAN(1920)

}  // namespace b2n
