// b2n_fft_plans_e.cu -- instantiates the compile-time planned FFT passes for lengths 160, 200, 240, 400, 800
// (see b2n_fft_fast_kernels.cuh; the plans are spread over several translation units so that they compile in parallel).
#include "b2n_fft_fast_kernels.cuh"

namespace b2n {

B2N_DEFINE_PLAN(160)
B2N_DEFINE_PLAN(200)
B2N_DEFINE_PLAN(240)
B2N_DEFINE_PLAN(400)
B2N_DEFINE_PLAN(800)

}  // namespace b2n
