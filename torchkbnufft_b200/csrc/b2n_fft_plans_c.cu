// b2n_fft_plans_c.cu -- instantiates the compile-time planned FFT passes for lengths 480, 512, 576, 640, 768
// (see b2n_fft_fast_kernels.cuh; the plans are spread over several translation units so that they compile in parallel).
#include "b2n_fft_fast_kernels.cuh"

namespace b2n {

B2N_DEFINE_PLAN(480)
B2N_DEFINE_PLAN(512)
B2N_DEFINE_PLAN(576)
B2N_DEFINE_PLAN(640)
B2N_DEFINE_PLAN(768)

}  // namespace b2n
