// TEST SCAFFOLDING: runs the engine's __host__ __device__ Stockham FFT building blocks
// (torchkbnufft_b200/csrc/b2n_fft_core.cuh) on the CPU so tests/test_fft_core.py can check
// them against numpy without a GPU.  Not part of libb200nufft.so.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "../torchkbnufft_b200/csrc/b2n_fft_core.cuh"

// returns the number of stages, or -1 when n is not supported; radix_out gets the radices
extern "C" int host_fft(int n, int inverse, const float *in, float *out, int *radix_out) {
  b2n::FftPlan plan;
  if (!b2n::fft_factorize(n, &plan)) return -1;
  float2 *tw = (float2 *)malloc(sizeof(float2) * n);
  for (int t = 0; t < n; ++t) {
    const double a = -2.0 * M_PI * (double)t / (double)n;
    tw[t].x = (float)cos(a);
    tw[t].y = (float)sin(a);
  }
  float2 *x = (float2 *)malloc(sizeof(float2) * n), *y = (float2 *)malloc(sizeof(float2) * n);
  memcpy(x, in, sizeof(float2) * n);
  int Ns = 1;
  for (int s = 0; s < plan.n_stages; ++s) {
    const int R = plan.radix[s];
    radix_out[s] = R;
    for (int j = 0; j < n / R; ++j) {
      auto load = [&](int i) { return x[i]; };
      auto store = [&](int i, float2 v) { y[i] = v; };
      if (inverse) b2n::fft_stage_item_any<true>(R, n, Ns, j, tw, load, store);
      else b2n::fft_stage_item_any<false>(R, n, Ns, j, tw, load, store);
    }
    float2 *t = x; x = y; y = t;
    Ns *= R;
  }
  memcpy(out, x, sizeof(float2) * n);
  free(tw); free(x); free(y);
  return plan.n_stages;
}
