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

// ---- compile-time planned passes (b2n_fft_fast.cuh): the kernel's phase sequence, one "thread"
// at a time, with a plain array standing in for the shared exchange buffer ----------------------
#include <vector>

#include "../torchkbnufft_b200/csrc/b2n_fft_fast.cuh"

template <class P, bool INV, bool HIN = false>
static void run_fast(const float2 *a, const float2 *b, float2 *oa, float2 *ob) {
  using namespace b2n::fast;
  std::vector<float2> tws(P::N);
  for (int e = 0; e < P::TW_COUNT; ++e) {
    int r, k, period;
    staged_twiddle_index<P>(e, &r, &k, &period);
    const double ang = -2.0 * M_PI * (double)((long long)r * k % period) / (double)period;
    tws[e].x = (float)cos(ang);
    tws[e].y = (float)sin(ang);
  }
  std::vector<float4> sm(P::NP), regs((size_t)P::T * P::RMAX);
  auto loadg = [&](int i) { return v4(a[i].x, a[i].y, b[i].x, b[i].y); };
  auto storeg = [&](int i, float4 v) {
    oa[i].x = v.x; oa[i].y = v.y; ob[i].x = v.z; ob[i].y = v.w;
  };
  for (int t = 0; t < P::I0; ++t) {
    float4 v[P::RMAX];
    if constexpr (HIN) {  // inputs beyond N/2 are zero padding: never read
      for (int r = 0; r < P::R0 / 2; ++r) v[r] = loadg(t + r * P::I0);
      dft_half_in<P::R0, INV>(v);
    } else {
      for (int r = 0; r < P::R0; ++r) v[r] = loadg(t + r * P::I0);
      dft<P::R0, INV>(v);
    }
    for (int r = 0; r < P::R0; ++r) sm[t * (P::R0 + 1) + r] = v[r];
  }
  for (int t = 0; t < P::I1; ++t) {  // all loads before any store: the single-buffer hazard
    float4 *v = &regs[(size_t)t * P::RMAX];
    for (int r = 0; r < P::R1; ++r) v[r] = sm[P::pad(t) + P::padc(r * P::I1)];
    stage_compute<P::R1, INV>(v, tws.data(), P::R0, t & (P::R0 - 1));
  }
  for (int t = 0; t < P::I1; ++t) {
    float4 *v = &regs[(size_t)t * P::RMAX];
    const int k1 = t & (P::R0 - 1), o1 = (t - k1) * P::R1 + k1;
    for (int r = 0; r < P::R1; ++r) {
      if (P::NS == 2) storeg(o1 + r * P::R0, v[r]);
      else sm[P::pad(o1) + P::padc(r * P::R0)] = v[r];
    }
  }
  if constexpr (P::NS == 3) {
    constexpr int Ns2 = P::R0 * P::R1;
    for (int t = 0; t < P::I2; ++t) {
      float4 v[P::RMAX];
      for (int r = 0; r < P::R2; ++r) v[r] = sm[P::pad(t) + P::padc(r * P::I2)];
      stage_compute<P::R2, INV>(v, tws.data() + P::TW2, Ns2, t);
      for (int r = 0; r < P::R2; ++r) storeg(t + r * Ns2, v[r]);
    }
  }
}

// two lines in, two lines out; returns 0 when length n has no compile-time plan.
// half_in: the caller guarantees a[i] = b[i] = 0 for i >= n/2 and the pruned first stage is used.
extern "C" int host_fft_fast(int n, int inverse, int half_in, const float *a, const float *b, float *oa, float *ob) {
  const float2 *pa = (const float2 *)a, *pb = (const float2 *)b;
  float2 *qa = (float2 *)oa, *qb = (float2 *)ob;
  B2N_FAST_PLAN_SWITCH(n,
                       (half_in ? (inverse ? run_fast<P, true, true>(pa, pb, qa, qb) : run_fast<P, false, true>(pa, pb, qa, qb))
                                : (inverse ? run_fast<P, true>(pa, pb, qa, qb) : run_fast<P, false>(pa, pb, qa, qb))),
                       return 0)
  return 1;
}
