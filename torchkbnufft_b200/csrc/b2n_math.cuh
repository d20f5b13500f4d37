// b2n_math.cuh -- coordinate / index arithmetic of the table-interpolation path.
//
// These few operations decide which table entry and which grid cell a sample
// touches, and the reference evaluates them in the trajectory's own precision
// (float32 for complex64 data).  To stay within the 1e-5 forward tolerance the
// engine must reproduce them BIT FOR BIT, so every step is a
// single correctly-rounded IEEE operation: no FMA contraction, no fast division.
// The functions are __host__ __device__ so tests/test_index_math.py can compile
// them with the host compiler (csrc/b2n_math_host.cpp, built with
// -ffp-contract=off) and compare against the reference's indices on the CPU
// build machine.  That host build is test scaffolding, not an execution path.
//
// reference: torchkbnufft/_nufft/interp.py
//   :171/:663  tm   = omega / (2*pi / K)          (reciprocal of K, then * 2pi)
//   :177/:670  base = 1 + floor(tm - J/2)
//   :129-132   dist = round((tm - float(base + j)) * L)   (half to even)
//   :174       centre = floor(J*L/2);  table index = dist + centre
//   :146       cell = (base + j) mod K  (non-negative)
//   :200-203   phase = exp(i * sum_d omega_d * n_shift_d)
#pragma once
#include <math.h>

#include "b2n_common.cuh"

namespace b2n {

#ifdef __CUDA_ARCH__
B2N_D float rn_mul(float a, float b) { return __fmul_rn(a, b); }
B2N_D float rn_add(float a, float b) { return __fadd_rn(a, b); }
B2N_D float rn_sub(float a, float b) { return __fsub_rn(a, b); }
B2N_D float rn_div(float a, float b) { return __fdiv_rn(a, b); }
B2N_D double rn_mul(double a, double b) { return __dmul_rn(a, b); }
B2N_D double rn_add(double a, double b) { return __dadd_rn(a, b); }
B2N_D double rn_sub(double a, double b) { return __dsub_rn(a, b); }
B2N_D double rn_div(double a, double b) { return __ddiv_rn(a, b); }
#else
// host build: compiled with -ffp-contract=off, plain operators are single IEEE ops
template <typename T> inline T rn_mul(T a, T b) { volatile T r = a * b; return r; }
template <typename T> inline T rn_add(T a, T b) { volatile T r = a + b; return r; }
template <typename T> inline T rn_sub(T a, T b) { volatile T r = a - b; return r; }
template <typename T> inline T rn_div(T a, T b) { volatile T r = a / b; return r; }
#endif

B2N_HD float fl_floor(float v) { return floorf(v); }
B2N_HD double fl_floor(double v) { return floor(v); }
B2N_HD float fl_rint(float v) { return rintf(v); }   // round half to even
B2N_HD double fl_rint(double v) { return rint(v); }

// gam = (1/K) * 2pi in T
template <typename T> B2N_HD T grid_spacing(int64_t K) {
  const T recip = rn_div(T(1), T(K));
  return rn_mul(recip, T(6.283185307179586476925286766559));
}

// normalised coordinate tm and (unwrapped) base cell of the J-point footprint
template <typename T> B2N_HD void locate(T omega, int64_t K, int J, T &tm, int64_t &base) {
  tm = rn_div(omega, grid_spacing<T>(K));
  const T half = T(J) / T(2);  // exact
  base = 1 + (int64_t)fl_floor(rn_sub(tm, half));
}

// table index of neighbour cell g (unwrapped) incl. the table centre
template <typename T> B2N_HD int64_t table_index(T tm, int64_t g, int J, int L) {
  const T diff = rn_sub(tm, T(g));
  const int64_t dist = (int64_t)fl_rint(rn_mul(diff, T(L)));
  return dist + (int64_t)((J * L) / 2);
}

// python-style modulo, result in [0, K)
B2N_HD int64_t wrap_cell(int64_t g, int64_t K) {
  int64_t r = g % K;
  return r < 0 ? r + K : r;
}

// fftshift phase argument: products rounded one by one, summed left to right
template <typename T> B2N_HD T phase_arg(const T *omega_d, int ndim, const T *n_shift) {
  T arg = rn_mul(omega_d[0], n_shift[0]);
  for (int d = 1; d < ndim; ++d) arg = rn_add(arg, rn_mul(omega_d[d], n_shift[d]));
  return arg;
}

}  // namespace b2n
