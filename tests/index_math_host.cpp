// TEST SCAFFOLDING: compiles the engine's __host__ __device__ index arithmetic
// (torchkbnufft_b200/csrc/b2n_math.cuh) with the host compiler so that
// tests/test_index_math.py can check it against the reference's integer indices on
// a machine without a GPU.  Not linked into libb200nufft.so, never used by the product.
#include <stdint.h>

#include "../torchkbnufft_b200/csrc/b2n_math.cuh"

template <typename T>
static void indices(int ndim, int64_t M, const T *omega, const int64_t *K, const int64_t *J, const int64_t *L,
                    int64_t *arr_ind, int64_t *tab_idx, T *phase_arg_out, const T *n_shift) {
  int64_t W = 1;
  for (int d = 0; d < ndim; ++d) W *= J[d];
  for (int64_t m = 0; m < M; ++m) {
    T tm[3], om[3];
    int64_t base[3];
    for (int d = 0; d < ndim; ++d) {
      om[d] = omega[d * M + m];
      b2n::locate<T>(om[d], K[d], (int)J[d], tm[d], base[d]);
    }
    phase_arg_out[m] = b2n::phase_arg<T>(om, ndim, n_shift);
    for (int64_t w = 0; w < W; ++w) {
      int64_t rem = w, j[3], flat = 0;
      for (int d = ndim - 1; d >= 0; --d) { j[d] = rem % J[d]; rem /= J[d]; }
      for (int d = 0; d < ndim; ++d) {
        const int64_t g = base[d] + j[d];
        tab_idx[(w * ndim + d) * M + m] = b2n::table_index<T>(tm[d], g, (int)J[d], (int)L[d]);
        flat = flat * K[d] + b2n::wrap_cell(g, K[d]);
      }
      arr_ind[w * M + m] = flat;
    }
  }
}

extern "C" void host_indices_f32(int ndim, int64_t M, const float *omega, const int64_t *K, const int64_t *J,
                                 const int64_t *L, int64_t *arr_ind, int64_t *tab_idx, float *phase_arg,
                                 const float *n_shift) {
  indices<float>(ndim, M, omega, K, J, L, arr_ind, tab_idx, phase_arg, n_shift);
}
extern "C" void host_indices_f64(int ndim, int64_t M, const double *omega, const int64_t *K, const int64_t *J,
                                 const int64_t *L, int64_t *arr_ind, int64_t *tab_idx, double *phase_arg,
                                 const double *n_shift) {
  indices<double>(ndim, M, omega, K, J, L, arr_ind, tab_idx, phase_arg, n_shift);
}
