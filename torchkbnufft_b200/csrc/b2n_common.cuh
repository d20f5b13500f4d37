// b2n_common.cuh -- shared helpers for libb200nufft (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/b200nufft.h"

#ifdef __CUDACC__
#define B2N_HD __host__ __device__ __forceinline__
#define B2N_D __device__ __forceinline__
#else
#define B2N_HD inline
#define B2N_D inline
#endif

namespace b2n {

// thread-local error text returned by b2n_last_error()
void set_error(const char *fmt, ...);
int fail_arg(int code, const char *fmt, ...);
int check_cuda(cudaError_t err, const char *what);
void count_launch();

#define B2N_CUDA_OK(expr)                                  \
  do {                                                     \
    int _rc = ::b2n::check_cuda((expr), #expr);            \
    if (_rc != 0) return _rc;                              \
  } while (0)

// after every launch of one of the library's own kernels: error check + the process-wide launch counter
// (b2n_launch_count(); benchmarks report it as evidence that the native kernels ran)
#define B2N_LAUNCH_OK(name)                                \
  do {                                                     \
    int _rc = ::b2n::check_cuda(cudaGetLastError(), name); \
    if (_rc != 0) return _rc;                              \
    ::b2n::count_launch();                                 \
  } while (0)

// Opt a kernel in to `bytes` of dynamic shared memory once per (kernel instantiation, device): the attribute call
// costs about a microsecond of host time, which matters next to 15-100 us kernels.
#define B2N_SMEM_OPT_IN(kern, bytes)                                                                       \
  do {                                                                                                     \
    static int _b2n_done[16] = {0};                                                                        \
    int _b2n_dev = 0;                                                                                      \
    if (cudaGetDevice(&_b2n_dev) != cudaSuccess || _b2n_dev < 0 || _b2n_dev >= 16) _b2n_dev = -1;          \
    if (_b2n_dev < 0 || _b2n_done[_b2n_dev] < (int)(bytes)) {                                              \
      B2N_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));  \
      if (_b2n_dev >= 0) _b2n_done[_b2n_dev] = (int)(bytes);                                               \
    }                                                                                                      \
  } while (0)

// ---- programmatic dependent launch (sm_90+) ----------------------------------------------------------------
// A kernel launched with launch_pdl() and B2N_OPT_PDL set may be scheduled while the tail of the preceding kernel in
// the stream is still running: its CTAs execute their prologue and block in griddep_wait() until the preceding grid
// has completed and its writes are visible.  griddep_launch() in the preceding kernel lets that happen as soon as all
// of ITS CTAs have been scheduled.  Without the launch attribute both instructions are no-ops.
extern int g_pdl;
extern int g_prefetch;
extern int g_zero_kernel;
extern int g_counters_early;
#ifdef __CUDACC__
B2N_D void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// L2 prefetch of a 16-byte aligned span (multiple of 16 bytes) / of the cache line holding p
B2N_D void prefetch_l2_bulk(const void *p, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
B2N_D void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
B2N_D void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
#endif

template <typename T> struct cplx { T x, y; };
template <> struct __align__(8) cplx<float> { float x, y; };
template <> struct __align__(16) cplx<double> { double x, y; };

template <typename T> B2N_HD cplx<T> cmul(cplx<T> a, cplx<T> b) {
  cplx<T> r;
  r.x = a.x * b.x - a.y * b.y;
  r.y = a.x * b.y + a.y * b.x;
  return r;
}
template <typename T> B2N_HD cplx<T> cconj(cplx<T> a) {
  cplx<T> r;
  r.x = a.x;
  r.y = -a.y;
  return r;
}
// acc += a * b
template <typename T> B2N_HD void cmac(cplx<T> &acc, cplx<T> a, cplx<T> b) {
  acc.x += a.x * b.x - a.y * b.y;
  acc.y += a.x * b.y + a.y * b.x;
}

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace b2n
