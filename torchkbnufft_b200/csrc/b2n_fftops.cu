// b2n_fftops.cu -- the element-wise steps around the oversampled FFT, fused:
//   b2n_apod_pad           image * smaps * scaling -> zero-padded grid
//   b2n_crop_apod_coilsum  crop(grid) * conj(scaling) * conj(smaps) -> (sum over coils) image
//   b2n_spectrum_mul       Toeplitz filter multiply, in place
// reference: torchkbnufft/_nufft/fft.py:36-118 (fft_and_scale / ifft_and_scale /
// crop_dims), :121-173 (fft_filter), modules/kbnufft.py:182-183 and :404-405 (SENSE).
// The reference spends one full pass over the grid per operator (mul, F.pad,
// index_select per dim, mul, sum); here each direction is ONE pass.
//
// Layouts: coil-major grids (B, C, *K) use coil-major smaps (Bs, C, *N);
// channel-last grids (B, *K, C) use channel-last smaps (Bs, *N, C).
#include "b2n_common.cuh"

namespace b2n {

struct PadGeom {
  int64_t N[3], K[3];  // padded to 3 dims with leading 1s
  int64_t Nprod, Kprod;
  int64_t B, C, Ci, Bs;
};

static int make_pad_geom(int ndim, const int64_t *im_size, const int64_t *grid_size, int64_t B, int64_t C, int64_t Ci,
                         int64_t Bs, PadGeom *g) {
  if (ndim < 1 || ndim > 3 || !im_size || !grid_size) return fail_arg(B2N_E_ARG, "bad ndim/im_size/grid_size");
  if (B < 1 || C < 1) return fail_arg(B2N_E_ARG, "n_batch=%lld n_coils=%lld", (long long)B, (long long)C);
  g->Nprod = g->Kprod = 1;
  for (int d = 0; d < 3; ++d) g->N[d] = g->K[d] = 1;
  for (int d = 0; d < ndim; ++d) {
    const int at = 3 - ndim + d;
    if (im_size[d] < 1 || grid_size[d] < im_size[d])
      return fail_arg(B2N_E_ARG, "im_size[%d]=%lld grid_size[%d]=%lld", d, (long long)im_size[d], d,
                      (long long)grid_size[d]);
    g->N[at] = im_size[d];
    g->K[at] = grid_size[d];
    g->Nprod *= im_size[d];
    g->Kprod *= grid_size[d];
  }
  g->B = B;
  g->C = C;
  g->Ci = Ci;
  g->Bs = Bs;
  return 0;
}

// ---- image -> padded grid ------------------------------------------------------
// coil-major: one block row per (b, c, k0, k1) -- no per-element index decode -- and each thread
// writes four 16-byte units spread over the row (all four stores in flight together), so
// the grid is written exactly once with full 128-bit stores.
template <typename T, typename I>
__global__ void __launch_bounds__(256) k_apod_pad_cm(PadGeom g, const cplx<T> *__restrict__ image,
                                                     const cplx<T> *__restrict__ smaps,
                                                     const cplx<T> *__restrict__ scaling, T scale,
                                                     cplx<T> *__restrict__ grid) {
  constexpr int VEC = 16 / (int)sizeof(cplx<T>);
  constexpr int UNROLL = 4;
  const I K0 = (I)g.K[0], K1 = (I)g.K[1], K2 = (I)g.K[2], N0 = (I)g.N[0], N1 = (I)g.N[1], N2 = (I)g.N[2];
  const I C = (I)g.C, Ci = (I)g.Ci, Bs = (I)g.Bs, Np = (I)g.Nprod;
  const I n_rows = (I)g.B * C * K0 * K1;
  const I units_per_row = (K2 + VEC - 1) / VEC;
  for (I row = blockIdx.x; row < n_rows; row += gridDim.x) {
    I t = row / K1;
    const I k1 = row - t * K1;
    I r2 = t;
    t = r2 / K0;
    const I k0 = r2 - t * K0;
    const I b = t / C, c = t - b * C;
    const bool row_inside = k0 < N0 && k1 < N1;
    const I nrow = (k0 * N1 + k1) * N2;
    const cplx<T> *img = image + (b * Ci + (Ci == 1 ? 0 : c)) * Np + nrow;
    const cplx<T> *smp = smaps ? smaps + ((Bs == 1 ? 0 : b) * C + c) * Np + nrow : nullptr;
    const cplx<T> *scl = scaling ? scaling + nrow : nullptr;
    cplx<T> *out = grid + (I)row * K2;
    for (I u0 = threadIdx.x; u0 < units_per_row; u0 += blockDim.x * UNROLL) {
      cplx<T> v[UNROLL][VEC];
#pragma unroll
      for (int q = 0; q < UNROLL; ++q) {
        const I k2 = (u0 + q * blockDim.x) * VEC;
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          v[q][e].x = T(0);
          v[q][e].y = T(0);
          if (row_inside && k2 + e < N2) {
            cplx<T> w = img[k2 + e];
            if (smp) w = cmul(w, smp[k2 + e]);
            if (scl) w = cmul(w, scl[k2 + e]);
            v[q][e].x = w.x * scale;
            v[q][e].y = w.y * scale;
          }
        }
      }
#pragma unroll
      for (int q = 0; q < UNROLL; ++q) {
        const I u = u0 + q * blockDim.x;
        if (u >= units_per_row) break;
        const I k2 = u * VEC;
        if (VEC == 2 && (K2 & 1) == 0) {
          *reinterpret_cast<float4 *>(out + k2) = *reinterpret_cast<const float4 *>(v[q]);  // 16-byte aligned: K2 even
        } else {
#pragma unroll
          for (int e = 0; e < VEC; ++e)
            if (k2 + e < K2) out[k2 + e] = v[q][e];
        }
      }
    }
  }
}

// channel-last: one block row per (b, k0, k1), threads along (k2, c)
template <typename T>
__global__ void __launch_bounds__(256) k_apod_pad_cl(PadGeom g, const cplx<T> *__restrict__ image,
                                                     const cplx<T> *__restrict__ smaps,
                                                     const cplx<T> *__restrict__ scaling, T scale,
                                                     cplx<T> *__restrict__ grid) {
  int64_t row = blockIdx.x;
  const int64_t k1 = row % g.K[1]; row /= g.K[1];
  const int64_t k0 = row % g.K[0];
  const int64_t b = row / g.K[0];
  cplx<T> *out = grid + (((b * g.K[0] + k0) * g.K[1] + k1) * g.K[2]) * g.C;
  const bool row_inside = k0 < g.N[0] && k1 < g.N[1];
  const int64_t nrow = (k0 * g.N[1] + k1) * g.N[2];
  const int64_t inner = g.K[2] * g.C;
  for (int64_t i = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; i < inner; i += (int64_t)gridDim.y * blockDim.x) {
    const int64_t k2 = i / g.C, c = i - k2 * g.C;
    cplx<T> v = {T(0), T(0)};
    if (row_inside && k2 < g.N[2]) {
      const int64_t n = nrow + k2;
      v = image[(b * g.Ci + (g.Ci == 1 ? 0 : c)) * g.Nprod + n];
      if (smaps) v = cmul(v, smaps[(((g.Bs == 1 ? 0 : b) * g.Nprod) + n) * g.C + c]);
      if (scaling) v = cmul(v, scaling[n]);
      v.x *= scale;
      v.y *= scale;
    }
    out[i] = v;
  }
}

// ---- grid -> cropped (coil-combined) image -------------------------------------
// one block row per (b, co, n0, n1) with co = output coil (1 when combining), threads along n2
template <typename T, bool CL>
__global__ void __launch_bounds__(256) k_crop_coilsum(PadGeom g, const cplx<T> *__restrict__ grid,
                                                      const cplx<T> *__restrict__ smaps,
                                                      const cplx<T> *__restrict__ scaling, T scale,
                                                      cplx<T> *__restrict__ image) {
  const int64_t Co = smaps ? 1 : g.C;
  int64_t row = blockIdx.x;
  const int64_t n1 = row % g.N[1]; row /= g.N[1];
  const int64_t n0 = row % g.N[0]; row /= g.N[0];
  const int64_t co = row % Co;
  const int64_t b = row / Co;
  const int64_t nrow = (n0 * g.N[1] + n1) * g.N[2];
  const int64_t krow = (n0 * g.K[1] + n1) * g.K[2];
  cplx<T> *out = image + (b * Co + co) * g.Nprod + nrow;
  const int64_t bs = g.Bs == 1 ? 0 : b;
  for (int64_t n2 = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; n2 < g.N[2]; n2 += (int64_t)gridDim.y * blockDim.x) {
    cplx<T> acc = {T(0), T(0)};
    if (smaps) {
      auto gv = [&](int64_t c) {
        return CL ? grid[(b * g.Kprod + krow + n2) * g.C + c] : grid[(b * g.C + c) * g.Kprod + krow + n2];
      };
      auto sv = [&](int64_t c) {
        return CL ? smaps[(bs * g.Nprod + nrow + n2) * g.C + c] : smaps[(bs * g.C + c) * g.Nprod + nrow + n2];
      };
      int64_t c = 0;
      for (; c + 4 <= g.C; c += 4) {  // eight independent loads in flight per thread; summation order unchanged
        const cplx<T> v0 = gv(c), v1 = gv(c + 1), v2 = gv(c + 2), v3 = gv(c + 3);
        const cplx<T> s0 = sv(c), s1 = sv(c + 1), s2 = sv(c + 2), s3 = sv(c + 3);
        cmac(acc, v0, cconj(s0));
        cmac(acc, v1, cconj(s1));
        cmac(acc, v2, cconj(s2));
        cmac(acc, v3, cconj(s3));
      }
      for (; c < g.C; ++c) cmac(acc, gv(c), cconj(sv(c)));
    } else {
      acc = CL ? grid[(b * g.Kprod + krow + n2) * g.C + co] : grid[(b * g.C + co) * g.Kprod + krow + n2];
    }
    if (scaling) acc = cmul(acc, cconj(scaling[nrow + n2]));
    acc.x *= scale;
    acc.y *= scale;
    out[n2] = acc;
  }
}

// ---- Toeplitz filter multiply -----------------------------------------------------
template <typename T, bool CL>
__global__ void __launch_bounds__(256) k_spectrum_mul(cplx<T> *__restrict__ spec, const cplx<T> *__restrict__ kernel,
                                                      int64_t B, int64_t C, int64_t Kprod, int64_t kernel_batch, T scale) {
  const int64_t total = B * C * Kprod;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t b, k;
    if (CL) {
      const int64_t bk = i / C;  // (b, k)
      b = bk / Kprod;
      k = bk - b * Kprod;
    } else {
      const int64_t bc = i / Kprod;
      k = i - bc * Kprod;
      b = bc / C;
    }
    cplx<T> kv = kernel[(kernel_batch == 1 ? 0 : b) * Kprod + k];
    kv.x *= scale;
    kv.y *= scale;
    spec[i] = cmul(spec[i], kv);
  }
}

static unsigned chunks(int64_t n, int threads) {
  int64_t c = ceil_div(n, threads);
  return (unsigned)(c < 1 ? 1 : (c > 64 ? 64 : c));
}

template <typename T>
static int apod_pad_t(const PadGeom &g, const void *image, const void *smaps, const void *scaling, double scale,
                      int layout, void *grid, cudaStream_t st) {
  if (layout == B2N_CHANNEL_LAST) {
    dim3 gd((unsigned)(g.B * g.K[0] * g.K[1]), chunks(g.K[2] * g.C, 256));
    k_apod_pad_cl<T><<<gd, 256, 0, st>>>(g, (const cplx<T> *)image, (const cplx<T> *)smaps, (const cplx<T> *)scaling,
                                         (T)scale, (cplx<T> *)grid);
  } else {
    int64_t blocks = g.B * g.C * g.K[0] * g.K[1];  // one row per CTA iteration, at most 32 CTAs per SM
    if (blocks > 148 * 32) blocks = 148 * 32;
    if (blocks < 1) blocks = 1;
    dim3 gd((unsigned)blocks);
    if (g.B * g.C * g.Kprod < ((int64_t)1 << 30))
      k_apod_pad_cm<T, int><<<gd, 128, 0, st>>>(g, (const cplx<T> *)image, (const cplx<T> *)smaps,
                                                (const cplx<T> *)scaling, (T)scale, (cplx<T> *)grid);
    else
      k_apod_pad_cm<T, int64_t><<<gd, 128, 0, st>>>(g, (const cplx<T> *)image, (const cplx<T> *)smaps,
                                                    (const cplx<T> *)scaling, (T)scale, (cplx<T> *)grid);
  }
  B2N_LAUNCH_OK("k_apod_pad");
  return 0;
}

template <typename T>
static int crop_t(const PadGeom &g, const void *grid, int layout, const void *smaps, const void *scaling, double scale,
                  void *image, cudaStream_t st) {
  const int64_t Co = smaps ? 1 : g.C;
  dim3 gd((unsigned)(g.B * Co * g.N[0] * g.N[1]), chunks(g.N[2], 256));
  if (layout == B2N_CHANNEL_LAST)
    k_crop_coilsum<T, true><<<gd, 256, 0, st>>>(g, (const cplx<T> *)grid, (const cplx<T> *)smaps,
                                                (const cplx<T> *)scaling, (T)scale, (cplx<T> *)image);
  else
    k_crop_coilsum<T, false><<<gd, 256, 0, st>>>(g, (const cplx<T> *)grid, (const cplx<T> *)smaps,
                                                 (const cplx<T> *)scaling, (T)scale, (cplx<T> *)image);
  B2N_LAUNCH_OK("k_crop_coilsum");
  return 0;
}

// complex64 coil-major entry used by the fused FFT path (b2n_fft.cu) for its final coil combination
int crop_apod_coilsum_c64(int ndim, const int64_t *im_size, const int64_t *grid_size, int64_t B, int64_t C,
                          const void *grid, const void *smaps, int64_t Bs, const void *scaling, double scale,
                          void *image, cudaStream_t st) {
  PadGeom g;
  int rc = make_pad_geom(ndim, im_size, grid_size, B, C, C, smaps ? Bs : 1, &g);
  if (rc) return rc;
  return crop_t<float>(g, grid, B2N_COIL_MAJOR, smaps, scaling, scale, image, st);
}

}  // namespace b2n

using namespace b2n;

static int check_common(int dtype, int layout) {
  if (dtype != B2N_C64 && dtype != B2N_C128) return fail_arg(B2N_E_ARG, "bad dtype %d", dtype);
  if (layout != B2N_COIL_MAJOR && layout != B2N_CHANNEL_LAST) return fail_arg(B2N_E_ARG, "bad layout %d", layout);
  return 0;
}

extern "C" int b2n_apod_pad(int ndim, int dtype, const int64_t *im_size, const int64_t *grid_size, int64_t n_batch,
                            int64_t n_coils, const void *image_dev, int64_t image_coils, const void *smaps_dev,
                            int64_t smaps_batch, const void *scaling_dev, double scale, int grid_layout,
                            void *grid_dev, void *stream) {
  int rc = check_common(dtype, grid_layout);
  if (rc) return rc;
  if (!image_dev || !grid_dev) return fail_arg(B2N_E_ARG, "NULL image/grid");
  if (image_coils != 1 && image_coils != n_coils) return fail_arg(B2N_E_ARG, "image_coils must be 1 or n_coils");
  if (smaps_dev && smaps_batch != 1 && smaps_batch != n_batch) return fail_arg(B2N_E_ARG, "smaps_batch must be 1 or n_batch");
  PadGeom g;
  rc = make_pad_geom(ndim, im_size, grid_size, n_batch, n_coils, image_coils, smaps_dev ? smaps_batch : 1, &g);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  return dtype == B2N_C64 ? apod_pad_t<float>(g, image_dev, smaps_dev, scaling_dev, scale, grid_layout, grid_dev, st)
                          : apod_pad_t<double>(g, image_dev, smaps_dev, scaling_dev, scale, grid_layout, grid_dev, st);
}

extern "C" int b2n_crop_apod_coilsum(int ndim, int dtype, const int64_t *im_size, const int64_t *grid_size,
                                     int64_t n_batch, int64_t n_coils, const void *grid_dev, int grid_layout,
                                     const void *smaps_dev, int64_t smaps_batch, const void *scaling_dev, double scale,
                                     void *image_dev, void *stream) {
  int rc = check_common(dtype, grid_layout);
  if (rc) return rc;
  if (!image_dev || !grid_dev) return fail_arg(B2N_E_ARG, "NULL image/grid");
  if (smaps_dev && smaps_batch != 1 && smaps_batch != n_batch) return fail_arg(B2N_E_ARG, "smaps_batch must be 1 or n_batch");
  PadGeom g;
  rc = make_pad_geom(ndim, im_size, grid_size, n_batch, n_coils, n_coils, smaps_dev ? smaps_batch : 1, &g);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  return dtype == B2N_C64 ? crop_t<float>(g, grid_dev, grid_layout, smaps_dev, scaling_dev, scale, image_dev, st)
                          : crop_t<double>(g, grid_dev, grid_layout, smaps_dev, scaling_dev, scale, image_dev, st);
}

extern "C" int b2n_spectrum_mul(int dtype, void *spectrum_dev, const void *kernel_dev, int64_t n_batch, int64_t n_coils,
                                int64_t n_grid, int64_t kernel_batch, int grid_layout, double scale, void *stream) {
  int rc = check_common(dtype, grid_layout);
  if (rc) return rc;
  if (!spectrum_dev || !kernel_dev || n_batch < 1 || n_coils < 1 || n_grid < 1) return fail_arg(B2N_E_ARG, "bad spectrum args");
  if (kernel_batch != 1 && kernel_batch != n_batch) return fail_arg(B2N_E_ARG, "kernel_batch must be 1 or n_batch");
  const int64_t total = n_batch * n_coils * n_grid;
  int64_t blocks = ceil_div(total, 256);
  if (blocks > 148 * 16) blocks = 148 * 16;  // grid-stride, 16 CTAs per SM
  cudaStream_t st = (cudaStream_t)stream;
  const bool cl = grid_layout == B2N_CHANNEL_LAST;
  if (dtype == B2N_C64) {
    if (cl) k_spectrum_mul<float, true><<<(unsigned)blocks, 256, 0, st>>>((cplx<float> *)spectrum_dev, (const cplx<float> *)kernel_dev, n_batch, n_coils, n_grid, kernel_batch, (float)scale);
    else k_spectrum_mul<float, false><<<(unsigned)blocks, 256, 0, st>>>((cplx<float> *)spectrum_dev, (const cplx<float> *)kernel_dev, n_batch, n_coils, n_grid, kernel_batch, (float)scale);
  } else {
    if (cl) k_spectrum_mul<double, true><<<(unsigned)blocks, 256, 0, st>>>((cplx<double> *)spectrum_dev, (const cplx<double> *)kernel_dev, n_batch, n_coils, n_grid, kernel_batch, scale);
    else k_spectrum_mul<double, false><<<(unsigned)blocks, 256, 0, st>>>((cplx<double> *)spectrum_dev, (const cplx<double> *)kernel_dev, n_batch, n_coils, n_grid, kernel_batch, scale);
  }
  B2N_LAUNCH_OK("k_spectrum_mul");
  return 0;
}
