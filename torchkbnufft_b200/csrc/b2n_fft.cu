// b2n_fft.cu -- pruned, fused FFT passes around the interpolation (complex64).
//
// What this replaces.  The reference multiplies by the apodisation, zero-pads the image to
// the oversampled grid and runs a full fftn over it (torchkbnufft/_nufft/fft.py:66-76); on
// the way back it runs a full ifftn, crops and multiplies (:113-118); the SENSE multiply /
// coil sum are further passes (modules/kbnufft.py:182-183, :404-405).  With cuFFT that was
// 4-5 full passes over the K-grid per direction (measured 70-95 us each way at BASELINE
// config 2, profiles/r01_d).  Here the transform is done one dimension per pass by our own
// shared-memory Stockham kernels (b2n_fft_core.cuh), which lets every pass skip what the
// zero-padding / cropping makes redundant and fuse the element-wise work:
//   forward : rows  -- read image*smaps*scaling (N_x values), FFT_x            -> T [.., N_y, K_x]
//             cols  -- read N_y rows only,                    FFT_y            -> grid [.., K_y, K_x]
//   adjoint : cols  -- read the grid, IFFT_y, keep the first N_y rows           -> T [.., N_y, K_x]
//             rows  -- IFFT_x, keep N_x, * conj(scaling) * conj(smaps), sum over coils -> image
// (3-D adds one more column pass; 1-D is the row pass alone.)  For 2x oversampling the
// traffic per direction drops from ~5 grid passes to ~2.25.  The Toeplitz filter
// (fft.py:121-173) is the forward passes, the inverse passes with the kernel multiply fused
// into the first inverse pass's loads, and no separate multiply pass.
//
// Row pass: a CTA transforms L contiguous lines; column pass: a CTA transforms 8 adjacent
// columns (64-byte global segments), shared layout [element][column].  Ping-pong buffers,
// one __syncthreads per stage; the first stage loads from global, the last stores to global.
// Sizes with a prime factor > 13 are not handled (the Python layer then uses cuFFT).
#include "b2n_common.cuh"
#include "b2n_fft_core.cuh"

namespace b2n {

constexpr int kFftThreads = 256;
constexpr int kColsPerCta = 8;
constexpr int kFftMaxN = 8192;

B2N_HD int fft_pad(int i) { return i + (i >> 3); }  // 1 slot of padding per 8: stride-R stores stay conflict-free

struct FftStages {
  int n, n_stages, radix[kFftMaxStages];
};

enum RowMode { ROW_PLAIN = 0, ROW_FWD_FIRST = 1, ROW_ADJ_LAST = 2 };

struct RowArgs {
  FftStages st;
  int n_in, n_out;       // nonzero inputs / kept outputs per line
  int L;                 // lines per CTA
  int64_t lines;         // lines (ROW_ADJ_LAST: (batch, image row) pairs)
  int64_t rows_per_img;  // image rows per (batch, coil) = prod of the slower image dims
  int C, Ci, Bs;         // coils, image coils (1 or C), smaps batch (1 or B)
  const float2 *in;      // ROW_PLAIN / ROW_ADJ_LAST: [..][n_in]
  float2 *out;           // [..][n_out]
  const float2 *image, *smaps, *scaling;
  float scale;
};

// twiddle table exp(-2 pi i t / n), built once per CTA in double precision
B2N_D void build_twiddles(float2 *tw, int n) {
  for (int t = threadIdx.x; t < n; t += blockDim.x) {
    double s, c;
    sincospi(-2.0 * (double)t / (double)n, &s, &c);
    tw[t] = f2((float)c, (float)s);
  }
}

template <bool INV, int MODE>
__global__ void __launch_bounds__(kFftThreads) k_fft_rows(RowArgs a) {
  extern __shared__ __align__(16) float2 fsm[];
  const int n = a.st.n, NP = fft_pad(n) + 1;
  float2 *tw = fsm;                 // [n]
  float2 *bufA = tw + n;            // [L][NP]
  float2 *bufB = bufA + a.L * NP;   // [L][NP]
  float2 *acc = bufB + a.L * NP;    // ROW_ADJ_LAST: [L][n_out] coil-sum accumulators
  build_twiddles(tw, n);
  const int64_t line0 = (int64_t)blockIdx.x * a.L;
  const int nl = (int)min((int64_t)a.L, a.lines - line0);
  const int n_coil_iter = MODE == ROW_ADJ_LAST ? a.C : 1;
  if (MODE == ROW_ADJ_LAST)
    for (int e = threadIdx.x; e < a.L * a.n_out; e += kFftThreads) acc[e] = f2(0.f, 0.f);
  __syncthreads();

  for (int coil = 0; coil < n_coil_iter; ++coil) {
    // global accessors of this CTA's lines
    auto gload = [&](int line, int i) -> float2 {
      if (i >= a.n_in) return f2(0.f, 0.f);  // zero padding is never read
      const int64_t l = line0 + line;
      if (MODE == ROW_FWD_FIRST) {
        const int64_t bc = l / a.rows_per_img, row = l - bc * a.rows_per_img;
        const int64_t b = bc / a.C, c = bc - b * a.C;
        const int64_t pix = row * a.n_in + i;
        float2 v = a.image[((b * a.Ci + (a.Ci == 1 ? 0 : c)) * a.rows_per_img) * a.n_in + pix];
        if (a.smaps) v = cmul2(v, a.smaps[(((a.Bs == 1 ? 0 : b) * a.C + c) * a.rows_per_img) * a.n_in + pix]);
        if (a.scaling) v = cmul2(v, a.scaling[pix]);
        return f2(v.x * a.scale, v.y * a.scale);
      } else if (MODE == ROW_ADJ_LAST) {
        const int64_t b = l / a.rows_per_img, row = l - b * a.rows_per_img;
        return a.in[(((b * a.C + coil) * a.rows_per_img) + row) * a.n_in + i];
      } else {
        return a.in[l * a.n_in + i];
      }
    };
    auto gstore = [&](int line, int i, float2 v) {
      if (i >= a.n_out) return;  // cropped outputs are never written
      const int64_t l = line0 + line;
      if (MODE == ROW_ADJ_LAST) {
        const int64_t b = l / a.rows_per_img, row = l - b * a.rows_per_img;
        const int64_t pix = row * a.n_out + i;
        if (a.smaps) {
          const float2 s = a.smaps[(((a.Bs == 1 ? 0 : b) * a.C + coil) * a.rows_per_img) * a.n_out + pix];
          v = cmul2(v, f2(s.x, -s.y));
        }
        float2 &t = acc[line * a.n_out + i];  // one owner thread per element: no race
        t = cadd(t, v);
      } else if (MODE == ROW_PLAIN && a.scaling) {
        const int64_t row = l % a.rows_per_img;
        const float2 s = a.scaling[row * a.n_out + i];
        v = cmul2(v, f2(s.x, -s.y));
        a.out[l * a.n_out + i] = f2(v.x * a.scale, v.y * a.scale);
      } else if (MODE == ROW_FWD_FIRST) {
        a.out[l * a.n_out + i] = v;  // scale was applied with the apodisation at load time
      } else {
        a.out[l * a.n_out + i] = f2(v.x * a.scale, v.y * a.scale);
      }
    };

    if (a.st.n_stages == 0) {  // n == 1
      for (int line = threadIdx.x; line < nl; line += kFftThreads) gstore(line, 0, gload(line, 0));
    }
    float2 *src = bufA, *dst = bufB;
    int Ns = 1;
    for (int s = 0; s < a.st.n_stages; ++s) {
      const int R = a.st.radix[s], per_line = n / R;
      const bool first = s == 0, last = s == a.st.n_stages - 1;
      for (int it = threadIdx.x; it < nl * per_line; it += kFftThreads) {
        const int line = it / per_line, j = it - line * per_line;
        auto load = [&](int i) -> float2 { return first ? gload(line, i) : src[line * NP + fft_pad(i)]; };
        auto store = [&](int i, float2 v) {
          if (last) gstore(line, i, v);
          else dst[line * NP + fft_pad(i)] = v;
        };
        fft_stage_item_any<INV>(R, n, Ns, j, tw, load, store);
      }
      __syncthreads();
      float2 *t = src;
      src = dst;
      dst = t;
      Ns *= R;
    }
  }
  if (MODE == ROW_ADJ_LAST) {
    // apodisation and store of the coil-combined rows
    for (int e = threadIdx.x; e < nl * a.n_out; e += kFftThreads) {
      const int line = e / a.n_out, i = e - line * a.n_out;
      const int64_t l = line0 + line;
      const int64_t row = l % a.rows_per_img;
      float2 v = acc[e];
      if (a.scaling) {
        const float2 s = a.scaling[row * a.n_out + i];
        v = cmul2(v, f2(s.x, -s.y));
      }
      a.out[l * a.n_out + i] = f2(v.x * a.scale, v.y * a.scale);
    }
  }
}

struct ColArgs {
  FftStages st;
  int n_in, n_out;   // rows read / rows written along the transformed dimension
  int64_t A, X;      // outer count, inner (contiguous) extent
  const float2 *in;  // [A][n_in][X]
  float2 *out;       // [A][n_out][X]
  const float2 *mul; // optional [mul_batch][n][X] factor applied to the inputs (Toeplitz kernel)
  int64_t a_per_mul; // outer indices per mul batch entry (0: single kernel)
  float scale;
};

template <bool INV>
__global__ void __launch_bounds__(kFftThreads) k_fft_cols(ColArgs a) {
  extern __shared__ __align__(16) float2 fsm[];
  constexpr int LX = kColsPerCta;
  const int n = a.st.n, NP = fft_pad(n) + 1;
  float2 *tw = fsm;            // [n]
  float2 *bufA = tw + n;       // [NP][LX]
  float2 *bufB = bufA + NP * LX;
  build_twiddles(tw, n);
  const int64_t xblocks = (a.X + LX - 1) / LX;
  const int64_t oa = blockIdx.x / xblocks;
  const int64_t x0 = (blockIdx.x - oa * xblocks) * LX;
  const float2 *in = a.in + oa * a.n_in * a.X + x0;
  float2 *out = a.out + oa * a.n_out * a.X + x0;
  const float2 *mul = a.mul ? a.mul + (a.a_per_mul ? (oa / a.a_per_mul) * (int64_t)n * a.X : 0) + x0 : nullptr;
  const int nx = (int)min((int64_t)LX, a.X - x0);
  __syncthreads();

  auto gload = [&](int l, int i) -> float2 {
    if (i >= a.n_in || l >= nx) return f2(0.f, 0.f);
    float2 v = in[(int64_t)i * a.X + l];
    if (mul) v = cmul2(v, mul[(int64_t)i * a.X + l]);
    return v;
  };
  auto gstore = [&](int l, int i, float2 v) {
    if (i < a.n_out && l < nx) out[(int64_t)i * a.X + l] = f2(v.x * a.scale, v.y * a.scale);
  };
  if (a.st.n_stages == 0) {
    for (int l = threadIdx.x; l < nx; l += kFftThreads) gstore(l, 0, gload(l, 0));
    return;
  }
  float2 *src = bufA, *dst = bufB;
  int Ns = 1;
  for (int s = 0; s < a.st.n_stages; ++s) {
    const int R = a.st.radix[s], per_line = n / R;
    const bool first = s == 0, last = s == a.st.n_stages - 1;
    for (int it = threadIdx.x; it < LX * per_line; it += kFftThreads) {
      const int j = it / LX, l = it - j * LX;  // column index fastest: 64-byte global segments
      auto load = [&](int i) -> float2 { return first ? gload(l, i) : src[fft_pad(i) * LX + l]; };
      auto store = [&](int i, float2 v) {
        if (last) gstore(l, i, v);
        else dst[fft_pad(i) * LX + l] = v;
      };
      fft_stage_item_any<INV>(R, n, Ns, j, tw, load, store);
    }
    __syncthreads();
    float2 *t = src;
    src = dst;
    dst = t;
    Ns *= R;
  }
}

// ---- host side ------------------------------------------------------------------------------
static bool make_stages(int64_t n, FftStages *st) {
  FftPlan p;
  if (n < 1 || n > kFftMaxN || !fft_factorize((int)n, &p)) return false;
  st->n = p.n;
  st->n_stages = p.n_stages;
  for (int i = 0; i < p.n_stages; ++i) st->radix[i] = p.radix[i];
  return true;
}

static int rows_lines_per_cta(int n) { return n <= 1024 ? 4 : (n <= 2048 ? 2 : 1); }

template <bool INV, int MODE> static int launch_rows(RowArgs &a, cudaStream_t st) {
  if (a.lines <= 0) return 0;
  a.L = rows_lines_per_cta(a.st.n);
  const int NP = fft_pad(a.st.n) + 1;
  size_t smem = sizeof(float2) * ((size_t)a.st.n + 2 * (size_t)a.L * NP + (MODE == ROW_ADJ_LAST ? (size_t)a.L * a.n_out : 0));
  auto kern = k_fft_rows<INV, MODE>;
  B2N_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<(unsigned)ceil_div(a.lines, a.L), kFftThreads, smem, st>>>(a);
  B2N_LAUNCH_OK("k_fft_rows");
  return 0;
}

template <bool INV> static int launch_cols(ColArgs &a, cudaStream_t st) {
  if (a.A <= 0 || a.X <= 0) return 0;
  const int NP = fft_pad(a.st.n) + 1;
  size_t smem = sizeof(float2) * ((size_t)a.st.n + 2 * (size_t)NP * kColsPerCta);
  auto kern = k_fft_cols<INV>;
  B2N_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t blocks = a.A * ceil_div(a.X, kColsPerCta);
  kern<<<(unsigned)blocks, kFftThreads, smem, st>>>(a);
  B2N_LAUNCH_OK("k_fft_cols");
  return 0;
}

struct FusedGeom {
  int ndim;
  int64_t N[3], K[3];  // image / grid sizes in dimension order (slowest first)
  int64_t B, C;
  FftStages st[3];
};

static int make_fused_geom(int ndim, const int64_t *im_size, const int64_t *grid_size, int64_t B, int64_t C,
                           FusedGeom *g) {
  if (ndim < 1 || ndim > 3 || !im_size || !grid_size) return fail_arg(B2N_E_ARG, "bad ndim/im_size/grid_size");
  if (B < 1 || C < 1) return fail_arg(B2N_E_ARG, "n_batch=%lld n_coils=%lld", (long long)B, (long long)C);
  g->ndim = ndim;
  g->B = B;
  g->C = C;
  for (int d = 0; d < ndim; ++d) {
    if (im_size[d] < 1 || grid_size[d] < im_size[d]) return fail_arg(B2N_E_ARG, "grid_size[%d] < im_size[%d]", d, d);
    g->N[d] = im_size[d];
    g->K[d] = grid_size[d];
    if (!make_stages(grid_size[d], &g->st[d]))
      return fail_arg(B2N_E_UNSUPPORTED, "FFT length %lld is not supported (prime factor > 13 or > %d)",
                      (long long)grid_size[d], kFftMaxN);
  }
  return 0;
}

// intermediate buffers: after the row pass [B*C][N0..N_{d-2}][K_last]; 3-D adds [B*C][N0][K1][K2]
static size_t fused_work_elems(const FusedGeom &g, size_t *t1_elems) {
  size_t t1 = (size_t)g.B * g.C * g.K[g.ndim - 1];
  for (int d = 0; d < g.ndim - 1; ++d) t1 *= g.N[d];
  size_t t2 = g.ndim == 3 ? (size_t)g.B * g.C * g.N[0] * g.K[1] * g.K[2] : 0;
  if (g.ndim == 1) t1 = 0;
  if (t1_elems) *t1_elems = t1;
  return t1 + t2;
}

static int fused_forward(const FusedGeom &g, const float2 *image, int64_t Ci, const float2 *smaps, int64_t Bs,
                         const float2 *scaling, float scale, float2 *grid, float2 *work, cudaStream_t st) {
  const int d = g.ndim;
  size_t t1e;
  fused_work_elems(g, &t1e);
  float2 *T1 = work, *T2 = work + t1e;
  RowArgs r;
  memset(&r, 0, sizeof(r));
  r.st = g.st[d - 1];
  r.n_in = (int)g.N[d - 1];
  r.n_out = (int)g.K[d - 1];
  r.rows_per_img = 1;
  for (int k = 0; k < d - 1; ++k) r.rows_per_img *= g.N[k];
  r.lines = g.B * g.C * r.rows_per_img;
  r.C = (int)g.C;
  r.Ci = (int)Ci;
  r.Bs = (int)Bs;
  r.image = image;
  r.smaps = smaps;
  r.scaling = scaling;
  r.scale = scale;
  r.out = d == 1 ? grid : T1;
  int rc = launch_rows<false, ROW_FWD_FIRST>(r, st);
  if (rc || d == 1) return rc;
  ColArgs c;
  memset(&c, 0, sizeof(c));
  c.scale = 1.f;
  if (d == 2) {
    c.st = g.st[0];
    c.n_in = (int)g.N[0];
    c.n_out = (int)g.K[0];
    c.A = g.B * g.C;
    c.X = g.K[1];
    c.in = T1;
    c.out = grid;
    return launch_cols<false>(c, st);
  }
  c.st = g.st[1];  // 3-D: y pass on the N0 non-zero planes, then z pass
  c.n_in = (int)g.N[1];
  c.n_out = (int)g.K[1];
  c.A = g.B * g.C * g.N[0];
  c.X = g.K[2];
  c.in = T1;
  c.out = T2;
  rc = launch_cols<false>(c, st);
  if (rc) return rc;
  c.st = g.st[0];
  c.n_in = (int)g.N[0];
  c.n_out = (int)g.K[0];
  c.A = g.B * g.C;
  c.X = g.K[1] * g.K[2];
  c.in = T2;
  c.out = grid;
  return launch_cols<false>(c, st);
}

// kernel (optional): Toeplitz factor multiplied into the loads of the first inverse pass
static int fused_adjoint(const FusedGeom &g, const float2 *grid, const float2 *kernel, int64_t kernel_batch,
                         const float2 *smaps, int64_t Bs, const float2 *scaling, float scale, float2 *image,
                         float2 *work, cudaStream_t st) {
  const int d = g.ndim;
  size_t t1e;
  fused_work_elems(g, &t1e);
  float2 *T1 = work, *T2 = work + t1e;
  const float2 *rows_in = grid;
  ColArgs c;
  memset(&c, 0, sizeof(c));
  c.scale = 1.f;
  if (d == 3) {
    c.st = g.st[0];
    c.n_in = (int)g.K[0];
    c.n_out = (int)g.N[0];
    c.A = g.B * g.C;
    c.X = g.K[1] * g.K[2];
    c.in = grid;
    c.out = T2;
    c.mul = kernel;
    c.a_per_mul = kernel_batch > 1 ? g.C : 0;
    int rc = launch_cols<true>(c, st);
    if (rc) return rc;
    c.mul = nullptr;
    c.st = g.st[1];
    c.n_in = (int)g.K[1];
    c.n_out = (int)g.N[1];
    c.A = g.B * g.C * g.N[0];
    c.X = g.K[2];
    c.in = T2;
    c.out = T1;
    rc = launch_cols<true>(c, st);
    if (rc) return rc;
    rows_in = T1;
  } else if (d == 2) {
    c.st = g.st[0];
    c.n_in = (int)g.K[0];
    c.n_out = (int)g.N[0];
    c.A = g.B * g.C;
    c.X = g.K[1];
    c.in = grid;
    c.out = T1;
    c.mul = kernel;
    c.a_per_mul = kernel_batch > 1 ? g.C : 0;
    int rc = launch_cols<true>(c, st);
    if (rc) return rc;
    rows_in = T1;
  } else if (kernel) {
    return fail_arg(B2N_E_UNSUPPORTED, "1-D Toeplitz filtering goes through the unfused path");
  }
  RowArgs r;
  memset(&r, 0, sizeof(r));
  r.st = g.st[d - 1];
  r.n_in = (int)g.K[d - 1];
  r.n_out = (int)g.N[d - 1];
  r.rows_per_img = 1;
  for (int k = 0; k < d - 1; ++k) r.rows_per_img *= g.N[k];
  r.C = (int)g.C;
  r.Bs = (int)Bs;
  r.in = rows_in;
  r.out = image;
  r.smaps = smaps;
  r.scaling = scaling;
  r.scale = scale;
  if (smaps) {
    r.lines = g.B * r.rows_per_img;
    return launch_rows<true, ROW_ADJ_LAST>(r, st);
  }
  r.lines = g.B * g.C * r.rows_per_img;
  return launch_rows<true, ROW_PLAIN>(r, st);
}

}  // namespace b2n

using namespace b2n;

extern "C" int b2n_fft_supported(int64_t n) {
  FftStages st;
  return make_stages(n, &st) ? 1 : 0;
}

extern "C" int b2n_fft_work_bytes(int ndim, const int64_t *im_size, const int64_t *grid_size, int64_t n_batch,
                                  int64_t n_coils, size_t *bytes) {
  FusedGeom g;
  int rc = make_fused_geom(ndim, im_size, grid_size, n_batch, n_coils, &g);
  if (rc) return rc;
  if (!bytes) return fail_arg(B2N_E_ARG, "bytes is NULL");
  *bytes = sizeof(float2) * fused_work_elems(g, nullptr);
  return 0;
}

extern "C" int b2n_fft_forward_fused(int ndim, const int64_t *im_size, const int64_t *grid_size, int64_t n_batch,
                                     int64_t n_coils, const void *image_dev, int64_t image_coils,
                                     const void *smaps_dev, int64_t smaps_batch, const void *scaling_dev, double scale,
                                     void *grid_dev, void *work_dev, void *stream) {
  FusedGeom g;
  int rc = make_fused_geom(ndim, im_size, grid_size, n_batch, n_coils, &g);
  if (rc) return rc;
  if (!image_dev || !grid_dev || (ndim > 1 && !work_dev)) return fail_arg(B2N_E_ARG, "NULL image/grid/work");
  if (image_coils != 1 && image_coils != n_coils) return fail_arg(B2N_E_ARG, "image_coils must be 1 or n_coils");
  if (smaps_dev && smaps_batch != 1 && smaps_batch != n_batch) return fail_arg(B2N_E_ARG, "smaps_batch must be 1 or n_batch");
  return fused_forward(g, (const float2 *)image_dev, image_coils, (const float2 *)smaps_dev, smaps_dev ? smaps_batch : 1,
                       (const float2 *)scaling_dev, (float)scale, (float2 *)grid_dev, (float2 *)work_dev,
                       (cudaStream_t)stream);
}

extern "C" int b2n_fft_adjoint_fused(int ndim, const int64_t *im_size, const int64_t *grid_size, int64_t n_batch,
                                     int64_t n_coils, const void *grid_dev, const void *kernel_dev,
                                     int64_t kernel_batch, const void *smaps_dev, int64_t smaps_batch,
                                     const void *scaling_dev, double scale, void *image_dev, void *work_dev,
                                     void *stream) {
  FusedGeom g;
  int rc = make_fused_geom(ndim, im_size, grid_size, n_batch, n_coils, &g);
  if (rc) return rc;
  if (!image_dev || !grid_dev || (ndim > 1 && !work_dev)) return fail_arg(B2N_E_ARG, "NULL image/grid/work");
  if (smaps_dev && smaps_batch != 1 && smaps_batch != n_batch) return fail_arg(B2N_E_ARG, "smaps_batch must be 1 or n_batch");
  if (kernel_dev && kernel_batch != 1 && kernel_batch != n_batch) return fail_arg(B2N_E_ARG, "kernel_batch must be 1 or n_batch");
  return fused_adjoint(g, (const float2 *)grid_dev, (const float2 *)kernel_dev, kernel_dev ? kernel_batch : 1,
                       (const float2 *)smaps_dev, smaps_dev ? smaps_batch : 1, (const float2 *)scaling_dev, (float)scale,
                       (float2 *)image_dev, (float2 *)work_dev, (cudaStream_t)stream);
}
