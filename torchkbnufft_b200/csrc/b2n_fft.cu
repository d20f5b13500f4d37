// b2n_fft.cu -- pruned, fused FFT passes around the interpolation (complex64).
//
// What this replaces.  The reference multiplies by the apodisation, zero-pads the image to
// the oversampled grid and runs a full fftn over it (torchkbnufft/_nufft/fft.py:66-76); on
// the way back it runs a full ifftn, crops and multiplies (:113-118); the SENSE multiply /
// coil sum are further passes (modules/kbnufft.py:182-183, :404-405).  With cuFFT that was
// 4-5 full passes over the K-grid per direction (measured 70-95 us each way at BASELINE
// config 2, profiles/r01_d).  Here the transform is done one dimension per pass by our own
// kernels, which lets every pass skip what the zero-padding / cropping makes redundant and
// fuse the element-wise work:
//   forward : rows  -- read image*smaps*scaling (N_x values), FFT_x            -> T [.., N_y, K_x]
//             cols  -- read N_y rows only,                    FFT_y            -> grid [.., K_y, K_x]
//   adjoint : cols  -- read the grid, IFFT_y, keep the first N_y rows           -> T [.., N_y, K_x]
//             rows  -- IFFT_x, keep N_x, * conj(scaling) * conj(smaps), sum over coils -> image
// (3-D adds one more column pass; 1-D is the row pass alone.)  For 2x oversampling the
// traffic per direction drops from ~5 grid passes to ~2.25.  The Toeplitz filter
// (fft.py:121-173) is the forward passes, the inverse passes with the kernel multiply fused
// into the first inverse pass's loads, and no separate multiply pass.
//
// Two sets of passes:
//  * compile-time planned (b2n_fft_fast.cuh; lengths 64 ... 1024 that 2x-oversampled imaging uses):
//    k_fft_rows_fast / k_fft_cols_fast / k_fft_rows_sense -- register-resident butterflies, two lines
//    per thread, one shared exchange buffer.  46 / 51 us per direction at BASELINE config 2
//    (profiles/r01_g), the default whenever every grid length has a plan;
//  * run-time Stockham (b2n_fft_core.cuh; any length <= 8192 with prime factors <= 13): k_fft_rows /
//    k_fft_cols -- a CTA transforms 4 contiguous lines / 8 adjacent columns through ping-pong shared
//    buffers, one __syncthreads per stage.  Slower than cuFFT (116 / 132 us), kept for lengths
//    without a plan inside a mixed transform and for A/B runs.
// Sizes with a prime factor > 13 are not handled (the Python layer then uses cuFFT).
#include "b2n_fft_args.cuh"
#include "b2n_fft_fast.cuh"

namespace b2n {

constexpr int kFftThreads = 256;
constexpr int kColThreads = 512;  // column pass: 64 threads per column, radices <= 8 (64 registers per thread)
constexpr int kColsPerCta = 8;
constexpr int kRowsPerCta = 4;
constexpr int kFftMaxN = 8192;

B2N_HD int fft_pad(int i) { return i + (i >> 3); }  // 1 slot of padding per 8: stride-R stores stay conflict-free

// One Stockham stage over this CTA's lines.  LOADF(line, i) / STOREF(line, i, v) access the
// stage input / output (global memory in the first / last stage, shared memory otherwise).
// Thread layout: `lanes_per_line` consecutive threads share a line and stride over its items,
// so no per-item division is needed.
template <int R, bool INV, typename LoadF, typename StoreF>
B2N_D void fft_stage_lines(int n, int Ns, const float2 *tw, int line, int j_begin, int j_step, bool line_on,
                           LoadF loadf, StoreF storef) {
  const int per_line = n / R;
  const int tmul = n / (Ns * R);
  const bool pow2 = (Ns & (Ns - 1)) == 0;
  if (!line_on) return;
  for (int j = j_begin; j < per_line; j += j_step) {
    const int k = pow2 ? (j & (Ns - 1)) : (j % Ns);
    float2 v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = loadf(line, j + r * per_line);
    if (Ns > 1) {
      const int tstep = k * tmul;
#pragma unroll
      for (int r = 1; r < R; ++r) v[r] = cmul2(v[r], twid<INV>(tw[r * tstep]));
    }
    Dft<R, INV>::run(v, tw, n);
    const int j0 = (j - k) * R + k;
#pragma unroll
    for (int r = 0; r < R; ++r) storef(line, j0 + r * Ns, v[r]);
  }
}

#define B2N_FFT_RADIX_SWITCH(R_, CALL)          \
  switch (R_) {                                 \
    case 16: { constexpr int RR = 16; CALL; } break; \
    case 8: { constexpr int RR = 8; CALL; } break;   \
    case 4: { constexpr int RR = 4; CALL; } break;   \
    case 2: { constexpr int RR = 2; CALL; } break;   \
    case 3: { constexpr int RR = 3; CALL; } break;   \
    case 5: { constexpr int RR = 5; CALL; } break;   \
    case 7: { constexpr int RR = 7; CALL; } break;   \
    case 11: { constexpr int RR = 11; CALL; } break; \
    default: { constexpr int RR = 13; CALL; } break; \
  }

// column pass: radices <= 8 only (its factorisation never contains 16), so that the 64-register
// budget of 512-thread CTAs is not blown by an unused radix-16 code path
#define B2N_FFT_RADIX_SWITCH_COLS(R_, CALL)     \
  switch (R_) {                                 \
    case 8: { constexpr int RR = 8; CALL; } break;   \
    case 4: { constexpr int RR = 4; CALL; } break;   \
    case 2: { constexpr int RR = 2; CALL; } break;   \
    case 3: { constexpr int RR = 3; CALL; } break;   \
    case 5: { constexpr int RR = 5; CALL; } break;   \
    case 7: { constexpr int RR = 7; CALL; } break;   \
    case 11: { constexpr int RR = 11; CALL; } break; \
    default: { constexpr int RR = 13; CALL; } break; \
  }

// -----------------------------------------------------------------------------------------
// row pass: kRowsPerCta contiguous lines per CTA, 64 threads per line
// -----------------------------------------------------------------------------------------
template <bool INV, int MODE>
__global__ void __launch_bounds__(kFftThreads, 3) k_fft_rows(RowArgs a) {
  extern __shared__ __align__(16) float2 fsm[];
  constexpr int L = kRowsPerCta, TPL = kFftThreads / L;
  const int n = a.st.n, NP = fft_pad(n) + 1;
  float2 *tw = fsm;             // [n]
  float2 *bufA = tw + n;        // [L][NP]
  float2 *bufB = bufA + L * NP; // [L][NP]
  __shared__ const float2 *s_in[L], *s_sm[L], *s_sc[L];
  __shared__ float2 *s_out[L];
  for (int t = threadIdx.x; t < n; t += kFftThreads) tw[t] = a.tw[t];
  const int64_t line0 = (int64_t)blockIdx.x * L;
  if (threadIdx.x < L) {
    const int64_t l = line0 + threadIdx.x;
    if (l < a.lines) {
      if (MODE == ROW_FWD_FIRST) {
        const int64_t bc = l / a.rows_per_img, row = l - bc * a.rows_per_img;
        const int64_t b = bc / a.C, c = bc - b * a.C;
        s_in[threadIdx.x] = a.image + ((b * a.Ci + (a.Ci == 1 ? 0 : c)) * a.rows_per_img + row) * a.n_in;
        s_sm[threadIdx.x] = a.smaps ? a.smaps + (((a.Bs == 1 ? 0 : b) * a.C + c) * a.rows_per_img + row) * a.n_in : nullptr;
        s_sc[threadIdx.x] = a.scaling ? a.scaling + row * a.n_in : nullptr;
      } else {
        s_in[threadIdx.x] = a.in + l * a.n_in;
        s_sm[threadIdx.x] = nullptr;
        s_sc[threadIdx.x] = a.scaling ? a.scaling + (l % a.rows_per_img) * a.n_out : nullptr;
      }
      s_out[threadIdx.x] = a.out + l * a.n_out;
    }
  }
  __syncthreads();
  const int line = threadIdx.x / TPL, jb = threadIdx.x - line * TPL;
  const bool line_on = line0 + line < a.lines;
  const float2 *pin = line_on ? s_in[line] : nullptr, *psm = line_on ? s_sm[line] : nullptr,
               *psc = line_on ? s_sc[line] : nullptr;
  float2 *pout = line_on ? s_out[line] : nullptr;
  const int n_in = a.n_in, n_out = a.n_out;
  const float scale = a.scale;

  auto gload = [&](int, int i) -> float2 {
    if (i >= n_in) return f2(0.f, 0.f);  // zero padding is never read
    float2 v = pin[i];
    if (MODE == ROW_FWD_FIRST) {
      if (psm) v = cmul2(v, psm[i]);
      if (psc) v = cmul2(v, psc[i]);
      v = f2(v.x * scale, v.y * scale);
    }
    return v;
  };
  auto gstore = [&](int, int i, float2 v) {
    if (i >= n_out) return;  // cropped outputs are never written
    if (MODE == ROW_PLAIN) {
      if (psc) v = cmul2(v, f2(psc[i].x, -psc[i].y));
      v = f2(v.x * scale, v.y * scale);
    }
    pout[i] = v;
  };

  if (a.st.n_stages == 0) {  // n == 1
    if (line_on && jb == 0) gstore(line, 0, gload(line, 0));
    return;
  }
  float2 *src = bufA, *dst = bufB;
  int Ns = 1;
  for (int s = 0; s < a.st.n_stages; ++s) {
    const int R = a.st.radix[s];
    const bool first = s == 0, last = s == a.st.n_stages - 1;
    auto sload = [&](int ln, int i) -> float2 { return src[ln * NP + fft_pad(i)]; };
    auto sstore = [&](int ln, int i, float2 v) { dst[ln * NP + fft_pad(i)] = v; };
    if (first && last) {
      B2N_FFT_RADIX_SWITCH(R, (fft_stage_lines<RR, INV>(n, Ns, tw, line, jb, TPL, line_on, gload, gstore)))
    } else if (first) {
      B2N_FFT_RADIX_SWITCH(R, (fft_stage_lines<RR, INV>(n, Ns, tw, line, jb, TPL, line_on, gload, sstore)))
    } else if (last) {
      B2N_FFT_RADIX_SWITCH(R, (fft_stage_lines<RR, INV>(n, Ns, tw, line, jb, TPL, line_on, sload, gstore)))
    } else {
      B2N_FFT_RADIX_SWITCH(R, (fft_stage_lines<RR, INV>(n, Ns, tw, line, jb, TPL, line_on, sload, sstore)))
    }
    __syncthreads();
    float2 *t = src;
    src = dst;
    dst = t;
    Ns *= R;
  }
}

// -----------------------------------------------------------------------------------------
// column pass: 8 adjacent columns per CTA (64-byte global segments), 32 threads per column
// -----------------------------------------------------------------------------------------
template <bool INV>
__global__ void __launch_bounds__(kColThreads, 2) k_fft_cols(ColArgs a) {
  extern __shared__ __align__(16) float2 fsm[];
  constexpr int LX = kColsPerCta, TPL = kColThreads / LX;
  const int n = a.st.n, NP = fft_pad(n) + 1;
  float2 *tw = fsm;          // [n]
  float2 *bufA = tw + n;     // [NP][LX]
  float2 *bufB = bufA + NP * LX;
  for (int t = threadIdx.x; t < n; t += kColThreads) tw[t] = a.tw[t];
  const int64_t xblocks = (a.X + LX - 1) / LX;
  const int64_t oa = blockIdx.x / xblocks;
  const int64_t x0 = (blockIdx.x - oa * xblocks) * LX;
  const int l = threadIdx.x & (LX - 1), jb = threadIdx.x / LX;  // column index fastest: 64-byte global segments
  const bool col_on = x0 + l < a.X;
  const int X = (int)a.X, n_in = a.n_in, n_out = a.n_out;
  const float2 *in = a.in + oa * a.n_in * a.X + x0 + l;
  float2 *out = a.out + oa * a.n_out * a.X + x0 + l;
  const float2 *mul = a.mul ? a.mul + (a.a_per_mul ? (oa / a.a_per_mul) * (int64_t)n * a.X : 0) + x0 + l : nullptr;
  const float scale = a.scale;
  __syncthreads();

  auto gload = [&](int, int i) -> float2 {
    if (i >= n_in) return f2(0.f, 0.f);
    float2 v = in[i * X];
    if (mul) v = cmul2(v, mul[i * X]);
    return v;
  };
  auto gstore = [&](int, int i, float2 v) {
    if (i < n_out) out[i * X] = f2(v.x * scale, v.y * scale);
  };
  if (a.st.n_stages == 0) {
    if (col_on && jb == 0) gstore(l, 0, gload(l, 0));
    return;
  }
  float2 *src = bufA, *dst = bufB;
  int Ns = 1;
  for (int s = 0; s < a.st.n_stages; ++s) {
    const int R = a.st.radix[s];
    const bool first = s == 0, last = s == a.st.n_stages - 1;
    auto sload = [&](int ln, int i) -> float2 { return src[fft_pad(i) * LX + ln]; };
    auto sstore = [&](int ln, int i, float2 v) { dst[fft_pad(i) * LX + ln] = v; };
    if (first && last) {
      B2N_FFT_RADIX_SWITCH_COLS(R, (fft_stage_lines<RR, INV>(n, Ns, tw, l, jb, TPL, col_on, gload, gstore)))
    } else if (first) {
      B2N_FFT_RADIX_SWITCH_COLS(R, (fft_stage_lines<RR, INV>(n, Ns, tw, l, jb, TPL, col_on, gload, sstore)))
    } else if (last) {
      B2N_FFT_RADIX_SWITCH_COLS(R, (fft_stage_lines<RR, INV>(n, Ns, tw, l, jb, TPL, col_on, sload, gstore)))
    } else {
      B2N_FFT_RADIX_SWITCH_COLS(R, (fft_stage_lines<RR, INV>(n, Ns, tw, l, jb, TPL, col_on, sload, sstore)))
    }
    __syncthreads();
    float2 *t = src;
    src = dst;
    dst = t;
    Ns *= R;
  }
}

// staged twiddle tables of the fast plans: entry e = exp(-2 pi i r k / period)
__global__ void k_fft_twiddles_staged(float2 *tw, int R0, int R1, int R2) {
  const int tw2 = (R1 - 1) * R0, count = tw2 + (R2 > 1 ? (R2 - 1) * R0 * R1 : 0);
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= count) return;
  int r, k, period;
  if (e < tw2) {
    r = e / R0 + 1;
    k = e % R0;
    period = R0 * R1;
  } else {
    const int f = e - tw2;
    r = f / (R0 * R1) + 1;
    k = f % (R0 * R1);
    period = R0 * R1 * R2;
  }
  double s, c;
  sincospi(-2.0 * (double)((int64_t)r * k % period) / (double)period, &s, &c);
  tw[e] = f2((float)c, (float)s);
}

// twiddle table exp(-2 pi i t / n) in double precision, rounded once
__global__ void k_fft_twiddles(float2 *tw, int n) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) {
    double s, c;
    sincospi(-2.0 * (double)t / (double)n, &s, &c);
    tw[t] = f2((float)c, (float)s);
  }
}

// ---- host side ------------------------------------------------------------------------------
static bool make_stages(int64_t n, FftStages *st, int max_pow2_bits = 4) {
  FftPlan p;
  if (n < 1 || n > kFftMaxN || !fft_factorize((int)n, &p, max_pow2_bits)) return false;
  st->n = p.n;
  st->n_stages = p.n_stages;
  for (int i = 0; i < p.n_stages; ++i) st->radix[i] = p.radix[i];
  return true;
}

// the run-time passes keep two ping-pong line buffers in shared memory: the column pass (8 columns) is the larger one
constexpr size_t kSmemLimit = 227 * 1024;
static bool runtime_pass_fits(int64_t n) {
  return sizeof(float2) * ((size_t)n + 2 * (size_t)(fft_pad((int)n) + 1) * kColsPerCta) <= kSmemLimit;
}

int g_fast_fft = 1;  // B2N_OPT_FAST_FFT: compile-time planned passes where a plan exists
int g_fft_stream = 1;  // B2N_OPT_FFT_STREAM: mask of the planned passes that run as streamed persistent kernels
int g_counters_early = 1;  // 0 (B2N_OPT_PDL = 3, for A/B): counters zeroed right before k_fft_rows_sense

// entry points of the compile-time planned passes, defined in b2n_fft_plans_*.cu
#define B2N_DECLARE_PLAN_X(N, R0, R1, R2, ...)                      \
  int fast_rows_fwd_##N(RowArgs &a, cudaStream_t st);                \
  int fast_rows_inv_##N(RowArgs &a, cudaStream_t st);                \
  int fast_cols_##N(bool inverse, ColArgs &a, cudaStream_t st);      \
  int fast_cols_toep_##N(ColArgs &a, cudaStream_t st);               \
  int fast_rows_sense_##N(RowArgs &a, int64_t B, cudaStream_t st);
B2N_FAST_PLANS(B2N_DECLARE_PLAN_X, 0)

#define B2N_CASE_ROWS_FWD(N, R0, R1, R2, ...) case N: return fast_rows_fwd_##N(a, st);
#define B2N_CASE_ROWS_INV(N, R0, R1, R2, ...) case N: return fast_rows_inv_##N(a, st);
#define B2N_CASE_COLS(N, R0, R1, R2, ...) case N: return fast_cols_##N(inverse, a, st);
#define B2N_CASE_SENSE(N, R0, R1, R2, ...) case N: return fast_rows_sense_##N(a, B, st);
#define B2N_CASE_TOEP(N, R0, R1, R2, ...) case N: return fast_cols_toep_##N(a, st);

// each returns -1 when the length has no compile-time plan (the caller takes the run-time / unfused route)
static int fast_rows(bool inverse, RowArgs &a, cudaStream_t st) {
  if (!g_fast_fft) return -1;
  if (inverse) {
    switch (a.st.n) { B2N_FAST_PLANS(B2N_CASE_ROWS_INV, 0) default: break; }
  } else {
    switch (a.st.n) { B2N_FAST_PLANS(B2N_CASE_ROWS_FWD, 0) default: break; }
  }
  return -1;
}
static int fast_cols(bool inverse, ColArgs &a, cudaStream_t st) {
  if (!g_fast_fft) return -1;
  switch (a.st.n) { B2N_FAST_PLANS(B2N_CASE_COLS, 0) default: break; }
  return -1;
}
static int fast_cols_toep(ColArgs &a, cudaStream_t st) {
  if (!g_fast_fft) return -1;
  switch (a.st.n) { B2N_FAST_PLANS(B2N_CASE_TOEP, 0) default: break; }
  return -1;
}
static int launch_rows_sense(RowArgs &a, int64_t B, cudaStream_t st) {
  if (!g_fast_fft) return -1;
  switch (a.st.n) { B2N_FAST_PLANS(B2N_CASE_SENSE, 0) default: break; }
  return -1;
}

static bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <bool INV, int MODE> static int launch_rows(RowArgs &a, cudaStream_t st) {
  if (a.lines <= 0) return 0;
  if ((INV && MODE == ROW_PLAIN) || (!INV && MODE == ROW_FWD_FIRST)) {  // the two combinations the fused passes use
    const int rc = fast_rows(INV, a, st);
    if (rc >= 0) return rc;
  }
  const int NP = fft_pad(a.st.n) + 1;
  const size_t smem = sizeof(float2) * ((size_t)a.st.n + 2 * (size_t)kRowsPerCta * NP);
  if (smem > kSmemLimit) return fail_arg(B2N_E_UNSUPPORTED, "FFT length %d is too long for the run-time row pass", a.st.n);
  auto kern = k_fft_rows<INV, MODE>;
  B2N_SMEM_OPT_IN(kern, smem);
  kern<<<(unsigned)ceil_div(a.lines, kRowsPerCta), kFftThreads, smem, st>>>(a);
  B2N_LAUNCH_OK("k_fft_rows");
  return 0;
}

template <bool INV> static int launch_cols(ColArgs &a, cudaStream_t st) {
  if (a.A <= 0 || a.X <= 0) return 0;
  if (a.X % 2 == 0 && aligned16(a.in) && aligned16(a.out) && aligned16(a.mul)) {
    const int rc = fast_cols(INV, a, st);
    if (rc >= 0) return rc;
  }
  const int NP = fft_pad(a.st.n) + 1;
  const size_t smem = sizeof(float2) * ((size_t)a.st.n + 2 * (size_t)NP * kColsPerCta);
  if (smem > kSmemLimit) return fail_arg(B2N_E_UNSUPPORTED, "FFT length %d is too long for the run-time column pass", a.st.n);
  auto kern = k_fft_cols<INV>;
  B2N_SMEM_OPT_IN(kern, smem);
  const int64_t blocks = a.A * ceil_div(a.X, kColsPerCta);
  kern<<<(unsigned)blocks, kColThreads, smem, st>>>(a);
  B2N_LAUNCH_OK("k_fft_cols");
  return 0;
}

struct FusedGeom {
  int ndim;
  int64_t N[3], K[3];  // image / grid sizes in dimension order (slowest first)
  int64_t B, C;
  FftStages st[3];      // row-pass factorisation (radices up to 16)
  FftStages st_col[3];  // column-pass factorisation (radices up to 8: 64 registers per thread at 512 threads)
  const float2 *tw[3];
};

static int make_fused_geom(int ndim, const int64_t *im_size, const int64_t *grid_size, int64_t B, int64_t C,
                           const void *const *twiddle_dev, FusedGeom *g) {
  if (ndim < 1 || ndim > 3 || !im_size || !grid_size) return fail_arg(B2N_E_ARG, "bad ndim/im_size/grid_size");
  if (B < 1 || C < 1) return fail_arg(B2N_E_ARG, "n_batch=%lld n_coils=%lld", (long long)B, (long long)C);
  g->ndim = ndim;
  g->B = B;
  g->C = C;
  for (int d = 0; d < ndim; ++d) {
    if (im_size[d] < 1 || grid_size[d] < im_size[d]) return fail_arg(B2N_E_ARG, "grid_size[%d] < im_size[%d]", d, d);
    g->N[d] = im_size[d];
    g->K[d] = grid_size[d];
    if (!make_stages(grid_size[d], &g->st[d]) || !make_stages(grid_size[d], &g->st_col[d], 3))
      return fail_arg(B2N_E_UNSUPPORTED, "FFT length %lld is not supported (prime factor > 13 or > %d)",
                      (long long)grid_size[d], kFftMaxN);
    if (twiddle_dev) {
      if (!twiddle_dev[d]) return fail_arg(B2N_E_ARG, "twiddle_dev[%d] is NULL", d);
      g->tw[d] = (const float2 *)twiddle_dev[d];
    }
  }
  if (B * C * grid_size[0] * (ndim > 1 ? grid_size[1] : 1) * (ndim > 2 ? grid_size[2] : 1) >= ((int64_t)1 << 31))
    return fail_arg(B2N_E_RANGE, "grid too large for the 32-bit indexing of the fused FFT passes");
  return 0;
}

// scratch: T1 [B*C][N0..N_{d-2}][K_last] (ndim > 1), T2 [B*C][N0][K1][K2] (3-D), T3 [B*C][prod N] (adjoint
// with smaps: cropped, un-combined image before the coil sum)
static void fused_work_layout(const FusedGeom &g, size_t *t1, size_t *t2, size_t *t3, size_t *cnt = nullptr) {
  size_t e1 = (size_t)g.B * g.C * g.K[g.ndim - 1];
  for (int d = 0; d < g.ndim - 1; ++d) e1 *= g.N[d];
  if (g.ndim == 1) e1 = 0;
  const size_t e2 = g.ndim == 3 ? (size_t)g.B * g.C * g.N[0] * g.K[1] * g.K[2] : 0;
  size_t e3 = (size_t)g.B * g.C;
  for (int d = 0; d < g.ndim; ++d) e3 *= g.N[d];
  *t1 = e1;
  *t2 = e2;
  *t3 = e3;
  // per-image-row arrival counters of k_fft_rows_sense (4-byte, in float2 slots), behind T3
  size_t rows = (size_t)g.B;
  for (int d = 0; d < g.ndim - 1; ++d) rows *= g.N[d];
  if (cnt) *cnt = (rows + 1) / 2;
}

static int fused_forward(const FusedGeom &g, const float2 *image, int64_t Ci, const float2 *smaps, int64_t Bs,
                         const float2 *scaling, float scale, float2 *grid, float2 *work, cudaStream_t st) {
  const int d = g.ndim;
  size_t t1e, t2e, t3e;
  fused_work_layout(g, &t1e, &t2e, &t3e);
  float2 *T1 = work, *T2 = work + t1e;
  RowArgs r;
  memset(&r, 0, sizeof(r));
  r.st = g.st[d - 1];
  r.tw = g.tw[d - 1];
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
    c.st = g.st_col[0];
    c.tw = g.tw[0];
    c.n_in = (int)g.N[0];
    c.n_out = (int)g.K[0];
    c.A = g.B * g.C;
    c.X = g.K[1];
    c.in = T1;
    c.out = grid;
    return launch_cols<false>(c, st);
  }
  c.st = g.st_col[1];  // 3-D: y pass on the N0 non-zero planes, then z pass
  c.tw = g.tw[1];
  c.n_in = (int)g.N[1];
  c.n_out = (int)g.K[1];
  c.A = g.B * g.C * g.N[0];
  c.X = g.K[2];
  c.in = T1;
  c.out = T2;
  rc = launch_cols<false>(c, st);
  if (rc) return rc;
  c.st = g.st_col[0];
  c.tw = g.tw[0];
  c.n_in = (int)g.N[0];
  c.n_out = (int)g.K[0];
  c.A = g.B * g.C;
  c.X = g.K[1] * g.K[2];
  c.in = T2;
  c.out = grid;
  return launch_cols<false>(c, st);
}

int crop_apod_coilsum_c64(int ndim, const int64_t *im_size, const int64_t *grid_size, int64_t B, int64_t C,
                          const void *grid, const void *smaps, int64_t Bs, const void *scaling, double scale,
                          void *image, cudaStream_t st);  // b2n_fftops.cu

// The arrival counters of k_fft_rows_sense are zeroed at the START of a fused sequence, not right before that kernel:
// a memset between two kernels would break the programmatic dependent launch of the second one.
static int zero_sense_counters(const FusedGeom &g, float2 *T3, size_t t3e, cudaStream_t st) {
  if (!g_counters_early) return 0;
  size_t rows = (size_t)g.B;
  for (int d = 0; d < g.ndim - 1; ++d) rows *= g.N[d];
  B2N_CUDA_OK(cudaMemsetAsync(T3 + t3e, 0, sizeof(unsigned int) * rows, st));
  return 0;
}

// last pass of the adjoint: inverse transform of the contiguous dimension of rows_in [B*C][N0..N_{d-2}][K_last],
// crop, * conj(scaling) * scale, and the SENSE coil combination when smaps is given
// `peer` (optional): sum all-reduce of the output image over peer memory -- inside the row pass where it can carry
// the exchange (k_fft_rows_sense with its whole grid resident), as one more kernel behind the last pass otherwise
static int adjoint_rows_local(const FusedGeom &g, const float2 *rows_in, const float2 *smaps, int64_t Bs,
                              const float2 *scaling, float scale, float2 *image, float2 *T3, size_t t3e,
                              cudaStream_t st, const b2n_peer_comm *peer, int *peer_fused);

static int adjoint_rows(const FusedGeom &g, const float2 *rows_in, const float2 *smaps, int64_t Bs,
                        const float2 *scaling, float scale, float2 *image, float2 *T3, size_t t3e, cudaStream_t st,
                        const b2n_peer_comm *peer = nullptr) {
  int fused = 0;
  const int rc = adjoint_rows_local(g, rows_in, smaps, Bs, scaling, scale, image, T3, t3e, st, peer, &fused);
  if (rc || !peer || peer->world <= 1 || fused) return rc;
  int64_t n = 2 * g.B * (smaps ? 1 : g.C);
  for (int k = 0; k < g.ndim; ++k) n *= g.N[k];
  return peer_allreduce_launch(peer, image, image, n, st);
}

static int adjoint_rows_local(const FusedGeom &g, const float2 *rows_in, const float2 *smaps, int64_t Bs,
                              const float2 *scaling, float scale, float2 *image, float2 *T3, size_t t3e,
                              cudaStream_t st, const b2n_peer_comm *peer, int *peer_fused) {
  const int d = g.ndim;
  RowArgs r;
  memset(&r, 0, sizeof(r));
  r.st = g.st[d - 1];
  r.tw = g.tw[d - 1];
  r.n_in = (int)g.K[d - 1];
  r.n_out = (int)g.N[d - 1];
  r.rows_per_img = 1;
  for (int k = 0; k < d - 1; ++k) r.rows_per_img *= g.N[k];
  r.lines = g.B * g.C * r.rows_per_img;
  r.C = (int)g.C;
  r.in = rows_in;
  if (!smaps) {  // no coil combination: crop * conj(scaling) * scale straight to the caller's image
    r.out = image;
    r.scaling = scaling;
    r.scale = scale;
    return launch_rows<true, ROW_PLAIN>(r, st);
  }
  // SENSE, compile-time planned length: coil combination fused into the row pass (T3 holds the partial rows)
  r.out = image;
  r.smaps = smaps;
  r.Bs = (int)Bs;
  r.scaling = scaling;
  r.scale = scale;
  r.partial = T3;
  r.counter = reinterpret_cast<unsigned int *>(T3 + t3e);
  if (peer && peer->world > 1) {
    const int rc = peer_args_from_comm(peer, &r.peer);
    if (rc) return rc;
  }
  {
    const int rc = launch_rows_sense(r, g.B, st);
    if (rc >= 0) {
      *peer_fused = r.peer_fused;
      return rc;
    }
  }
  r.peer.world = 0;
  // otherwise: cropped per-coil rows to scratch, then one pass multiplies conj(smaps) * conj(scaling)
  // and sums the coils (every line of the row pass stays independent: full parallelism)
  r.smaps = nullptr;
  r.scaling = nullptr;
  r.out = T3;
  r.scale = 1.f;
  int rc = launch_rows<true, ROW_PLAIN>(r, st);
  if (rc) return rc;
  return crop_apod_coilsum_c64(d, g.N, g.N, g.B, g.C, T3, smaps, Bs, scaling, scale, image, st);
}

// Toeplitz normal operator in three passes (2-D, compile-time planned column length, grid >= 2 x image):
//   rows: image * smaps -> FFT_x (pruned inputs)                         -> T1 [B*C][N0][K1]
//   cols: FFT_y (pruned inputs) * kernel -> IFFT_y (cropped outputs), in place on T1 (k_fft_cols_toep)
//   rows: IFFT_x (cropped outputs) * conj(smaps), coil sum               -> image
// Returns -1 when the shape does not qualify (the caller runs fused_forward + fused_adjoint instead).
static int fused_toeplitz(const FusedGeom &g, const float2 *image, int64_t Ci, const float2 *smaps, int64_t Bs,
                          const float2 *kernel, int64_t kernel_batch, float scale, float2 *out, float2 *work,
                          cudaStream_t st) {
  if (g.ndim != 2 || !g_fast_fft || 2 * g.N[0] > g.K[0] || g.K[1] % 2) return -1;
  if (b2n_fft_supported(g.K[0]) != 2) return -1;
  size_t t1e, t2e, t3e;
  fused_work_layout(g, &t1e, &t2e, &t3e);
  float2 *T1 = work, *T3 = work + t1e + t2e;
  if (!aligned16(T1) || !aligned16(kernel)) return -1;
  ColArgs c;
  memset(&c, 0, sizeof(c));
  c.st = g.st_col[0];
  c.tw = g.tw[0];
  c.n_in = (int)g.N[0];
  c.n_out = (int)g.N[0];
  c.A = g.B * g.C;
  c.X = g.K[1];
  c.in = T1;
  c.out = T1;
  c.mul = kernel;
  c.a_per_mul = kernel_batch > 1 ? g.C : 0;
  c.scale = 1.f;
  {  // probe the column kernel's limits (shared memory) before anything is launched
    const size_t need = sizeof(float4) * 4 * (size_t)(2 * g.K[0] + (g.K[0] >> 3) + 1);
    if (need > kSmemLimit) return -1;
  }
  RowArgs r;
  memset(&r, 0, sizeof(r));
  r.st = g.st[1];
  r.tw = g.tw[1];
  r.n_in = (int)g.N[1];
  r.n_out = (int)g.K[1];
  r.rows_per_img = g.N[0];
  r.lines = g.B * g.C * r.rows_per_img;
  r.C = (int)g.C;
  r.Ci = (int)Ci;
  r.Bs = (int)Bs;
  r.image = image;
  r.smaps = smaps;
  r.scale = 1.f;
  r.out = T1;
  int rc = smaps ? zero_sense_counters(g, T3, t3e, st) : 0;
  if (rc) return rc;
  rc = launch_rows<false, ROW_FWD_FIRST>(r, st);
  if (rc) return rc;
  rc = fast_cols_toep(c, st);
  if (rc) return rc < 0 ? fail_arg(B2N_E_UNSUPPORTED, "Toeplitz column pass refused length %d", c.st.n) : rc;
  return adjoint_rows(g, T1, smaps, Bs, nullptr, scale, out, T3, t3e, st);
}

// kernel (optional): Toeplitz factor multiplied into the loads of the first inverse pass
static int fused_adjoint(const FusedGeom &g, const float2 *grid, const float2 *kernel, int64_t kernel_batch,
                         const float2 *smaps, int64_t Bs, const float2 *scaling, float scale, float2 *image,
                         float2 *work, cudaStream_t st, const b2n_peer_comm *peer = nullptr) {
  const int d = g.ndim;
  size_t t1e, t2e, t3e;
  fused_work_layout(g, &t1e, &t2e, &t3e);
  float2 *T1 = work, *T2 = work + t1e, *T3 = work + t1e + t2e;
  const float2 *rows_in = grid;
  if (smaps) {
    const int rc = zero_sense_counters(g, T3, t3e, st);
    if (rc) return rc;
  }
  ColArgs c;
  memset(&c, 0, sizeof(c));
  c.scale = 1.f;
  if (d == 3) {
    c.st = g.st_col[0];
    c.tw = g.tw[0];
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
    c.st = g.st_col[1];
    c.tw = g.tw[1];
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
    c.st = g.st_col[0];
    c.tw = g.tw[0];
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
  return adjoint_rows(g, rows_in, smaps, Bs, scaling, scale, image, T3, t3e, st, peer);
}

}  // namespace b2n

using namespace b2n;

extern "C" int b2n_fft_supported(int64_t n) {
  FftStages st;
  if (!make_stages(n, &st)) return 0;
  B2N_FAST_PLAN_SWITCH((int)n, return 2, (void)0)
  return runtime_pass_fits(n) ? 1 : 0;
}

extern "C" int b2n_fft_twiddles(int64_t n, void *twiddle_dev, void *stream) {
  if (n < 1 || n > kFftMaxN || !twiddle_dev) return fail_arg(B2N_E_ARG, "bad twiddle request");
  k_fft_twiddles<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>((float2 *)twiddle_dev, (int)n);
  B2N_LAUNCH_OK("k_fft_twiddles");
  // second half: per-stage tables of the compile-time plan for this length, when there is one
  B2N_FAST_PLAN_SWITCH((int)n,
                       (k_fft_twiddles_staged<<<(unsigned)ceil_div(P::TW_COUNT, 256), 256, 0, (cudaStream_t)stream>>>(
                           (float2 *)twiddle_dev + n, P::R0, P::R1, P::R2)),
                       (void)0)
  B2N_LAUNCH_OK("k_fft_twiddles_staged");
  return 0;
}

extern "C" int b2n_fft_work_bytes(int ndim, const int64_t *im_size, const int64_t *grid_size, int64_t n_batch,
                                  int64_t n_coils, size_t *bytes) {
  FusedGeom g;
  int rc = make_fused_geom(ndim, im_size, grid_size, n_batch, n_coils, nullptr, &g);
  if (rc) return rc;
  if (!bytes) return fail_arg(B2N_E_ARG, "bytes is NULL");
  size_t t1, t2, t3, cnt;
  fused_work_layout(g, &t1, &t2, &t3, &cnt);
  *bytes = sizeof(float2) * (t1 + t2 + t3 + cnt);
  return 0;
}

extern "C" int b2n_fft_forward_fused(int ndim, const int64_t *im_size, const int64_t *grid_size, int64_t n_batch,
                                     int64_t n_coils, const void *image_dev, int64_t image_coils,
                                     const void *smaps_dev, int64_t smaps_batch, const void *scaling_dev, double scale,
                                     const void *const *twiddle_dev, void *grid_dev, void *work_dev, void *stream) {
  FusedGeom g;
  if (!twiddle_dev) return fail_arg(B2N_E_ARG, "twiddle_dev is NULL");
  int rc = make_fused_geom(ndim, im_size, grid_size, n_batch, n_coils, twiddle_dev, &g);
  if (rc) return rc;
  if (!image_dev || !grid_dev || (ndim > 1 && !work_dev)) return fail_arg(B2N_E_ARG, "NULL image/grid/work");
  if (image_coils != 1 && image_coils != n_coils) return fail_arg(B2N_E_ARG, "image_coils must be 1 or n_coils");
  if (smaps_dev && smaps_batch != 1 && smaps_batch != n_batch) return fail_arg(B2N_E_ARG, "smaps_batch must be 1 or n_batch");
  return fused_forward(g, (const float2 *)image_dev, image_coils, (const float2 *)smaps_dev, smaps_dev ? smaps_batch : 1,
                       (const float2 *)scaling_dev, (float)scale, (float2 *)grid_dev, (float2 *)work_dev,
                       (cudaStream_t)stream);
}

extern "C" int b2n_fft_toeplitz_fused(int ndim, const int64_t *im_size, const int64_t *grid_size, int64_t n_batch,
                                      int64_t n_coils, const void *image_dev, int64_t image_coils,
                                      const void *smaps_dev, int64_t smaps_batch, const void *kernel_dev,
                                      int64_t kernel_batch, double scale, const void *const *twiddle_dev,
                                      void *out_dev, void *work_dev, void *stream) {
  FusedGeom g;
  if (!twiddle_dev) return fail_arg(B2N_E_ARG, "twiddle_dev is NULL");
  int rc = make_fused_geom(ndim, im_size, grid_size, n_batch, n_coils, twiddle_dev, &g);
  if (rc) return rc;
  if (!image_dev || !kernel_dev || !out_dev || !work_dev) return fail_arg(B2N_E_ARG, "NULL image/kernel/out/work");
  if (image_coils != 1 && image_coils != n_coils) return fail_arg(B2N_E_ARG, "image_coils must be 1 or n_coils");
  if (smaps_dev && smaps_batch != 1 && smaps_batch != n_batch) return fail_arg(B2N_E_ARG, "smaps_batch must be 1 or n_batch");
  if (kernel_batch != 1 && kernel_batch != n_batch) return fail_arg(B2N_E_ARG, "kernel_batch must be 1 or n_batch");
  rc = fused_toeplitz(g, (const float2 *)image_dev, image_coils, (const float2 *)smaps_dev, smaps_dev ? smaps_batch : 1,
                      (const float2 *)kernel_dev, kernel_batch, (float)scale, (float2 *)out_dev, (float2 *)work_dev,
                      (cudaStream_t)stream);
  if (rc < 0)
    return fail_arg(B2N_E_UNSUPPORTED, "no three-pass Toeplitz route for this shape (2-D, planned column length, "
                                       "grid >= 2 x image): use b2n_fft_forward_fused + b2n_fft_adjoint_fused");
  return rc;
}

extern "C" int b2n_fft_adjoint_fused(int ndim, const int64_t *im_size, const int64_t *grid_size, int64_t n_batch,
                                     int64_t n_coils, const void *grid_dev, const void *kernel_dev,
                                     int64_t kernel_batch, const void *smaps_dev, int64_t smaps_batch,
                                     const void *scaling_dev, double scale, const void *const *twiddle_dev,
                                     void *image_dev, void *work_dev, void *stream) {
  FusedGeom g;
  if (!twiddle_dev) return fail_arg(B2N_E_ARG, "twiddle_dev is NULL");
  int rc = make_fused_geom(ndim, im_size, grid_size, n_batch, n_coils, twiddle_dev, &g);
  if (rc) return rc;
  if (!image_dev || !grid_dev || !work_dev) return fail_arg(B2N_E_ARG, "NULL image/grid/work");
  if (smaps_dev && smaps_batch != 1 && smaps_batch != n_batch) return fail_arg(B2N_E_ARG, "smaps_batch must be 1 or n_batch");
  if (kernel_dev && kernel_batch != 1 && kernel_batch != n_batch) return fail_arg(B2N_E_ARG, "kernel_batch must be 1 or n_batch");
  return fused_adjoint(g, (const float2 *)grid_dev, (const float2 *)kernel_dev, kernel_dev ? kernel_batch : 1,
                       (const float2 *)smaps_dev, smaps_dev ? smaps_batch : 1, (const float2 *)scaling_dev, (float)scale,
                       (float2 *)image_dev, (float2 *)work_dev, (cudaStream_t)stream);
}

extern "C" int b2n_fft_adjoint_fused_allreduce(int ndim, const int64_t *im_size, const int64_t *grid_size,
                                               int64_t n_batch, int64_t n_coils, const void *grid_dev,
                                               const void *kernel_dev, int64_t kernel_batch, const void *smaps_dev,
                                               int64_t smaps_batch, const void *scaling_dev, double scale,
                                               const void *const *twiddle_dev, void *image_dev, void *work_dev,
                                               const b2n_peer_comm *comm, void *stream) {
  FusedGeom g;
  if (!twiddle_dev) return fail_arg(B2N_E_ARG, "twiddle_dev is NULL");
  if (!comm) return fail_arg(B2N_E_ARG, "comm is NULL");
  int rc = make_fused_geom(ndim, im_size, grid_size, n_batch, n_coils, twiddle_dev, &g);
  if (rc) return rc;
  if (!image_dev || !grid_dev || !work_dev) return fail_arg(B2N_E_ARG, "NULL image/grid/work");
  if (smaps_dev && smaps_batch != 1 && smaps_batch != n_batch) return fail_arg(B2N_E_ARG, "smaps_batch must be 1 or n_batch");
  if (kernel_dev && kernel_batch != 1 && kernel_batch != n_batch) return fail_arg(B2N_E_ARG, "kernel_batch must be 1 or n_batch");
  int64_t n = 2 * n_batch * (smaps_dev ? 1 : n_coils);
  for (int k = 0; k < ndim; ++k) n *= im_size[k];
  if (n > comm->max_floats) return fail_arg(B2N_E_RANGE, "image of %lld floats, window sized for %lld", (long long)n, (long long)comm->max_floats);
  return fused_adjoint(g, (const float2 *)grid_dev, (const float2 *)kernel_dev, kernel_dev ? kernel_batch : 1,
                       (const float2 *)smaps_dev, smaps_dev ? smaps_batch : 1, (const float2 *)scaling_dev, (float)scale,
                       (float2 *)image_dev, (float2 *)work_dev, (cudaStream_t)stream, comm);
}
