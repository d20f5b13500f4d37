// b2n_fft_fast_kernels.cuh -- kernels and launchers of the compile-time planned FFT passes (b2n_fft_fast.cuh).
// Included by the b2n_fft_plans_*.cu translation units, each of which instantiates a few plans (they compile in
// parallel); b2n_fft.cu only sees the per-plan entry points declared by B2N_DECLARE_PLAN.
#pragma once
#include "b2n_fft_args.cuh"
#include "b2n_fft_fast.cuh"
#include "b2n_tiled_common.cuh"

namespace b2n {

// -----------------------------------------------------------------------------------------
// fast passes: compile-time plans of b2n_fft_fast.cuh, two lines per thread
// -----------------------------------------------------------------------------------------
template <class P> struct FastCfg {
  // row pass: LP line pairs per CTA (about 160 threads: small CTAs keep the last wave of a launch short -- a
  // launch is typically 1-3 waves of resident line pairs); column pass: PAIRS column pairs per CTA
#ifndef B2N_FFT_ROW_TARGET
#define B2N_FFT_ROW_TARGET 160  // threads per row-pass CTA (A/B: profiles/scripts/fft_cfg_ab.sh)
#endif
#ifndef B2N_FFT_COL_PAIRS
#define B2N_FFT_COL_PAIRS 4     // column pairs per column-pass CTA for 64 <= T < 256
#endif
  static constexpr int LP = P::T >= B2N_FFT_ROW_TARGET ? 1 : B2N_FFT_ROW_TARGET / P::T;
  static constexpr int ROW_THREADS = LP * P::T;
  // row pass + coil sum: twice as many line pairs per CTA, i.e. fewer coil groups to meet through global memory
  static constexpr int LPS = P::T >= 160 ? 1 : 320 / P::T;
  static constexpr int SENSE_THREADS = LPS * P::T;
  static constexpr int PAIRS = P::T >= 256 ? 2 : (P::T >= 64 ? B2N_FFT_COL_PAIRS : 256 / P::T);
  static constexpr int COL_THREADS = PAIRS * P::T;
  // register budget: 64 per thread (128 for radix-16 butterflies on pairs) -> resident CTAs per SM
  static constexpr int REG_THREADS = P::RMAX >= 15 ? 512 : 1024;
  static constexpr int ROW_MINB = REG_THREADS / ROW_THREADS > 0 ? REG_THREADS / ROW_THREADS : 1;
  static constexpr int SENSE_MINB = REG_THREADS / SENSE_THREADS > 0 ? REG_THREADS / SENSE_THREADS : 1;
  static constexpr int COL_MINB = REG_THREADS / COL_THREADS > 0 ? REG_THREADS / COL_THREADS : 1;
};

B2N_D float2 row_operand(const float2 *in, const float2 *sm, const float2 *sc, int i, float scale) {
  float2 v = in[i];
  if (sm) v = cmul2(v, sm[i]);
  if (sc) v = cmul2(v, sc[i]);
  return f2(v.x * scale, v.y * scale);
}

// HALF: the padded half of the inputs (forward) / the cropped half of the outputs (inverse) is skipped at
// compile time (n_in <= N/2 resp. n_out <= N/2, the 2x-oversampled case).
template <class P, bool INV, int MODE, bool HALF>
__global__ void __launch_bounds__(FastCfg<P>::ROW_THREADS, FastCfg<P>::ROW_MINB) k_fft_rows_fast(RowArgs a) {
  extern __shared__ __align__(16) float4 fsm4[];
  griddep_launch();
  griddep_wait();  // inputs may come from the preceding kernel
  constexpr int LP = FastCfg<P>::LP;
  const int lp = threadIdx.x / P::T, t = threadIdx.x - lp * P::T;
  const int64_t lA = ((int64_t)blockIdx.x * LP + lp) * 2;
  const bool onA = lA < a.lines, onB = lA + 1 < a.lines;
  const float2 *inA = nullptr, *inB = nullptr, *smA = nullptr, *smB = nullptr, *scA = nullptr, *scB = nullptr;
  float2 *outA = nullptr, *outB = nullptr;
  // 32-bit index arithmetic: the launchers guarantee every array here has < 2^31 elements
  auto setup = [&](uint32_t l, const float2 *&in, const float2 *&sm, const float2 *&sc, float2 *&out) {
    const uint32_t rpi = (uint32_t)a.rows_per_img, n_in = (uint32_t)a.n_in, n_out = (uint32_t)a.n_out;
    const uint32_t bc = l / rpi, row = l - bc * rpi;
    if (MODE == ROW_FWD_FIRST) {
      const uint32_t C = (uint32_t)a.C, b = bc / C, c = bc - b * C;
      in = a.image + ((b * (uint32_t)a.Ci + (a.Ci == 1 ? 0u : c)) * rpi + row) * n_in;
      sm = a.smaps ? a.smaps + (((a.Bs == 1 ? 0u : b) * C + c) * rpi + row) * n_in : nullptr;
      sc = a.scaling ? a.scaling + row * n_in : nullptr;
    } else {
      in = a.in + l * n_in;
      sc = a.scaling ? a.scaling + row * n_out : nullptr;
    }
    out = a.out + l * n_out;
  };
  if (onA) setup((uint32_t)lA, inA, smA, scA, outA);
  if (onB) setup((uint32_t)lA + 1, inB, smB, scB, outB);
  const int n_in = a.n_in, n_out = a.n_out;
  const float scale = a.scale;
  if (a.prefetch && t < 2) {  // the operand rows of the CTA that takes this one's place in the next wave -> L2
    const int64_t lF = lA + (int64_t)a.prefetch * LP * 2 + t;
    if (lF < a.lines) {
      const float2 *in = nullptr, *sm = nullptr, *sc = nullptr;
      float2 *out = nullptr;
      setup((uint32_t)lF, in, sm, sc, out);
      const unsigned bytes = (unsigned)n_in * 8u;
      if (!(bytes & 15u) && !(reinterpret_cast<uintptr_t>(in) & 15)) prefetch_l2_bulk(in, bytes);
      if (sm && !(bytes & 15u) && !(reinterpret_cast<uintptr_t>(sm) & 15)) prefetch_l2_bulk(sm, bytes);
    }
  }
  auto loadg = [&](int i, int) -> float4 {
    float4 v = fast::v4(0.f, 0.f, 0.f, 0.f);
    if (i < n_in) {  // zero padding is never read
      if (MODE == ROW_FWD_FIRST) {
        if (onA) { const float2 x = row_operand(inA, smA, scA, i, scale); v.x = x.x; v.y = x.y; }
        if (onB) { const float2 x = row_operand(inB, smB, scB, i, scale); v.z = x.x; v.w = x.y; }
      } else {
        if (onA) { const float2 x = inA[i]; v.x = x.x; v.y = x.y; }
        if (onB) { const float2 x = inB[i]; v.z = x.x; v.w = x.y; }
      }
    }
    return v;
  };
  auto storeg = [&](int i, float4 v) {
    if (i >= n_out) return;  // cropped outputs are never written
    if (MODE == ROW_PLAIN) {
      if (onA) {
        float2 x = f2(v.x, v.y);
        if (scA) x = cmul2(x, f2(scA[i].x, -scA[i].y));
        outA[i] = f2(x.x * scale, x.y * scale);
      }
      if (onB) {
        float2 x = f2(v.z, v.w);
        if (scB) x = cmul2(x, f2(scB[i].x, -scB[i].y));
        outB[i] = f2(x.x * scale, x.y * scale);
      }
    } else {
      if (onA) outA[i] = f2(v.x, v.y);
      if (onB) outB[i] = f2(v.z, v.w);
    }
  };
  fast::fft_line_pair<P, INV, HALF && !INV, HALF && INV>(t, fsm4 + lp * P::NP, 1, a.tw + P::N, loadg, storeg);
}

template <class P, bool INV, bool HALF>
__global__ void __launch_bounds__(FastCfg<P>::COL_THREADS, FastCfg<P>::COL_MINB) k_fft_cols_fast(ColArgs a) {
  extern __shared__ __align__(16) float4 fsm4[];
  griddep_launch();
  griddep_wait();  // the input rows come from the preceding pass
  constexpr int PAIRS = FastCfg<P>::PAIRS;
  const int p = threadIdx.x % PAIRS, t = threadIdx.x / PAIRS;  // pair index fastest: contiguous global segments
  const int X = (int)a.X, X2 = X >> 1, n_in = a.n_in, n_out = a.n_out;
  // grid: x = block of 2*PAIRS columns, (y, z) = outer index
  const int64_t oa = (int64_t)blockIdx.z * gridDim.y + blockIdx.y;
  const int x = ((int)blockIdx.x * PAIRS + p) * 2;
  const bool on = x < X && oa < a.A;
  const float4 *in = reinterpret_cast<const float4 *>(a.in + oa * n_in * a.X + x);
  float4 *out = reinterpret_cast<float4 *>(a.out + oa * n_out * a.X + x);
  const float4 *mul =
      a.mul ? reinterpret_cast<const float4 *>(a.mul + (a.a_per_mul ? (oa / a.a_per_mul) * (int64_t)P::N * a.X : 0) + x)
            : nullptr;
  const float scale = a.scale;
  if (a.prefetch) {  // the input columns of the CTA that takes this one's place in the next wave -> L2
    const int64_t lin = oa * gridDim.x + blockIdx.x + a.prefetch;
    const int64_t oaF = lin / gridDim.x;
    const int xF = (int)(lin - oaF * gridDim.x) * PAIRS * 2;
    if (oaF < a.A) {
      const float2 *base = a.in + oaF * n_in * a.X + xF;
      for (int i = threadIdx.x; i < n_in; i += FastCfg<P>::COL_THREADS) prefetch_l2(base + (int64_t)i * X);
    }
  }
  auto loadg = [&](int i, int) -> float4 {
    if (!on || i >= n_in) return fast::v4(0.f, 0.f, 0.f, 0.f);
    float4 v = in[i * X2];
    if (mul) v = fast::vmul2(v, mul[i * X2]);
    return v;
  };
  auto storeg = [&](int i, float4 v) {
    if (on && i < n_out) out[i * X2] = fast::vscale(v, scale);
  };
  fast::fft_line_pair<P, INV, HALF && !INV, HALF && INV>(t, fsm4 + p, PAIRS, a.tw + P::N, loadg, storeg);
}

// Streamed column pass: a persistent CTA walks tiles (blocks of 2*PAIRS columns of one outer index) and keeps the
// first-stage operands of its NEXT tile in flight while it transforms the current one.  Every thread copies the
// operands of its own first butterfly (cp.async, 16 bytes each, zero-filled beyond n_in / X) into slots only it reads,
// so consuming them needs no barrier -- just cp.async.wait_group -- and no registers are held across the tile (a
// register prefetch cost a resident CTA: profiles/r02_fft_persist_ab.log).  The copies are issued behind the first
// barrier of the tile, when the previous operands have been consumed, and land during the remaining stages and stores.
template <class P, bool INV, bool HALF, int PAIRS_ = FastCfg<P>::PAIRS> struct StreamCfg {
  static constexpr int PAIRS = PAIRS_, NT = PAIRS_ * P::T;
  static constexpr int NL = (HALF && !INV) ? P::R0 / 2 : P::R0;  // first-stage operands per thread
  static constexpr size_t SMEM = sizeof(float4) * ((size_t)PAIRS * P::NP + (size_t)NL * NT);
  static constexpr int BY_SMEM = (int)((size_t)226 * 1024 / (SMEM + 1024));
  static constexpr int BY_REGS = FastCfg<P>::REG_THREADS / NT > 0 ? FastCfg<P>::REG_THREADS / NT : 1;
  static constexpr int MINB = BY_SMEM < 1 ? 1 : (BY_SMEM < BY_REGS ? BY_SMEM : BY_REGS);
};

template <class P, bool INV, bool HALF, int PAIRS_>
__global__ void __launch_bounds__(StreamCfg<P, INV, HALF, PAIRS_>::NT, StreamCfg<P, INV, HALF, PAIRS_>::MINB)
    k_fft_cols_stream(ColArgs a, uint32_t tiles_x, uint32_t tiles) {
  extern __shared__ __align__(16) float4 fsm4[];
  using S = StreamCfg<P, INV, HALF, PAIRS_>;
  constexpr int PAIRS = S::PAIRS, NT = S::NT, NL = S::NL;
  griddep_launch();
  griddep_wait();  // the input rows come from the preceding pass
  const int p = threadIdx.x % PAIRS, t = threadIdx.x / PAIRS;
  const int X = (int)a.X, X2 = X >> 1, n_in = a.n_in, n_out = a.n_out;
  float4 *slot = fsm4 + PAIRS * P::NP + threadIdx.x;  // leg r of this thread at slot[r * NT]
  const float scale = a.scale;
  auto issue = [&](uint32_t tile) {
    if (t < P::I0) {
      const uint32_t oa = tile / tiles_x, bx = tile - oa * tiles_x;
      const int x = ((int)bx * PAIRS + p) * 2;
      const bool on = x < X;
      const float4 *in = reinterpret_cast<const float4 *>(a.in + (int64_t)oa * n_in * a.X + (on ? x : 0));
#pragma unroll
      for (int r = 0; r < NL; ++r) {
        const int i = t + r * P::I0;
        const bool ok = on && i < n_in;
        cp_async16z(slot + r * NT, in + (ok ? i * X2 : 0), ok);
      }
    }
    cp_async_commit();
  };
  uint32_t tile = blockIdx.x;
  if (tile < tiles) issue(tile);
  for (; tile < tiles; tile += gridDim.x) {
    const uint32_t oa = tile / tiles_x, bx = tile - oa * tiles_x;
    const int x = ((int)bx * PAIRS + p) * 2;
    const bool on = x < X;
    float4 *out = reinterpret_cast<float4 *>(a.out + (int64_t)oa * n_out * a.X + (on ? x : 0));
    cp_async_wait_all();  // this thread's own operands have landed
    auto loadg = [&](int, int r) -> float4 { return slot[r * NT]; };
    auto storeg = [&](int i, float4 v) {
      if (on && i < n_out) out[i * X2] = fast::vscale(v, scale);
    };
    auto after0 = [&]() {
      if (tile + gridDim.x < tiles) issue(tile + gridDim.x);
    };
    fast::fft_line_pair<P, INV, HALF && !INV, HALF && INV>(t, fsm4 + p, PAIRS, a.tw + P::N, loadg, storeg, after0);
    __syncthreads();  // the exchange buffer is rewritten by the next tile's first stage
  }
}

// Toeplitz column pass: forward transform of the columns, multiply by the kernel spectrum, inverse transform, all
// inside the CTA -- the full spectrum [A][N][X] never exists in global memory.  In and out are the half-height
// arrays [A][n_in][X] / [A][n_out][X] (n_in, n_out <= N/2: the Toeplitz grid is twice the image), and may be the
// same buffer: a CTA reads all of its columns before it writes any of them.
template <class P>
__global__ void __launch_bounds__(FastCfg<P>::COL_THREADS, FastCfg<P>::COL_MINB) k_fft_cols_toep(ColArgs a) {
  extern __shared__ __align__(16) float4 fsm4[];
  griddep_launch();
  griddep_wait();  // the input rows come from the preceding pass
  constexpr int PAIRS = FastCfg<P>::PAIRS;
  const int p = threadIdx.x % PAIRS, t = threadIdx.x / PAIRS;
  const int X = (int)a.X, X2 = X >> 1, n_in = a.n_in, n_out = a.n_out;
  const int64_t oa = (int64_t)blockIdx.z * gridDim.y + blockIdx.y;
  const int x = ((int)blockIdx.x * PAIRS + p) * 2;
  const bool on = x < X && oa < a.A;
  const float4 *in = reinterpret_cast<const float4 *>(a.in + oa * n_in * a.X + x);
  float4 *out = reinterpret_cast<float4 *>(a.out + oa * n_out * a.X + x);
  const float4 *mul =
      reinterpret_cast<const float4 *>(a.mul + (a.a_per_mul ? (oa / a.a_per_mul) * (int64_t)P::N * a.X : 0) + x);
  const float scale = a.scale;
  if (a.prefetch) {  // the input columns of the CTA that takes this one's place in the next wave -> L2
    const int64_t lin = oa * gridDim.x + blockIdx.x + a.prefetch;
    const int64_t oaF = lin / gridDim.x;
    const int xF = (int)(lin - oaF * gridDim.x) * PAIRS * 2;
    if (oaF < a.A) {
      const float2 *base = a.in + oaF * n_in * a.X + xF;
      for (int i = threadIdx.x; i < n_in; i += FastCfg<P>::COL_THREADS) prefetch_l2(base + (int64_t)i * X);
    }
  }
  float4 *spec = fsm4 + PAIRS * P::NP;  // the filtered spectrum of this CTA's columns, natural order [N][PAIRS]
  auto load_in = [&](int i, int) -> float4 {
    return (!on || i >= n_in) ? fast::v4(0.f, 0.f, 0.f, 0.f) : in[i * X2];
  };
  auto store_spec = [&](int i, float4 v) {
    spec[i * PAIRS + p] = on ? fast::vmul2(v, mul[i * X2]) : fast::v4(0.f, 0.f, 0.f, 0.f);
  };
  fast::fft_line_pair<P, false, true, false>(t, fsm4 + p, PAIRS, a.tw + P::N, load_in, store_spec);
  __syncthreads();  // spectrum complete; the exchange buffer is free again
  auto load_spec = [&](int i, int) -> float4 { return spec[i * PAIRS + p]; };
  auto store_out = [&](int i, float4 v) {
    if (on && i < n_out) out[i * X2] = fast::vscale(v, scale);
  };
  fast::fft_line_pair<P, true, false, true>(t, fsm4 + p, PAIRS, a.tw + P::N, load_spec, store_out);
}

// Inverse row pass + SENSE coil combination (the last pass of the SENSE adjoint):
//   image[b, row, :] = scale * conj(scaling[row, :]) * sum_c conj(smaps[b, c, row, :]) * IFFT_x(in[b, c, row, :])[:n_out]
// CTA = one image row x 2*LP coils.  The LP pair sums meet in shared memory; when the coils span several
// CTAs each writes its partial row to scratch and the CTA that arrives last (one atomic ticket per row) adds
// the partial rows in coil-group order -- a fixed summation order, so the result is bit-reproducible.
template <class P, bool HALF>
__global__ void __launch_bounds__(FastCfg<P>::SENSE_THREADS, FastCfg<P>::SENSE_MINB) k_fft_rows_sense(RowArgs a) {
  extern __shared__ __align__(16) float4 fsm4[];
  griddep_launch();
  griddep_wait();  // the input rows come from the preceding pass
  constexpr int LP = FastCfg<P>::LPS, NT = FastCfg<P>::SENSE_THREADS;
  __shared__ int s_last;
  __shared__ uint32_t s_epoch, s_prev;
  const bool peer_on = a.peer.world > 1;  // the finished rows go through the peer-memory all-reduce
  if (peer_on && threadIdx.x == 0) {      // read behind griddep_wait: the preceding call has advanced the counter
    const uint32_t *hdr = reinterpret_cast<const uint32_t *>(a.peer.window[a.peer.rank]);
    const uint32_t e = *reinterpret_cast<const volatile uint32_t *>(hdr) + 1u;
    s_epoch = e;
    s_prev = *reinterpret_cast<const volatile uint32_t *>(hdr + 4 + (e + 2u) % 3u);
  }
  const int n_in = a.n_in, n_out = a.n_out;
  float2 *red = reinterpret_cast<float2 *>(fsm4 + LP * P::NP);  // [LP][n_out] pair sums
  const int lp = threadIdx.x / P::T, t = threadIdx.x - lp * P::T;
  const uint32_t G = (uint32_t)a.coil_groups, rpi = (uint32_t)a.rows_per_img, C = (uint32_t)a.C;
  const uint32_t br = blockIdx.x / G, g = blockIdx.x - br * G;  // (batch, row), coil group
  const uint32_t b = br / rpi, row = br - b * rpi;
  const uint32_t cA = (g * LP + lp) * 2, cB = cA + 1;
  const bool onA = cA < C, onB = cB < C;
  const uint32_t bs = a.Bs == 1 ? 0u : b;
  const float2 *inA = a.in + ((b * C + (onA ? cA : 0u)) * rpi + row) * (uint32_t)n_in;
  const float2 *inB = a.in + ((b * C + (onB ? cB : 0u)) * rpi + row) * (uint32_t)n_in;
  const float2 *smA = a.smaps + ((bs * C + (onA ? cA : 0u)) * rpi + row) * (uint32_t)n_out;
  const float2 *smB = a.smaps + ((bs * C + (onB ? cB : 0u)) * rpi + row) * (uint32_t)n_out;
  if (a.prefetch && t < 2 && blockIdx.x + (unsigned)a.prefetch < gridDim.x) {  // next wave's rows -> L2
    const uint32_t blkF = blockIdx.x + (uint32_t)a.prefetch, brF = blkF / G, gF = blkF - brF * G;
    const uint32_t bF = brF / rpi, rowF = brF - bF * rpi, cF = (gF * LP + lp) * 2 + t;
    if (cF < C) {
      const float2 *in = a.in + ((bF * C + cF) * rpi + rowF) * (uint32_t)n_in;
      const float2 *sm = a.smaps + (((a.Bs == 1 ? 0u : bF) * C + cF) * rpi + rowF) * (uint32_t)n_out;
      if (!(n_in & 1) && !(reinterpret_cast<uintptr_t>(in) & 15)) prefetch_l2_bulk(in, (unsigned)n_in * 8u);
      if (!(n_out & 1) && !(reinterpret_cast<uintptr_t>(sm) & 15)) prefetch_l2_bulk(sm, (unsigned)n_out * 8u);
    }
  }
  auto loadg = [&](int i, int) -> float4 {
    float4 v = fast::v4(0.f, 0.f, 0.f, 0.f);
    if (i < n_in) {
      if (onA) { const float2 x = inA[i]; v.x = x.x; v.y = x.y; }
      if (onB) { const float2 x = inB[i]; v.z = x.x; v.w = x.y; }
    }
    return v;
  };
  auto storeg = [&](int i, float4 v) {
    if (i >= n_out) return;
    float2 p = f2(0.f, 0.f);
    if (onA) { const float2 m = smA[i]; p = f2(fmaf(v.x, m.x, v.y * m.y), fmaf(v.y, m.x, -(v.x * m.y))); }
    if (onB) { const float2 m = smB[i]; p = f2(p.x + fmaf(v.z, m.x, v.w * m.y), p.y + fmaf(v.w, m.x, -(v.z * m.y))); }
    red[lp * n_out + i] = p;
  };
  fast::fft_line_pair<P, true, false, HALF>(t, fsm4 + lp * P::NP, 1, a.tw + P::N, loadg, storeg);
  __syncthreads();
  const float2 *sc = a.scaling ? a.scaling + row * (uint32_t)n_out : nullptr;
  float2 *out = a.out + br * (uint32_t)n_out;
  unsigned spins = 0;
  unsigned long long t0 = 0;
  auto finish = [&](int j, float2 sum) {
    if (sc) sum = cmul2(sum, f2(sc[j].x, -sc[j].y));
    float2 val = f2(sum.x * a.scale, sum.y * a.scale);
    // s_epoch was written before the barriers of the transform above
    if (peer_on) val = peer_exchange(a.peer, s_epoch, (int64_t)br * (uint32_t)n_out + j, val, spins, t0);
    out[j] = val;
  };
  // the CTA that finishes a row: after its last value, reset what an earlier, larger call left beyond this image
  // (row 0 only) and take a ticket; the last ticket advances the window's call counter
  auto peer_done = [&]() {
    if (!peer_on) return;
    const uint32_t epoch = s_epoch, gclr = (epoch + 2u) % 3u;
    float *mine_w = reinterpret_cast<float *>(a.peer.window[a.peer.rank] + a.peer.data_off);
    if (br == 0 && (int64_t)s_prev > a.peer_floats) {
      const float fill = __uint_as_float(kPeerFill);
      for (int r = 0; r < a.peer.world; ++r) {
        if (r == a.peer.rank) continue;
        float *dst = mine_w + ((size_t)gclr * a.peer.world + r) * a.peer.slot_floats;
        for (int64_t i = a.peer_floats + threadIdx.x; i < (int64_t)s_prev; i += NT) dst[i] = fill;
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t *hdr = reinterpret_cast<uint32_t *>(a.peer.window[a.peer.rank]);
      __threadfence();
      if (atomicAdd(hdr + 1, 1u) == gridDim.x / G - 1) {
        hdr[1] = 0;
        hdr[4 + epoch % 3u] = (uint32_t)a.peer_floats;
        __threadfence();
        *reinterpret_cast<volatile uint32_t *>(hdr) = epoch;
      }
    }
  };
  float2 *mine = a.partial + ((size_t)g * gridDim.x / G + br) * (uint32_t)n_out;  // [G][B*rows][n_out]
  for (int j = threadIdx.x; j < n_out; j += NT) {
    float2 sum = red[j];
#pragma unroll
    for (int q = 1; q < LP; ++q) sum = cadd(sum, red[q * n_out + j]);
    if (G == 1) finish(j, sum);
    else __stcg(&mine[j], sum);
  }
  if (G == 1) {
    peer_done();
    return;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(&a.counter[br], 1u) == G - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int j = threadIdx.x; j < n_out; j += NT) {
    float2 sum = __ldcg(&a.partial[(size_t)br * (uint32_t)n_out + j]);
    for (uint32_t q = 1; q < G; ++q) sum = cadd(sum, __ldcg(&a.partial[((size_t)q * gridDim.x / G + br) * (uint32_t)n_out + j]));
    finish(j, sum);
  }
  if (threadIdx.x == 0) a.counter[br] = 0;  // leave the counters zero for the next call
  peer_done();
}

// B2N_OPT_FFT_PREFETCH is a mask over the passes: 1 forward rows, 2 forward columns, 4 inverse columns, 8 inverse rows
// (+ coil sum), 16 Toeplitz columns; 32 lifts the size threshold.  A pass prefetches when its bit is set and it reads
// at least 32 MB (it then streams from HBM; smaller passes find their input in L2 and the prefetch only costs issue
// slots).  Default 19 = forward rows + forward columns + Toeplitz columns: -12 us on the 384^2 x 32-coil forward, -8 us
// on its Toeplitz apply; on the inverse passes the prefetch is neutral to slightly harmful (+2..6 us), so their bits
// are off (profiles/r01_h_opts_ab.log).
enum { PF_ROWS_FWD = 1, PF_COLS_FWD = 2, PF_COLS_INV = 4, PF_ROWS_INV = 8, PF_COLS_TOEP = 16, PF_ANY_SIZE = 32 };
static inline bool want_prefetch(size_t input_bytes, int pass) {
  return (g_prefetch & pass) && ((g_prefetch & PF_ANY_SIZE) || input_bytes >= ((size_t)32 << 20));
}

// CTAs of `kern` resident on the whole device at once (one wave).  Cached per kernel entry point and thread: kernels
// with the same signature share a function-pointer type, so the key is the pointer, not the type.
template <class K> static int resident_ctas(K kern, int threads, size_t smem) {
  struct Entry {
    const void *kern;
    int ctas;
  };
  static thread_local Entry cache[64];
  static thread_local int used = 0;
  const void *key = reinterpret_cast<const void *>(kern);
  for (int i = 0; i < used; ++i)
    if (cache[i].kern == key) return cache[i].ctas;
  int per_sm = 0, dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, key, threads, smem);
  const int ctas = per_sm > 0 && sms > 0 ? per_sm * sms : 1;
  if (used < 64) cache[used++] = {key, ctas};
  return ctas;
}

template <class P, bool INV, int MODE, bool HALF> int launch_rows_fast_h(RowArgs &a, cudaStream_t st) {
  using Cfg = FastCfg<P>;
  const size_t smem = sizeof(float4) * (size_t)Cfg::LP * P::NP;
  auto kern = k_fft_rows_fast<P, INV, MODE, HALF>;
  B2N_SMEM_OPT_IN(kern, smem);
  a.prefetch = want_prefetch(sizeof(float2) * (size_t)a.lines * a.n_in, INV ? PF_ROWS_INV : PF_ROWS_FWD) ? resident_ctas(kern, Cfg::ROW_THREADS, smem) : 0;
  B2N_CUDA_OK(launch_pdl(kern, dim3((unsigned)ceil_div(a.lines, 2 * Cfg::LP)), dim3(Cfg::ROW_THREADS), smem, st, a));
  B2N_LAUNCH_OK("k_fft_rows_fast");
  return 0;
}
template <class P, bool INV, int MODE> int launch_rows_fast(RowArgs &a, cudaStream_t st) {
  const bool half = 2 * (INV ? a.n_out : a.n_in) <= P::N;
  return half ? launch_rows_fast_h<P, INV, MODE, true>(a, st) : launch_rows_fast_h<P, INV, MODE, false>(a, st);
}

template <class P, bool INV, bool HALF> int launch_cols_fast_h(ColArgs &a, cudaStream_t st) {
  using Cfg = FastCfg<P>;
  const size_t smem = sizeof(float4) * (size_t)Cfg::PAIRS * P::NP;
  auto kern = k_fft_cols_fast<P, INV, HALF>;
  B2N_SMEM_OPT_IN(kern, smem);
  a.prefetch = want_prefetch(sizeof(float2) * (size_t)a.A * a.n_in * a.X, INV ? PF_COLS_INV : PF_COLS_FWD) ? resident_ctas(kern, Cfg::COL_THREADS, smem) : 0;
  const int64_t gy = a.A < 32768 ? a.A : 32768;
  const dim3 grid((unsigned)ceil_div(a.X, 2 * Cfg::PAIRS), (unsigned)gy, (unsigned)ceil_div(a.A, gy));
  B2N_CUDA_OK(launch_pdl(kern, grid, dim3(Cfg::COL_THREADS), smem, st, a));
  B2N_LAUNCH_OK("k_fft_cols_fast");
  return 0;
}
// B2N_OPT_FFT_STREAM: mask of the passes that run as streamed persistent kernels (1 forward columns, 2 inverse columns)
extern int g_fft_stream;
template <class P, bool INV, bool HALF, int PAIRS_> int launch_cols_stream_h(ColArgs &a, cudaStream_t st) {
  using S = StreamCfg<P, INV, HALF, PAIRS_>;
  auto kern = k_fft_cols_stream<P, INV, HALF, PAIRS_>;
  B2N_SMEM_OPT_IN(kern, S::SMEM);
  const int64_t tiles_x = ceil_div(a.X, 2 * S::PAIRS), tiles = tiles_x * a.A;
  const int resident = resident_ctas(kern, S::NT, S::SMEM);
  a.prefetch = 0;
  B2N_CUDA_OK(launch_pdl(kern, dim3((unsigned)(tiles < resident ? tiles : resident)), dim3(S::NT), S::SMEM, st, a,
                         (uint32_t)tiles_x, (uint32_t)tiles));
  B2N_LAUNCH_OK("k_fft_cols_stream");
  return 0;
}
template <class P, bool INV> int launch_cols_fast(ColArgs &a, cudaStream_t st) {
  const bool half = 2 * (INV ? a.n_out : a.n_in) <= P::N;
  // streamed variant: no fused multiply, 32-bit tile arithmetic, operand slots + exchange buffer within one CTA's share
  if ((g_fft_stream & (INV ? 2 : 1)) && !a.mul && half && !(a.X & 1) && ceil_div(a.X, 2 * FastCfg<P>::PAIRS) * a.A < ((int64_t)1 << 31) &&
      a.A * a.n_in * a.X < ((int64_t)1 << 31) && StreamCfg<P, INV, true>::SMEM <= (size_t)113 * 1024) {
    constexpr int PF = FastCfg<P>::PAIRS, PH = PF >= 2 ? PF / 2 : 1;
    if (g_fft_stream & (INV ? 32 : 16)) return launch_cols_stream_h<P, INV, true, PH>(a, st);  // A/B: narrower tiles
    return launch_cols_stream_h<P, INV, true, PF>(a, st);
  }
  return half ? launch_cols_fast_h<P, INV, true>(a, st) : launch_cols_fast_h<P, INV, false>(a, st);
}

template <class P, bool HALF> int launch_rows_sense_h(RowArgs &a, int64_t B, cudaStream_t st) {
  using Cfg = FastCfg<P>;
  a.coil_groups = (int)ceil_div(a.C, 2 * Cfg::LPS);
  const int64_t rows = B * a.rows_per_img;
  const size_t smem = sizeof(float4) * (size_t)Cfg::LPS * P::NP + sizeof(float2) * (size_t)Cfg::LPS * a.n_out;
  auto kern = k_fft_rows_sense<P, HALF>;
  B2N_SMEM_OPT_IN(kern, smem);
  a.prefetch = want_prefetch(sizeof(float2) * (size_t)a.lines * a.n_in, PF_ROWS_INV) ? resident_ctas(kern, Cfg::SENSE_THREADS, smem) : 0;
  // the fused all-reduce makes CTAs wait for the same rows of the peers: only with the whole grid resident (no CTA
  // of a peer can be kept off its GPU by CTAs that wait for it) and the image inside the window
  a.peer_fused = 0;
  if (a.peer.world > 1) {
    a.peer_floats = 2 * rows * a.n_out;
    // ... and in the latency-bound regime of the one-shot exchange (larger images: the stand-alone kernel's two-shot form)
    const bool one_shot = 4 * a.peer_floats * (a.peer.world - 1) <= ((int64_t)8 << 20);
    if (one_shot && rows * a.coil_groups <= resident_ctas(kern, Cfg::SENSE_THREADS, smem) && a.peer_floats <= a.peer.slot_floats)
      a.peer_fused = 1;
    else a.peer.world = 0;
  }
  if (a.coil_groups > 1 && !g_counters_early) B2N_CUDA_OK(cudaMemsetAsync(a.counter, 0, sizeof(unsigned int) * (size_t)rows, st));
  B2N_CUDA_OK(launch_pdl(kern, dim3((unsigned)(rows * a.coil_groups)), dim3(Cfg::SENSE_THREADS), smem, st, a));
  B2N_LAUNCH_OK("k_fft_rows_sense");
  return 0;
}

template <class P> int launch_cols_toep(ColArgs &a, cudaStream_t st) {
  using Cfg = FastCfg<P>;
  if (2 * a.n_in > P::N || 2 * a.n_out > P::N) return -1;
  const size_t smem = sizeof(float4) * (size_t)Cfg::PAIRS * (P::NP + P::N);
  if (smem > (size_t)227 * 1024) return -1;  // spectrum + exchange buffer must fit one CTA
  auto kern = k_fft_cols_toep<P>;
  B2N_SMEM_OPT_IN(kern, smem);
  a.prefetch = want_prefetch(sizeof(float2) * (size_t)a.A * a.n_in * a.X, PF_COLS_TOEP) ? resident_ctas(kern, Cfg::COL_THREADS, smem) : 0;
  const int64_t gy = a.A < 32768 ? a.A : 32768;
  const dim3 grid((unsigned)ceil_div(a.X, 2 * Cfg::PAIRS), (unsigned)gy, (unsigned)ceil_div(a.A, gy));
  B2N_CUDA_OK(launch_pdl(kern, grid, dim3(Cfg::COL_THREADS), smem, st, a));
  B2N_LAUNCH_OK("k_fft_cols_toep");
  return 0;
}

template <class P> int launch_rows_sense_any(RowArgs &a, int64_t B, cudaStream_t st) {
  // a CTA carries FastCfg<P>::LPS coil pairs of one image row: with fewer coils than half of that most of its threads
  // would idle through the barriers (short lines, few coils: 256^3 x 8 coils) -- take the unfused route instead
  // ... unless the pass is to carry the all-reduce of a coil-sharded adjoint (2 coils per rank at 8 ranks): as a plain
  // adjoint the mostly idle CTAs cost 0.8 us against the unfused route (profiles/r02_few_coil_sense_ab.log), but they
  // save the all-reduce kernel behind it
  if (2 * ((a.C + 1) / 2) < FastCfg<P>::LPS && !((a.peer.world > 1 || (g_fft_stream & 128)) && a.C >= 2)) return -1;
  return 2 * a.n_out <= P::N ? launch_rows_sense_h<P, true>(a, B, st) : launch_rows_sense_h<P, false>(a, B, st);
}

// per-plan entry points: forward first row pass, inverse row pass, column pass, inverse row pass + coil sum
#define B2N_DECLARE_PLAN(N)                                                        \
  int fast_rows_fwd_##N(RowArgs &a, cudaStream_t st);                              \
  int fast_rows_inv_##N(RowArgs &a, cudaStream_t st);                              \
  int fast_cols_##N(bool inverse, ColArgs &a, cudaStream_t st);                    \
  int fast_cols_toep_##N(ColArgs &a, cudaStream_t st);                             \
  int fast_rows_sense_##N(RowArgs &a, int64_t B, cudaStream_t st);

#define B2N_DEFINE_PLAN(N)                                                                                  \
  int fast_rows_fwd_##N(RowArgs &a, cudaStream_t st) {                                                      \
    return launch_rows_fast<fast::Plan##N, false, ROW_FWD_FIRST>(a, st);                                    \
  }                                                                                                         \
  int fast_rows_inv_##N(RowArgs &a, cudaStream_t st) { return launch_rows_fast<fast::Plan##N, true, ROW_PLAIN>(a, st); } \
  int fast_cols_##N(bool inverse, ColArgs &a, cudaStream_t st) {                                            \
    return inverse ? launch_cols_fast<fast::Plan##N, true>(a, st) : launch_cols_fast<fast::Plan##N, false>(a, st); \
  }                                                                                                         \
  int fast_cols_toep_##N(ColArgs &a, cudaStream_t st) { return launch_cols_toep<fast::Plan##N>(a, st); }    \
  int fast_rows_sense_##N(RowArgs &a, int64_t B, cudaStream_t st) {                                         \
    return launch_rows_sense_any<fast::Plan##N>(a, B, st);                                                  \
  }

}  // namespace b2n
