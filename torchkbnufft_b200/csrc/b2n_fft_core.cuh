// b2n_fft_core.cuh -- mixed-radix Stockham FFT building blocks (complex64), host + device.
//
// One "item" of a Stockham autosort stage with radix R, for a transform of length n whose
// already-processed radices multiply to Ns:
//     k  = j mod Ns
//     v[r] = x[j + r*n/R] * W_n^(r*k*n/(Ns*R))        r = 0..R-1,  j = 0..n/R-1
//     v  = DFT_R(v)
//     y[(j div Ns)*Ns*R + k + r*Ns] = v[r]
// After all stages the output is in natural order.  W_n = exp(-2 pi i / n) (forward) or its
// conjugate (inverse, unnormalised).  Twiddles come from a table tw[t] = exp(-2 pi i t / n).
// The functions are __host__ __device__ so tests/test_fft_core.py can run the exact same
// arithmetic on the CPU build host against numpy (test scaffolding, not an execution path).
#pragma once
#include "b2n_common.cuh"

namespace b2n {

constexpr int kFftMaxStages = 12;

struct FftPlan {
  int n;
  int n_stages;
  int radix[kFftMaxStages];
};

// radices: powers of two in balanced chunks of <= 4 bits (largest first), then odd primes.
// Returns false when n has a prime factor > 13 (caller falls back to cuFFT).
static inline bool fft_factorize(int n, FftPlan *p, int max_pow2_bits = 4) {
  p->n = n;
  p->n_stages = 0;
  if (n < 1) return false;
  int e = 0;
  while (n % 2 == 0) { n /= 2; ++e; }
  int chunks = (e + max_pow2_bits - 1) / max_pow2_bits;
  for (int i = 0; i < chunks; ++i) {
    const int bits = (e + (chunks - 1 - i)) / chunks;  // balanced split, larger chunks first
    p->radix[p->n_stages++] = 1 << bits;
  }
  const int odd[5] = {3, 5, 7, 11, 13};
  for (int i = 0; i < 5; ++i)
    while (n % odd[i] == 0) {
      if (p->n_stages >= kFftMaxStages) return false;
      p->radix[p->n_stages++] = odd[i];
      n /= odd[i];
    }
  return n == 1;
}

B2N_HD float2 f2(float x, float y) {
  float2 r;
  r.x = x;
  r.y = y;
  return r;
}
B2N_HD float2 cadd(float2 a, float2 b) { return f2(a.x + b.x, a.y + b.y); }
B2N_HD float2 csub(float2 a, float2 b) { return f2(a.x - b.x, a.y - b.y); }
B2N_HD float2 cmul2(float2 a, float2 b) { return f2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// multiply by -i (forward) or +i (inverse)
template <bool INV> B2N_HD float2 rot90(float2 a) { return INV ? f2(-a.y, a.x) : f2(a.y, -a.x); }
template <bool INV> B2N_HD float2 twid(float2 w) { return INV ? f2(w.x, -w.y) : w; }

template <bool INV> B2N_HD void dft2(float2 &a, float2 &b) {
  const float2 t = a;
  a = cadd(t, b);
  b = csub(t, b);
}

template <bool INV> B2N_HD void dft4(float2 &v0, float2 &v1, float2 &v2, float2 &v3) {
  const float2 t0 = cadd(v0, v2), t1 = csub(v0, v2), t2 = cadd(v1, v3), t3 = rot90<INV>(csub(v1, v3));
  v0 = cadd(t0, t2);
  v1 = cadd(t1, t3);
  v2 = csub(t0, t2);
  v3 = csub(t1, t3);
}

template <int R, bool INV> struct Dft;

template <bool INV> struct Dft<2, INV> {
  B2N_HD static void run(float2 *v, const float2 *, int) { dft2<INV>(v[0], v[1]); }
};
template <bool INV> struct Dft<4, INV> {
  B2N_HD static void run(float2 *v, const float2 *, int) { dft4<INV>(v[0], v[1], v[2], v[3]); }
};
template <bool INV> struct Dft<3, INV> {
  B2N_HD static void run(float2 *v, const float2 *, int) {
    const float s3 = 0.86602540378443864676f;  // sin(pi/3)
    const float2 s = cadd(v[1], v[2]), d = csub(v[1], v[2]);
    const float2 m = f2(v[0].x - 0.5f * s.x, v[0].y - 0.5f * s.y);
    const float2 nn = rot90<INV>(f2(s3 * d.x, s3 * d.y));
    v[0] = cadd(v[0], s);
    v[1] = cadd(m, nn);
    v[2] = csub(m, nn);
  }
};
template <bool INV> struct Dft<5, INV> {
  B2N_HD static void run(float2 *v, const float2 *, int) {
    const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;  // cos(2pi/5), cos(4pi/5)
    const float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;   // sin(2pi/5), sin(4pi/5)
    const float2 a = cadd(v[1], v[4]), b = cadd(v[2], v[3]), c = csub(v[1], v[4]), d = csub(v[2], v[3]);
    const float2 m1 = f2(v[0].x + c1 * a.x + c2 * b.x, v[0].y + c1 * a.y + c2 * b.y);
    const float2 m2 = f2(v[0].x + c2 * a.x + c1 * b.x, v[0].y + c2 * a.y + c1 * b.y);
    const float2 n1 = rot90<INV>(f2(s1 * c.x + s2 * d.x, s1 * c.y + s2 * d.y));
    const float2 n2 = rot90<INV>(f2(s2 * c.x - s1 * d.x, s2 * c.y - s1 * d.y));
    v[0] = cadd(v[0], cadd(a, b));
    v[1] = cadd(m1, n1);
    v[4] = csub(m1, n1);
    v[2] = cadd(m2, n2);
    v[3] = csub(m2, n2);
  }
};
template <bool INV> struct Dft<8, INV> {
  B2N_HD static void run(float2 *v, const float2 *, int) {
    const float h = 0.70710678118654752440f;
    float2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6], o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];
    dft4<INV>(e0, e1, e2, e3);
    dft4<INV>(o0, o1, o2, o3);
    // W8^k * o_k, W8 = exp(-+ 2 pi i / 8)
    const float2 w1 = INV ? f2(h * (o1.x - o1.y), h * (o1.x + o1.y)) : f2(h * (o1.x + o1.y), h * (o1.y - o1.x));
    const float2 w2 = rot90<INV>(o2);
    const float2 w3 = INV ? f2(-h * (o3.x + o3.y), h * (o3.x - o3.y)) : f2(h * (o3.y - o3.x), -h * (o3.x + o3.y));
    v[0] = cadd(e0, o0);
    v[4] = csub(e0, o0);
    v[1] = cadd(e1, w1);
    v[5] = csub(e1, w1);
    v[2] = cadd(e2, w2);
    v[6] = csub(e2, w2);
    v[3] = cadd(e3, w3);
    v[7] = csub(e3, w3);
  }
};
template <bool INV> struct Dft<16, INV> {
  B2N_HD static void run(float2 *v, const float2 *, int) {
    // 4 x 4 Cooley-Tukey inside registers: columns q = 0..3 hold v[q + 4m]
    const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f;  // cos, sin of pi/8
    const float h = 0.70710678118654752440f;
    float2 a[4][4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      a[q][0] = v[q];
      a[q][1] = v[q + 4];
      a[q][2] = v[q + 8];
      a[q][3] = v[q + 12];
      dft4<INV>(a[q][0], a[q][1], a[q][2], a[q][3]);
    }
    // twiddles W16^(q*k), W16 = exp(-+ 2 pi i / 16); table of W16^t for t = 0..9 (forward)
    const float2 w16[10] = {f2(1.f, 0.f), f2(c1, -s1), f2(h, -h),   f2(s1, -c1),  f2(0.f, -1.f),
                            f2(-s1, -c1), f2(-h, -h),  f2(-c1, -s1), f2(-1.f, 0.f), f2(-c1, s1)};
#pragma unroll
    for (int q = 1; q < 4; ++q)
#pragma unroll
      for (int k = 1; k < 4; ++k) a[q][k] = cmul2(a[q][k], twid<INV>(w16[q * k]));
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      dft4<INV>(a[0][k], a[1][k], a[2][k], a[3][k]);
      v[k] = a[0][k];
      v[k + 4] = a[1][k];
      v[k + 8] = a[2][k];
      v[k + 12] = a[3][k];
    }
  }
};
// generic odd prime radix (7, 11, 13): O(R^2) with twiddles from the length-n table
template <int R, bool INV> struct Dft {
  B2N_HD static void run(float2 *v, const float2 *tw, int n) {
    float2 out[R];
    const int step = n / R;
#pragma unroll
    for (int p = 0; p < R; ++p) {
      float2 acc = v[0];
#pragma unroll
      for (int q = 1; q < R; ++q) acc = cadd(acc, cmul2(v[q], twid<INV>(tw[((p * q) % R) * step])));
      out[p] = acc;
    }
#pragma unroll
    for (int p = 0; p < R; ++p) v[p] = out[p];
  }
};

// One Stockham item.  LOAD(i) returns element i of the stage input, STORE(i, value) writes
// element i of the stage output.
template <int R, bool INV, typename Load, typename Store>
B2N_HD void fft_stage_item(int n, int Ns, int j, const float2 *tw, Load load, Store store) {
  const int stride = n / R;
  const int k = j % Ns;
  float2 v[R];
#pragma unroll
  for (int r = 0; r < R; ++r) v[r] = load(j + r * stride);
  if (Ns > 1) {
    const int tstep = k * (n / (Ns * R));  // < n / R
#pragma unroll
    for (int r = 1; r < R; ++r) v[r] = cmul2(v[r], twid<INV>(tw[r * tstep]));
  }
  Dft<R, INV>::run(v, tw, n);
  const int j0 = (j / Ns) * Ns * R + k;
#pragma unroll
  for (int r = 0; r < R; ++r) store(j0 + r * Ns, v[r]);
}

// runtime radix dispatch
template <bool INV, typename Load, typename Store>
B2N_HD void fft_stage_item_any(int radix, int n, int Ns, int j, const float2 *tw, Load load, Store store) {
  switch (radix) {
    case 16: fft_stage_item<16, INV>(n, Ns, j, tw, load, store); break;
    case 8: fft_stage_item<8, INV>(n, Ns, j, tw, load, store); break;
    case 4: fft_stage_item<4, INV>(n, Ns, j, tw, load, store); break;
    case 2: fft_stage_item<2, INV>(n, Ns, j, tw, load, store); break;
    case 3: fft_stage_item<3, INV>(n, Ns, j, tw, load, store); break;
    case 5: fft_stage_item<5, INV>(n, Ns, j, tw, load, store); break;
    case 7: fft_stage_item<7, INV>(n, Ns, j, tw, load, store); break;
    case 11: fft_stage_item<11, INV>(n, Ns, j, tw, load, store); break;
    default: fft_stage_item<13, INV>(n, Ns, j, tw, load, store); break;
  }
}

}  // namespace b2n
