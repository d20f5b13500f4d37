// b2n_fft_fast.cuh -- compile-time planned, register-resident FFT lines (complex64).
//
// The generic Stockham passes in b2n_fft.cu take any factorisation at run time and pay for it
// (radix switch, index arithmetic, shared-memory twiddle gathers: ~50 instructions per element
// per stage, profiles/r01_g).  For the grid sizes 2x-oversampled MRI actually uses this header
// fixes the plan at compile time: N = R0*R1[*R2], thread t of a line owns item t of every stage
// (one radix-R butterfly held in registers), stages exchange through ONE shared-memory buffer,
// and every thread carries TWO lines side by side in float4 values -- two adjacent columns in
// the column pass (128-bit global and shared accesses), two rows in the row pass -- so index
// arithmetic and twiddles are shared by the pair.
//
// Stockham autosort with radices R0, R1, R2 (Ns = product of the radices already done):
//   stage input  index of (item j, leg r) : j + r * N/R
//   twiddle                               : W_{Ns*R}^(r * (j mod Ns))
//   stage output index                    : (j - j mod Ns) * R + (j mod Ns) + r * Ns
// Twiddles come from a per-stage table laid out [r-1][k] (consecutive threads read consecutive
// entries): stage 1 at offset 0 ((R1-1)*R0 entries), stage 2 behind it ((R2-1)*R0*R1 entries).
//
// Everything except the __syncthreads() choreography is __host__ __device__ so that
// tests/test_fft_core.py runs the same index / twiddle / butterfly code on the CPU.
#pragma once
#include "b2n_common.cuh"

namespace b2n {
namespace fast {

// ---- a pair of complex numbers: (x, y) is line A, (z, w) is line B -----------------------
B2N_HD float4 v4(float x, float y, float z, float w) {
  float4 r;
  r.x = x;
  r.y = y;
  r.z = z;
  r.w = w;
  return r;
}
B2N_HD float4 vadd(float4 a, float4 b) { return v4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
B2N_HD float4 vsub(float4 a, float4 b) { return v4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
B2N_HD float4 vscale(float4 a, float s) { return v4(a.x * s, a.y * s, a.z * s, a.w * s); }
// a + s * b
B2N_HD float4 vaxpy(float4 a, float s, float4 b) {
  return v4(fmaf(s, b.x, a.x), fmaf(s, b.y, a.y), fmaf(s, b.z, a.z), fmaf(s, b.w, a.w));
}
// multiply by -i (forward transform) or +i (inverse)
template <bool INV> B2N_HD float4 vrot(float4 a) { return INV ? v4(-a.y, a.x, -a.w, a.z) : v4(a.y, -a.x, a.w, -a.z); }
// multiply both lines by w = (wx, wy) (forward) or by conj(w) (inverse)
template <bool INV> B2N_HD float4 vmulw(float4 a, float wx, float wy) {
  const float s = INV ? -wy : wy;
  return v4(fmaf(a.x, wx, -(a.y * s)), fmaf(a.x, s, a.y * wx), fmaf(a.z, wx, -(a.w * s)), fmaf(a.z, s, a.w * wx));
}
B2N_HD float4 vmul2(float4 a, float4 b) {  // line-wise complex product
  return v4(fmaf(a.x, b.x, -(a.y * b.y)), fmaf(a.x, b.y, a.y * b.x), fmaf(a.z, b.z, -(a.w * b.w)),
            fmaf(a.z, b.w, a.w * b.z));
}
B2N_HD float4 vmul2_conj(float4 a, float4 b) {  // a * conj(b), line-wise
  return v4(fmaf(a.x, b.x, a.y * b.y), fmaf(a.y, b.x, -(a.x * b.y)), fmaf(a.z, b.z, a.w * b.w),
            fmaf(a.w, b.z, -(a.z * b.w)));
}

// ---- butterflies ---------------------------------------------------------------------------
B2N_HD void dft2(float4 &a, float4 &b) {
  const float4 t = a;
  a = vadd(t, b);
  b = vsub(t, b);
}
template <bool INV> B2N_HD void dft3(float4 &v0, float4 &v1, float4 &v2) {
  const float s3 = 0.86602540378443864676f;
  const float4 s = vadd(v1, v2), d = vsub(v1, v2);
  const float4 m = vaxpy(v0, -0.5f, s);
  const float4 n = vrot<INV>(vscale(d, s3));
  v0 = vadd(v0, s);
  v1 = vadd(m, n);
  v2 = vsub(m, n);
}
template <bool INV> B2N_HD void dft4(float4 &v0, float4 &v1, float4 &v2, float4 &v3) {
  const float4 t0 = vadd(v0, v2), t1 = vsub(v0, v2), t2 = vadd(v1, v3), t3 = vrot<INV>(vsub(v1, v3));
  v0 = vadd(t0, t2);
  v1 = vadd(t1, t3);
  v2 = vsub(t0, t2);
  v3 = vsub(t1, t3);
}
template <bool INV> B2N_HD void dft5(float4 &v0, float4 &v1, float4 &v2, float4 &v3, float4 &v4_) {
  const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;  // cos(2pi/5), cos(4pi/5)
  const float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;   // sin(2pi/5), sin(4pi/5)
  const float4 a = vadd(v1, v4_), b = vadd(v2, v3), c = vsub(v1, v4_), d = vsub(v2, v3);
  const float4 m1 = vaxpy(vaxpy(v0, c1, a), c2, b);
  const float4 m2 = vaxpy(vaxpy(v0, c2, a), c1, b);
  const float4 n1 = vrot<INV>(vaxpy(vscale(c, s1), s2, d));
  const float4 n2 = vrot<INV>(vaxpy(vscale(c, s2), -s1, d));
  v0 = vadd(v0, vadd(a, b));
  v1 = vadd(m1, n1);
  v4_ = vsub(m1, n1);
  v2 = vadd(m2, n2);
  v3 = vsub(m2, n2);
}
template <bool INV> B2N_HD void dft7(float4 *v) {
  const float c1 = 0.62348980185873353053f, c2 = -0.22252093395631440429f, c3 = -0.90096886790241912624f;  // cos(2 pi j/7)
  const float s1 = 0.78183148246802980871f, s2 = 0.97492791218182360702f, s3 = 0.43388373911755812048f;   // sin(2 pi j/7)
  const float4 a1 = vadd(v[1], v[6]), a2 = vadd(v[2], v[5]), a3 = vadd(v[3], v[4]);
  const float4 b1 = vsub(v[1], v[6]), b2 = vsub(v[2], v[5]), b3 = vsub(v[3], v[4]);
  const float4 x0 = v[0];
  const float4 m1 = vaxpy(vaxpy(vaxpy(x0, c1, a1), c2, a2), c3, a3);
  const float4 m2 = vaxpy(vaxpy(vaxpy(x0, c2, a1), c3, a2), c1, a3);
  const float4 m3 = vaxpy(vaxpy(vaxpy(x0, c3, a1), c1, a2), c2, a3);
  const float4 n1 = vrot<INV>(vaxpy(vaxpy(vscale(b1, s1), s2, b2), s3, b3));
  const float4 n2 = vrot<INV>(vaxpy(vaxpy(vscale(b1, s2), -s3, b2), -s1, b3));
  const float4 n3 = vrot<INV>(vaxpy(vaxpy(vscale(b1, s3), -s1, b2), s2, b3));
  v[0] = vadd(x0, vadd(a1, vadd(a2, a3)));
  v[1] = vadd(m1, n1);
  v[6] = vsub(m1, n1);
  v[2] = vadd(m2, n2);
  v[5] = vsub(m2, n2);
  v[3] = vadd(m3, n3);
  v[4] = vsub(m3, n3);
}
// second half of the radix-8 butterfly: e = DFT4 of the even inputs, o = DFT4 of the odd inputs
template <bool INV>
B2N_HD void dft8_finish(float4 e0, float4 e1, float4 e2, float4 e3, float4 o0, float4 o1, float4 o2, float4 o3,
                        float4 *v) {
  const float h = 0.70710678118654752440f;
  const float4 w1 = vmulw<INV>(o1, h, -h);  // W8^1
  const float4 w2 = vrot<INV>(o2);          // W8^2
  const float4 w3 = vmulw<INV>(o3, -h, -h); // W8^3
  v[0] = vadd(e0, o0);
  v[4] = vsub(e0, o0);
  v[1] = vadd(e1, w1);
  v[5] = vsub(e1, w1);
  v[2] = vadd(e2, w2);
  v[6] = vsub(e2, w2);
  v[3] = vadd(e3, w3);
  v[7] = vsub(e3, w3);
}
template <bool INV> B2N_HD void dft8(float4 *v) {
  float4 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6], o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];
  dft4<INV>(e0, e1, e2, e3);
  dft4<INV>(o0, o1, o2, o3);
  dft8_finish<INV>(e0, e1, e2, e3, o0, o1, o2, o3, v);
}
// DFT4 of (a, b, 0, 0)
template <bool INV> B2N_HD void dft4_half_in(float4 a, float4 b, float4 &y0, float4 &y1, float4 &y2, float4 &y3) {
  const float4 rb = vrot<INV>(b);
  y0 = vadd(a, b);
  y1 = vadd(a, rb);
  y2 = vsub(a, b);
  y3 = vsub(a, rb);
}
// radix 8 with inputs 4..7 known to be zero (they lie in the zero padding)
template <bool INV> B2N_HD void dft8_half_in(float4 *v) {
  float4 e0, e1, e2, e3, o0, o1, o2, o3;
  dft4_half_in<INV>(v[0], v[2], e0, e1, e2, e3);
  dft4_half_in<INV>(v[1], v[3], o0, o1, o2, o3);
  dft8_finish<INV>(e0, e1, e2, e3, o0, o1, o2, o3, v);
}
// 10 = 2 x 5: DFT5 over the even and the odd inputs, twiddle W10^k, DFT2
template <bool INV> B2N_HD void dft10(float4 *v) {
  const float ca = 0.80901699437494742410f, sa = 0.58778525229247312917f;  // cos, sin of 36 deg
  const float cb = 0.30901699437494742410f, sb = 0.95105651629515357212f;  // cos, sin of 72 deg
  float4 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6], e4 = v[8];
  float4 o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7], o4 = v[9];
  dft5<INV>(e0, e1, e2, e3, e4);
  dft5<INV>(o0, o1, o2, o3, o4);
  o1 = vmulw<INV>(o1, ca, -sa);
  o2 = vmulw<INV>(o2, cb, -sb);
  o3 = vmulw<INV>(o3, -cb, -sb);
  o4 = vmulw<INV>(o4, -ca, -sa);
  v[0] = vadd(e0, o0);
  v[5] = vsub(e0, o0);
  v[1] = vadd(e1, o1);
  v[6] = vsub(e1, o1);
  v[2] = vadd(e2, o2);
  v[7] = vsub(e2, o2);
  v[3] = vadd(e3, o3);
  v[8] = vsub(e3, o3);
  v[4] = vadd(e4, o4);
  v[9] = vsub(e4, o4);
}
// 12 = 4 x 3: n = 4*n2 + n1, k = k2 + 3*k1: DFT3 over n2, twiddle W12^(n1*k2), DFT4 over n1
template <bool INV> B2N_HD void dft12(float4 *v) {
  const float s3 = 0.86602540378443864676f;
  float4 y[4][3];
#pragma unroll
  for (int n1 = 0; n1 < 4; ++n1) {
    y[n1][0] = v[n1];
    y[n1][1] = v[n1 + 4];
    y[n1][2] = v[n1 + 8];
    dft3<INV>(y[n1][0], y[n1][1], y[n1][2]);
  }
  y[1][1] = vmulw<INV>(y[1][1], s3, -0.5f);   // W12^1
  y[1][2] = vmulw<INV>(y[1][2], 0.5f, -s3);   // W12^2
  y[2][1] = vmulw<INV>(y[2][1], 0.5f, -s3);   // W12^2
  y[2][2] = vmulw<INV>(y[2][2], -0.5f, -s3);  // W12^4
  y[3][1] = vrot<INV>(y[3][1]);               // W12^3
  y[3][2] = vscale(y[3][2], -1.f);            // W12^6
#pragma unroll
  for (int k2 = 0; k2 < 3; ++k2) {
    dft4<INV>(y[0][k2], y[1][k2], y[2][k2], y[3][k2]);
    v[k2] = y[0][k2];
    v[k2 + 3] = y[1][k2];
    v[k2 + 6] = y[2][k2];
    v[k2 + 9] = y[3][k2];
  }
}
// 9 = 3 x 3; y[n1][k2] = DFT3 over n2 of v[3*n2 + n1], twiddles W9^(n1*k2), then DFT3 over n1
template <bool INV> B2N_HD void dft9(float4 *v) {
  const float c1 = 0.76604444311897803520f, s1 = 0.64278760968653932632f;   // 2 pi / 9
  const float c2 = 0.17364817766693034885f, s2 = 0.98480775301220805937f;   // 4 pi / 9
  const float c4 = -0.93969262078590838405f, s4 = 0.34202014332566873304f;  // 8 pi / 9
  float4 y[3][3];
#pragma unroll
  for (int n1 = 0; n1 < 3; ++n1) {
    y[n1][0] = v[n1];
    y[n1][1] = v[n1 + 3];
    y[n1][2] = v[n1 + 6];
    dft3<INV>(y[n1][0], y[n1][1], y[n1][2]);
  }
  y[1][1] = vmulw<INV>(y[1][1], c1, -s1);  // W9^1
  y[1][2] = vmulw<INV>(y[1][2], c2, -s2);  // W9^2
  y[2][1] = vmulw<INV>(y[2][1], c2, -s2);  // W9^2
  y[2][2] = vmulw<INV>(y[2][2], c4, -s4);  // W9^4
#pragma unroll
  for (int k2 = 0; k2 < 3; ++k2) {
    dft3<INV>(y[0][k2], y[1][k2], y[2][k2]);
    v[k2] = y[0][k2];
    v[k2 + 3] = y[1][k2];
    v[k2 + 6] = y[2][k2];
  }
}
// 15 = 5 x 3; y[n1][k2] = DFT3 over n2 of v[5*n2 + n1], twiddles W15^(n1*k2), then DFT5 over n1
template <bool INV> B2N_HD void dft15(float4 *v) {
  // cos / sin of 2 pi k / 15 for the exponents k = n1 * k2 that occur
  const float c1 = 0.91354545764260089550f, s1 = 0.40673664307580020775f;
  const float c2 = 0.66913060635885821383f, s2 = 0.74314482547739423501f;
  const float c3 = 0.30901699437494742410f, s3 = 0.95105651629515357212f;
  const float c4 = -0.10452846326765347140f, s4 = 0.99452189536827333692f;
  const float c6 = -0.80901699437494742410f, s6 = 0.58778525229247312917f;
  const float c8 = -0.97814760073380563793f, s8 = -0.20791169081775933710f;
  float4 y[5][3];
#pragma unroll
  for (int n1 = 0; n1 < 5; ++n1) {
    y[n1][0] = v[n1];
    y[n1][1] = v[n1 + 5];
    y[n1][2] = v[n1 + 10];
    dft3<INV>(y[n1][0], y[n1][1], y[n1][2]);
  }
  y[1][1] = vmulw<INV>(y[1][1], c1, -s1);
  y[1][2] = vmulw<INV>(y[1][2], c2, -s2);
  y[2][1] = vmulw<INV>(y[2][1], c2, -s2);
  y[2][2] = vmulw<INV>(y[2][2], c4, -s4);
  y[3][1] = vmulw<INV>(y[3][1], c3, -s3);
  y[3][2] = vmulw<INV>(y[3][2], c6, -s6);
  y[4][1] = vmulw<INV>(y[4][1], c4, -s4);
  y[4][2] = vmulw<INV>(y[4][2], c8, -s8);
#pragma unroll
  for (int k2 = 0; k2 < 3; ++k2) {
    dft5<INV>(y[0][k2], y[1][k2], y[2][k2], y[3][k2], y[4][k2]);
    v[k2] = y[0][k2];
    v[k2 + 3] = y[1][k2];
    v[k2 + 6] = y[2][k2];
    v[k2 + 9] = y[3][k2];
    v[k2 + 12] = y[4][k2];
  }
}
// 16 = 4 x 4; y[n1][k2] = DFT4 over n2 of v[4*n2 + n1]
template <bool INV> B2N_HD void dft16_finish(float4 (*y)[4], float4 *v) {
  const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f;  // cos, sin of pi/8
  const float h = 0.70710678118654752440f;
  y[1][1] = vmulw<INV>(y[1][1], c1, -s1);   // W16^1
  y[1][2] = vmulw<INV>(y[1][2], h, -h);     // W16^2
  y[1][3] = vmulw<INV>(y[1][3], s1, -c1);   // W16^3
  y[2][1] = vmulw<INV>(y[2][1], h, -h);     // W16^2
  y[2][2] = vrot<INV>(y[2][2]);             // W16^4
  y[2][3] = vmulw<INV>(y[2][3], -h, -h);    // W16^6
  y[3][1] = vmulw<INV>(y[3][1], s1, -c1);   // W16^3
  y[3][2] = vmulw<INV>(y[3][2], -h, -h);    // W16^6
  y[3][3] = vmulw<INV>(y[3][3], -c1, s1);   // W16^9
#pragma unroll
  for (int k2 = 0; k2 < 4; ++k2) {
    dft4<INV>(y[0][k2], y[1][k2], y[2][k2], y[3][k2]);
    v[k2] = y[0][k2];
    v[k2 + 4] = y[1][k2];
    v[k2 + 8] = y[2][k2];
    v[k2 + 12] = y[3][k2];
  }
}
template <bool INV> B2N_HD void dft16(float4 *v) {
  float4 y[4][4];
#pragma unroll
  for (int n1 = 0; n1 < 4; ++n1) {
    y[n1][0] = v[n1];
    y[n1][1] = v[n1 + 4];
    y[n1][2] = v[n1 + 8];
    y[n1][3] = v[n1 + 12];
    dft4<INV>(y[n1][0], y[n1][1], y[n1][2], y[n1][3]);
  }
  dft16_finish<INV>(y, v);
}
// radix 16 with inputs 8..15 known to be zero
template <bool INV> B2N_HD void dft16_half_in(float4 *v) {
  float4 y[4][4];
#pragma unroll
  for (int n1 = 0; n1 < 4; ++n1) dft4_half_in<INV>(v[n1], v[n1 + 4], y[n1][0], y[n1][1], y[n1][2], y[n1][3]);
  dft16_finish<INV>(y, v);
}

template <int R, bool INV> B2N_HD void dft(float4 *v) {
  static_assert(R == 2 || R == 3 || R == 4 || R == 5 || R == 7 || R == 8 || R == 9 || R == 10 || R == 12 || R == 15 ||
                    R == 16,
                "radix");
  if constexpr (R == 2) dft2(v[0], v[1]);
  if constexpr (R == 3) dft3<INV>(v[0], v[1], v[2]);
  if constexpr (R == 4) dft4<INV>(v[0], v[1], v[2], v[3]);
  if constexpr (R == 5) dft5<INV>(v[0], v[1], v[2], v[3], v[4]);
  if constexpr (R == 7) dft7<INV>(v);
  if constexpr (R == 8) dft8<INV>(v);
  if constexpr (R == 9) dft9<INV>(v);
  if constexpr (R == 10) dft10<INV>(v);
  if constexpr (R == 12) dft12<INV>(v);
  if constexpr (R == 15) dft15<INV>(v);
  if constexpr (R == 16) dft16<INV>(v);
}

// first-stage butterfly when legs r >= R/2 are zero
template <int R, bool INV> B2N_HD void dft_half_in(float4 *v) {
  static_assert(R == 8 || R == 16, "first radix");
  if constexpr (R == 8) dft8_half_in<INV>(v);
  if constexpr (R == 16) dft16_half_in<INV>(v);
}

// ---- compile-time plan ---------------------------------------------------------------------
constexpr int cmax(int a, int b) { return a > b ? a : b; }

template <int N_, int R0_, int R1_, int R2_> struct Plan {
  static_assert(R0_ * R1_ * R2_ == N_, "radices must multiply to N");
  static_assert((R0_ & (R0_ - 1)) == 0, "first radix must be a power of two (padding, k = t & (R0-1))");
  static constexpr int N = N_, R0 = R0_, R1 = R1_, R2 = R2_;
  static constexpr int NS = R2 > 1 ? 3 : 2;
  static constexpr int I0 = N / R0, I1 = N / R1, I2 = R2 > 1 ? N / R2 : 0;  // butterflies per stage
  static constexpr int T = cmax(I0, cmax(I1, I2));                          // threads per line pair
  static constexpr int RMAX = cmax(R0, cmax(R1, R2));
  static constexpr int PSH = R0 >= 16 ? 4 : 3;                // one padding slot per R0 = 2^PSH elements:
  static constexpr int NP = N + (N >> PSH) + 1;               // stride-R0 stores stay conflict-free
  static_assert(R0 == (1 << PSH), "first radix must be 8 or 16");
  static_assert(I1 % R0 == 0 && I2 % R0 == 0, "leg strides must be multiples of the padding period");
  static constexpr int TW2 = (R1 - 1) * R0;                   // offset of the stage-2 twiddle table
  static constexpr int TW_COUNT = TW2 + (R2 > 1 ? (R2 - 1) * R0 * R1 : 0);
  B2N_HD static int pad(int i) { return i + (i >> PSH); }
  // pad(i + c) == pad(i) + padc(c) when c is a multiple of R0
  static constexpr int padc(int c) { return c + (c >> PSH); }
};

// staged twiddle table entry e of plan P (double precision, rounded once)
template <class P> inline void staged_twiddle_index(int e, int *r, int *k, int *period) {
  if (e < P::TW2) {
    *r = e / P::R0 + 1;
    *k = e % P::R0;
    *period = P::R0 * P::R1;
  } else {
    e -= P::TW2;
    *r = e / (P::R0 * P::R1) + 1;
    *k = e % (P::R0 * P::R1);
    *period = P::N;
  }
}

B2N_HD float2 tw_load(const float2 *p) {
#ifdef __CUDA_ARCH__
  return __ldg(p);
#else
  return *p;
#endif
}

// twiddle + butterfly of one item: v[r] *= W^(r*k), v = DFT_R(v); tw is the stage table [r-1][k]
template <int R, bool INV> B2N_HD void stage_compute(float4 *v, const float2 *tw, int Ns, int k) {
#pragma unroll
  for (int r = 1; r < R; ++r) {
    const float2 w = tw_load(tw + (r - 1) * Ns + k);
    v[r] = vmulw<INV>(v[r], w.x, w.y);
  }
  dft<R, INV>(v);
}

#ifdef __CUDACC__
// One line pair through all stages.  `t` = this thread's item index (0 <= t < P::T), `sm` = the
// pair's shared exchange buffer, element i at sm[P::pad(i) * es].  loadg(i, r) returns input element i = t + r*I0
// of both lines (r = the leg: a staged pass keeps leg r in its own slot), storeg(i, v) takes output element i,
// after0() runs once behind the first barrier (every first-stage operand has been consumed: a streamed pass
// issues the asynchronous copies of its next tile there).  Every thread of the CTA must call this
// (it contains CTA-wide barriers); threads with nothing to do pass functors that read zeros
// and drop stores.
// HIN : input elements i >= N/2 are zero padding (never loaded, first butterfly simplified).
// HOUT: output elements i >= N/2 are cropped (never computed: only the legs r < R/2 of the last
//       stage are stored and the compiler drops the rest of that butterfly).
struct NoHook {
  __device__ __forceinline__ void operator()() const {}
};
template <class P, bool INV, bool HIN, bool HOUT, class LoadG, class StoreG, class After0 = NoHook>
__device__ __forceinline__ void fft_line_pair(int t, float4 *sm, int es, const float2 *tws, LoadG loadg,
                                              StoreG storeg, After0 after0 = After0()) {
  float4 v[P::RMAX];
  if (t < P::I0) {
    if constexpr (HIN) {
#pragma unroll
      for (int r = 0; r < P::R0 / 2; ++r) v[r] = loadg(t + r * P::I0, r);
      dft_half_in<P::R0, INV>(v);
    } else {
#pragma unroll
      for (int r = 0; r < P::R0; ++r) v[r] = loadg(t + r * P::I0, r);
      dft<P::R0, INV>(v);
    }
    float4 *dst = sm + t * (P::R0 + 1) * es;  // pad(t*R0 + r) == t*(R0+1) + r
#pragma unroll
    for (int r = 0; r < P::R0; ++r) dst[r * es] = v[r];
  }
  __syncthreads();
  after0();
  const int k1 = t & (P::R0 - 1);
  const int o1 = (t - k1) * P::R1 + k1;
  const float4 *src = sm + P::pad(t) * es;
  if (t < P::I1) {
#pragma unroll
    for (int r = 0; r < P::R1; ++r) v[r] = src[P::padc(r * P::I1) * es];
    stage_compute<P::R1, INV>(v, tws, P::R0, k1);
  }
  if constexpr (P::NS == 2) {
    constexpr int LEGS = HOUT ? (P::R1 + 1) / 2 : P::R1;
    if (t < P::I1) {
#pragma unroll
      for (int r = 0; r < LEGS; ++r) storeg(o1 + r * P::R0, v[r]);
    }
  } else {
    __syncthreads();  // everyone holds its stage-1 inputs: the buffer may be overwritten
    if (t < P::I1) {
      float4 *dst = sm + P::pad(o1) * es;
#pragma unroll
      for (int r = 0; r < P::R1; ++r) dst[P::padc(r * P::R0) * es] = v[r];
    }
    __syncthreads();
    constexpr int Ns2 = P::R0 * P::R1;  // == P::I2, so k2 == t and the output index is t + r*Ns2
    constexpr int LEGS = HOUT ? (P::R2 + 1) / 2 : P::R2;
    if (t < P::I2) {
#pragma unroll
      for (int r = 0; r < P::R2; ++r) v[r] = src[P::padc(r * P::I2) * es];
      stage_compute<P::R2, INV>(v, tws + P::TW2, Ns2, t);
#pragma unroll
      for (int r = 0; r < LEGS; ++r) storeg(t + r * Ns2, v[r]);
    }
  }
}
#endif  // __CUDACC__

// Plans (N, R0, R1, R2) for the grid lengths 2x-oversampled acquisitions use (BASELINE configs: 256, 512, 640,
// 768; the reference notebooks' 400 -> 800; 1152, 1536 and other 5-smooth sizes).  R0 is 8 or 16 (padding period), the leg strides N/R1 and N/R2 are multiples of R0, R2 = 1 means two
// stages.  Every other length takes the run-time passes or cuFFT.
#define B2N_FAST_PLANS(X, ...)                                                                                  \
  X(64, 8, 8, 1, __VA_ARGS__) X(96, 8, 12, 1, __VA_ARGS__) X(128, 8, 16, 1, __VA_ARGS__)                       \
  X(192, 8, 8, 3, __VA_ARGS__) X(224, 8, 4, 7, __VA_ARGS__) X(256, 16, 16, 1, __VA_ARGS__)                     \
  X(288, 8, 12, 3, __VA_ARGS__) X(320, 8, 8, 5, __VA_ARGS__) X(384, 8, 16, 3, __VA_ARGS__)                     \
  X(448, 8, 8, 7, __VA_ARGS__) X(480, 8, 12, 5, __VA_ARGS__) X(512, 8, 8, 8, __VA_ARGS__)                      \
  X(576, 16, 12, 3, __VA_ARGS__) X(640, 8, 8, 10, __VA_ARGS__) X(768, 8, 8, 12, __VA_ARGS__)                   \
  X(896, 16, 8, 7, __VA_ARGS__) X(960, 16, 12, 5, __VA_ARGS__) X(1024, 8, 8, 16, __VA_ARGS__)                  \
  X(1280, 16, 8, 10, __VA_ARGS__) X(2048, 16, 16, 8, __VA_ARGS__)                                               \
  X(160, 8, 4, 5, __VA_ARGS__) X(200, 8, 5, 5, __VA_ARGS__) X(240, 8, 10, 3, __VA_ARGS__)                      \
  X(400, 8, 10, 5, __VA_ARGS__) X(800, 8, 10, 10, __VA_ARGS__) X(1152, 8, 12, 12, __VA_ARGS__)                 \
  X(1536, 16, 8, 12, __VA_ARGS__) X(1600, 16, 10, 10, __VA_ARGS__) X(1920, 16, 12, 10, __VA_ARGS__)              \
  X(72, 8, 9, 1, __VA_ARGS__) X(120, 8, 15, 1, __VA_ARGS__) X(144, 16, 9, 1, __VA_ARGS__)                      \
  X(360, 8, 5, 9, __VA_ARGS__) X(600, 8, 5, 15, __VA_ARGS__) X(720, 16, 5, 9, __VA_ARGS__)                     \
  X(1200, 16, 5, 15, __VA_ARGS__) X(1440, 16, 10, 9, __VA_ARGS__)

#define B2N_FAST_PLAN_USING(N, R0, R1, R2, ...) using Plan##N = Plan<N, R0, R1, R2>;
B2N_FAST_PLANS(B2N_FAST_PLAN_USING, 0)

// switch over the planned lengths: CALL is evaluated with `P` naming the plan type
#define B2N_FAST_PLAN_CASE(N, R0, R1, R2, CALL) \
  case N: {                                     \
    using P = ::b2n::fast::Plan##N;             \
    CALL;                                       \
  } break;
#define B2N_FAST_PLAN_SWITCH(n_, CALL, DEFAULT)  \
  switch (n_) {                                  \
    B2N_FAST_PLANS(B2N_FAST_PLAN_CASE, CALL)     \
    default: {                                   \
      DEFAULT;                                   \
    } break;                                     \
  }

}  // namespace fast
}  // namespace b2n
