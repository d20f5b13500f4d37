// b2n_interp.cu -- table interpolation: forward gather (grid -> points) and adjoint
// spread (points -> grid) over a b2n_points plan.
//
// reference hot loops replaced here:
//   forward  torchkbnufft/_nufft/interp.py:185-203  (W passes of index + mul + add_)
//   adjoint  torchkbnufft/_nufft/interp.py:689-724  (W x B*C index_add_ launches)
// One launch covers all W = prod(J_d) neighbour offsets, every batch/coil row and
// the fftshift phase.  The accumulation order over offsets is the reference's
// (row-major offsets; coefficient = running product over dimensions).
#include <type_traits>

#include "b2n_common.cuh"
#include "b2n_interp.cuh"

namespace b2n {

template <bool CL> B2N_D int64_t grid_addr(int64_t b, int64_t c, int64_t cell, int64_t C, int64_t Kprod) {
  return CL ? (b * Kprod + cell) * C + c : (b * C + c) * Kprod + cell;
}

B2N_D int64_t wrap_up(int64_t g, int64_t K) { return g >= K ? g % K : g; }

B2N_D void atomic_add_cplx(cplx<float> *addr, cplx<float> v) {
  atomicAdd(reinterpret_cast<float2 *>(addr), make_float2(v.x, v.y));  // one 8-byte L2 reduction
}
B2N_D void atomic_add_cplx(cplx<double> *addr, cplx<double> v) {
  atomicAdd(&addr->x, v.x);
  atomicAdd(&addr->y, v.y);
}

// -----------------------------------------------------------------------------
// generic forward gather: one thread per (sorted point, batch/coil row)
// -----------------------------------------------------------------------------
template <typename T, int ND, bool CL>
__global__ void __launch_bounds__(128) k_fwd_generic(InterpArgs<T> a, const cplx<T> *__restrict__ grid,
                                                     cplx<T> *__restrict__ kdata) {
  const int64_t total = a.n_traj * a.M;
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= total) return;
  const int32_t *bs = a.base + s * ND;
  const cplx<T> *rec = a.coef + s * a.coef_stride;
  const int64_t m = a.perm[s];
  const int64_t rows = a.n_traj == 1 ? a.B * a.C : a.C;
  for (int64_t r = blockIdx.y; r < rows; r += gridDim.y) {
    const int64_t b = a.n_traj == 1 ? r / a.C : s / a.M;
    const int64_t c = a.n_traj == 1 ? r - b * a.C : r;
    cplx<T> acc = {T(0), T(0)};
    for (int j0 = 0; j0 < a.J[0]; ++j0) {
      const int64_t g0 = wrap_up(bs[0] + j0, a.K[0]);
      const cplx<T> c0 = rec[j0];
      if (ND == 1) {
        cmac(acc, c0, grid[grid_addr<CL>(b, c, g0, a.C, a.Kprod)]);
      } else {
        for (int j1 = 0; j1 < a.J[1]; ++j1) {
          const int64_t g1 = wrap_up(bs[1] + j1, a.K[1]);
          const cplx<T> c01 = cmul(c0, rec[a.coef_off[1] + j1]);
          if (ND == 2) {
            cmac(acc, c01, grid[grid_addr<CL>(b, c, g0 * a.K[1] + g1, a.C, a.Kprod)]);
          } else {
            for (int j2 = 0; j2 < a.J[2]; ++j2) {
              const int64_t g2 = wrap_up(bs[2] + j2, a.K[2]);
              const cplx<T> c012 = cmul(c01, rec[a.coef_off[2] + j2]);
              cmac(acc, c012, grid[grid_addr<CL>(b, c, (g0 * a.K[1] + g1) * a.K[2] + g2, a.C, a.Kprod)]);
            }
          }
        }
      }
    }
    kdata[(b * a.C + c) * a.M + m] = acc;  // the fftshift phase is folded into the dim-0 weights
  }
}

// -----------------------------------------------------------------------------
// generic adjoint, atomic mode: one thread per (sorted point, row), L2 reductions
// -----------------------------------------------------------------------------
template <typename T, int ND, bool CL>
__global__ void __launch_bounds__(128) k_adj_atomic_generic(InterpArgs<T> a, const cplx<T> *__restrict__ kdata,
                                                            cplx<T> *__restrict__ grid) {
  const int64_t total = a.n_traj * a.M;
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= total) return;
  const int32_t *bs = a.base + s * ND;
  const cplx<T> *rec = a.coef + s * a.coef_stride;
  const int64_t m = a.perm[s];
  const int64_t rows = a.n_traj == 1 ? a.B * a.C : a.C;
  for (int64_t r = blockIdx.y; r < rows; r += gridDim.y) {
    const int64_t b = a.n_traj == 1 ? r / a.C : s / a.M;
    const int64_t c = a.n_traj == 1 ? r - b * a.C : r;
    const cplx<T> val = kdata[(b * a.C + c) * a.M + m];
    for (int j0 = 0; j0 < a.J[0]; ++j0) {
      const int64_t g0 = wrap_up(bs[0] + j0, a.K[0]);
      const cplx<T> c0 = rec[j0];
      if (ND == 1) {
        atomic_add_cplx(&grid[grid_addr<CL>(b, c, g0, a.C, a.Kprod)], cmul(cconj(c0), val));
      } else {
        for (int j1 = 0; j1 < a.J[1]; ++j1) {
          const int64_t g1 = wrap_up(bs[1] + j1, a.K[1]);
          const cplx<T> c01 = cmul(c0, rec[a.coef_off[1] + j1]);
          if (ND == 2) {
            atomic_add_cplx(&grid[grid_addr<CL>(b, c, g0 * a.K[1] + g1, a.C, a.Kprod)], cmul(cconj(c01), val));
          } else {
            for (int j2 = 0; j2 < a.J[2]; ++j2) {
              const int64_t g2 = wrap_up(bs[2] + j2, a.K[2]);
              const cplx<T> c012 = cmul(c01, rec[a.coef_off[2] + j2]);
              atomic_add_cplx(&grid[grid_addr<CL>(b, c, (g0 * a.K[1] + g1) * a.K[2] + g2, a.C, a.Kprod)],
                              cmul(cconj(c012), val));
            }
          }
        }
      }
    }
  }
}

// -----------------------------------------------------------------------------
// 2-D, J = 6, complex64, coil-major point kernels: the generic kernels above with everything the
// common case fixes resolved at compile time (36 unrolled taps, 32-bit indexing, 16-byte weight
// loads, wrap by compare-and-subtract).  Used for a single (batch, coil) row, where one thread per
// point beats the tiled kernels' one warp per point (profiles/r01_g_coil_sweep.log).
// -----------------------------------------------------------------------------
struct Point6 {
  float2 cy[6], cx[6];
  int row[6], col[6];
};
B2N_D Point6 load_point6(const InterpArgs<float> &a, int64_t s) {
  Point6 p;
  const int2 bs = reinterpret_cast<const int2 *>(a.base)[s];
  const float4 *rec = reinterpret_cast<const float4 *>(a.coef + s * 12);
  const int Ky = (int)a.K[0], Kx = (int)a.K[1];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float4 wy = rec[k], wx = rec[3 + k];
    p.cy[2 * k] = make_float2(wy.x, wy.y);
    p.cy[2 * k + 1] = make_float2(wy.z, wy.w);
    p.cx[2 * k] = make_float2(wx.x, wx.y);
    p.cx[2 * k + 1] = make_float2(wx.z, wx.w);
  }
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    const int gy = bs.x + j, gx = bs.y + j;
    p.row[j] = (gy >= Ky ? gy - Ky : gy) * Kx;
    p.col[j] = gx >= Kx ? gx - Kx : gx;
  }
  return p;
}

__global__ void __launch_bounds__(128) k_fwd_point6_2d(InterpArgs<float> a, const float2 *__restrict__ grid,
                                                       float2 *__restrict__ kdata) {
  const int64_t total = a.n_traj * a.M;
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  griddep_launch();
  if (s >= total) return;
  const Point6 p = load_point6(a, s);  // plan records: independent of the kernel that produced the grid
  const int64_t m = a.perm[s];
  const int64_t rows = a.n_traj == 1 ? a.B * a.C : a.C;
  griddep_wait();
  for (int64_t r = blockIdx.y; r < rows; r += gridDim.y) {
    const int64_t bc = a.n_traj == 1 ? r : (s / a.M) * a.C + r;
    const float2 *g = grid + bc * a.Kprod;
    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
    for (int jy = 0; jy < 6; ++jy) {
      float2 v[6];
#pragma unroll
      for (int jx = 0; jx < 6; ++jx) v[jx] = __ldg(&g[p.row[jy] + p.col[jx]]);
      float2 line = make_float2(0.f, 0.f);
#pragma unroll
      for (int jx = 0; jx < 6; ++jx) {
        line.x = fmaf(p.cx[jx].x, v[jx].x, fmaf(-p.cx[jx].y, v[jx].y, line.x));
        line.y = fmaf(p.cx[jx].x, v[jx].y, fmaf(p.cx[jx].y, v[jx].x, line.y));
      }
      acc.x = fmaf(p.cy[jy].x, line.x, fmaf(-p.cy[jy].y, line.y, acc.x));
      acc.y = fmaf(p.cy[jy].x, line.y, fmaf(p.cy[jy].y, line.x, acc.y));
    }
    kdata[bc * a.M + m] = acc;  // the fftshift phase is folded into the dim-0 weights
  }
}

__global__ void __launch_bounds__(128) k_adj_point6_2d(InterpArgs<float> a, const float2 *__restrict__ kdata,
                                                       float2 *__restrict__ grid) {
  const int64_t total = a.n_traj * a.M;
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= total) return;
  const Point6 p = load_point6(a, s);
  const int64_t m = a.perm[s];
  const int64_t rows = a.n_traj == 1 ? a.B * a.C : a.C;
  for (int64_t r = blockIdx.y; r < rows; r += gridDim.y) {
    const int64_t bc = a.n_traj == 1 ? r : (s / a.M) * a.C + r;
    const float2 val = kdata[bc * a.M + m];
    float2 *g = grid + bc * a.Kprod;
#pragma unroll
    for (int jy = 0; jy < 6; ++jy) {
      float2 u;  // conj(cy) * val
      u.x = fmaf(p.cy[jy].x, val.x, p.cy[jy].y * val.y);
      u.y = fmaf(p.cy[jy].x, val.y, -p.cy[jy].y * val.x);
#pragma unroll
      for (int jx = 0; jx < 6; ++jx) {
        float2 w;  // conj(cx) * u
        w.x = fmaf(p.cx[jx].x, u.x, p.cx[jx].y * u.y);
        w.y = fmaf(p.cx[jx].x, u.y, -p.cx[jx].y * u.x);
        atomicAdd(&g[p.row[jy] + p.col[jx]], w);  // one 8-byte L2 reduction
      }
    }
  }
}

static bool point6_eligible(const b2n_geom *g, const b2n_points *p, int layout) {
  return p && g->dtype == B2N_C64 && g->ndim == 2 && layout == B2N_COIL_MAJOR && g->numpoints[0] == 6 &&
         g->numpoints[1] == 6 && g->grid_size[0] >= 6 && g->grid_size[1] >= 6 &&
         g->grid_size[0] * g->grid_size[1] < ((int64_t)1 << 31) && p->n_points > 0;
}

// -----------------------------------------------------------------------------
// generic adjoint, sorted (deterministic) mode: one thread per (grid cell, row)
// gathers, in a fixed order, every sample whose footprint covers the cell:
// offsets row-major, samples of one base cell in plan order (stable cell sort =>
// original sample order).  No atomics, each cell is written exactly once.
// -----------------------------------------------------------------------------
B2N_D int64_t wrap_down(int64_t g, int64_t K) {
  if (g >= 0) return g;
  g %= K;
  return g < 0 ? g + K : g;
}

// Deterministic adjoint: one thread per grid cell gathers, in plan order, the samples whose footprint covers the
// cell.  The neighbour walk (36 / 216 CSR lookups + weights) is done once for a block of RB coils held in
// registers; the per-(cell, coil) summation order is fixed, so the result is bit-reproducible.
template <typename T, int ND, bool CL, int RB>
__global__ void __launch_bounds__(128) k_adj_sorted_generic(InterpArgs<T> a, const cplx<T> *__restrict__ kdata,
                                                            cplx<T> *__restrict__ grid) {
  const int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= a.Kprod) return;
  int64_t g[B2N_MAX_DIMS] = {0, 0, 0};
  {
    int64_t rem = cell;
    for (int d = ND - 1; d >= 0; --d) {
      g[d] = rem % a.K[d];
      rem /= a.K[d];
    }
  }
  const int64_t ncb = (a.C + RB - 1) / RB, nblk = a.B * ncb;  // row blocks = (batch element, RB coils)
  for (int64_t q = blockIdx.y; q < nblk; q += gridDim.y) {
    const int64_t b = q / ncb, c0 = (q - b * ncb) * RB;
    const int nc = (int)(a.C - c0 < RB ? a.C - c0 : RB);
    const int64_t t = a.n_traj == 1 ? 0 : b;
    const int32_t *cs = a.cell_start + t * a.tiling.n_cells;
    const cplx<T> *rows = kdata + (b * a.C + c0) * a.M;
    cplx<T> acc[RB];
#pragma unroll
    for (int k = 0; k < RB; ++k) acc[k] = {T(0), T(0)};
    for (int j0 = 0; j0 < a.J[0]; ++j0) {
      const int64_t b0 = wrap_down(g[0] - j0, a.K[0]);
      for (int j1 = 0; j1 < (ND > 1 ? a.J[1] : 1); ++j1) {
        const int64_t b1 = ND > 1 ? wrap_down(g[1] - j1, a.K[1]) : 0;
        for (int j2 = 0; j2 < (ND > 2 ? a.J[2] : 1); ++j2) {
          const int64_t b2 = ND > 2 ? wrap_down(g[2] - j2, a.K[2]) : 0;
          const int64_t bc[B2N_MAX_DIMS] = {b0, b1, b2};
          const int64_t key = tiled_cell(a.tiling, bc);
          const int32_t lo = cs[key], hi = cs[key + 1];
          for (int32_t s = lo; s < hi; ++s) {
            const cplx<T> *rec = a.coef + (int64_t)s * a.coef_stride;
            cplx<T> cc = rec[j0];
            if (ND > 1) cc = cmul(cc, rec[a.coef_off[1] + j1]);
            if (ND > 2) cc = cmul(cc, rec[a.coef_off[2] + j2]);
            cc = cconj(cc);
            const cplx<T> *v = rows + a.perm[s];
#pragma unroll
            for (int k = 0; k < RB; ++k)
              if (k < nc) cmac(acc[k], cc, v[(int64_t)k * a.M]);
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < RB; ++k)
      if (k < nc) grid[grid_addr<CL>(b, c0 + k, cell, a.C, a.Kprod)] = acc[k];
  }
}

// ---- dispatch -----------------------------------------------------------------
template <typename T, int ND, bool CL>
static int launch_forward(const InterpArgs<T> &a, const void *grid, void *kdata, cudaStream_t st) {
  const int64_t total = a.n_traj * a.M;
  if (total == 0) return 0;
  const int64_t rows = a.n_traj == 1 ? a.B * a.C : a.C;
  dim3 block(128), gridDim((unsigned)ceil_div(total, 128), (unsigned)(rows < 65535 ? rows : 65535));
  k_fwd_generic<T, ND, CL><<<gridDim, block, 0, st>>>(a, (const cplx<T> *)grid, (cplx<T> *)kdata);
  B2N_LAUNCH_OK("k_fwd_generic");
  return 0;
}

template <typename T, int ND, bool CL>
static int launch_adjoint(const InterpArgs<T> &a, const void *kdata, int mode, void *grid, cudaStream_t st) {
  const int64_t total = a.n_traj * a.M;
  const int64_t rows_all = a.B * a.C;
  if (mode == B2N_ADJ_ATOMIC) {
    B2N_CUDA_OK(cudaMemsetAsync(grid, 0, sizeof(cplx<T>) * (size_t)(rows_all * a.Kprod), st));
    if (total == 0) return 0;
    const int64_t rows = a.n_traj == 1 ? rows_all : a.C;
    dim3 block(128), gridDim((unsigned)ceil_div(total, 128), (unsigned)(rows < 65535 ? rows : 65535));
    k_adj_atomic_generic<T, ND, CL><<<gridDim, block, 0, st>>>(a, (const cplx<T> *)kdata, (cplx<T> *)grid);
    B2N_LAUNCH_OK("k_adj_atomic_generic");
    return 0;
  }
  // 2-D: 8 coils per thread (2.2 -> 1.5 ms at BASELINE config 2); 3-D: one (216 neighbour cells keep a thread busy
  // long enough, and blocking coils there measured slower: 568 -> 961 ms at config 4)
  constexpr int RB = ND == 3 ? 1 : 8;
  const int64_t nblk = a.B * ceil_div(a.C, RB);
  dim3 block(128), gridDim((unsigned)ceil_div(a.Kprod, 128), (unsigned)(nblk < 65535 ? nblk : 65535));
  k_adj_sorted_generic<T, ND, CL, RB><<<gridDim, block, 0, st>>>(a, (const cplx<T> *)kdata, (cplx<T> *)grid);
  B2N_LAUNCH_OK("k_adj_sorted_generic");
  return 0;
}

template <typename T>
static int forward_t(const b2n_geom *g, const b2n_points *p, const void *grid, int64_t B, int64_t C, int layout,
                     void *kdata, cudaStream_t st) {
  InterpArgs<T> a;
  int rc = make_args<T>(g, p, B, C, &a);
  if (rc) return rc;
  if constexpr (std::is_same<T, float>::value) {
    if (point6_eligible(g, p, layout)) {
      const int64_t total = a.n_traj * a.M, rows = a.n_traj == 1 ? a.B * a.C : a.C;
      dim3 gd((unsigned)ceil_div(total, 128), (unsigned)(rows < 65535 ? rows : 65535));
      B2N_CUDA_OK(launch_pdl(k_fwd_point6_2d, gd, dim3(128), 0, st, a, (const float2 *)grid, (float2 *)kdata));
      B2N_LAUNCH_OK("k_fwd_point6_2d");
      return 0;
    }
  }
  const bool cl = layout == B2N_CHANNEL_LAST;
  switch (g->ndim) {
    case 1: return cl ? launch_forward<T, 1, true>(a, grid, kdata, st) : launch_forward<T, 1, false>(a, grid, kdata, st);
    case 2: return cl ? launch_forward<T, 2, true>(a, grid, kdata, st) : launch_forward<T, 2, false>(a, grid, kdata, st);
    default: return cl ? launch_forward<T, 3, true>(a, grid, kdata, st) : launch_forward<T, 3, false>(a, grid, kdata, st);
  }
}

template <typename T>
static int adjoint_t(const b2n_geom *g, const b2n_points *p, const void *kdata, int64_t B, int64_t C, int layout,
                     int mode, void *grid, cudaStream_t st) {
  InterpArgs<T> a;
  int rc = make_args<T>(g, p, B, C, &a);
  if (rc) return rc;
  if constexpr (std::is_same<T, float>::value) {
    if (mode == B2N_ADJ_ATOMIC && point6_eligible(g, p, layout)) {
      B2N_CUDA_OK(cudaMemsetAsync(grid, 0, sizeof(float2) * (size_t)(a.B * a.C * a.Kprod), st));
      const int64_t total = a.n_traj * a.M, rows = a.n_traj == 1 ? a.B * a.C : a.C;
      dim3 gd((unsigned)ceil_div(total, 128), (unsigned)(rows < 65535 ? rows : 65535));
      k_adj_point6_2d<<<gd, 128, 0, st>>>(a, (const float2 *)kdata, (float2 *)grid);
      B2N_LAUNCH_OK("k_adj_point6_2d");
      return 0;
    }
  }
  const bool cl = layout == B2N_CHANNEL_LAST;
  switch (g->ndim) {
    case 1: return cl ? launch_adjoint<T, 1, true>(a, kdata, mode, grid, st) : launch_adjoint<T, 1, false>(a, kdata, mode, grid, st);
    case 2: return cl ? launch_adjoint<T, 2, true>(a, kdata, mode, grid, st) : launch_adjoint<T, 2, false>(a, kdata, mode, grid, st);
    default: return cl ? launch_adjoint<T, 3, true>(a, kdata, mode, grid, st) : launch_adjoint<T, 3, false>(a, kdata, mode, grid, st);
  }
}

int tiled_forward(const b2n_geom *g, const b2n_points *p, const void *grid, int64_t B, int64_t C, int layout,
                  void *kdata, cudaStream_t st);
int tiled_adjoint(const b2n_geom *g, const b2n_points *p, const void *kdata, int64_t B, int64_t C, int layout,
                  void *grid, cudaStream_t st);
int tiled_adjoint_cl(const b2n_geom *g, const b2n_points *p, const void *kdata, int64_t B, int64_t C, int layout,
                     void *grid, cudaStream_t st);
int tiled3_forward(const b2n_geom *g, const b2n_points *p, const void *grid, int64_t B, int64_t C, int layout,
                   void *kdata, cudaStream_t st);
int tiled3_adjoint(const b2n_geom *g, const b2n_points *p, const void *kdata, int64_t B, int64_t C, int layout,
                   void *grid, cudaStream_t st);

size_t tiled_adjoint_ordered_bytes(const b2n_geom *g, const b2n_points *p, int64_t B, int64_t C, int layout);
size_t tiled3_adjoint_ordered_bytes(const b2n_geom *g, const b2n_points *p, int64_t B, int64_t C, int layout);
int tiled3_adjoint_ordered(const b2n_geom *g, const b2n_points *p, const void *kdata, int64_t B, int64_t C, int layout,
                           void *scratch, size_t scratch_bytes, void *grid, cudaStream_t st);
int tiled_adjoint_ordered(const b2n_geom *g, const b2n_points *p, const void *kdata, int64_t B, int64_t C, int layout,
                          void *scratch, size_t scratch_bytes, void *grid, cudaStream_t st);

size_t own_adjoint_bytes(const b2n_geom *g, const b2n_points *p, int64_t B, int64_t C, int layout, int64_t n_slots,
                         size_t *zero_bytes);
int own_adjoint(const b2n_geom *g, const b2n_points *p, const void *kdata, int64_t B, int64_t C, int layout,
                void *scratch, size_t scratch_bytes, void *grid, cudaStream_t st);
extern int g_adj_owned;
extern int g_own_cap;

long long *g_trace_buffer = nullptr;
int64_t g_trace_capacity = 0;
extern int g_adj_rowwarp;
extern int g_fwd_chunk;
extern int g_adj_chunk;
extern int g_fast_fft;
int g_pdl = 1;
int g_prefetch = 19;
int g_zero_kernel = 1;
extern int g_fft_stream;
extern int g_peer_form;
static int g_options[B2N_OPT_COUNT] = {1, 0, 0, 0, 1, 1, 19, 1, 0, 1, 0};

// With very few 2-D (batch, coil) rows most coil lanes of a tiled gather CTA idle while its per-point cost stays the
// same: the one-thread-per-point kernel (k_fwd_point6_2d) wins up to 3 rows (16 vs 29 us for one row, 30 vs 41 us for
// three at BASELINE config 1, profiles/r01_g_coil_sweep.log).  The spread always stays on the tiled path, which has its
// own few-coil kernel (k_adj_taps_2d).  B2N_OPT_TILED_KERNELS = 2 forces the tiled kernels everywhere.
static bool use_tiled(const b2n_geom *geom, const b2n_points *pts, int64_t n_batch, int64_t n_coils, bool forward) {
  const int opt = g_options[B2N_OPT_TILED_KERNELS];
  if (!opt || !pts) return false;
  if (opt == 2 || geom->ndim != 2 || !forward) return true;
  return n_batch * n_coils > 3;
}

}  // namespace b2n

using namespace b2n;

extern "C" int b2n_set_option(int option, int value) {
  if (option < 0 || option >= B2N_OPT_COUNT) return fail_arg(B2N_E_ARG, "unknown option %d", option);
  g_options[option] = value;
  if (option == B2N_OPT_ADJ_ROW_OWNERSHIP) g_adj_rowwarp = value;
  if (option == B2N_OPT_FWD_COIL_CHUNK) g_fwd_chunk = value;
  if (option == B2N_OPT_ADJ_COIL_CHUNK) g_adj_chunk = value;
  if (option == B2N_OPT_FAST_FFT) g_fast_fft = value;
  if (option == B2N_OPT_PDL) {
    g_pdl = value != 0;
    g_zero_kernel = value != 2;  // 2: dependent launches, but the adjoint grid is zeroed by cudaMemsetAsync (A/B)
    g_counters_early = value != 3;  // 3: ... but the coil-sum counters are zeroed right before their kernel (A/B)
  }
  if (option == B2N_OPT_FFT_PREFETCH) g_prefetch = value;
  if (option == B2N_OPT_ADJ_OWNED) g_adj_owned = value;
  if (option == B2N_OPT_OWN_CAP) g_own_cap = value;
  if (option == B2N_OPT_FFT_STREAM) g_fft_stream = value;
  if (option == B2N_OPT_PEER_FORM) g_peer_form = value;
  return 0;
}

extern "C" int b2n_set_trace_buffer(void *records_dev, int64_t capacity) {
  g_trace_buffer = (long long *)records_dev;
  g_trace_capacity = records_dev ? capacity : 0;
  return 0;
}

extern "C" int b2n_get_option(int option) {
  if (option < 0 || option >= B2N_OPT_COUNT) return fail_arg(B2N_E_ARG, "unknown option %d", option);
  return g_options[option];
}

extern "C" int b2n_interp_forward(const b2n_geom *geom, const b2n_points *pts, const void *grid_dev, int64_t n_batch,
                                  int64_t n_coils, int grid_layout, void *kdata_dev, void *stream) {
  if (!geom || !grid_dev || !kdata_dev) return fail_arg(B2N_E_ARG, "NULL geom/grid/kdata");
  if (grid_layout != B2N_COIL_MAJOR && grid_layout != B2N_CHANNEL_LAST) return fail_arg(B2N_E_ARG, "bad layout");
  cudaStream_t st = (cudaStream_t)stream;
  if (use_tiled(geom, pts, n_batch, n_coils, true)) {
    int rc = tiled_forward(geom, pts, grid_dev, n_batch, n_coils, grid_layout, kdata_dev, st);
    if (rc != 1) return rc;  // 1 = not eligible, use the generic kernel
    rc = tiled3_forward(geom, pts, grid_dev, n_batch, n_coils, grid_layout, kdata_dev, st);
    if (rc != 1) return rc;
  }
  if (geom->dtype == B2N_C64) return forward_t<float>(geom, pts, grid_dev, n_batch, n_coils, grid_layout, kdata_dev, st);
  if (geom->dtype == B2N_C128) return forward_t<double>(geom, pts, grid_dev, n_batch, n_coils, grid_layout, kdata_dev, st);
  return fail_arg(B2N_E_ARG, "bad dtype");
}

extern "C" int b2n_interp_adjoint(const b2n_geom *geom, const b2n_points *pts, const void *kdata_dev, int64_t n_batch,
                                  int64_t n_coils, int grid_layout, int mode, void *grid_dev, void *stream) {
  if (!geom || !grid_dev || !kdata_dev) return fail_arg(B2N_E_ARG, "NULL geom/grid/kdata");
  if (grid_layout != B2N_COIL_MAJOR && grid_layout != B2N_CHANNEL_LAST) return fail_arg(B2N_E_ARG, "bad layout");
  if (mode != B2N_ADJ_ATOMIC && mode != B2N_ADJ_SORTED) return fail_arg(B2N_E_ARG, "bad adjoint mode %d", mode);
  cudaStream_t st = (cudaStream_t)stream;
  if (use_tiled(geom, pts, n_batch, n_coils, false) && mode == B2N_ADJ_ATOMIC) {
    int rc = tiled_adjoint(geom, pts, kdata_dev, n_batch, n_coils, grid_layout, grid_dev, st);
    if (rc != 1) return rc;
    rc = tiled_adjoint_cl(geom, pts, kdata_dev, n_batch, n_coils, grid_layout, grid_dev, st);
    if (rc != 1) return rc;
    rc = tiled3_adjoint(geom, pts, kdata_dev, n_batch, n_coils, grid_layout, grid_dev, st);
    if (rc != 1) return rc;
  }
  if (geom->dtype == B2N_C64)
    return adjoint_t<float>(geom, pts, kdata_dev, n_batch, n_coils, grid_layout, mode, grid_dev, st);
  if (geom->dtype == B2N_C128)
    return adjoint_t<double>(geom, pts, kdata_dev, n_batch, n_coils, grid_layout, mode, grid_dev, st);
  return fail_arg(B2N_E_ARG, "bad dtype");
}

extern "C" int b2n_interp_adjoint_ordered_bytes(const b2n_geom *geom, const b2n_points *pts, int64_t n_batch,
                                                int64_t n_coils, int grid_layout, size_t *bytes) {
  if (!geom || !pts || !bytes) return fail_arg(B2N_E_ARG, "NULL geom/pts/bytes");
  if (n_batch < 1 || n_coils < 1)
    return fail_arg(B2N_E_ARG, "n_batch=%lld n_coils=%lld", (long long)n_batch, (long long)n_coils);
  *bytes = 0;
  if (g_options[B2N_OPT_TILED_KERNELS]) {
    *bytes = own_adjoint_bytes(geom, pts, n_batch, n_coils, grid_layout, 0, nullptr);
    if (!*bytes) *bytes = tiled_adjoint_ordered_bytes(geom, pts, n_batch, n_coils, grid_layout);
    if (!*bytes) *bytes = tiled3_adjoint_ordered_bytes(geom, pts, n_batch, n_coils, grid_layout);
  }
  return 0;
}

extern "C" int b2n_interp_adjoint_ordered_layout(const b2n_geom *geom, const b2n_points *pts, int64_t n_batch,
                                                 int64_t n_coils, int grid_layout, int64_t n_slots, size_t *bytes,
                                                 size_t *zero_bytes) {
  if (!geom || !pts || !bytes || !zero_bytes) return fail_arg(B2N_E_ARG, "NULL geom/pts/bytes/zero_bytes");
  if (n_batch < 1 || n_coils < 1)
    return fail_arg(B2N_E_ARG, "n_batch=%lld n_coils=%lld", (long long)n_batch, (long long)n_coils);
  *bytes = *zero_bytes = 0;
  if (!g_options[B2N_OPT_TILED_KERNELS]) return 0;
  *bytes = own_adjoint_bytes(geom, pts, n_batch, n_coils, grid_layout, n_slots, zero_bytes);
  if (*bytes) return 0;
  return b2n_interp_adjoint_ordered_bytes(geom, pts, n_batch, n_coils, grid_layout, bytes);
}

extern "C" int b2n_interp_adjoint_ordered(const b2n_geom *geom, const b2n_points *pts, const void *kdata_dev,
                                          int64_t n_batch, int64_t n_coils, int grid_layout, void *scratch_dev,
                                          size_t scratch_bytes, void *grid_dev, void *stream) {
  if (!geom || !pts || !grid_dev || !kdata_dev) return fail_arg(B2N_E_ARG, "NULL geom/pts/grid/kdata");
  {
    const int rc = own_adjoint(geom, pts, kdata_dev, n_batch, n_coils, grid_layout, scratch_dev, scratch_bytes, grid_dev,
                               (cudaStream_t)stream);
    if (rc != 1) return rc;
  }
  if (geom->ndim == 3) {
    const int rc = tiled3_adjoint_ordered(geom, pts, kdata_dev, n_batch, n_coils, grid_layout, scratch_dev, scratch_bytes,
                                          grid_dev, (cudaStream_t)stream);
    return rc == 1 ? fail_arg(B2N_E_UNSUPPORTED, "ordered adjoint: 3-D complex64 J=6 coil-major grids of at least 13 cells only")
                   : rc;
  }
  return tiled_adjoint_ordered(geom, pts, kdata_dev, n_batch, n_coils, grid_layout, scratch_dev, scratch_bytes, grid_dev,
                               (cudaStream_t)stream);
}
