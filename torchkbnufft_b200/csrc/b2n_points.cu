// b2n_points.cu -- trajectory plan: coordinate arithmetic, cell sort, per-point records.
//
// Everything that depends on the trajectory alone is computed here once and kept in
// a caller-owned workspace (b2n_points): the reference redoes it on every call
// (torchkbnufft/_nufft/interp.py:171-177, :129-148, :552-584, :663-686).
#include <cub/device/device_radix_sort.cuh>

#include <stdarg.h>
#include <string.h>

#include <atomic>

#include "b2n_common.cuh"
#include "b2n_math.cuh"
#include "b2n_tiling.cuh"
#include <cub/device/device_scan.cuh>

namespace b2n {

static thread_local char g_error[512] = "";
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int fail_arg(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
  return code;
}

int check_cuda(cudaError_t err, const char *what) {
  if (err == cudaSuccess) return 0;
  set_error("CUDA error %d (%s) at %s", (int)err, cudaGetErrorString(err), what);
  return (int)err;
}

// device-side copy of the geometry in the working precision
template <typename T> struct GeomT {
  int ndim;
  int J[B2N_MAX_DIMS];
  int L[B2N_MAX_DIMS];
  int coef_off[B2N_MAX_DIMS];
  int coef_stride;
  int64_t K[B2N_MAX_DIMS];
  int64_t Kprod;
  int64_t table_len[B2N_MAX_DIMS];
  const cplx<T> *table[B2N_MAX_DIMS];
  T n_shift[B2N_MAX_DIMS];
  const T *rtable[B2N_MAX_DIMS];      // real kernel values (b2n_geom.rtable_dev) or NULL
  double table_phase[B2N_MAX_DIMS];   // slope of the tables' linear phase
};

template <typename T> static GeomT<T> make_geom(const b2n_geom *g) {
  GeomT<T> r;
  memset(&r, 0, sizeof(r));
  r.ndim = g->ndim;
  r.Kprod = 1;
  int off = 0;
  for (int d = 0; d < g->ndim; ++d) {
    r.J[d] = g->numpoints[d];
    r.L[d] = g->table_oversamp[d];
    r.K[d] = g->grid_size[d];
    r.Kprod *= g->grid_size[d];
    r.table_len[d] = g->table_len[d];
    r.table[d] = (const cplx<T> *)g->table_dev[d];
    r.n_shift[d] = (T)g->n_shift[d];
    r.rtable[d] = (const T *)g->rtable_dev[d];
    r.table_phase[d] = g->table_phase[d];
    r.coef_off[d] = off;
    off += g->numpoints[d];
  }
  r.coef_stride = off;
  return r;
}

int validate_geom(const b2n_geom *g, bool need_tables) {
  if (!g) return fail_arg(B2N_E_ARG, "geom is NULL");
  if (g->ndim < 1 || g->ndim > B2N_MAX_DIMS) return fail_arg(B2N_E_RANGE, "ndim=%d not in 1..3", g->ndim);
  if (g->dtype != B2N_C64 && g->dtype != B2N_C128) return fail_arg(B2N_E_ARG, "bad dtype %d", g->dtype);
  for (int d = 0; d < g->ndim; ++d) {
    if (g->grid_size[d] < 1) return fail_arg(B2N_E_ARG, "grid_size[%d]=%lld", d, (long long)g->grid_size[d]);
    if (g->numpoints[d] < 1 || g->numpoints[d] > B2N_MAX_NUMPOINTS)
      return fail_arg(B2N_E_RANGE, "numpoints[%d]=%d not in 1..%d", d, g->numpoints[d], B2N_MAX_NUMPOINTS);
    if (g->table_oversamp[d] < 1) return fail_arg(B2N_E_ARG, "table_oversamp[%d]=%d", d, g->table_oversamp[d]);
    if (need_tables) {
      if (!g->table_dev[d]) return fail_arg(B2N_E_ARG, "table_dev[%d] is NULL", d);
      if (g->table_len[d] < (int64_t)g->numpoints[d] * g->table_oversamp[d] + 1)
        return fail_arg(B2N_E_ARG, "table_len[%d]=%lld shorter than J*L+1", d, (long long)g->table_len[d]);
    }
  }
  return 0;
}

// ---- kernels ----------------------------------------------------------------
template <typename T>
__global__ void k_point_keys(GeomT<T> g, Tiling tl, const T *__restrict__ omega, int64_t M, int64_t total,
                             uint32_t *__restrict__ keys, uint32_t *__restrict__ idx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int64_t t = i / M, m = i - t * M;
  const T *om = omega + t * g.ndim * M + m;
  int64_t cell[B2N_MAX_DIMS] = {0, 0, 0};
  for (int d = 0; d < g.ndim; ++d) {
    T tm;
    int64_t base;
    locate<T>(om[d * M], g.K[d], g.J[d], tm, base);
    cell[d] = wrap_cell(base, g.K[d]);
  }
  keys[i] = (uint32_t)(t * tl.n_cells + tiled_cell(tl, cell));
  idx[i] = (uint32_t)i;
}

// sub-problems: per (trajectory, tile) the sorted points are cut into chunks of <= cap
__global__ void k_tile_chunks(const int32_t *__restrict__ cell_start, int64_t n_tiles_all, int TT, int cap,
                              int32_t *__restrict__ n_chunks) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tiles_all) return;
  const int32_t n = cell_start[(t + 1) * TT] - cell_start[t * TT];
  n_chunks[t] = (n + cap - 1) / cap;
}

__global__ void k_fill_subs(const int32_t *__restrict__ cell_start, const int32_t *__restrict__ n_chunks,
                            const int32_t *__restrict__ offsets, int64_t n_tiles_all, int TT, int cap,
                            int32_t *__restrict__ sub_tile, int32_t *__restrict__ sub_start,
                            int32_t *__restrict__ sub_count, int32_t *__restrict__ n_sub) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tiles_all) return;
  const int32_t lo = cell_start[t * TT], n = cell_start[(t + 1) * TT] - lo;
  const int32_t off = offsets[t], nc = n_chunks[t];
  for (int32_t j = 0; j < nc; ++j) {
    sub_tile[off + j] = (int32_t)t;
    sub_start[off + j] = lo + j * cap;
    sub_count[off + j] = min(cap, n - j * cap);
  }
  if (t == n_tiles_all - 1) *n_sub = off + nc;
}

template <typename T>
__global__ void k_point_records(GeomT<T> g, const T *__restrict__ omega, int64_t M, int64_t total,
                                const uint32_t *__restrict__ sorted_idx, int32_t *__restrict__ perm,
                                int32_t *__restrict__ inv_perm, int32_t *__restrict__ base_out, cplx<T> *__restrict__ coef,
                                cplx<T> *__restrict__ phase, float *__restrict__ hw, int hw_stride,
                                float2 *__restrict__ fac, unsigned char *__restrict__ exc_flag) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= total) return;
  const int64_t i = sorted_idx[s];
  const int64_t t = i / M, m = i - t * M;
  const T *om = omega + t * g.ndim * M + m;
  perm[s] = (int32_t)m;
  inv_perm[i] = (int32_t)s;
  T omv[B2N_MAX_DIMS];
  for (int d = 0; d < g.ndim; ++d) omv[d] = om[d * M];
  const T arg = phase_arg<T>(omv, g.ndim, g.n_shift);
  T sn, cs;
  sincos(arg, &sn, &cs);
  cplx<T> ph;
  ph.x = cs;
  ph.y = sn;
  phase[s] = ph;
  double fac_arg = 0.0;  // sum_d table_phase[d] * (x_d(neighbour 0) + wrapped base_d)
  bool exc = false;      // a neighbour whose table index is not that of neighbour 0 minus j L (see b2n_points.own_exc)
  for (int d = 0; d < g.ndim; ++d) {
    T tm;
    int64_t base;
    locate<T>(omv[d], g.K[d], g.J[d], tm, base);
    base_out[s * g.ndim + d] = (int32_t)wrap_cell(base, g.K[d]);
    if (hw) {
      // real-weight records (2-D / 3-D, J = 6 on every axis; see b2n_geom.rtable_dev): neighbour j of this point has the
      // table index t0 - j L, so its complex weight is r(t0 - j L) exp(-1i p (x0 - j)); the part that does not
      // depend on j goes into the point's factor
      const int64_t t0 = table_index<T>(tm, base, g.J[d], g.L[d]);
      fac_arg += g.table_phase[d] * ((double)t0 / g.L[d] - 0.5 * g.J[d] + (double)wrap_cell(base, g.K[d]));
      for (int j = 0; j < 6; ++j) {
        int64_t ti = table_index<T>(tm, base + j, g.J[d], g.L[d]);
        exc |= ti != t0 - (int64_t)j * g.L[d] || ti < 0 || ti >= g.table_len[d];
        if (ti < 0) ti += g.table_len[d];
        ti = ti < 0 ? 0 : (ti >= g.table_len[d] ? g.table_len[d] - 1 : ti);
        const float r = (float)g.rtable[d][ti];
        hw[s * hw_stride + 6 * d + j] = r;
        if (d == 2) hw[s * hw_stride + 18 + j] = -r;  // 3-D: the staging copies cannot negate (wrap sign), so both
      }
    }
    cplx<T> *rec = coef + s * g.coef_stride + g.coef_off[d];
    for (int j = 0; j < g.J[d]; ++j) {
      int64_t ti = table_index<T>(tm, base + j, g.J[d], g.L[d]);
      if (ti < 0) ti += g.table_len[d];  // torch indexing wraps a negative index
      ti = ti < 0 ? 0 : (ti >= g.table_len[d] ? g.table_len[d] - 1 : ti);
      const cplx<T> tv = g.table[d][ti];
      // the fftshift phase multiplies every weight of the point: fold it into dimension 0, so
      // the gather needs no epilogue and the spread no prologue (conj of the product is used)
      rec[j] = d == 0 ? cmul(tv, ph) : tv;
    }
  }
  if (hw) {
    exc_flag[s] = exc ? 1 : 0;
    if (exc)
      for (int k = 0; k < hw_stride; ++k) hw[s * hw_stride + k] = 0.f;
  }
  if (fac) {
    // conj(fftshift phase * exp(-1i fac_arg)): the angle reaches thousands of radians, so it is reduced in double
    double sn2, cs2;
    sincos(fac_arg - (double)arg, &sn2, &cs2);
    fac[s] = make_float2((float)cs2, (float)sn2);
  }
}

// sorted slots of the exception points, ascending: flags counted per block, then every block writes its flagged
// slots behind those of the blocks before it
__global__ void __launch_bounds__(1024) k_own_exc_count(const unsigned char *__restrict__ flag, int64_t total,
                                                        int32_t *__restrict__ block_count) {
  const int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x;
  const int n = __syncthreads_count(i < total && flag[i]);
  if (threadIdx.x == 0) block_count[blockIdx.x] = n;
}

__global__ void __launch_bounds__(1024) k_own_exc_write(const unsigned char *__restrict__ flag, int64_t total,
                                                        const int32_t *__restrict__ block_count,
                                                        int32_t *__restrict__ list, int32_t *__restrict__ count) {
  __shared__ int s_warp[32];
  __shared__ int s_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // slots flagged in the blocks before this one (every block: all of them, for the total)
  int before = 0, all = 0;
  for (int b = tid; b < (int)gridDim.x; b += 1024) {
    const int c = block_count[b];
    all += c;
    if (b < (int)blockIdx.x) before += c;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    before += __shfl_xor_sync(0xffffffffu, before, off);
    all += __shfl_xor_sync(0xffffffffu, all, off);
  }
  if (lane == 0) s_warp[warp] = before;
  __syncthreads();
  if (tid == 0) {
    int t = 0;
    for (int w = 0; w < 32; ++w) t += s_warp[w];
    s_base = t;
  }
  __syncthreads();
  if (blockIdx.x == 0) {  // the total, by the same reduction
    __syncthreads();
    if (lane == 0) s_warp[warp] = all;
    __syncthreads();
    if (tid == 0) {
      int t = 0;
      for (int w = 0; w < 32; ++w) t += s_warp[w];
      *count = t;
    }
    __syncthreads();
  }
  const int64_t i = (int64_t)blockIdx.x * 1024 + tid;
  const bool f = i < total && flag[i];
  const unsigned m = __ballot_sync(0xffffffffu, f);
  __syncthreads();
  if (lane == 0) s_warp[warp] = __popc(m);
  __syncthreads();
  int at = s_base;
  for (int w = 0; w < warp; ++w) at += s_warp[w];
  if (f) list[at + __popc(m & ((1u << lane) - 1))] = (int32_t)i;
}

// exp(-1i * slope_d * cell) for every row and column of the grid: the per-cell factor of the real-weight adjoint
__global__ void k_own_cell_phase(int K0, int K1, int K2, double p0, double p1, double p2, float2 *__restrict__ q) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K0 + K1 + K2) return;
  const double a = i < K0 ? p0 * i : (i < K0 + K1 ? p1 * (i - K0) : p2 * (i - K0 - K1));
  double sn, cs;
  sincos(a, &sn, &cs);
  q[i] = make_float2((float)cs, (float)-sn);
}

// Longest-processing-time-first order of the sub-problems: CTAs are dispatched in block
// index order, so the expensive ones (full chunks; tiles on the periodic boundary, which
// cannot use TMA) must come first or they form the tail of every launch (measured:
// profiles/r01_d, per-SM end times 36..59 us before this ordering).
__global__ void k_sub_keys(const int32_t *__restrict__ sub_tile, const int32_t *__restrict__ sub_count,
                           const int32_t *__restrict__ n_sub, int64_t n_slots, Tiling tl, int64_t K0, int64_t K1,
                           int64_t K2, int cap, uint32_t *__restrict__ keys, uint32_t *__restrict__ idx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_slots) return;
  idx[i] = (uint32_t)i;
  if (i >= *n_sub) {
    keys[i] = 0xFFFFFFFFu;
    return;
  }
  int64_t tid = sub_tile[i] % tl.n_tiles;
  const int64_t t2 = tid % tl.nt[2]; tid /= tl.nt[2];
  const int64_t t1 = tid % tl.nt[1]; tid /= tl.nt[1];
  const int64_t t0 = tid;
  // halo of J-1 (<= 15) cells past the tile must stay inside the grid for the TMA path
  const bool interior = (t0 + 1) * tl.T[0] + 6 <= K0 && (tl.ndim < 2 || (t1 + 1) * tl.T[1] + 6 <= K1) &&
                        (tl.ndim < 3 || (t2 + 1) * tl.T[2] + 6 <= K2);
  keys[i] = (interior ? (uint32_t)(cap + 1) : 0u) + (uint32_t)(cap - sub_count[i]);
}

__global__ void k_sub_gather(const uint32_t *__restrict__ order, int64_t n_slots, const int32_t *__restrict__ in_tile,
                             const int32_t *__restrict__ in_start, const int32_t *__restrict__ in_count,
                             int32_t *__restrict__ sub_tile, int32_t *__restrict__ sub_start,
                             int32_t *__restrict__ sub_count) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_slots) return;
  const uint32_t j = order[i];
  sub_tile[i] = in_tile[j];
  sub_start[i] = in_start[j];
  sub_count[i] = in_count[j];
}

// cell_start[c] = first sorted slot whose key >= c  (c in [0, n_cells])
__global__ void k_cell_start(const uint32_t *__restrict__ keys, int64_t total, int64_t n_cells,
                             int32_t *__restrict__ cell_start) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c > n_cells) return;
  int64_t lo = 0, hi = total;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if ((int64_t)keys[mid] < c) lo = mid + 1; else hi = mid;
  }
  cell_start[c] = (int32_t)lo;
}

template <typename T>
__global__ void k_export_indices(GeomT<T> g, const T *__restrict__ omega, int64_t M, int64_t W,
                                 int64_t *__restrict__ arr_ind, int32_t *__restrict__ tab_idx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= W * M) return;
  const int64_t w = i / M, m = i - w * M;
  int j[B2N_MAX_DIMS];
  int64_t rem = w;
  for (int d = g.ndim - 1; d >= 0; --d) {
    j[d] = (int)(rem % g.J[d]);
    rem /= g.J[d];
  }
  int64_t flat = 0;
  for (int d = 0; d < g.ndim; ++d) {
    T tm;
    int64_t base;
    locate<T>(omega[d * M + m], g.K[d], g.J[d], tm, base);
    const int64_t cell = base + j[d];
    if (tab_idx) tab_idx[(w * g.ndim + d) * M + m] = (int32_t)table_index<T>(tm, cell, g.J[d], g.L[d]);
    flat = flat * g.K[d] + wrap_cell(cell, g.K[d]);
  }
  arr_ind[i] = flat;
}


// ---- owner-tile visit lists (output-stationary spread, b2n_interp_own.cu) ----------------------------------------
// The spread kernel gives every own_tile x own_tile OUTPUT tile of the grid to one warp, which keeps the tile in
// registers and visits every point whose J x J footprint intersects it.  The lists are derived from the cell-sorted
// plan without a second sort and without atomics on the visit order: the base cells that can touch a tile form a
// (T+J-1)^2 window; cell_start gives the sorted slots of each window cell, and the window is walked in a fixed order.
constexpr int kOwnBuckets = 16;  // LPT buckets of the work items by size (+ one for empty tiles)

struct OwnGeom {
  int Ky, Kx, nty, ntx, J, Ty, Tx, cap;  // Ty x Tx cells per output tile
  int neg_y, neg_x;  // 1: a footprint that wraps around this axis flips the sign of the real-weight form
  int64_t n_traj, n_own_tiles;  // tiles per trajectory
  Tiling tl;                    // tiling of the cell-sorted plan (cell_start index)
};

B2N_D int own_wrap(int v, int K) { return v < 0 ? v + K : (v >= K ? v - K : v); }

// sorted-slot range of window cell w (row-major over (th+J-1) x (tw+J-1)) of tile (y0, x0) of trajectory traj
B2N_D void own_window_cell(const OwnGeom &g, const int32_t *__restrict__ cell_start, int64_t traj, int y0, int x0, int ww,
                           int w, int &s0, int &s1, int &ry, int &rx) {
  const int wy = w / ww, wx = w - wy * ww;
  ry = wy - (g.J - 1);
  rx = wx - (g.J - 1);
  int64_t cell[B2N_MAX_DIMS] = {own_wrap(y0 + ry, g.Ky), own_wrap(x0 + rx, g.Kx), 0};
  const int64_t idx = traj * g.tl.n_cells + tiled_cell(g.tl, cell);
  s0 = cell_start[idx];
  s1 = cell_start[idx + 1];
}

B2N_D int own_bucket(int size, int cap) { return size <= 0 ? kOwnBuckets : (cap - size) * kOwnBuckets / cap; }

// one warp per output tile: number of visits, number of work items, LPT histogram
__global__ void __launch_bounds__(256) k_own_count(OwnGeom g, const int32_t *__restrict__ cell_start,
                                                   int4 *__restrict__ tiles, int32_t *__restrict__ hist) {
  const int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (t >= g.n_traj * g.n_own_tiles) return;
  const int64_t traj = t / g.n_own_tiles, tid = t - traj * g.n_own_tiles;
  const int ty = (int)(tid / g.ntx), tx = (int)(tid - (int64_t)ty * g.ntx);
  const int y0 = ty * g.Ty, x0 = tx * g.Tx;
  const int th = min(g.Ty, g.Ky - y0), tw = min(g.Tx, g.Kx - x0);
  const int wh = th + g.J - 1, ww = tw + g.J - 1;
  int n = 0;
  for (int w = lane; w < wh * ww; w += 32) {
    int s0, s1, ry, rx;
    own_window_cell(g, cell_start, traj, y0, x0, ww, w, s0, s1, ry, rx);
    n += s1 - s0;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) n += __shfl_xor_sync(0xffffffffu, n, off);
  if (lane == 0) {
    const int nch = n > 0 ? (n + g.cap - 1) / g.cap : 1;
    tiles[t] = make_int4(n, 0, nch, -1);
    if (nch > 1) atomicAdd(&hist[0], nch - 1);  // full chunks
    // the short last chunk of a multi-chunk tile goes with the full ones: it must not be the item that arrives last
    // and merges the tile's partial sums at the very end of the launch
    atomicAdd(&hist[nch > 1 ? 0 : own_bucket(n, g.cap)], 1);
  }
}

// single CTA: exclusive scans over the tiles (first visit, first partial slot), bucket bases, totals; 1024 tiles per
// round with coalesced loads (the next round's are issued before this round's scan), warp shuffles inside the round
__global__ void __launch_bounds__(1024) k_own_scan(int64_t n_tiles_all, int4 *__restrict__ tiles,
                                                   int32_t *__restrict__ hist, int32_t *__restrict__ bucket_base,
                                                   int32_t *__restrict__ bucket_fill, int32_t *__restrict__ counts) {
  __shared__ int s_v[32], s_s[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    int acc = 0;
    for (int b = 0; b <= kOwnBuckets; ++b) {
      bucket_base[b] = acc;
      acc += hist[b];
      bucket_fill[b] = 0;
    }
    counts[0] = acc;
  }
  int run_v = 0, run_s = 0;
  int4 nxt = make_int4(0, 0, 0, -1);
  if (tid < n_tiles_all) nxt = tiles[tid];
  for (int64_t base = 0; base < n_tiles_all; base += 1024) {
    const int64_t t = base + tid;
    int4 ti = nxt;
    nxt = make_int4(0, 0, 0, -1);
    if (t + 1024 < n_tiles_all) nxt = tiles[t + 1024];
    const int v = ti.x, sl = ti.z > 1 ? ti.z : 0;
    int iv = v, is = sl;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int a = __shfl_up_sync(0xffffffffu, iv, off), b = __shfl_up_sync(0xffffffffu, is, off);
      if (lane >= off) {
        iv += a;
        is += b;
      }
    }
    if (lane == 31) {
      s_v[warp] = iv;
      s_s[warp] = is;
    }
    __syncthreads();
    int bv = 0, bs = 0, tv = 0, ts = 0;
    for (int w = 0; w < 32; ++w) {
      if (w < warp) {
        bv += s_v[w];
        bs += s_s[w];
      }
      tv += s_v[w];
      ts += s_s[w];
    }
    if (t < n_tiles_all) {
      ti.y = run_v + bv + iv - v;
      ti.w = sl ? run_s + bs + is - sl : -1;
      tiles[t] = ti;
    }
    run_v += tv;
    run_s += ts;
    __syncthreads();
  }
  if (tid == 0) counts[1] = run_s;
}

// one CTA per WORK ITEM (at most cap visits of one tile -- the dense tiles at the centre of a radial trajectory hold
// thousands of visits and would serialise on one CTA): the tile's window (at most 9 x 13 base cells) is scanned into
// shared memory; then every group of 4 lanes takes visits of the item, finds the window cell of each by bisection of
// the cell offsets and writes the four 16-byte parts of its record.
constexpr int kOwnFillThreads = 128, kOwnWinMax = 128;
__global__ void __launch_bounds__(kOwnFillThreads) k_own_fill(OwnGeom g, const int32_t *__restrict__ cell_start,
                                                  const int32_t *__restrict__ perm, const float *__restrict__ hw,
                                                  const int4 *__restrict__ tiles, const int32_t *__restrict__ counts,
                                                  const int4 *__restrict__ items, float4 *__restrict__ visits) {
  __shared__ int s_first[kOwnWinMax + 1], s_s0[kOwnWinMax], s_warp[kOwnFillThreads / 32];
  if ((int)blockIdx.x >= counts[0]) return;
  const int4 item = items[blockIdx.x];
  const int64_t t = item.x;
  const int tid_in = threadIdx.x, lane = tid_in & 31, warp = tid_in >> 5;
  const int64_t traj = t / g.n_own_tiles, tid = t - traj * g.n_own_tiles;
  const int ty = (int)(tid / g.ntx), tx = (int)(tid - (int64_t)ty * g.ntx);
  const int y0 = ty * g.Ty, x0 = tx * g.Tx;
  const int th = min(g.Ty, g.Ky - y0), tw = min(g.Tx, g.Kx - x0);
  const int wh = th + g.J - 1, ww = tw + g.J - 1, nw = wh * ww;  // nw <= 117
  const int4 ti = tiles[t];
  {
    // thread w holds window cell w: exclusive scan of the point counts over the CTA
    int s0 = 0, s1 = 0, ry, rx;
    if (tid_in < nw) own_window_cell(g, cell_start, traj, y0, x0, ww, tid_in, s0, s1, ry, rx);
    const int cnt = s1 - s0;
    int incl = cnt;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int up = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += up;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int before = 0;
    for (int w = 0; w < warp; ++w) before += s_warp[w];
    s_first[tid_in] = before + incl - cnt;
    s_s0[tid_in] = s0;
    if (tid_in == kOwnFillThreads - 1) s_first[kOwnWinMax] = before + incl;  // = ti.x
    __syncthreads();
  }
  // this item's visits of the tile: [chunk * cap, chunk * cap + size)
  const int v_lo = (item.z >> 12) * g.cap, n = v_lo + (item.z & 0xfff), part = tid_in & 3;
#pragma unroll 2
  for (int v = v_lo + (tid_in >> 2); v < n; v += kOwnFillThreads / 4) {
    int lo = 0, hi = nw - 1;  // last window cell whose first visit is <= v (empty cells share their successor's offset)
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (s_first[mid] <= v) lo = mid; else hi = mid - 1;
    }
    const int wy = lo / ww, ry = wy - (g.J - 1), rx = lo - wy * ww - (g.J - 1);
    const int slot = s_s0[lo] + (v - s_first[lo]);
    // 64-byte visit record: the four row weights hy[k] = r_y[k - ry], the eight column weights hx[x] = +-r_x[x - rx]
    // (zero outside the footprint), the sample index.  (ry, rx) are unwrapped: a negative base cell means the
    // footprint reached this tile around the grid's edge, which flips the sign where exp(1i table_phase K) = -1.
    float4 rec;
    if (part == 3) {
      rec = make_float4(__int_as_float(perm[slot]), 0.f, 0.f, 0.f);
    } else {
      const float sgn = (((y0 + ry < 0) & g.neg_y) ^ ((x0 + rx < 0) & g.neg_x)) ? -1.f : 1.f;
      // part 0: rows 0..3 against ry; parts 1, 2: columns 0..3 / 4..7 against rx
      const int j0 = part == 0 ? -ry : 4 * (part - 1) - rx;
      const float *hp = hw + (int64_t)slot * 12 + (part == 0 ? 0 : 6);
      const float sc = part == 0 ? 1.f : sgn;
      float e[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) e[q] = (unsigned)(j0 + q) < 6u ? sc * hp[j0 + q] : 0.f;
      rec = make_float4(e[0], e[1], e[2], e[3]);
    }
    visits[(int64_t)(ti.y + v) * 4 + part] = rec;
  }
}

// one thread per output tile: its work items (chunks of at most cap visits) go into their LPT bucket
__global__ void __launch_bounds__(256) k_own_items(int64_t n_tiles_all, int64_t n_own_tiles, int nt1, int nt2, int ndim,
                                                   int cap, const int4 *__restrict__ tiles,
                                                   const int32_t *__restrict__ bucket_base,
                                                   int32_t *__restrict__ bucket_fill, int4 *__restrict__ items) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tiles_all) return;
  const int4 ti = tiles[t];
  const int64_t tid = t % n_own_tiles;
  // 2-D: tile row << 16 | tile column; 3-D: the tile's row-major index
  const int where = ndim == 2 ? (int)(((tid / nt1) << 16) | (tid % nt1)) : (int)tid;
  for (int j = 0; j < ti.z; ++j) {
    const int size = max(0, min(cap, ti.x - j * cap));
    const int b = ti.z > 1 ? 0 : own_bucket(size, cap);
    const int pos = bucket_base[b] + atomicAdd(&bucket_fill[b], 1);
    items[pos] = make_int4((int)t, ti.y + j * cap, size | (j << 12), where);
  }
}

// ---- owner-tile visit lists, 3-D (4 x 4 x 8 output tiles, k_adj_own_3d) ---------------------------------------------
// Same construction one dimension up: the base cells that can touch a tile form a (4+5) x (4+5) x (8+5) window.  A
// visit is an index record {sorted slot, sample index, footprint origin relative to the tile, sign}; the spread
// kernel forms the window weights from the per-point records while staging (a 64-byte record per visit as in 2-D
// would be 4.4 GB at BASELINE config 4).
struct Own3Geom {
  int K[3], nt[3], T[3], neg[3], cap;
  int64_t n_traj, n_own_tiles;
  Tiling tl;
};
constexpr int kOwn3WinMax = 9 * 9 * 13;

B2N_D void own3_tile(const Own3Geom &g, int64_t t, int64_t &traj, int *o, int *wd, int &tile_id) {
  traj = t / g.n_own_tiles;
  int64_t tid = t - traj * g.n_own_tiles;
  tile_id = (int)tid;
  const int t2 = (int)(tid % g.nt[2]);
  tid /= g.nt[2];
  const int t1 = (int)(tid % g.nt[1]), t0 = (int)(tid / g.nt[1]);
  o[0] = t0 * g.T[0];
  o[1] = t1 * g.T[1];
  o[2] = t2 * g.T[2];
  for (int d = 0; d < 3; ++d) wd[d] = min(g.T[d], g.K[d] - o[d]) + 5;
}

// sorted-slot range of window cell w (row-major over wd[0] x wd[1] x wd[2])
B2N_D void own3_window_cell(const Own3Geom &g, const int32_t *__restrict__ cell_start, int64_t traj, const int *o,
                            const int *wd, int w, int &s0, int &s1, int *r) {
  const int w2 = w % wd[2], w01 = w / wd[2], w1 = w01 % wd[1], w0 = w01 / wd[1];
  r[0] = w0 - 5;
  r[1] = w1 - 5;
  r[2] = w2 - 5;
  int64_t cell[B2N_MAX_DIMS] = {own_wrap(o[0] + r[0], g.K[0]), own_wrap(o[1] + r[1], g.K[1]),
                                own_wrap(o[2] + r[2], g.K[2])};
  const int64_t idx = traj * g.tl.n_cells + tiled_cell(g.tl, cell);
  s0 = cell_start[idx];
  s1 = cell_start[idx + 1];
}

// one warp per output tile: number of visits, number of work items, LPT histogram
__global__ void __launch_bounds__(256) k_own3_count(Own3Geom g, const int32_t *__restrict__ cell_start,
                                                    int4 *__restrict__ tiles, int32_t *__restrict__ hist) {
  const int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (t >= g.n_traj * g.n_own_tiles) return;
  int64_t traj;
  int o[3], wd[3], tile_id;
  own3_tile(g, t, traj, o, wd, tile_id);
  int n = 0;
  for (int w = lane; w < wd[0] * wd[1] * wd[2]; w += 32) {
    int s0, s1, r[3];
    own3_window_cell(g, cell_start, traj, o, wd, w, s0, s1, r);
    n += s1 - s0;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) n += __shfl_xor_sync(0xffffffffu, n, off);
  if (lane == 0) {
    const int nch = n > 0 ? (n + g.cap - 1) / g.cap : 1;
    tiles[t] = make_int4(n, 0, nch, -1);
    if (nch > 1) atomicAdd(&hist[0], nch - 1);
    atomicAdd(&hist[nch > 1 ? 0 : own_bucket(n, g.cap)], 1);
  }
}

// one CTA per work item: index records of its visits (window order)
__global__ void __launch_bounds__(256) k_own3_fill(Own3Geom g, const int32_t *__restrict__ cell_start,
                                                   const int32_t *__restrict__ perm, const int4 *__restrict__ tiles,
                                                   const int32_t *__restrict__ counts, const int4 *__restrict__ items,
                                                   int4 *__restrict__ visits) {
  __shared__ int s_first[kOwn3WinMax + 1], s_s0[kOwn3WinMax], s_warp[8], s_run;
  if ((int)blockIdx.x >= counts[0]) return;
  const int4 item = items[blockIdx.x];
  const int64_t t = item.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int64_t traj;
  int o[3], wd[3], tile_id;
  own3_tile(g, t, traj, o, wd, tile_id);
  const int nw = wd[0] * wd[1] * wd[2];
  const int4 ti = tiles[t];
  if (tid == 0) s_run = 0;
  __syncthreads();
  for (int w0 = 0; w0 < nw; w0 += 256) {  // exclusive scan of the point counts over the window, 256 cells a round
    const int w = w0 + tid;
    int s0 = 0, s1 = 0, r[3];
    if (w < nw) own3_window_cell(g, cell_start, traj, o, wd, w, s0, s1, r);
    const int cnt = s1 - s0;
    int incl = cnt;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int up = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += up;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int before = s_run;
    for (int k = 0; k < warp; ++k) before += s_warp[k];
    if (w < nw) {
      s_first[w] = before + incl - cnt;
      s_s0[w] = s0;
    }
    __syncthreads();
    if (tid == 255) s_run = before + incl;
    __syncthreads();
  }
  const int v_lo = (item.z >> 12) * g.cap, n = v_lo + (item.z & 0xfff);  // this item's visits of the tile
#pragma unroll 2
  for (int v = v_lo + tid; v < n; v += 256) {
    int lo = 0, hi = nw - 1;  // last window cell whose first visit is <= v
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (s_first[mid] <= v) lo = mid; else hi = mid - 1;
    }
    const int w2 = lo % wd[2], w01 = lo / wd[2], w1 = w01 % wd[1], w0 = w01 / wd[1];
    const int r0 = w0 - 5, r1 = w1 - 5, r2 = w2 - 5;
    const int slot = s_s0[lo] + (v - s_first[lo]);
    // a negative (unwrapped) base cell: the footprint reached this tile around the grid's edge
    const int neg = ((o[0] + r0 < 0) & g.neg[0]) ^ ((o[1] + r1 < 0) & g.neg[1]) ^ ((o[2] + r2 < 0) & g.neg[2]);
    visits[(int64_t)ti.y + v] = make_int4(slot, perm[slot], (r0 + 16) | ((r1 + 16) << 8) | ((r2 + 16) << 16), neg);
  }
}

// Exception points per output tile (b2n_points.own_xt / own_xv): one warp per tile tests every exception point (a few
// thousand at most on real trajectories) for a footprint that reaches the tile, counts, reserves a segment of own_xv
// and writes the pairs in ascending slot order.  The segment's position depends on the order of the reservations, its
// contents do not.
template <int ND>
__global__ void __launch_bounds__(256) k_own_exc_tiles(int K0, int K1, int K2, int nt0, int nt1, int nt2, int64_t n_tiles,
                                                       int64_t n_traj, int64_t M, const int32_t *__restrict__ exc,
                                                       const int32_t *__restrict__ counts, const int32_t *__restrict__ base,
                                                       int2 *__restrict__ xt, int2 *__restrict__ xv, int64_t xcap,
                                                       int32_t *__restrict__ cursor) {
  const int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t traj = t / n_tiles;
  if (traj >= n_traj) return;
  int64_t tid = t - traj * n_tiles;
  const int K[3] = {K0, K1, K2}, nt[3] = {nt0, nt1, nt2};
  int o[3] = {0, 0, 0}, T[3] = {1, 1, 1};
  for (int d = ND - 1; d >= 0; --d) {
    T[d] = d == ND - 1 ? 8 : 4;
    o[d] = (int)(tid % nt[d]) * T[d];
    tid /= nt[d];
  }
  const int n_exc = counts[2];
  // footprint origin relative to the tile along every axis, or "no overlap"
  auto rel = [&](int e, int &packed) {
    const int64_t s = exc[e];
    if (s / M != traj) return false;
    packed = 0;
    for (int d = 0; d < ND; ++d) {
      int r = base[s * ND + d] - o[d];           // wrapped base in [0, K): bring r into [-(J-1), T-1] modulo K
      if (r > T[d] - 1) r -= K[d];
      if (r < -5 || r > min(T[d], K[d] - o[d]) - 1) return false;
      packed |= (r + 16) << (8 * d);
    }
    return true;
  };
  int n = 0;
  for (int e0 = 0; e0 < n_exc; e0 += 32) {
    int packed;
    const bool hit = e0 + lane < n_exc && rel(e0 + lane, packed);
    n += __popc(__ballot_sync(0xffffffffu, hit));
  }
  int first = 0;
  if (lane == 0 && n > 0) first = atomicAdd(cursor, n);
  first = __shfl_sync(0xffffffffu, first, 0);
  if (lane == 0) xt[t] = make_int2(first, n);
  if (n == 0 || (int64_t)first + n > xcap) return;  // over capacity: the fix-up kernel traps on counts[3] > capacity
  int at = first;
  for (int e0 = 0; e0 < n_exc; e0 += 32) {
    int packed = 0;
    const bool hit = e0 + lane < n_exc && rel(e0 + lane, packed);
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (hit) xv[at + __popc(m & ((1u << lane) - 1))] = make_int2(exc[e0 + lane], packed);
    at += __popc(m);
  }
}

// ---- workspace carving --------------------------------------------------------
struct Carve {
  size_t perm, inv_perm, base, coef, phase, cell_start, keys, sub_tile, sub_start, sub_count, n_sub;
  size_t keys_in, idx_in, idx_out, chunks, offsets, tmp_tile, tmp_start, tmp_count, sub_keys, sub_keys_out, sub_idx,
      sub_order, cub, total, cub_bytes;
  size_t own_visits, own_items, own_tiles, own_counts, own_hist, own_hw, own_fac, own_q, own_excf, own_excb, own_exc, own_xt, own_xv;
  int64_t n_own_xv_max;
  bool own_real;
  int64_t n_own_items_max, n_own_tiles;  // per trajectory; 0 = no visit lists for this geometry
  int own_nt[3], own_cap, own_rows;
  int64_t n_sub_max;
  int sub_cap;
  Tiling tiling;
};

static int default_sub_cap(int ndim) { return ndim == 3 ? 128 : 128; }

constexpr int kOwnTileRows = 4, kOwnTileCols = 8;
int g_own_cap = 0;  // visits per work item of the owner-tile spread: 0 = 128 (2-D) / 1024 (3-D); B2N_OPT_OWN_CAP for A/B
// Visit lists are built for what the owner-tile spreads handle: 2-D / 3-D complex64, J = 6, K_d >= 16, tables of the
// reference's form with the real kernel supplied by the caller (own_real).  Tiles: 4 x 8 cells (2-D), 4 x 4 x 8 (3-D).
static int own_tile_edge(int ndim, int d) { return d == ndim - 1 ? kOwnTileCols : kOwnTileRows; }
static bool own_eligible(const b2n_geom *g) {
  if ((g->ndim != 2 && g->ndim != 3) || g->dtype != B2N_C64) return false;
  for (int d = 0; d < g->ndim; ++d)
    if (g->numpoints[d] != 6 || g->grid_size[d] < 16) return false;
  return true;
}
// upper bound on the tiles one footprint touches: ceil((T + J - 1) / T) per axis, one more where the last tile of the
// axis is a partial one of fewer than J - 1 cells (the footprint can then cover it and reach around the grid's edge)
static int own_visits_per_point(const b2n_geom *g) {
  int v = 1;
  for (int d = 0; d < g->ndim; ++d) {
    const int T = own_tile_edge(g->ndim, d), rem = (int)(g->grid_size[d] % T);
    v *= (T + 4) / T + 1 + (rem != 0 && rem < 5 ? 1 : 0);
  }
  return v;
}
// real-weight records: the caller vouches for the tables' form, and the table step per neighbour must be exact
static bool own_real(const b2n_geom *g) {
  if (!own_eligible(g)) return false;
  for (int d = 0; d < g->ndim; ++d) {
    const int L = g->table_oversamp[d];
    if (!g->rtable_dev[d] || (L & (L - 1)) != 0) return false;
  }
  return true;
}

static int sort_bits(int64_t n_keys) {
  int bits = 1;
  while (bits < 32 && ((int64_t)1 << bits) < n_keys) ++bits;
  return bits;
}

static int carve(const b2n_geom *g, int64_t M, int64_t n_traj, Carve *c) {
  int stride = 0;
  for (int d = 0; d < g->ndim; ++d) stride += g->numpoints[d];
  c->tiling = make_tiling(g->ndim, g->grid_size);
  c->sub_cap = default_sub_cap(g->ndim);
  const int64_t total = M * n_traj, n_cells = c->tiling.n_cells * n_traj;
  const int64_t n_tiles_all = c->tiling.n_tiles * n_traj;
  c->n_sub_max = n_tiles_all + total / c->sub_cap + 1;
  if (total >= ((int64_t)1 << 31) - 1 || n_cells >= ((int64_t)1 << 32) - 1)
    return fail_arg(B2N_E_RANGE, "n_traj*M=%lld or n_traj*prod(K)=%lld exceeds the 32-bit plan limit",
                    (long long)total, (long long)n_cells);
  const size_t csz = g->dtype == B2N_C64 ? 8 : 16;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t at = off;
    off = align_up(off + bytes, 256);
    return at;
  };
  c->perm = take(sizeof(int32_t) * total);
  c->inv_perm = take(sizeof(int32_t) * total);
  c->base = take(sizeof(int32_t) * total * g->ndim);
  c->coef = take(csz * total * stride);
  c->phase = take(csz * total);
  c->cell_start = take(sizeof(int32_t) * (n_cells + 1));
  c->keys = take(sizeof(uint32_t) * total);
  c->sub_tile = take(sizeof(int32_t) * c->n_sub_max);
  c->sub_start = take(sizeof(int32_t) * c->n_sub_max);
  c->sub_count = take(sizeof(int32_t) * c->n_sub_max);
  c->n_sub = take(sizeof(int32_t));
  c->chunks = take(sizeof(int32_t) * n_tiles_all);
  c->offsets = take(sizeof(int32_t) * n_tiles_all);
  c->tmp_tile = take(sizeof(int32_t) * c->n_sub_max);
  c->tmp_start = take(sizeof(int32_t) * c->n_sub_max);
  c->tmp_count = take(sizeof(int32_t) * c->n_sub_max);
  c->sub_keys = take(sizeof(uint32_t) * c->n_sub_max);
  c->sub_keys_out = take(sizeof(uint32_t) * c->n_sub_max);
  c->sub_idx = take(sizeof(uint32_t) * c->n_sub_max);
  c->sub_order = take(sizeof(uint32_t) * c->n_sub_max);
  c->keys_in = take(sizeof(uint32_t) * total);
  c->idx_in = take(sizeof(uint32_t) * total);
  c->idx_out = take(sizeof(uint32_t) * total);
  size_t cub_bytes = 0;
  cudaError_t err = cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr,
                                                    (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)total, 0,
                                                    sort_bits(n_cells));
  if (err != cudaSuccess) return check_cuda(err, "cub::DeviceRadixSort::SortPairs(size query)");
  size_t scan_bytes = 0;
  err = cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (const int32_t *)nullptr, (int32_t *)nullptr, (int)n_tiles_all);
  if (err != cudaSuccess) return check_cuda(err, "cub::DeviceScan::ExclusiveSum(size query)");
  if (scan_bytes > cub_bytes) cub_bytes = scan_bytes;
  size_t sub_sort_bytes = 0;
  err = cub::DeviceRadixSort::SortPairs(nullptr, sub_sort_bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr,
                                        (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)c->n_sub_max, 0, 32);
  if (err != cudaSuccess) return check_cuda(err, "cub::DeviceRadixSort::SortPairs(sub-problem size query)");
  if (sub_sort_bytes > cub_bytes) cub_bytes = sub_sort_bytes;
  c->cub_bytes = cub_bytes;
  c->cub = take(cub_bytes + 256);
  c->n_own_tiles = 0;
  c->n_own_items_max = 0;
  const int64_t vpp = own_eligible(g) ? own_visits_per_point(g) : 0;
  c->own_hw = c->own_fac = c->own_q = 0;
  c->own_real = false;
  if (own_real(g) && vpp * total < ((int64_t)1 << 31) - 1) {
    c->own_rows = kOwnTileRows;
    c->n_own_tiles = 1;
    c->own_nt[0] = c->own_nt[1] = c->own_nt[2] = 1;
    for (int d = 0; d < g->ndim; ++d) {
      c->own_nt[d] = (int)ceil_div(g->grid_size[d], own_tile_edge(g->ndim, d));
      c->n_own_tiles *= c->own_nt[d];
    }
    c->own_cap = g_own_cap <= 0 ? (g->ndim == 2 ? 128 : 1024) : (g_own_cap < 8 ? 8 : (g_own_cap > 4095 ? 4095 : g_own_cap));
    c->n_own_items_max = c->n_own_tiles * n_traj + vpp * total / c->own_cap + 1;
    c->own_visits = take((g->ndim == 2 ? 64 : sizeof(int4)) * (size_t)(vpp * total + 1));
    c->own_items = take(sizeof(int4) * (size_t)c->n_own_items_max);
    c->own_tiles = take(sizeof(int4) * (size_t)(c->n_own_tiles * n_traj));
    c->own_counts = take(sizeof(int32_t) * 4);
    c->own_hist = take(sizeof(int32_t) * 3 * (kOwnBuckets + 1));
    c->own_real = true;
    c->own_hw = take(sizeof(float) * (g->ndim == 2 ? 12 : 24) * (size_t)total);
    c->own_fac = take(sizeof(float2) * (size_t)total);
    c->own_q = take(sizeof(float2) * (size_t)(g->grid_size[0] + g->grid_size[1] + (g->ndim == 3 ? g->grid_size[2] : 0)));
    c->own_excf = take((size_t)total);
    c->own_excb = take(sizeof(int32_t) * (size_t)ceil_div(total > 0 ? total : 1, 1024));
    c->own_exc = take(sizeof(int32_t) * (size_t)total);
    c->own_xt = take(sizeof(int2) * (size_t)(c->n_own_tiles * n_traj));
    c->n_own_xv_max = vpp * (total < 262144 ? total : 262144) + 1;
    c->own_xv = take(sizeof(int2) * (size_t)c->n_own_xv_max);
  }
  c->total = off;
  return 0;
}

template <typename T>
static int build_impl(const b2n_geom *geom, const void *omega, int64_t M, int64_t n_traj, char *ws, const Carve &c,
                      b2n_points *out, cudaStream_t st) {
  GeomT<T> g = make_geom<T>(geom);
  const Tiling &tl = c.tiling;
  const int64_t total = M * n_traj, n_cells = tl.n_cells * n_traj, n_tiles_all = tl.n_tiles * n_traj;
  out->n_points = M;
  out->n_traj = n_traj;
  out->ndim = geom->ndim;
  out->dtype = geom->dtype;
  out->coef_stride = g.coef_stride;
  out->sub_cap = c.sub_cap;
  for (int d = 0; d < B2N_MAX_DIMS; ++d) {
    out->tile[d] = tl.T[d];
    out->n_tiles[d] = tl.nt[d];
  }
  out->n_cells = tl.n_cells;
  out->n_sub_max = c.n_sub_max;
  out->sub_tile = (int32_t *)(ws + c.sub_tile);
  out->sub_start = (int32_t *)(ws + c.sub_start);
  out->sub_count = (int32_t *)(ws + c.sub_count);
  out->n_sub = (int32_t *)(ws + c.n_sub);
  out->sub_slot = (int32_t *)(ws + c.sub_order);      // LPT position -> tile-major rank (output of the LPT sort)
  out->tile_sub_start = (int32_t *)(ws + c.offsets);  // exclusive scan of the chunk counts per tile
  out->perm = (int32_t *)(ws + c.perm);
  out->inv_perm = (int32_t *)(ws + c.inv_perm);
  out->base = (int32_t *)(ws + c.base);
  out->coef = ws + c.coef;
  out->phase = ws + c.phase;
  out->cell_start = (int32_t *)(ws + c.cell_start);
  out->keys = (uint32_t *)(ws + c.keys);
  uint32_t *keys_in = (uint32_t *)(ws + c.keys_in), *idx_in = (uint32_t *)(ws + c.idx_in),
           *idx_out = (uint32_t *)(ws + c.idx_out);
  const int threads = 256;
  if (total > 0) {
    k_point_keys<T><<<(unsigned)ceil_div(total, threads), threads, 0, st>>>(g, tl, (const T *)omega, M, total,
                                                                            keys_in, idx_in);
    B2N_LAUNCH_OK("k_point_keys");
    size_t cub_bytes = c.cub_bytes;
    // stable LSD radix sort on the integer cell key: points of one cell keep their original order
    B2N_CUDA_OK(cub::DeviceRadixSort::SortPairs(ws + c.cub, cub_bytes, keys_in, out->keys, idx_in, idx_out, (int)total,
                                                0, sort_bits(n_cells), st));
    k_point_records<T><<<(unsigned)ceil_div(total, threads), threads, 0, st>>>(
        g, (const T *)omega, M, total, idx_out, out->perm, out->inv_perm, out->base, (cplx<T> *)out->coef,
        (cplx<T> *)out->phase, c.own_real ? (float *)(ws + c.own_hw) : nullptr, geom->ndim == 2 ? 12 : 24,
        c.own_real ? (float2 *)(ws + c.own_fac) : nullptr, c.own_real ? (unsigned char *)(ws + c.own_excf) : nullptr);
    B2N_LAUNCH_OK("k_point_records");
  }
  k_cell_start<<<(unsigned)ceil_div(n_cells + 1, threads), threads, 0, st>>>(out->keys, total, n_cells,
                                                                             out->cell_start);
  B2N_LAUNCH_OK("k_cell_start");
  int32_t *chunks = (int32_t *)(ws + c.chunks), *offsets = (int32_t *)(ws + c.offsets);
  k_tile_chunks<<<(unsigned)ceil_div(n_tiles_all, threads), threads, 0, st>>>(out->cell_start, n_tiles_all, tl.TT,
                                                                              c.sub_cap, chunks);
  B2N_LAUNCH_OK("k_tile_chunks");
  size_t scan_bytes = c.cub_bytes;
  B2N_CUDA_OK(cub::DeviceScan::ExclusiveSum(ws + c.cub, scan_bytes, chunks, offsets, (int)n_tiles_all, st));
  int32_t *tmp_tile = (int32_t *)(ws + c.tmp_tile), *tmp_start = (int32_t *)(ws + c.tmp_start),
          *tmp_count = (int32_t *)(ws + c.tmp_count);
  k_fill_subs<<<(unsigned)ceil_div(n_tiles_all, threads), threads, 0, st>>>(
      out->cell_start, chunks, offsets, n_tiles_all, tl.TT, c.sub_cap, tmp_tile, tmp_start, tmp_count, out->n_sub);
  B2N_LAUNCH_OK("k_fill_subs");
  uint32_t *skeys = (uint32_t *)(ws + c.sub_keys), *skeys_out = (uint32_t *)(ws + c.sub_keys_out),
           *sidx = (uint32_t *)(ws + c.sub_idx), *sorder = (uint32_t *)(ws + c.sub_order);
  k_sub_keys<<<(unsigned)ceil_div(c.n_sub_max, threads), threads, 0, st>>>(tmp_tile, tmp_count, out->n_sub, c.n_sub_max,
                                                                           tl, g.K[0], g.ndim > 1 ? g.K[1] : 1,
                                                                           g.ndim > 2 ? g.K[2] : 1, c.sub_cap, skeys, sidx);
  B2N_LAUNCH_OK("k_sub_keys");
  size_t sort_bytes = c.cub_bytes;
  B2N_CUDA_OK(cub::DeviceRadixSort::SortPairs(ws + c.cub, sort_bytes, skeys, skeys_out, sidx, sorder, (int)c.n_sub_max, 0,
                                              32, st));
  k_sub_gather<<<(unsigned)ceil_div(c.n_sub_max, threads), threads, 0, st>>>(sorder, c.n_sub_max, tmp_tile, tmp_start,
                                                                             tmp_count, out->sub_tile, out->sub_start,
                                                                             out->sub_count);
  B2N_LAUNCH_OK("k_sub_gather");
  out->own_tile = 0;
  out->own_cap = 0;
  out->n_own_tiles[0] = out->n_own_tiles[1] = out->n_own_tiles[2] = 0;
  out->n_own_items_max = 0;
  out->own_visits = out->own_items = out->own_tiles = nullptr;
  out->own_counts = nullptr;
  out->own_hw = nullptr;
  out->own_fac = out->own_q = nullptr;
  out->own_exc = nullptr;
  out->n_own_exc_max = 0;
  out->own_xt = out->own_xv = nullptr;
  out->n_own_xv_max = 0;
  if (c.n_own_tiles > 0) {
    const int nd = geom->ndim;
    // exp(1i table_phase K) = (-1)^(N - 1): the sign a neighbour picks up when it wraps around the grid
    int neg[3] = {0, 0, 0};
    for (int d = 0; d < nd; ++d)
      neg[d] = (int)(llround(geom->table_phase[d] * (double)g.K[d] / 3.14159265358979323846) & 1);
    out->own_hw = (float *)(ws + c.own_hw);
    out->own_fac = ws + c.own_fac;
    out->own_q = ws + c.own_q;
    const int64_t Kq = g.K[0] + g.K[1] + (nd == 3 ? g.K[2] : 0);
    k_own_cell_phase<<<(unsigned)ceil_div(Kq, threads), threads, 0, st>>>(
        (int)g.K[0], (int)g.K[1], nd == 3 ? (int)g.K[2] : 0, geom->table_phase[0], geom->table_phase[1],
        nd == 3 ? geom->table_phase[2] : 0.0, (float2 *)out->own_q);
    B2N_LAUNCH_OK("k_own_cell_phase");
    out->own_exc = (int32_t *)(ws + c.own_exc);
    out->n_own_exc_max = total;
    const unsigned nb = (unsigned)ceil_div(total > 0 ? total : 1, 1024);
    k_own_exc_count<<<nb, 1024, 0, st>>>((const unsigned char *)(ws + c.own_excf), total, (int32_t *)(ws + c.own_excb));
    B2N_LAUNCH_OK("k_own_exc_count");
    k_own_exc_write<<<nb, 1024, 0, st>>>((const unsigned char *)(ws + c.own_excf), total,
                                         (const int32_t *)(ws + c.own_excb), out->own_exc,
                                         (int32_t *)(ws + c.own_counts) + 2);
    B2N_LAUNCH_OK("k_own_exc_write");
    const int64_t nt_all = c.n_own_tiles * n_traj;
    {
      out->own_xt = ws + c.own_xt;
      out->own_xv = ws + c.own_xv;
      out->n_own_xv_max = c.n_own_xv_max;
      int32_t *cnts = (int32_t *)(ws + c.own_counts);
      B2N_CUDA_OK(cudaMemsetAsync(cnts + 3, 0, sizeof(int32_t), st));
      const unsigned gx = (unsigned)ceil_div(c.n_own_tiles * 32 * n_traj, threads);
      if (nd == 2)
        k_own_exc_tiles<2><<<gx, threads, 0, st>>>((int)g.K[0], (int)g.K[1], 1, c.own_nt[0], c.own_nt[1], 1, c.n_own_tiles,
                                                   n_traj, M, out->own_exc, cnts, out->base, (int2 *)out->own_xt,
                                                   (int2 *)out->own_xv, c.n_own_xv_max, cnts + 3);
      else
        k_own_exc_tiles<3><<<gx, threads, 0, st>>>((int)g.K[0], (int)g.K[1], (int)g.K[2], c.own_nt[0], c.own_nt[1],
                                                   c.own_nt[2], c.n_own_tiles, n_traj, M, out->own_exc, cnts, out->base,
                                                   (int2 *)out->own_xt, (int2 *)out->own_xv, c.n_own_xv_max, cnts + 3);
      B2N_LAUNCH_OK("k_own_exc_tiles");
    }
    int4 *tiles = (int4 *)(ws + c.own_tiles);
    int32_t *hist = (int32_t *)(ws + c.own_hist), *bucket_base = hist + (kOwnBuckets + 1),
            *bucket_fill = bucket_base + (kOwnBuckets + 1);
    B2N_CUDA_OK(cudaMemsetAsync(hist, 0, sizeof(int32_t) * 3 * (kOwnBuckets + 1), st));
    if (nd == 2) {
      OwnGeom og;
      og.Ky = (int)g.K[0];
      og.Kx = (int)g.K[1];
      og.nty = c.own_nt[0];
      og.ntx = c.own_nt[1];
      og.J = g.J[0];
      og.Ty = c.own_rows;
      og.Tx = kOwnTileCols;
      og.cap = c.own_cap;
      og.neg_y = neg[0];
      og.neg_x = neg[1];
      og.n_traj = n_traj;
      og.n_own_tiles = c.n_own_tiles;
      og.tl = tl;
      k_own_count<<<(unsigned)ceil_div(nt_all * 32, threads), threads, 0, st>>>(og, out->cell_start, tiles, hist);
      B2N_LAUNCH_OK("k_own_count");
      k_own_scan<<<1, 1024, 0, st>>>(nt_all, tiles, hist, bucket_base, bucket_fill, (int32_t *)(ws + c.own_counts));
      B2N_LAUNCH_OK("k_own_scan");
      k_own_items<<<(unsigned)ceil_div(nt_all, threads), threads, 0, st>>>(nt_all, c.n_own_tiles, c.own_nt[1], 1, 2, c.own_cap,
                                                                           tiles, bucket_base, bucket_fill,
                                                                           (int4 *)(ws + c.own_items));
      B2N_LAUNCH_OK("k_own_items");
      k_own_fill<<<(unsigned)c.n_own_items_max, kOwnFillThreads, 0, st>>>(
          og, out->cell_start, out->perm, (const float *)(ws + c.own_hw), tiles, (const int32_t *)(ws + c.own_counts),
          (const int4 *)(ws + c.own_items), (float4 *)(ws + c.own_visits));
      B2N_LAUNCH_OK("k_own_fill");
    } else {
      Own3Geom og;
      for (int d = 0; d < 3; ++d) {
        og.K[d] = (int)g.K[d];
        og.nt[d] = c.own_nt[d];
        og.T[d] = own_tile_edge(3, d);
        og.neg[d] = neg[d];
      }
      og.cap = c.own_cap;
      og.n_traj = n_traj;
      og.n_own_tiles = c.n_own_tiles;
      og.tl = tl;
      k_own3_count<<<(unsigned)ceil_div(nt_all * 32, threads), threads, 0, st>>>(og, out->cell_start, tiles, hist);
      B2N_LAUNCH_OK("k_own3_count");
      k_own_scan<<<1, 1024, 0, st>>>(nt_all, tiles, hist, bucket_base, bucket_fill, (int32_t *)(ws + c.own_counts));
      B2N_LAUNCH_OK("k_own_scan");
      k_own_items<<<(unsigned)ceil_div(nt_all, threads), threads, 0, st>>>(nt_all, c.n_own_tiles, c.own_nt[1], c.own_nt[2], 3,
                                                                           c.own_cap, tiles, bucket_base, bucket_fill,
                                                                           (int4 *)(ws + c.own_items));
      B2N_LAUNCH_OK("k_own_items");
      k_own3_fill<<<(unsigned)c.n_own_items_max, 256, 0, st>>>(og, out->cell_start, out->perm, tiles,
                                                               (const int32_t *)(ws + c.own_counts),
                                                               (const int4 *)(ws + c.own_items), (int4 *)(ws + c.own_visits));
      B2N_LAUNCH_OK("k_own3_fill");
    }
    out->own_tile = c.own_rows;
    out->own_cap = c.own_cap;
    for (int d = 0; d < 3; ++d) out->n_own_tiles[d] = c.own_nt[d];
    out->n_own_items_max = c.n_own_items_max;
    out->own_visits = ws + c.own_visits;
    out->own_items = ws + c.own_items;
    out->own_tiles = tiles;
    out->own_counts = (int32_t *)(ws + c.own_counts);
  }
  return 0;
}

}  // namespace b2n

using namespace b2n;

extern "C" int b2n_abi_version(void) { return B2N_ABI_VERSION; }
extern "C" int b2n_struct_sizes(size_t *geom_bytes, size_t *points_bytes) {
  if (!geom_bytes || !points_bytes) return fail_arg(B2N_E_ARG, "NULL output");
  *geom_bytes = sizeof(b2n_geom);
  *points_bytes = sizeof(b2n_points);
  return 0;
}
extern "C" const char *b2n_last_error(void) { return g_error; }
extern "C" long long b2n_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int b2n_device_count(void) {
  int n = 0;
  cudaError_t err = cudaGetDeviceCount(&n);
  if (err != cudaSuccess) {
    check_cuda(err, "cudaGetDeviceCount");  // records the reason for b2n_last_error()
    (void)cudaGetLastError();
    return 0;
  }
  return n;
}

extern "C" int b2n_points_workspace_bytes(const b2n_geom *geom, int64_t n_points, int64_t n_traj, size_t *bytes) {
  int rc = validate_geom(geom, false);
  if (rc) return rc;
  if (!bytes || n_points < 0 || n_traj < 1) return fail_arg(B2N_E_ARG, "bad n_points/n_traj/bytes");
  Carve c;
  rc = carve(geom, n_points, n_traj, &c);
  if (rc) return rc;
  *bytes = c.total;
  return 0;
}

extern "C" int b2n_points_build(const b2n_geom *geom, const void *omega_dev, int64_t n_points, int64_t n_traj,
                                void *workspace_dev, size_t workspace_bytes, b2n_points *out, void *stream) {
  int rc = validate_geom(geom, true);
  if (rc) return rc;
  if (!out || n_points < 0 || n_traj < 1) return fail_arg(B2N_E_ARG, "bad n_points/n_traj/out");
  if (n_points > 0 && !omega_dev) return fail_arg(B2N_E_ARG, "omega_dev is NULL");
  Carve c;
  rc = carve(geom, n_points, n_traj, &c);
  if (rc) return rc;
  if (!workspace_dev || workspace_bytes < c.total || ((uintptr_t)workspace_dev & 255))
    return fail_arg(B2N_E_WORKSPACE, "workspace needs %zu bytes, 256-byte aligned (got %zu)", c.total,
                    workspace_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  if (geom->dtype == B2N_C64) return build_impl<float>(geom, omega_dev, n_points, n_traj, (char *)workspace_dev, c, out, st);
  return build_impl<double>(geom, omega_dev, n_points, n_traj, (char *)workspace_dev, c, out, st);
}

extern "C" int b2n_export_indices(const b2n_geom *geom, const void *omega_dev, int64_t n_points, int64_t *arr_ind_dev,
                                  int32_t *tab_idx_dev, void *stream) {
  int rc = validate_geom(geom, false);
  if (rc) return rc;
  if (!omega_dev || !arr_ind_dev || n_points < 0) return fail_arg(B2N_E_ARG, "bad omega/arr_ind/n_points");
  int64_t W = 1;
  for (int d = 0; d < geom->ndim; ++d) W *= geom->numpoints[d];
  if (n_points == 0) return 0;
  const int threads = 256;
  cudaStream_t st = (cudaStream_t)stream;
  if (geom->dtype == B2N_C64) {
    GeomT<float> g = make_geom<float>(geom);
    k_export_indices<float><<<(unsigned)ceil_div(W * n_points, threads), threads, 0, st>>>(
        g, (const float *)omega_dev, n_points, W, arr_ind_dev, tab_idx_dev);
  } else {
    GeomT<double> g = make_geom<double>(geom);
    k_export_indices<double><<<(unsigned)ceil_div(W * n_points, threads), threads, 0, st>>>(
        g, (const double *)omega_dev, n_points, W, arr_ind_dev, tab_idx_dev);
  }
  B2N_LAUNCH_OK("k_export_indices");
  return 0;
}
