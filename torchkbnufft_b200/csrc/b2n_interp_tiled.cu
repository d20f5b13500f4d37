// b2n_interp_tiled.cu -- shared-memory tiled gather / spread kernels (complex64, 2-D, J=6).
//
// Measured motivation (profiles/r01_a..c): the per-point kernels pull every point's J^d-cell
// footprint through L2 (590 MB for BASELINE config 2 against a 52 MB grid) and sit on the
// L2->SM bandwidth; earlier tiled versions were bound first by the latency of dependent
// global loads, then by instruction issue (index arithmetic of element-wise tile staging).
//
// Design:
//   * a CTA owns one sub-problem of the trajectory plan: <= sub_cap consecutive points whose
//     base cell lies in one 16x16 tile.  The tile plus its J-1 halo, for 16 coils, is staged
//     in shared memory coil-major [coil][21 rows][22 cols] by TMA (two 8-plane
//     cp.async.bulk.tensor boxes on an mbarrier); boundary tiles, which need the periodic
//     wrap TMA cannot express, and odd/unaligned grids use a cp.async element path;
//   * a warp works on ONE point with lanes = (coil, cell slot).  With the plane stride
//     21*22 = 462 (= 14 mod 16 in 8-byte banks) and lanes ordered (coil_hi, x parity,
//     coil_lo) every half-warp access of two x-adjacent cells x 8 coils hits 16 distinct
//     bank pairs: conflict-free without padding, which is what lets TMA write the tile;
//   * per-point plan records (separable weights cy[jy] -- with the fftshift phase folded in --
//     and cx[jx], base cell, original sample index) are contiguous in plan order and are
//     bulk-staged with cp.async; nothing in the hot loops reads global memory;
//   * forward: warps take different points; results are stored straight to the caller's
//     (B, C, M) array (8-byte scattered stores, merged in L2);
//   * adjoint: k-space samples are gathered with cp.async one round ahead (sample indices two
//     rounds ahead), each of the 8 warps owns the tile rows r with r mod 8 == warp, so the
//     accumulation is a plain shared-memory read-modify-write with a fixed per-cell order
//     (shared float atomics are CAS loops on sm_100a); the tile is merged into the global
//     grid by TMA reduce-add (cp.reduce.async.bulk.tensor ... .add, FP32) -- element-wise
//     RED.ADD.F32x2 on boundary tiles.
//
// reference loops replaced: torchkbnufft/_nufft/interp.py:185-203 and :689-724.
#include "b2n_tiled_common.cuh"

namespace b2n {

constexpr int kTile = 16;   // must match make_tiling() for ndim == 2
constexpr int kWarps = 8;   // warps per CTA (== row-ownership modulus of the adjoint)
constexpr int kThreads = kWarps * 32;
// forward gather: its own CTA shape (A/B: profiles/scripts/gather_cfg_ab.sh)
#ifndef B2N_FWD2_WARPS
#define B2N_FWD2_WARPS 8
#endif
#ifndef B2N_FWD2_MINB
#define B2N_FWD2_MINB 3
#endif
constexpr int kFwdWarps = B2N_FWD2_WARPS, kFwdThreads = kFwdWarps * 32, kFwdMinB = B2N_FWD2_MINB;
constexpr int kCap = 128;   // max points per sub-problem the forward kernel stages at once
constexpr int kRound = 32;  // points per staging round of the adjoint kernel
constexpr int kJ = 6;       // neighbours per dimension handled here
constexpr int kSY = kTile + kJ - 1;      // 21 staged rows
constexpr int kSX = kTile + kJ - 1 + 1;  // 22 staged columns (even: 16-byte rows for TMA)
constexpr int kPS = kSY * kSX;           // plane stride, 462 float2
constexpr int kNC = 2 * kJ;              // complex weights per point record
constexpr int kBoxPlanes = 8;            // coil planes per TMA box

struct SubProblem {
  int b, c0, y0, x0, start, count;
  bool valid, interior;
};

// decode blockIdx -> (sub-problem, coil chunk, batch element)
template <int CC> B2N_D SubProblem decode(const InterpArgs<float> &a) {
  SubProblem sp;
  // four independent loads (the descriptor arrays are allocated to gridDim.x entries), so the
  // CTA pays one global-memory latency here instead of two dependent ones
  const int n_sub = *a.n_sub;
  const int tile_all = a.sub_tile[blockIdx.x];
  sp.start = a.sub_start[blockIdx.x];
  sp.count = a.sub_count[blockIdx.x];
  sp.valid = (int)blockIdx.x < n_sub;
  if (!sp.valid) return sp;
  const int n_tiles = (int)a.tiling.n_tiles;
  const int traj = tile_all / n_tiles, tid = tile_all - traj * n_tiles;
  const int ty = tid / a.tiling.nt[1], tx = tid - ty * a.tiling.nt[1];
  sp.y0 = ty * kTile;
  sp.x0 = tx * kTile;
  sp.c0 = blockIdx.y * CC;
  sp.b = a.n_traj == 1 ? (int)blockIdx.z : traj;
  sp.interior = sp.y0 + kSY <= (int)a.K[0] && sp.x0 + kSX <= (int)a.K[1];
  return sp;
}

// lane -> (coil, cell slot).  CC == 16: (coil_hi, slot, coil_lo) so that a half-warp is two
// x-adjacent cells x 8 coils (see the bank argument in the header); otherwise (slot, coil).
template <int CC> B2N_D void lane_map(int lane, int &c, int &q) {
  if (CC == 16) {
    c = ((lane >> 4) << 3) | (lane & 7);
    q = (lane >> 3) & 1;
  } else {
    c = lane % CC;
    q = lane / CC;
  }
}
template <int CC> constexpr int planes() { return CC < kBoxPlanes ? kBoxPlanes : CC; }

// element-wise staging of the tile with periodic wrap (boundary tiles / no tensor map)
template <int CC, int NT = kThreads>
B2N_D void stage_tile_elementwise(float2 *tile, const float2 *__restrict__ grid, const SubProblem &sp, int C, int Ky,
                                  int Kx) {
  for (int e = threadIdx.x; e < CC * kPS; e += NT) {
    const int c = e / kPS, rem = e - c * kPS;
    const int r = rem / kSX, x = rem - r * kSX;
    int gy = sp.y0 + r, gx = sp.x0 + x;
    gy = gy < Ky ? gy : gy % Ky;
    gx = gx < Kx ? gx : gx % Kx;
    const bool on = sp.c0 + c < C;
    cp_async8(&tile[e], &grid[((int64_t)(sp.b * C + (on ? sp.c0 + c : 0)) * Ky + gy) * Kx + gx], on);
  }
}

// -----------------------------------------------------------------------------------------
// forward gather
// -----------------------------------------------------------------------------------------
template <int CC, int QY, int QX>
__global__ void __launch_bounds__(kFwdThreads, kFwdMinB) k_fwd_tiled_2d(InterpArgs<float> a, const float2 *__restrict__ grid,
                                                           float2 *__restrict__ kdata,
                                                           const __grid_constant__ CUtensorMap tmap, int use_tma) {
  constexpr int Q = QY * QX, NY = kJ / QY, NX = kJ / QX;
  static_assert(kJ % QY == 0 && kJ % QX == 0 && Q * CC <= 32, "bad lane mapping");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float2 *tile = reinterpret_cast<float2 *>(smem_raw);                 // [planes][kSY][kSX]
  float2 *s_coef = tile + planes<CC>() * kPS;                          // [kCap][kNC]
  int2 *s_base = reinterpret_cast<int2 *>(s_coef + kCap * kNC);        // [kCap]
  int *s_perm = reinterpret_cast<int *>(s_base + kCap);                // [kCap]
  uint64_t *bar = reinterpret_cast<uint64_t *>(s_perm + kCap);
  const long long t_start = a.trace ? gtime() : 0;
  griddep_launch();
  const SubProblem sp = decode<CC>(a);
  if (!sp.valid) return;
  const int Ky = (int)a.K[0], Kx = (int)a.K[1];
  const int C = (int)a.C;
  const bool tma = use_tma && sp.interior;

  // plan records of this sub-problem (contiguous in plan order): they do not depend on the kernel that produced the
  // grid, so under programmatic dependent launch they are fetched while that kernel's tail is still running
  {
    const float4 *src =
        reinterpret_cast<const float4 *>(reinterpret_cast<const float2 *>(a.coef) + (int64_t)sp.start * kNC);
    float4 *dst = reinterpret_cast<float4 *>(s_coef);
    for (int e = threadIdx.x; e < sp.count * (kNC / 2); e += kFwdThreads) cp_async16(&dst[e], &src[e]);
    const int2 *bsrc = reinterpret_cast<const int2 *>(a.base) + sp.start;
    for (int e = threadIdx.x; e < sp.count; e += kFwdThreads) {
      cp_async8(&s_base[e], &bsrc[e], true);
      cp_async4(&s_perm[e], &a.perm[sp.start + e]);
    }
  }
  if (tma && threadIdx.x == 0) mbar_init(bar, 1);
  griddep_wait();  // the grid is complete and visible from here on
  // the tile: one thread arms the mbarrier and issues the TMA boxes
  if (tma) {
    if (threadIdx.x == 0) {
      mbar_expect_tx(bar, (unsigned)(planes<CC>() * kPS * sizeof(float2)));
      for (int p = 0; p < planes<CC>(); p += kBoxPlanes)
        tma_load_4d(tile + p * kPS, &tmap, 2 * sp.x0, sp.y0, sp.c0 + p, sp.b, bar);
    }
  } else {
    stage_tile_elementwise<CC, kFwdThreads>(tile, grid, sp, C, Ky, Kx);
  }
  cp_async_commit();
  const long long t_issued = a.trace ? gtime() : 0;
  cp_async_wait_all();
  const long long t_records = a.trace ? gtime() : 0;
  __syncthreads();  // everyone's copies landed; the mbarrier initialisation is visible to all
  if (tma) mbar_wait(bar, 0);
  const long long t_staged = a.trace ? gtime() : 0;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int c, q;
  lane_map<CC>(lane, c, q);
  const bool lane_on = q < Q;  // lanes beyond the QY x QX block idle on cell (0, 0) and contribute zero
  const int qy = lane_on ? q / QX : 0, qx = lane_on ? q - (q / QX) * QX : 0;
  const float2 *tplane = tile + c * kPS + qy * kSX + qx;
  float2 *out = kdata + (int64_t)(sp.b * C + sp.c0 + c) * a.M;
  const bool store = q < 2 && sp.c0 + c < C;
  // PU points per warp iteration: their shared-memory loads are independent, which gives the
  // scheduler the ILP that one point alone (a chain LDS -> FFMA -> FFMA -> SHFL) lacks
  constexpr int PU = 2;
  for (int i0 = warp * PU; i0 < sp.count; i0 += kFwdWarps * PU) {
    float2 acc[PU];
#pragma unroll
    for (int u = 0; u < PU; ++u) {
      const int i = min(i0 + u, sp.count - 1);
      const int2 bs = s_base[i];
      const float2 *rec = s_coef + i * kNC;
      // this lane's weights: cy[jy] for jy = ny*QY + qy, cx[jx] for jx = nx*QX + qx
      float2 cy[NY], cx[NX];
      if (QY == 1) {  // every lane needs all six: three 16-byte broadcast loads instead of six 8-byte ones
        const float4 *r4 = reinterpret_cast<const float4 *>(rec);
#pragma unroll
        for (int k = 0; k < NY / 2; ++k) {
          const float4 w = r4[k];
          cy[2 * k] = make_float2(w.x, w.y);
          cy[2 * k + 1] = make_float2(w.z, w.w);
        }
      } else {
#pragma unroll
        for (int ny = 0; ny < NY; ++ny) cy[ny] = rec[ny * QY + qy];
      }
#pragma unroll
      for (int nx = 0; nx < NX; ++nx) cx[nx] = rec[kJ + nx * QX + qx];
      const float2 *t0 = tplane + (bs.x - sp.y0) * kSX + (bs.y - sp.x0);
      // all footprint loads first, then NY independent row accumulators advanced together:
      // 2*NY independent FFMA chains instead of 2 (the kernel was latency-bound, r01_d profile)
      float2 g[NY][NX];
#pragma unroll
      for (int ny = 0; ny < NY; ++ny)
#pragma unroll
        for (int nx = 0; nx < NX; ++nx) g[ny][nx] = t0[ny * QY * kSX + nx * QX];
      float2 row[NY];
#pragma unroll
      for (int ny = 0; ny < NY; ++ny) row[ny] = make_float2(0.f, 0.f);
#pragma unroll
      for (int nx = 0; nx < NX; ++nx)
#pragma unroll
        for (int ny = 0; ny < NY; ++ny) cmacf(row[ny], cx[nx], g[ny][nx]);
      float2 even = make_float2(0.f, 0.f), odd = make_float2(0.f, 0.f);
#pragma unroll
      for (int ny = 0; ny < NY; ++ny) cmacf((ny & 1) ? odd : even, cy[ny], row[ny]);
      acc[u] = lane_on ? make_float2(even.x + odd.x, even.y + odd.y) : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < PU; ++u) {
      if (CC == 16) {
        acc[u].x += __shfl_xor_sync(0xffffffffu, acc[u].x, 8);
        acc[u].y += __shfl_xor_sync(0xffffffffu, acc[u].y, 8);
      } else {
#pragma unroll
        for (int off = CC; off < 32; off <<= 1) {
          acc[u].x += __shfl_xor_sync(0xffffffffu, acc[u].x, off);
          acc[u].y += __shfl_xor_sync(0xffffffffu, acc[u].y, off);
        }
      }
    }
    // after the reduction every lane of a coil holds the sums: lane slot q stores point i0 + q, so ONE store
    // instruction covers PU (mostly k-space-adjacent) samples of each coil -- half the L1 wavefronts of a
    // store per point
    static_assert(PU == 2 && Q >= PU, "store mapping");
    {
      const int iu = i0 + (q == 1 ? 1 : 0);
      const float2 mine = q == 1 ? acc[1] : acc[0];
      if (store && iu < sp.count) out[s_perm[iu]] = mine;
    }
  }
  if (a.trace) {
    __syncthreads();
    // flags: bit 0 = TMA path; bits 8.. = ns from start to "all copies issued"; bits 32.. = ns to "records landed"
    trace_write(a, sp.count, t_start, t_staged, gtime(),
                (tma ? 1 : 0) | (int)(((t_issued - t_start) & 0xFFFFF) << 8));
    if (a.trace && threadIdx.x == 0) {
      const int64_t id = ((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
      if (id < a.trace_cap) a.trace[id * 6 + 0] |= (long long)(t_records - t_start) << 32;
    }
  }
}

// -----------------------------------------------------------------------------------------
// forward gather, persistent: one CTA per SM walks the (longest-first) sub-problem list with
// TWO staging buffers, so the tile / records of item k+1 stream in (TMA + cp.async) while
// item k is computed.  Measured motivation (profiles/trace_cta.py, r01_e): with one
// sub-problem per CTA the staging phase (4 us) was as long as the compute phase (6 us) and,
// with three CTAs per SM, was hidden only part of the time.
// -----------------------------------------------------------------------------------------
constexpr int kPersistBufs = 3;   // staging buffers of the persistent forward kernel (prefetch distance 2)
constexpr int kPersistWarps = 24;
constexpr int kPersistThreads = kPersistWarps * 32;

struct StageBuf {
  float2 *tile, *coef;
  int2 *base;
  int *perm;
};

template <int CC> struct Item {
  int start, count, tile_all;
};

template <int CC, int QY, int QX>
__global__ void __launch_bounds__(kPersistThreads, 1)
    k_fwd_persist_2d(InterpArgs<float> a, const float2 *__restrict__ grid, float2 *__restrict__ kdata,
                     const __grid_constant__ CUtensorMap tmap, int use_tma, int n_chunks, int n_batch) {
  constexpr int Q = QY * QX, NY = kJ / QY, NX = kJ / QX;
  constexpr int TILE_F2 = planes<CC>() * kPS;
  constexpr int BUF_BYTES = TILE_F2 * 8 + kCap * kNC * 8 + kCap * 8 + kCap * 4;
  static_assert(BUF_BYTES % 128 == 0, "staging buffers must keep 128-byte alignment for TMA");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int NBUF = kPersistBufs;
  StageBuf buf[NBUF];
#pragma unroll
  for (int k = 0; k < NBUF; ++k) {
    unsigned char *p = smem_raw + k * BUF_BYTES;
    buf[k].tile = reinterpret_cast<float2 *>(p);
    buf[k].coef = buf[k].tile + TILE_F2;
    buf[k].base = reinterpret_cast<int2 *>(buf[k].coef + kCap * kNC);
    buf[k].perm = reinterpret_cast<int *>(buf[k].base + kCap);
  }
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + NBUF * BUF_BYTES);
  const int Ky = (int)a.K[0], Kx = (int)a.K[1];
  const int C = (int)a.C;
  const int n_sub = *a.n_sub;
  const int total = n_sub * n_chunks * n_batch;
  if (threadIdx.x == 0)
    for (int k = 0; k < NBUF; ++k) mbar_init(&bars[k], 1);
  __syncthreads();

  auto load_item = [&](int item, int &start, int &count, int &tile_all) {
    if (item < total) {
      const int x = item % n_sub;
      start = a.sub_start[x];
      count = a.sub_count[x];
      tile_all = a.sub_tile[x];
    } else {
      start = count = tile_all = 0;
    }
  };
  // decode an item into its sub-problem (coil chunk / batch element from the item index)
  auto make_sp = [&](int item, int start, int count, int tile_all) {
    SubProblem sp;
    sp.valid = item < total;
    const int yz = item / n_sub;
    const int n_tiles = (int)a.tiling.n_tiles;
    const int traj = tile_all / n_tiles, tid = tile_all - traj * n_tiles;
    const int ty = tid / a.tiling.nt[1], tx = tid - ty * a.tiling.nt[1];
    sp.y0 = ty * kTile;
    sp.x0 = tx * kTile;
    sp.c0 = (yz % n_chunks) * CC;
    sp.b = a.n_traj == 1 ? yz / n_chunks : traj;
    sp.start = start;
    sp.count = count;
    sp.interior = sp.y0 + kSY <= Ky && sp.x0 + kSX <= Kx;
    return sp;
  };
  auto issue = [&](const SubProblem &sp, StageBuf &b, uint64_t *bar) {
    if (sp.valid) {
      const bool tma = use_tma && sp.interior;
      if (tma) {
        if (threadIdx.x == 0) {
          mbar_expect_tx(bar, (unsigned)(TILE_F2 * sizeof(float2)));
          for (int p = 0; p < planes<CC>(); p += kBoxPlanes)
            tma_load_4d(b.tile + p * kPS, &tmap, 2 * sp.x0, sp.y0, sp.c0 + p, sp.b, bar);
        }
      } else {
        for (int e = threadIdx.x; e < CC * kPS; e += kPersistThreads) {
          const int c = e / kPS, rem = e - c * kPS;
          const int r = rem / kSX, x = rem - r * kSX;
          int gy = sp.y0 + r, gx = sp.x0 + x;
          gy = gy < Ky ? gy : gy % Ky;
          gx = gx < Kx ? gx : gx % Kx;
          const bool on = sp.c0 + c < C;
          cp_async8(&b.tile[e], &grid[((int64_t)(sp.b * C + (on ? sp.c0 + c : 0)) * Ky + gy) * Kx + gx], on);
        }
      }
      const float4 *src =
          reinterpret_cast<const float4 *>(reinterpret_cast<const float2 *>(a.coef) + (int64_t)sp.start * kNC);
      float4 *dst = reinterpret_cast<float4 *>(b.coef);
      for (int e = threadIdx.x; e < sp.count * (kNC / 2); e += kPersistThreads) cp_async16(&dst[e], &src[e]);
      const int2 *bsrc = reinterpret_cast<const int2 *>(a.base) + sp.start;
      for (int e = threadIdx.x; e < sp.count; e += kPersistThreads) {
        cp_async8(&b.base[e], &bsrc[e], true);
        cp_async4(&b.perm[e], &a.perm[sp.start + e]);
      }
    }
    cp_async_commit();  // one group per item, even when empty, so that wait_group counts stay aligned
  };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int c, q;
  lane_map<CC>(lane, c, q);
  const bool lane_on = q < Q;
  const int qy = lane_on ? q / QX : 0, qx = lane_on ? q - (q / QX) * QX : 0;

  // software pipeline of depth NBUF-1: items k+1 .. k+NBUF-1 are in flight while item k is computed
  const int G = gridDim.x;
  constexpr int DR = NBUF + 2;  // descriptor ring: descriptors are fetched two items before they are issued
  int it_id[DR], it_s[DR], it_n[DR], it_t[DR];  // slot = item ordinal % DR
  unsigned phase[NBUF];
#pragma unroll
  for (int k = 0; k < NBUF; ++k) phase[k] = 0u;
#pragma unroll
  for (int k = 0; k < DR; ++k) {
    it_id[k] = blockIdx.x + k * G;
    load_item(it_id[k], it_s[k], it_n[k], it_t[k]);
  }
#pragma unroll
  for (int k = 0; k < NBUF - 1; ++k) issue(make_sp(it_id[k], it_s[k], it_n[k], it_t[k]), buf[k], &bars[k]);
  for (int k = 0;; ++k) {
    const int cur = k % NBUF, dcur = k % DR;
    if (it_id[dcur] >= total) break;
    const SubProblem sp0 = make_sp(it_id[dcur], it_s[dcur], it_n[dcur], it_t[dcur]);
    {
      // refill the buffer whose item finished an iteration ago with item k + NBUF - 1
      const int nb = (k + NBUF - 1) % NBUF, nd = (k + NBUF - 1) % DR;
      issue(make_sp(it_id[nd], it_s[nd], it_n[nd], it_t[nd]), buf[nb], &bars[nb]);
    }
    // wait for the current item: all but the NBUF-1 newest cp.async groups, and the tile's mbarrier phase
    asm volatile("cp.async.wait_group %0;\n" ::"n"(NBUF - 1) : "memory");
    __syncthreads();
    if (use_tma && sp0.interior) {
      mbar_wait(&bars[cur], phase[cur] & 1u);
      ++phase[cur];
    }

    const StageBuf &b = buf[cur];
    const float2 *tplane = b.tile + c * kPS + qy * kSX + qx;
    float2 *out = kdata + (int64_t)(sp0.b * C + sp0.c0 + c) * a.M;
    const bool store = q == 0 && sp0.c0 + c < C;
    for (int i = warp; i < sp0.count; i += kPersistWarps) {
      const int2 bs = b.base[i];
      const float2 *rec = b.coef + i * kNC;
      float2 cy[NY], cx[NX];
#pragma unroll
      for (int ny = 0; ny < NY; ++ny) cy[ny] = rec[ny * QY + qy];
#pragma unroll
      for (int nx = 0; nx < NX; ++nx) cx[nx] = rec[kJ + nx * QX + qx];
      const float2 *tp = tplane + (bs.x - sp0.y0) * kSX + (bs.y - sp0.x0);
      float2 g[NY][NX];
#pragma unroll
      for (int ny = 0; ny < NY; ++ny)
#pragma unroll
        for (int nx = 0; nx < NX; ++nx) g[ny][nx] = tp[ny * QY * kSX + nx * QX];
      float2 row[NY];
#pragma unroll
      for (int ny = 0; ny < NY; ++ny) row[ny] = make_float2(0.f, 0.f);
#pragma unroll
      for (int nx = 0; nx < NX; ++nx)
#pragma unroll
        for (int ny = 0; ny < NY; ++ny) cmacf(row[ny], cx[nx], g[ny][nx]);
      float2 even = make_float2(0.f, 0.f), odd = make_float2(0.f, 0.f);
#pragma unroll
      for (int ny = 0; ny < NY; ++ny) cmacf((ny & 1) ? odd : even, cy[ny], row[ny]);
      float2 acc = lane_on ? make_float2(even.x + odd.x, even.y + odd.y) : make_float2(0.f, 0.f);
      if (CC == 16) {
        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 8);
        acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 8);
      } else {
#pragma unroll
        for (int off = CC; off < 32; off <<= 1) {
          acc.x += __shfl_xor_sync(0xffffffffu, acc.x, off);
          acc.y += __shfl_xor_sync(0xffffffffu, acc.y, off);
        }
      }
      if (store) out[b.perm[i]] = acc;
    }
    __syncthreads();  // everyone is done with this buffer before a later iteration refills it
    // this descriptor slot now describes item k + DR (issued two iterations after the load)
    it_id[dcur] += DR * G;
    load_item(it_id[dcur], it_s[dcur], it_n[dcur], it_t[dcur]);
  }
}

// -----------------------------------------------------------------------------------------
// adjoint spread
// -----------------------------------------------------------------------------------------
// float2 slots of one staging buffer of k_adj_tiled_2d: weights, samples ((kRound+1) per coil), base cells;
// even, so that the second buffer stays 16-byte aligned
template <int CC> constexpr int adj_stage_slots() { return (kRound * kNC + (kRound + 1) * CC + kRound + 1) & ~1; }

// NW warps per CTA: warp w owns the tile rows r with r mod NW == w.  NW = 8: one footprint row per
// (point, warp); NW = 4: one or two rows, with the per-point loads (base cell, sample, x-weights)
// shared by both -- fewer shared-memory wavefronts per point, fewer warps to hide latency with.
// ORDERED: instead of adding the tile into the grid (TMA reduce / atomics, arrival order), store it to this
// sub-problem's slot of `scratch` ([rank][batch][coil][kPS]); k_adj_merge_2d then adds the slots per cell in a
// fixed order.  The in-tile accumulation is order-fixed already (one warp per cell, points in plan order).
template <int CC, int NW, bool ORDERED = false>
__global__ void __launch_bounds__(NW * 32) k_adj_tiled_2d(InterpArgs<float> a, const float2 *__restrict__ kdata,
                                                           float2 *__restrict__ grid,
                                                           const __grid_constant__ CUtensorMap tmap, int use_tma,
                                                           float2 *__restrict__ scratch = nullptr) {
  constexpr int QX = 32 / CC, NX = (kJ + QX - 1) / QX;
  constexpr int NT = NW * 32;
  // staged samples are coil-major with an odd row stride: the gather writes 32 consecutive points of one
  // coil (conflict-free) and the spread reads 8 coils x one point per half-warp (conflict-free)
  constexpr int VS = kRound + 1;
  constexpr int STAGE = adj_stage_slots<CC>();  // float2 slots: coef, val, base
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float2 *tile = reinterpret_cast<float2 *>(smem_raw);  // [planes][kSY][kSX] accumulators
  float2 *stage0 = tile + planes<CC>() * kPS;           // 2 x STAGE
  int *s_perm = reinterpret_cast<int *>(stage0 + 2 * STAGE);  // 3 x kRound sample indices
  const long long t_start = a.trace ? gtime() : 0;
  // (no griddep_launch() here: letting the inverse column pass move in during this kernel's tail costs 4 us at
  // 320^2 x 16 coils and 15-30 us at 384^2 x 32, profiles/r01_h_opts_ab.log)
  const SubProblem sp = decode<CC>(a);
  if (!sp.valid) return;
  const int Ky = (int)a.K[0], Kx = (int)a.K[1];
  const int C = (int)a.C;
  const int64_t M = a.M;
  const float2 *pcoef = reinterpret_cast<const float2 *>(a.coef);
  const int rounds = (sp.count + kRound - 1) / kRound;

  auto issue_perm = [&](int round) {  // sample indices, two rounds ahead of their use
    if (round < rounds) {
      const int p0 = round * kRound, nb = min(kRound, sp.count - p0);
      int *dst = s_perm + (round % 3) * kRound;
      for (int e = threadIdx.x; e < nb; e += NT) cp_async4(&dst[e], &a.perm[sp.start + p0 + e]);
    }
  };
  auto issue_data = [&](int round) {  // weights, base cells and gathered samples, one round ahead
    if (round < rounds) {
      float2 *buf = stage0 + (round & 1) * STAGE;
      const int p0 = round * kRound, nb = min(kRound, sp.count - p0), s0 = sp.start + p0;
      const float4 *src = reinterpret_cast<const float4 *>(pcoef + (int64_t)s0 * kNC);
      float4 *dst = reinterpret_cast<float4 *>(buf);
      for (int e = threadIdx.x; e < nb * (kNC / 2); e += NT) cp_async16(&dst[e], &src[e]);
      float2 *val = buf + kRound * kNC;
      const int *perm = s_perm + (round % 3) * kRound;
      for (int e = threadIdx.x; e < kRound * CC; e += NT) {
        const int cc = e / kRound, i = e - cc * kRound;  // consecutive threads: consecutive points, one coil
        const bool on = sp.c0 + cc < C && i < nb;
        cp_async8(&val[cc * VS + i], &kdata[(int64_t)(sp.b * C + (on ? sp.c0 + cc : 0)) * M + (on ? perm[i] : 0)], on);
      }
      int2 *sb = reinterpret_cast<int2 *>(val + VS * CC);
      const int2 *bsrc = reinterpret_cast<const int2 *>(a.base) + s0;
      for (int e = threadIdx.x; e < nb; e += NT) cp_async8(&sb[e], &bsrc[e], true);
    }
  };

  issue_perm(0);
  issue_perm(1);
  cp_async_commit();
  for (int e = threadIdx.x; e < planes<CC>() * kPS; e += NT) tile[e] = make_float2(0.f, 0.f);
  cp_async_wait_all();
  __syncthreads();
  issue_data(0);
  cp_async_commit();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int c, qx;
  lane_map<CC>(lane, c, qx);
  float2 *tplane = tile + c * kPS;
  for (int round = 0; round < rounds; ++round) {
    cp_async_wait_all();
    __syncthreads();  // this round's data (and next round's indices) landed; buffers of round-1 are free
    issue_data(round + 1);
    issue_perm(round + 2);
    cp_async_commit();
    const float2 *buf = stage0 + (round & 1) * STAGE;
    const float2 *s_coef = buf, *s_val = buf + kRound * kNC;
    const int2 *s_base = reinterpret_cast<const int2 *>(s_val + VS * CC);
    const int nb = min(kRound, sp.count - round * kRound);
    for (int i = 0; i < nb; ++i) {
      const int2 bs = s_base[i];
      const int by = bs.x - sp.y0, bx = bs.y - sp.x0;
      const int jy0 = (warp - by) & (NW - 1);  // the first footprint row this warp owns, if any
      if (jy0 >= kJ) continue;
      const float2 v = s_val[c * VS + i];
      float2 cxv[NX];
#pragma unroll
      for (int nx = 0; nx < NX; ++nx) {
        const int jx = nx * QX + qx;
        cxv[nx] = s_coef[i * kNC + kJ + ((kJ % QX == 0 || jx < kJ) ? jx : 0)];
      }
#pragma unroll
      for (int rr = 0; rr < (kJ + NW - 1) / NW; ++rr) {  // the rows jy0, jy0 + NW, ... of this footprint
        const int jy = jy0 + rr * NW;
        if (rr > 0 && jy >= kJ) break;
        const float2 cyv = s_coef[i * kNC + jy];
        float2 *trow = tplane + (by + jy) * kSX + bx;
        float2 u;  // conj(cy) * v   (cy carries the fftshift phase)
        u.x = fmaf(cyv.x, v.x, cyv.y * v.y);
        u.y = fmaf(cyv.x, v.y, -cyv.y * v.x);
        float2 t[NX];
#pragma unroll
        for (int nx = 0; nx < NX; ++nx) {  // all loads first: independent, latency overlaps
          const int jx = nx * QX + qx;
          t[nx] = trow[(kJ % QX == 0 || jx < kJ) ? jx : 0];
        }
#pragma unroll
        for (int nx = 0; nx < NX; ++nx) {
          const int jx = nx * QX + qx;
          if (kJ % QX == 0 || jx < kJ) {
            cmacf_conj(t[nx], cxv[nx], u);  // += conj(cx) * conj(cy) * v
            trow[jx] = t[nx];
          }
        }
      }
      __syncwarp();
    }
  }
  const long long t_accum = a.trace ? gtime() : 0;
  if (ORDERED) {
    __syncthreads();
    const int ncoil = min(CC, C - sp.c0);
    const int64_t slot = a.sub_slot[blockIdx.x];
    float4 *dst = reinterpret_cast<float4 *>(scratch + ((slot * gridDim.z + blockIdx.z) * C + sp.c0) * kPS);
    const float4 *src = reinterpret_cast<const float4 *>(tile);
    for (int e = threadIdx.x; e < ncoil * (kPS / 2); e += NT) dst[e] = src[e];
    return;
  }
  // merge the tile into the global grid.  Under programmatic dependent launch the kernel before this one is
  // k_zero_grid: everything above overlapped it, the grid itself is first touched here.
  griddep_wait();
  if (use_tma && sp.interior) {
    fence_async_proxy();  // generic-proxy writes to shared memory -> visible to the TMA engine
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int p = 0; p < planes<CC>(); p += kBoxPlanes)
        tma_reduce_add_4d(&tmap, 2 * sp.x0, sp.y0, sp.c0 + p, sp.b, tile + p * kPS);
      tma_store_commit_wait();  // shared memory must stay valid until the engine has read it
    }
  } else {
    __syncthreads();
    for (int e = threadIdx.x; e < CC * kPS; e += NT) {
      const int cc = e / kPS, rem = e - cc * kPS;
      if (sp.c0 + cc >= C) break;
      const int r = rem / kSX, x = rem - r * kSX;
      if (x >= kSX - 1) continue;  // the 22nd column is padding for TMA, never written
      const float2 v = tile[e];
      if (v.x == 0.f && v.y == 0.f) continue;
      int gy = sp.y0 + r, gx = sp.x0 + x;
      gy = gy < Ky ? gy : gy % Ky;
      gx = gx < Kx ? gx : gx % Kx;
      atomicAdd(&grid[((int64_t)(sp.b * C + sp.c0 + cc) * Ky + gy) * Kx + gx], v);
    }
  }
  if (a.trace) {
    __syncthreads();
    trace_write(a, sp.count, t_start, t_accum, gtime(), (use_tma && sp.interior) ? 1 : 0);
  }
}

// -----------------------------------------------------------------------------------------
// adjoint spread, 16-coil chunks: warps partition the COILS (warp w owns coils p and p+4,
// p = (w & 3) + 8 * (w >> 2)), lanes = (16 footprint cells, 2 coils).  Every warp visits
// every point, so there is no ownership test and no divergence, and because coil planes
// are disjoint between warps the read-modify-write needs no atomics.  The 36 conjugated
// weight products of each staged point are formed once per round by the whole CTA.
// Banks: cell n = 6r + x sits at r*22 + x = n (mod 16) in its plane and the two coils of a
// warp are 4 planes = 8 (mod 16) apart, so a half-warp (8 cells x 2 coils) is conflict-free.
// -----------------------------------------------------------------------------------------
constexpr int kRoundC = 24;  // points per round of the coil-partitioned adjoint

__global__ void __launch_bounds__(kThreads, 3) k_adj_coilwarp_2d(InterpArgs<float> a, const float2 *__restrict__ kdata,
                                                              float2 *__restrict__ grid,
                                                              const __grid_constant__ CUtensorMap tmap, int use_tma) {
  constexpr int CC = 16, W = kJ * kJ;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float2 *tile = reinterpret_cast<float2 *>(smem_raw);     // [16][kSY][kSX] accumulators
  float2 *s_coef = tile + CC * kPS;                        // [kRoundC][kNC]   raw weights of the round
  float2 *s_w = s_coef + kRoundC * kNC;                    // [kRoundC][W]     conj(cy[jy] * cx[jx])
  float2 *s_val = s_w + kRoundC * W;                       // 2 x [kRoundC][CC] gathered samples
  int2 *s_base = reinterpret_cast<int2 *>(s_val + 2 * kRoundC * CC);  // 2 x [kRoundC]
  int *s_perm = reinterpret_cast<int *>(s_base + 2 * kRoundC);        // 3 x [kRoundC]
  const long long t_start = a.trace ? gtime() : 0;
  const SubProblem sp = decode<CC>(a);
  if (!sp.valid) return;
  const int Ky = (int)a.K[0], Kx = (int)a.K[1];
  const int C = (int)a.C;
  const int64_t M = a.M;
  const float2 *pcoef = reinterpret_cast<const float2 *>(a.coef);
  const int rounds = (sp.count + kRoundC - 1) / kRoundC;

  auto issue_perm = [&](int round) {  // sample indices, two rounds ahead of their use
    if (round < rounds) {
      const int p0 = round * kRoundC, nb = min(kRoundC, sp.count - p0);
      int *dst = s_perm + (round % 3) * kRoundC;
      for (int e = threadIdx.x; e < nb; e += kThreads) cp_async4(&dst[e], &a.perm[sp.start + p0 + e]);
    }
  };
  auto issue_coef = [&](int round) {
    if (round < rounds) {
      const int p0 = round * kRoundC, nb = min(kRoundC, sp.count - p0);
      const float4 *src = reinterpret_cast<const float4 *>(pcoef + (int64_t)(sp.start + p0) * kNC);
      float4 *dst = reinterpret_cast<float4 *>(s_coef);
      for (int e = threadIdx.x; e < nb * (kNC / 2); e += kThreads) cp_async16(&dst[e], &src[e]);
    }
  };
  auto issue_val = [&](int round) {  // gathered samples and base cells, one round ahead
    if (round < rounds) {
      const int p0 = round * kRoundC, nb = min(kRoundC, sp.count - p0);
      float2 *val = s_val + (round & 1) * kRoundC * CC;
      const int *perm = s_perm + (round % 3) * kRoundC;
      for (int e = threadIdx.x; e < kRoundC * CC; e += kThreads) {
        const int cc = e / kRoundC, i = e - cc * kRoundC;  // consecutive threads: consecutive points, one coil
        const bool on = sp.c0 + cc < C && i < nb;
        cp_async8(&val[i * CC + cc], &kdata[(int64_t)(sp.b * C + (on ? sp.c0 + cc : 0)) * M + (on ? perm[i] : 0)], on);
      }
      int2 *sb = s_base + (round & 1) * kRoundC;
      const int2 *bsrc = reinterpret_cast<const int2 *>(a.base) + sp.start + p0;
      for (int e = threadIdx.x; e < nb; e += kThreads) cp_async8(&sb[e], &bsrc[e], true);
    }
  };

  issue_perm(0);
  issue_perm(1);
  issue_coef(0);
  cp_async_commit();
  for (int e = threadIdx.x; e < CC * kPS; e += kThreads) tile[e] = make_float2(0.f, 0.f);
  cp_async_wait_all();
  __syncthreads();
  issue_val(0);
  cp_async_commit();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qs = lane >> 1;
  const int coil = (warp & 3) + 8 * (warp >> 2) + 4 * (lane & 1);
  float2 *tplane = tile + coil * kPS;
  int n_it[3], off_it[3];
#pragma unroll
  for (int it = 0; it < 3; ++it) {
    const int n = min(it * 16 + qs, W - 1);
    n_it[it] = n;
    off_it[it] = (n / kJ) * kSX + (n % kJ);
  }
  const bool last_on = 2 * 16 + qs < W;
  for (int round = 0; round < rounds; ++round) {
    const int nb = min(kRoundC, sp.count - round * kRoundC);
    cp_async_wait_all();
    __syncthreads();  // raw weights, samples and base cells of this round landed; last round is fully consumed
    for (int e = threadIdx.x; e < nb * W; e += kThreads) {  // conjugated separable products, once per point
      const int i = e / W, n = e - i * W;
      const float2 cy = s_coef[i * kNC + n / kJ], cx = s_coef[i * kNC + kJ + n % kJ];
      s_w[e] = make_float2(cy.x * cx.x - cy.y * cx.y, -(cy.x * cx.y + cy.y * cx.x));
    }
    __syncthreads();  // products visible; raw weight buffer free again
    issue_coef(round + 1);
    issue_val(round + 1);
    issue_perm(round + 2);
    cp_async_commit();
    const float2 *val = s_val + (round & 1) * kRoundC * CC;
    const int2 *sb = s_base + (round & 1) * kRoundC;
    // software pipeline: the next point's base cell, sample and weights (read-only) are fetched
    // while the current point's tile cells are updated (those must stay in program order)
    int2 bs_n = sb[0];
    float2 v_n = val[coil];
    float2 w0_n = s_w[n_it[0]], w1_n = s_w[n_it[1]], w2_n = s_w[n_it[2]];
    for (int i = 0; i < nb; ++i) {
      const int2 bs = bs_n;
      const float2 v = v_n, w0 = w0_n, w1 = w1_n, w2 = w2_n;
      if (i + 1 < nb) {
        bs_n = sb[i + 1];
        v_n = val[(i + 1) * CC + coil];
        const float2 *w = s_w + (i + 1) * W;
        w0_n = w[n_it[0]];
        w1_n = w[n_it[1]];
        w2_n = w[n_it[2]];
      }
      float2 *tp = tplane + (bs.x - sp.y0) * kSX + (bs.y - sp.x0);
      // lanes without a third cell re-read their own first one (a dummy read of another lane's cell would be an
      // intra-warp read/write hazard for compute-sanitizer racecheck)
      float2 t0 = tp[off_it[0]], t1 = tp[off_it[1]], t2 = tp[last_on ? off_it[2] : off_it[0]];
      cmacf(t0, w0, v);
      cmacf(t1, w1, v);
      cmacf(t2, w2, v);
      tp[off_it[0]] = t0;
      tp[off_it[1]] = t1;
      if (last_on) tp[off_it[2]] = t2;
      __syncwarp();
    }
  }
  const long long t_accum = a.trace ? gtime() : 0;
  // merge the tile into the global grid
  if (use_tma && sp.interior) {
    fence_async_proxy();  // generic-proxy writes to shared memory -> visible to the TMA engine
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int p = 0; p < CC; p += kBoxPlanes) tma_reduce_add_4d(&tmap, 2 * sp.x0, sp.y0, sp.c0 + p, sp.b, tile + p * kPS);
      tma_store_commit_wait();  // shared memory must stay valid until the engine has read it
    }
  } else {
    __syncthreads();
    for (int e = threadIdx.x; e < CC * kPS; e += kThreads) {
      const int cc = e / kPS, rem = e - cc * kPS;
      if (sp.c0 + cc >= C) break;
      const int r = rem / kSX, x = rem - r * kSX;
      if (x >= kSX - 1) continue;  // the 22nd column is padding for TMA, never written
      const float2 v = tile[e];
      if (v.x == 0.f && v.y == 0.f) continue;
      int gy = sp.y0 + r, gx = sp.x0 + x;
      gy = gy < Ky ? gy : gy % Ky;
      gx = gx < Kx ? gx : gx % Kx;
      atomicAdd(&grid[((int64_t)(sp.b * C + sp.c0 + cc) * Ky + gy) * Kx + gx], v);
    }
  }
  if (a.trace) {
    __syncthreads();
    trace_write(a, sp.count, t_start, t_accum, gtime(), (use_tma && sp.interior) ? 1 : 0);
  }
}

// -----------------------------------------------------------------------------------------
// adjoint spread, warp-private tiles: each warp of the CTA owns a private 8-coil accumulation
// tile and visits EVERY point of the sub-problem with lanes = (2x2 footprint cells, 8 coils),
// nine read-modify-write steps per point, all nine independent (distinct cells).  No two
// warps ever touch the same shared-memory word, so there are no atomics and no ownership
// tests, and the operand traffic per point (3 cy + 3 cx + sample + base cell) is the
// smallest of the three variants: measured shared-memory wavefronts per point
// (profiles/r01_e): warp-owned coils 167, warp-owned rows ~110, this one ~88 (floor 72 = the
// read-modify-write of 36 cells x 16 coils itself).  Small CTAs (NW warps) also mean many
// independent CTAs per SM, so one CTA's staging overlaps the others' accumulation.
// Banks: plane stride 462 = 14 (mod 16): 8 coils -> 8 distinct even bank pairs; the two
// x-adjacent cells of a half-warp take the even and the odd ones -> conflict-free.
// -----------------------------------------------------------------------------------------
template <int NW>
__global__ void __launch_bounds__(NW * 32) k_adj_warptile_2d(InterpArgs<float> a, const float2 *__restrict__ kdata,
                                                             float2 *__restrict__ grid,
                                                             const __grid_constant__ CUtensorMap tmap, int use_tma) {
  constexpr int CC = NW * 8, NT = NW * 32;
  constexpr int STAGE = kRound * kNC + kRound * CC + kRound;  // float2 slots: coef, val, base
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float2 *tile = reinterpret_cast<float2 *>(smem_raw);         // [CC][kSY][kSX] accumulators
  float2 *stage0 = tile + CC * kPS;                            // 2 x STAGE
  int *s_perm = reinterpret_cast<int *>(stage0 + 2 * STAGE);   // 3 x kRound sample indices
  const long long t_start = a.trace ? gtime() : 0;
  const SubProblem sp = decode<CC>(a);
  if (!sp.valid) return;
  const int Ky = (int)a.K[0], Kx = (int)a.K[1];
  const int C = (int)a.C;
  const int64_t M = a.M;
  const float2 *pcoef = reinterpret_cast<const float2 *>(a.coef);
  const int rounds = (sp.count + kRound - 1) / kRound;

  auto issue_perm = [&](int round) {  // sample indices, two rounds ahead of their use
    if (round < rounds) {
      const int p0 = round * kRound, nb = min(kRound, sp.count - p0);
      int *dst = s_perm + (round % 3) * kRound;
      for (int e = threadIdx.x; e < nb; e += NT) cp_async4(&dst[e], &a.perm[sp.start + p0 + e]);
    }
  };
  auto issue_data = [&](int round) {  // weights, base cells and gathered samples, one round ahead
    if (round < rounds) {
      float2 *buf = stage0 + (round & 1) * STAGE;
      const int p0 = round * kRound, nb = min(kRound, sp.count - p0), s0 = sp.start + p0;
      const float4 *src = reinterpret_cast<const float4 *>(pcoef + (int64_t)s0 * kNC);
      float4 *dst = reinterpret_cast<float4 *>(buf);
      for (int e = threadIdx.x; e < nb * (kNC / 2); e += NT) cp_async16(&dst[e], &src[e]);
      float2 *val = buf + kRound * kNC;
      const int *perm = s_perm + (round % 3) * kRound;
      for (int e = threadIdx.x; e < kRound * CC; e += NT) {
        const int cc = e / kRound, i = e - cc * kRound;  // consecutive threads: consecutive points, one coil
        const bool on = sp.c0 + cc < C && i < nb;
        cp_async8(&val[i * CC + cc], &kdata[(int64_t)(sp.b * C + (on ? sp.c0 + cc : 0)) * M + (on ? perm[i] : 0)], on);
      }
      int2 *sb = reinterpret_cast<int2 *>(val + kRound * CC);
      const int2 *bsrc = reinterpret_cast<const int2 *>(a.base) + s0;
      for (int e = threadIdx.x; e < nb; e += NT) cp_async8(&sb[e], &bsrc[e], true);
    }
  };

  issue_perm(0);
  issue_perm(1);
  cp_async_commit();
  {
    float4 *t4 = reinterpret_cast<float4 *>(tile);
    for (int e = threadIdx.x; e < CC * kPS / 2; e += NT) t4[e] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  cp_async_wait_all();
  __syncthreads();
  issue_data(0);
  cp_async_commit();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c8 = lane & 7, q = lane >> 3, qy = q >> 1, qx = q & 1;
  const int coil = warp * 8 + c8;
  float2 *tplane = tile + coil * kPS + qy * kSX + qx;
  for (int round = 0; round < rounds; ++round) {
    cp_async_wait_all();
    __syncthreads();  // this round's data (and next round's indices) landed; buffers of round-1 are free
    issue_data(round + 1);
    issue_perm(round + 2);
    cp_async_commit();
    const float2 *buf = stage0 + (round & 1) * STAGE;
    const float2 *s_coef = buf, *s_val = buf + kRound * kNC;
    const int2 *s_base = reinterpret_cast<const int2 *>(s_val + kRound * CC);
    const int nb = min(kRound, sp.count - round * kRound);
    for (int i = 0; i < nb; ++i) {
      const int2 bs = s_base[i];
      const float2 v = s_val[i * CC + coil];
      const float2 *rec = s_coef + i * kNC;
      float2 cyv[3], cxv[3], uy[3], t[3][3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        cyv[k] = rec[2 * k + qy];
        cxv[k] = rec[kJ + 2 * k + qx];
      }
      float2 *tp = tplane + (bs.x - sp.y0) * kSX + (bs.y - sp.x0);
#pragma unroll
      for (int ny = 0; ny < 3; ++ny)
#pragma unroll
        for (int nx = 0; nx < 3; ++nx) t[ny][nx] = tp[2 * ny * kSX + 2 * nx];
#pragma unroll
      for (int ny = 0; ny < 3; ++ny) {  // conj(cy) * v  (cy carries the fftshift phase)
        uy[ny].x = fmaf(cyv[ny].x, v.x, cyv[ny].y * v.y);
        uy[ny].y = fmaf(cyv[ny].x, v.y, -cyv[ny].y * v.x);
      }
#pragma unroll
      for (int ny = 0; ny < 3; ++ny)
#pragma unroll
        for (int nx = 0; nx < 3; ++nx) {
          cmacf_conj(t[ny][nx], cxv[nx], uy[ny]);
          tp[2 * ny * kSX + 2 * nx] = t[ny][nx];
        }
      __syncwarp();
    }
  }
  const long long t_accum = a.trace ? gtime() : 0;
  // merge the tile into the global grid
  if (use_tma && sp.interior) {
    fence_async_proxy();  // generic-proxy writes to shared memory -> visible to the TMA engine
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int p = 0; p < CC; p += kBoxPlanes) tma_reduce_add_4d(&tmap, 2 * sp.x0, sp.y0, sp.c0 + p, sp.b, tile + p * kPS);
      tma_store_commit_wait();  // shared memory must stay valid until the engine has read it
    }
  } else {
    __syncthreads();
    for (int e = threadIdx.x; e < CC * kPS; e += NT) {
      const int cc = e / kPS, rem = e - cc * kPS;
      if (sp.c0 + cc >= C) break;
      const int r = rem / kSX, x = rem - r * kSX;
      if (x >= kSX - 1) continue;  // the 22nd column is padding for TMA, never written
      const float2 v = tile[e];
      if (v.x == 0.f && v.y == 0.f) continue;
      int gy = sp.y0 + r, gx = sp.x0 + x;
      gy = gy < Ky ? gy : gy % Ky;
      gx = gx < Kx ? gx : gx % Kx;
      atomicAdd(&grid[((int64_t)(sp.b * C + sp.c0 + cc) * Ky + gy) * Kx + gx], v);
    }
  }
  if (a.trace) {
    __syncthreads();
    trace_write(a, sp.count, t_start, t_accum, gtime(), (use_tma && sp.interior) ? 1 : 0);
  }
}

// ---- host side ------------------------------------------------------------------------------
EncodeTiledFn tensor_map_encoder() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
    (void)cudaGetLastError();
  }
  return fn;
}

// 4-D FP32 view of a coil-major complex64 grid (B, C, Ky, Kx): dims (2*Kx, Ky, C, B), box
// (2*22, 21, 8, 1).  Returns false when TMA cannot describe it (odd Kx, unaligned base, ...).
static bool make_grid_tmap(CUtensorMap *map, const void *grid, int64_t B, int64_t C, int64_t Ky, int64_t Kx) {
  EncodeTiledFn fn = tensor_map_encoder();
  if (!fn || (Kx & 1) || ((uintptr_t)grid & 15) || Ky < kSY || Kx < kSX) return false;
  const cuuint64_t gdim[4] = {(cuuint64_t)(2 * Kx), (cuuint64_t)Ky, (cuuint64_t)C, (cuuint64_t)B};
  const cuuint64_t gstride[3] = {(cuuint64_t)(Kx * 8), (cuuint64_t)(Ky * Kx * 8), (cuuint64_t)(C * Ky * Kx * 8)};
  const cuuint32_t box[4] = {2 * kSX, kSY, kBoxPlanes, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void *>(grid), gdim, gstride, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int g_adj_rowwarp = 0;
int g_adj_chunk = 0;     // A/B switch: adjoint coil chunk per CTA for C > 8 (0 = 16, 8 = two 8-coil CTAs)
int g_fwd_chunk = 0;     // A/B switch for C > 8: 0 = one 16-coil CTA per sub-problem, 1 = persistent kernel, 8 = 8-coil CTAs  // A/B switch: 1 = row-ownership kernel also for 16-coil chunks

// -----------------------------------------------------------------------------------------
// deterministic merge of the per-sub-problem tiles written by k_adj_tiled_2d<.., ORDERED = true>:
// every grid cell adds the slots of the tiles whose 21 x 21 footprint covers it -- the tile the
// cell lies in and up to two predecessors per dimension (one normally; two when the last tile of a
// dimension is narrower than the halo) -- in a fixed order.  Writes every cell: no memset.
// -----------------------------------------------------------------------------------------
constexpr int kMergeCoils = 8;  // coils per thread of the merge: the tile / slot walk is shared by all of them
__global__ void __launch_bounds__(256) k_adj_merge_2d(InterpArgs<float> a, const float2 *__restrict__ scratch,
                                                      float2 *__restrict__ grid) {
  const int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= a.Kprod) return;
  const int Ky = (int)a.K[0], Kx = (int)a.K[1], C = (int)a.C;
  const int y = (int)(cell / Kx), x = (int)(cell - (int64_t)y * Kx);
  const int nty = a.tiling.nt[0], ntx = a.tiling.nt[1];
  const int64_t n_tiles = a.tiling.n_tiles, n_tiles_all = a.n_traj * n_tiles;
  const int n_sub = *a.n_sub;
  const int64_t Bz = a.n_traj == 1 ? a.B : 1;
  const int64_t ncb = (C + kMergeCoils - 1) / kMergeCoils, nblk = a.B * ncb;  // (batch element, coil block)
  for (int64_t q = blockIdx.y; q < nblk; q += gridDim.y) {
    const int64_t b = q / ncb, c0 = (q - b * ncb) * kMergeCoils;
    const int nc = (int)(C - c0 < kMergeCoils ? C - c0 : kMergeCoils);
    const int64_t traj = a.n_traj == 1 ? 0 : b, bz = a.n_traj == 1 ? b : 0;
    float2 acc[kMergeCoils];
#pragma unroll
    for (int k = 0; k < kMergeCoils; ++k) acc[k] = make_float2(0.f, 0.f);
    for (int ky = 0; ky < min(3, nty); ++ky) {
      int ty = y / kTile - ky;
      if (ty < 0) ty += nty;
      int ry = y - ty * kTile;
      if (ry < 0) ry += Ky;
      if (ry >= kSY) continue;
      for (int kx = 0; kx < min(3, ntx); ++kx) {
        int tx = x / kTile - kx;
        if (tx < 0) tx += ntx;
        int rx = x - tx * kTile;
        if (rx < 0) rx += Kx;
        if (rx >= kSY) continue;  // column kSX-1 of a plane is alignment padding, never accumulated
        const int64_t t = traj * n_tiles + (int64_t)ty * ntx + tx;
        const int s0 = a.tile_sub_start[t], s1 = t + 1 < n_tiles_all ? a.tile_sub_start[t + 1] : n_sub;
        for (int sl = s0; sl < s1; ++sl) {
          const float2 *src = scratch + (((int64_t)sl * Bz + bz) * C + c0) * kPS + ry * kSX + rx;
#pragma unroll
          for (int k = 0; k < kMergeCoils; ++k)
            if (k < nc) {
              const float2 v = src[(int64_t)k * kPS];
              acc[k].x += v.x;
              acc[k].y += v.y;
            }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < kMergeCoils; ++k)
      if (k < nc) grid[(b * C + c0 + k) * a.Kprod + cell] = acc[k];
  }
}

template <int CC>
static int launch_adj_ordered(const InterpArgs<float> &a, const void *kdata, void *scratch, void *grid, cudaStream_t st) {
  const size_t smem =
      sizeof(float2) * (planes<CC>() * kPS + 2 * adj_stage_slots<CC>()) + sizeof(int) * 3 * kRound;
  auto kern = k_adj_tiled_2d<CC, kWarps, true>;
  B2N_SMEM_OPT_IN(kern, smem);
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  dim3 gd((unsigned)a.n_sub_max, (unsigned)ceil_div(a.C, CC), (unsigned)(a.n_traj == 1 ? a.B : 1));
  kern<<<gd, kThreads, smem, st>>>(a, (const float2 *)kdata, (float2 *)grid, map, 0, (float2 *)scratch);
  B2N_LAUNCH_OK("k_adj_tiled_2d<ordered>");
  const int64_t nblk = a.B * ceil_div(a.C, kMergeCoils);
  dim3 gm((unsigned)ceil_div(a.Kprod, 256), (unsigned)(nblk < 65535 ? nblk : 65535));
  k_adj_merge_2d<<<gm, 256, 0, st>>>(a, (const float2 *)scratch, (float2 *)grid);
  B2N_LAUNCH_OK("k_adj_merge_2d");
  return 0;
}

static bool tiled_eligible(const b2n_geom *g, const b2n_points *p, int layout) {
  return g->dtype == B2N_C64 && g->ndim == 2 && layout == B2N_COIL_MAJOR && g->numpoints[0] == kJ &&
         g->numpoints[1] == kJ && p->tile[0] == kTile && p->tile[1] == kTile && p->n_points > 0 &&
         p->sub_cap <= kCap;
}

template <int CC, int QY, int QX>
static int launch_fwd(const InterpArgs<float> &a, const void *grid, void *kdata, cudaStream_t st) {
  const size_t smem = sizeof(float2) * (planes<CC>() * kPS + kCap * kNC) + sizeof(int2) * kCap + sizeof(int) * kCap + 16;
  auto kern = k_fwd_tiled_2d<CC, QY, QX>;
  B2N_SMEM_OPT_IN(kern, smem);
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  const int use_tma = make_grid_tmap(&map, grid, a.B, a.C, a.K[0], a.K[1]) ? 1 : 0;
  dim3 gd((unsigned)a.n_sub_max, (unsigned)ceil_div(a.C, CC), (unsigned)(a.n_traj == 1 ? a.B : 1));
  B2N_CUDA_OK(launch_pdl(kern, gd, dim3(kFwdThreads), smem, st, a, (const float2 *)grid, (float2 *)kdata, map, use_tma));
  B2N_LAUNCH_OK("k_fwd_tiled_2d");
  return 0;
}

template <int CC, int QY, int QX>
static int launch_fwd_persist(const InterpArgs<float> &a, const void *grid, void *kdata, cudaStream_t st) {
  const size_t buf_bytes = sizeof(float2) * (planes<CC>() * kPS + kCap * kNC) + sizeof(int2) * kCap + sizeof(int) * kCap;
  const size_t smem = kPersistBufs * buf_bytes + kPersistBufs * sizeof(uint64_t);
  auto kern = k_fwd_persist_2d<CC, QY, QX>;
  B2N_SMEM_OPT_IN(kern, smem);
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  const int use_tma = make_grid_tmap(&map, grid, a.B, a.C, a.K[0], a.K[1]) ? 1 : 0;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int n_chunks = (int)ceil_div(a.C, CC), n_batch = (int)(a.n_traj == 1 ? a.B : 1);
  kern<<<sms, kPersistThreads, smem, st>>>(a, (const float2 *)grid, (float2 *)kdata, map, use_tma, n_chunks, n_batch);
  B2N_LAUNCH_OK("k_fwd_persist_2d");
  return 0;
}

// Zeroes the adjoint grid as a kernel so that the spread can be its programmatic dependent: the spread's CTAs stage
// their samples and accumulate their tiles while this runs, and wait for it only before their write-back.  The wait
// comes BEFORE the trigger: the spread reads k-space data that the kernel before this one may still be writing, and
// the grid buffer may be memory that kernel still reads.
__global__ void __launch_bounds__(256) k_zero_grid(float4 *__restrict__ p, int64_t n4) {
  griddep_wait();
  griddep_launch();
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) p[i] = z;
}

template <int CC, int NW = kWarps>
static int launch_adj(const InterpArgs<float> &a, const void *kdata, void *grid, cudaStream_t st) {
  const size_t smem =
      sizeof(float2) * (planes<CC>() * kPS + 2 * adj_stage_slots<CC>()) + sizeof(int) * 3 * kRound;
  auto kern = k_adj_tiled_2d<CC, NW>;
  B2N_SMEM_OPT_IN(kern, smem);
  const int64_t n_cells = a.B * a.C * a.Kprod;
  // measured (profiles/r01_h_opts_ab.log): -2 us at 52 MB (320^2 x 16 coils), -4 us at 151 MB (384^2 x 32), but +46 us
  // at 268 MB (8 x 16 x 256^2): k_zero_grid's few CTAs write slower than the memset engine, so large grids keep it
  const bool overlap = g_pdl >= 1 && g_zero_kernel && n_cells % 2 == 0 && !(reinterpret_cast<uintptr_t>(grid) & 15) &&
                       n_cells * (int64_t)sizeof(float2) <= ((int64_t)192 << 20);
  if (overlap) {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t want = ceil_div(n_cells / 2, (int64_t)256 * 8);
    const unsigned blocks = (unsigned)(want < (int64_t)sms * 4 ? (want > 0 ? want : 1) : (int64_t)sms * 4);
    B2N_CUDA_OK(launch_pdl(k_zero_grid, dim3(blocks), dim3(256), 0, st, (float4 *)grid, n_cells / 2));
    B2N_LAUNCH_OK("k_zero_grid");
  } else {
    B2N_CUDA_OK(cudaMemsetAsync(grid, 0, sizeof(float2) * (size_t)n_cells, st));
  }
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  const int use_tma = make_grid_tmap(&map, grid, a.B, a.C, a.K[0], a.K[1]) ? 1 : 0;
  dim3 gd((unsigned)a.n_sub_max, (unsigned)ceil_div(a.C, CC), (unsigned)(a.n_traj == 1 ? a.B : 1));
  if (overlap) {
    B2N_CUDA_OK(launch_pdl(kern, gd, dim3(NW * 32), smem, st, a, (const float2 *)kdata, (float2 *)grid, map, use_tma,
                           (float2 *)nullptr));
  } else {
    kern<<<gd, NW * 32, smem, st>>>(a, (const float2 *)kdata, (float2 *)grid, map, use_tma, (float2 *)nullptr);
  }
  B2N_LAUNCH_OK("k_adj_tiled_2d");
  return 0;
}

static int launch_adj_coilwarp(const InterpArgs<float> &a, const void *kdata, void *grid, cudaStream_t st) {
  constexpr int CC = 16;
  const size_t smem = sizeof(float2) * (CC * kPS + kRoundC * kNC + kRoundC * kJ * kJ + 2 * kRoundC * CC) +
                      sizeof(int2) * 2 * kRoundC + sizeof(int) * 3 * kRoundC;
  B2N_SMEM_OPT_IN(k_adj_coilwarp_2d, smem);
  B2N_CUDA_OK(cudaMemsetAsync(grid, 0, sizeof(float2) * (size_t)(a.B * a.C * a.Kprod), st));
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  const int use_tma = make_grid_tmap(&map, grid, a.B, a.C, a.K[0], a.K[1]) ? 1 : 0;
  dim3 gd((unsigned)a.n_sub_max, (unsigned)ceil_div(a.C, CC), (unsigned)(a.n_traj == 1 ? a.B : 1));
  k_adj_coilwarp_2d<<<gd, kThreads, smem, st>>>(a, (const float2 *)kdata, (float2 *)grid, map, use_tma);
  B2N_LAUNCH_OK("k_adj_coilwarp_2d");
  return 0;
}

// -----------------------------------------------------------------------------------------
// adjoint spread for very few coils (CC <= 4): lanes = the 36 taps of one point (32 + 4), every warp
// accumulates ITS points into a warp-private tile (CC planes of 21 x 22 cells, 3.7 KB each), so the
// read-modify-write needs neither atomics nor ownership tests; the private tiles are summed at the end and
// added to the grid with RED.ADD.F32x2 (441 cells per coil per sub-problem instead of 36 per point).
// Banks: tap (jy, jx) sits at jy*22 + jx; within a half-warp those 16 offsets are distinct mod 16.
// -----------------------------------------------------------------------------------------
constexpr int kTapWarps = 4;
template <int CC>
__global__ void __launch_bounds__(kTapWarps * 32) k_adj_taps_2d(InterpArgs<float> a, const float2 *__restrict__ kdata,
                                                                float2 *__restrict__ grid) {
  constexpr int NT = kTapWarps * 32;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float2 *tiles = reinterpret_cast<float2 *>(smem_raw);      // [kTapWarps][CC][kPS] private accumulators
  float2 *s_coef = tiles + kTapWarps * CC * kPS;             // [kCap][kNC]
  float2 *s_val = s_coef + kCap * kNC;                       // [CC][kCap]
  int2 *s_base = reinterpret_cast<int2 *>(s_val + CC * kCap);  // [kCap]
  int *s_perm = reinterpret_cast<int *>(s_base + kCap);      // [kCap]
  const SubProblem sp = decode<CC>(a);
  if (!sp.valid) return;
  const int Ky = (int)a.K[0], Kx = (int)a.K[1];
  const int C = (int)a.C;
  const int64_t M = a.M;
  {  // plan records of this sub-problem (contiguous in plan order), then the samples they point at
    const float4 *src =
        reinterpret_cast<const float4 *>(reinterpret_cast<const float2 *>(a.coef) + (int64_t)sp.start * kNC);
    float4 *dst = reinterpret_cast<float4 *>(s_coef);
    for (int e = threadIdx.x; e < sp.count * (kNC / 2); e += NT) cp_async16(&dst[e], &src[e]);
    const int2 *bsrc = reinterpret_cast<const int2 *>(a.base) + sp.start;
    for (int e = threadIdx.x; e < sp.count; e += NT) {
      cp_async8(&s_base[e], &bsrc[e], true);
      cp_async4(&s_perm[e], &a.perm[sp.start + e]);
    }
    cp_async_commit();
  }
  for (int e = threadIdx.x; e < kTapWarps * CC * kPS; e += NT) tiles[e] = make_float2(0.f, 0.f);
  cp_async_wait_all();
  __syncthreads();
  for (int e = threadIdx.x; e < CC * sp.count; e += NT) {
    const int cc = e / sp.count, i = e - cc * sp.count;
    s_val[cc * kCap + i] = sp.c0 + cc < C ? kdata[(int64_t)(sp.b * C + sp.c0 + cc) * M + s_perm[i]] : make_float2(0.f, 0.f);
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int jy0 = lane / kJ, jx0 = lane - jy0 * kJ;  // taps 0..31
  const int t1 = 32 + lane, jy1 = t1 / kJ, jx1 = t1 - jy1 * kJ;  // taps 32..35 (lanes 0..3)
  const bool second = lane < kJ * kJ - 32;
  float2 *mine = tiles + warp * CC * kPS;
  for (int i = warp; i < sp.count; i += kTapWarps) {
    const int2 bs = s_base[i];
    float2 *cell = mine + (bs.x - sp.y0) * kSX + (bs.y - sp.x0);
    const float2 *rec = s_coef + i * kNC;
    // conj(cy * cx) for this lane's taps (cy carries the fftshift phase)
    const float2 cy0 = rec[jy0], cx0 = rec[kJ + jx0];
    const float2 w0 = make_float2(fmaf(cy0.x, cx0.x, -cy0.y * cx0.y), -fmaf(cy0.x, cx0.y, cy0.y * cx0.x));
    float2 w1 = make_float2(0.f, 0.f);
    if (second) {
      const float2 cy1 = rec[jy1], cx1 = rec[kJ + jx1];
      w1 = make_float2(fmaf(cy1.x, cx1.x, -cy1.y * cx1.y), -fmaf(cy1.x, cx1.y, cy1.y * cx1.x));
    }
#pragma unroll
    for (int cc = 0; cc < CC; ++cc) {
      const float2 v = s_val[cc * kCap + i];
      float2 *p0 = cell + cc * kPS + jy0 * kSX + jx0;
      float2 t = *p0;
      cmacf(t, w0, v);
      *p0 = t;
      if (second) {
        float2 *p1 = cell + cc * kPS + jy1 * kSX + jx1;
        float2 u = *p1;
        cmacf(u, w1, v);
        *p1 = u;
      }
    }
    __syncwarp();
  }
  __syncthreads();
  // sum the private tiles (fixed order) and add the result to the grid, with the periodic wrap
  for (int e = threadIdx.x; e < CC * kPS; e += NT) {
    const int cc = e / kPS, rem = e - cc * kPS;
    if (sp.c0 + cc >= C) break;
    const int r = rem / kSX, x = rem - r * kSX;
    if (x >= kSX - 1) continue;  // padding column
    float2 v = tiles[e];
#pragma unroll
    for (int w = 1; w < kTapWarps; ++w) {
      const float2 q = tiles[w * CC * kPS + e];
      v.x += q.x;
      v.y += q.y;
    }
    if (v.x == 0.f && v.y == 0.f) continue;
    int gy = sp.y0 + r, gx = sp.x0 + x;
    gy = gy < Ky ? gy : gy % Ky;
    gx = gx < Kx ? gx : gx % Kx;
    atomicAdd(&grid[((int64_t)(sp.b * C + sp.c0 + cc) * Ky + gy) * Kx + gx], v);
  }
}

template <int CC> static int launch_adj_taps(const InterpArgs<float> &a, const void *kdata, void *grid, cudaStream_t st) {
  const size_t smem = sizeof(float2) * (kTapWarps * CC * kPS + kCap * kNC + CC * kCap) + sizeof(int2) * kCap + sizeof(int) * kCap;
  auto kern = k_adj_taps_2d<CC>;
  B2N_SMEM_OPT_IN(kern, smem);
  B2N_CUDA_OK(cudaMemsetAsync(grid, 0, sizeof(float2) * (size_t)(a.B * a.C * a.Kprod), st));
  dim3 gd((unsigned)a.n_sub_max, (unsigned)ceil_div(a.C, CC), (unsigned)(a.n_traj == 1 ? a.B : 1));
  kern<<<gd, kTapWarps * 32, smem, st>>>(a, (const float2 *)kdata, (float2 *)grid);
  B2N_LAUNCH_OK("k_adj_taps_2d");
  return 0;
}

template <int NW>
static int launch_adj_warptile(const InterpArgs<float> &a, const void *kdata, void *grid, cudaStream_t st) {
  constexpr int CC = NW * 8;
  const size_t smem = sizeof(float2) * (CC * kPS + 2 * (kRound * kNC + kRound * CC + kRound)) + sizeof(int) * 3 * kRound;
  auto kern = k_adj_warptile_2d<NW>;
  B2N_SMEM_OPT_IN(kern, smem);
  B2N_CUDA_OK(cudaMemsetAsync(grid, 0, sizeof(float2) * (size_t)(a.B * a.C * a.Kprod), st));
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  const int use_tma = make_grid_tmap(&map, grid, a.B, a.C, a.K[0], a.K[1]) ? 1 : 0;
  dim3 gd((unsigned)a.n_sub_max, (unsigned)ceil_div(a.C, CC), (unsigned)(a.n_traj == 1 ? a.B : 1));
  kern<<<gd, NW * 32, smem, st>>>(a, (const float2 *)kdata, (float2 *)grid, map, use_tma);
  B2N_LAUNCH_OK("k_adj_warptile_2d");
  return 0;
}

// both return 1 when the tiled path does not apply (caller falls back to the generic kernels)
int tiled_forward(const b2n_geom *g, const b2n_points *p, const void *grid, int64_t B, int64_t C, int layout,
                  void *kdata, cudaStream_t st) {
  if (!tiled_eligible(g, p, layout)) return 1;
  InterpArgs<float> a;
  int rc = make_args<float>(g, p, B, C, &a);
  if (rc) return rc;
  if (C > 8 && g_fwd_chunk == 1) return launch_fwd_persist<16, 1, 2>(a, grid, kdata, st);
  if (C > 8 && g_fwd_chunk != 8) return launch_fwd<16, 1, 2>(a, grid, kdata, st);
  if (C > 8) return launch_fwd<8, 2, 2>(a, grid, kdata, st);
  if (C > 4) return launch_fwd<8, 2, 2>(a, grid, kdata, st);
  if (C > 2) return launch_fwd<4, 2, 3>(a, grid, kdata, st);
  if (C > 1) return launch_fwd<2, 2, 6>(a, grid, kdata, st);
  return launch_fwd<1, 3, 6>(a, grid, kdata, st);
}

int tiled_adjoint(const b2n_geom *g, const b2n_points *p, const void *kdata, int64_t B, int64_t C, int layout,
                  void *grid, cudaStream_t st) {
  if (!tiled_eligible(g, p, layout)) return 1;
  InterpArgs<float> a;
  int rc = make_args<float>(g, p, B, C, &a);
  if (rc) return rc;
  const bool wide = C > 8 && g_adj_chunk != 8;
  // variant 0 (auto): warp-owned tile rows for 16-coil CTAs (107 us vs 114 us at BASELINE config 2,
  // profiles/r01_g), warp-private tiles for <= 8 coils; 1 / 2 / 3 force rows / coils / private tiles
  const int v = g_adj_rowwarp;
  if (v == 0 || v == 6) {  // one or two coils: taps-per-lane kernel with warp-private tiles (32 vs 48-64 us at
    // BASELINE config 1; with 3-4 coils it is slower than the 8-coil private tiles: 103 vs 67 us, variant 6 for A/B)
    if (C == 1) return launch_adj_taps<1>(a, kdata, grid, st);
    if (C == 2) return launch_adj_taps<2>(a, kdata, grid, st);
    if (C <= 4 && v == 6) return launch_adj_taps<4>(a, kdata, grid, st);
  }
  if (v == 3 || (v == 0 && !wide))
    return wide ? launch_adj_warptile<2>(a, kdata, grid, st) : launch_adj_warptile<1>(a, kdata, grid, st);
  if (wide && v == 4) return launch_adj<16, 4>(a, kdata, grid, st);
  if (wide && v == 5) return launch_adj<16, 2>(a, kdata, grid, st);
  if (wide) return v == 2 ? launch_adj_coilwarp(a, kdata, grid, st) : launch_adj<16>(a, kdata, grid, st);
  if (C > 8) return launch_adj<8>(a, kdata, grid, st);
  if (C > 4) return launch_adj<8>(a, kdata, grid, st);
  if (C > 2) return launch_adj<4>(a, kdata, grid, st);
  if (C > 1) return launch_adj<2>(a, kdata, grid, st);
  return launch_adj<1>(a, kdata, grid, st);
}

static bool ordered_eligible(const b2n_geom *g, const b2n_points *p, int layout) {
  return tiled_eligible(g, p, layout) && g->grid_size[0] >= kSY && g->grid_size[1] >= kSY && p->sub_slot &&
         p->tile_sub_start;
}

// scratch bytes of the deterministic tiled adjoint, 0 when it does not apply
size_t tiled_adjoint_ordered_bytes(const b2n_geom *g, const b2n_points *p, int64_t B, int64_t C, int layout) {
  if (!ordered_eligible(g, p, layout)) return 0;
  const int64_t Bz = p->n_traj == 1 ? B : 1;
  return sizeof(float2) * (size_t)p->n_sub_max * (size_t)Bz * (size_t)C * kPS;
}

int tiled_adjoint_ordered(const b2n_geom *g, const b2n_points *p, const void *kdata, int64_t B, int64_t C, int layout,
                          void *scratch, size_t scratch_bytes, void *grid, cudaStream_t st) {
  if (!ordered_eligible(g, p, layout))
    return fail_arg(B2N_E_UNSUPPORTED, "ordered adjoint: 2-D complex64 J=6 coil-major grids of at least 21 cells only");
  if (!scratch || scratch_bytes < tiled_adjoint_ordered_bytes(g, p, B, C, layout))
    return fail_arg(B2N_E_ARG, "ordered adjoint: scratch too small");
  if ((reinterpret_cast<uintptr_t>(scratch) & 15) != 0)
    return fail_arg(B2N_E_ARG, "ordered adjoint: scratch must be 16-byte aligned");
  InterpArgs<float> a;
  int rc = make_args<float>(g, p, B, C, &a);
  if (rc) return rc;
  if (C > 8) return launch_adj_ordered<16>(a, kdata, scratch, grid, st);
  if (C > 4) return launch_adj_ordered<8>(a, kdata, scratch, grid, st);
  if (C > 2) return launch_adj_ordered<4>(a, kdata, scratch, grid, st);
  if (C > 1) return launch_adj_ordered<2>(a, kdata, scratch, grid, st);
  return launch_adj_ordered<1>(a, kdata, scratch, grid, st);
}

}  // namespace b2n
