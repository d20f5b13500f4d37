// b2n_interp_tiled3d.cu -- shared-memory tiled gather / spread kernels (complex64, 3-D, J=6).
//
// 3-D analogue of b2n_interp_tiled.cu (BASELINE config 4: 128^3 kooshball, 8 coils).  With
// 216 neighbours per sample the generic per-point kernels move 13.8 KB through L2 per
// (sample, 8 coils) and the adjoint issues 216 L2 reductions per (sample, coil); here a CTA
// owns a sub-problem of the plan (<= 128 samples whose base cell lies in one 8x8x8 tile) and
// stages the tile plus its halo, 13 x 14 x 14 cells for 4 coils (81.5 KB), in shared memory
// by ONE 5-D TMA box (boundary tiles: cp.async with periodic wrap).
//
// Lanes = (8 footprint cells as a 2x2x2 block, 4 coils).  With the coil-plane stride
// 13*14*14 = 2548 = 4 (mod 16 bank pairs) and the y stride 14, a half-warp -- the 2x2 (y, x)
// block of one z parity x 4 coils -- hits 16 distinct bank pairs: conflict-free without padding,
// which is what lets TMA write (forward) and reduce (adjoint) the tile as a dense box.
// Forward: warps take different samples; separable accumulation x -> y -> z (39 complex FMAs
// for 27 cells per lane).  Adjoint: warp w owns the tile z-planes z = w (mod 8), so the
// accumulation is a plain shared-memory read-modify-write with a fixed per-cell order;
// samples are gathered with cp.async a round ahead; the tile is merged into the grid with a
// TMA reduce-add (boundary tiles: RED.ADD.F32x2).
//
// reference loops replaced: torchkbnufft/_nufft/interp.py:185-203 and :689-724 (W = 216).
#include "b2n_tiled_common.cuh"

namespace b2n {

constexpr int k3Tile = 8;   // must match make_tiling() for ndim == 3
constexpr int k3Warps = 8;
constexpr int k3Threads = k3Warps * 32;
// forward gather's own CTA shape: 16 warps (two 100 KB CTAs per SM = 32 resident warps at 56 registers) measured 2.5 %
// faster than 8 at config 4 (7.75 vs 7.95 ms); in 2-D more warps per CTA lose (profiles/r02_gather_cta_shape_ab.log)
#ifndef B2N_FWD3_WARPS
#define B2N_FWD3_WARPS 16
#endif
constexpr int k3FwdWarps = B2N_FWD3_WARPS, k3FwdThreads = k3FwdWarps * 32;
constexpr int k3Cap = 128;  // max points per sub-problem staged at once (forward)
constexpr int k3Round = 32;
constexpr int k3J = 6;
constexpr int k3SZ = k3Tile + k3J - 1;      // 13
constexpr int k3SY = k3Tile + k3J - 1 + 1;  // 14 (one padding row: plane stride = 4 mod 16)
constexpr int k3SX = k3Tile + k3J - 1 + 1;  // 14 (even: 16-byte rows for TMA)
constexpr int k3ZS = k3SY * k3SX;           // z stride, 196
constexpr int k3PS = k3SZ * k3ZS;           // coil-plane stride, 2548
constexpr int k3CC = 4;                     // coils per CTA
constexpr int k3NC = 3 * k3J;               // complex weights per point record
constexpr int k3TileF2 = k3CC * k3PS;       // 10192 float2 = 81 536 B

struct Sub3 {
  int b, c0, z0, y0, x0, start, count;
  bool valid, interior;
};

B2N_D Sub3 decode3(const InterpArgs<float> &a) {
  Sub3 sp;
  const int n_sub = *a.n_sub;
  const int tile_all = a.sub_tile[blockIdx.x];
  sp.start = a.sub_start[blockIdx.x];
  sp.count = a.sub_count[blockIdx.x];
  sp.valid = (int)blockIdx.x < n_sub;
  if (!sp.valid) return sp;
  const int n_tiles = (int)a.tiling.n_tiles;
  const int traj = tile_all / n_tiles;
  int tid = tile_all - traj * n_tiles;
  const int tx = tid % a.tiling.nt[2];
  tid /= a.tiling.nt[2];
  const int ty = tid % a.tiling.nt[1], tz = tid / a.tiling.nt[1];
  sp.z0 = tz * k3Tile;
  sp.y0 = ty * k3Tile;
  sp.x0 = tx * k3Tile;
  sp.c0 = blockIdx.y * k3CC;
  sp.b = a.n_traj == 1 ? (int)blockIdx.z : traj;
  sp.interior = sp.z0 + k3SZ <= (int)a.K[0] && sp.y0 + k3SY <= (int)a.K[1] && sp.x0 + k3SX <= (int)a.K[2];
  return sp;
}

// global element of tile slot e (coil-major [c][z][y][x]) with periodic wrap
B2N_D int64_t tile_global_index(const Sub3 &sp, int e, int C, int Kz, int Ky, int Kx, bool &coil_on, bool &pad) {
  const int c = e / k3PS, rem = e - c * k3PS;
  const int z = rem / k3ZS, r2 = rem - z * k3ZS;
  const int y = r2 / k3SX, x = r2 - y * k3SX;
  pad = y >= k3SY - 1 || x >= k3SX - 1;  // padding row / column: loaded but never part of a footprint
  int gz = sp.z0 + z, gy = sp.y0 + y, gx = sp.x0 + x;
  gz = gz < Kz ? gz : gz % Kz;
  gy = gy < Ky ? gy : gy % Ky;
  gx = gx < Kx ? gx : gx % Kx;
  coil_on = sp.c0 + c < C;
  return (((int64_t)(sp.b * C + (coil_on ? sp.c0 + c : 0)) * Kz + gz) * Ky + gy) * Kx + gx;
}

// lane -> (coil, 2x2x2 cell): lane = qz*16 + qy*8 + qx*4 + c
B2N_D void lane_map3(int lane, int &c, int &qz, int &qy, int &qx) {
  c = lane & 3;
  qx = (lane >> 2) & 1;
  qy = (lane >> 3) & 1;
  qz = lane >> 4;
}

// -----------------------------------------------------------------------------------------
// forward gather
// -----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(k3FwdThreads, 2) k_fwd_tiled_3d(InterpArgs<float> a, const float2 *__restrict__ grid,
                                                               float2 *__restrict__ kdata,
                                                               const __grid_constant__ CUtensorMap tmap, int use_tma) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float2 *tile = reinterpret_cast<float2 *>(smem_raw);          // [4][13][14][14]
  float2 *s_coef = tile + k3TileF2;                             // [k3Cap][18]
  int *s_base = reinterpret_cast<int *>(s_coef + k3Cap * k3NC); // [k3Cap][3]
  int *s_perm = s_base + k3Cap * 3;                             // [k3Cap]
  uint64_t *bar = reinterpret_cast<uint64_t *>(s_perm + k3Cap);
  const Sub3 sp = decode3(a);
  if (!sp.valid) return;
  const int Kz = (int)a.K[0], Ky = (int)a.K[1], Kx = (int)a.K[2];
  const int C = (int)a.C;
  const bool tma = use_tma && sp.interior;

  if (tma) {
    if (threadIdx.x == 0) {
      mbar_init(bar, 1);
      mbar_expect_tx(bar, (unsigned)(k3TileF2 * sizeof(float2)));
      tma_load_5d(tile, &tmap, 2 * sp.x0, sp.y0, sp.z0, sp.c0, sp.b, bar);
    }
  } else {
    for (int e = threadIdx.x; e < k3TileF2; e += k3FwdThreads) {
      bool on, pad;
      const int64_t gi = tile_global_index(sp, e, C, Kz, Ky, Kx, on, pad);
      cp_async8(&tile[e], &grid[gi], on);
    }
  }
  {
    const float4 *src =
        reinterpret_cast<const float4 *>(reinterpret_cast<const float2 *>(a.coef) + (int64_t)sp.start * k3NC);
    float4 *dst = reinterpret_cast<float4 *>(s_coef);
    for (int e = threadIdx.x; e < sp.count * (k3NC / 2); e += k3FwdThreads) cp_async16(&dst[e], &src[e]);
    for (int e = threadIdx.x; e < sp.count * 3; e += k3FwdThreads) cp_async4(&s_base[e], &a.base[(int64_t)sp.start * 3 + e]);
    for (int e = threadIdx.x; e < sp.count; e += k3FwdThreads) cp_async4(&s_perm[e], &a.perm[sp.start + e]);
  }
  cp_async_commit();
  cp_async_wait_all();
  __syncthreads();
  if (tma) mbar_wait(bar, 0);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int c, qz, qy, qx;
  lane_map3(lane, c, qz, qy, qx);
  const float2 *tplane = tile + c * k3PS + qz * k3ZS + qy * k3SX + qx;
  float2 *out = kdata + (int64_t)(sp.b * C + sp.c0 + c) * a.M;
  const bool store = lane < 4 && sp.c0 + c < C;
  for (int i = warp; i < sp.count; i += k3FwdWarps) {
    const float2 *rec = s_coef + i * k3NC;
    float2 cz[3], cy[3], cx[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      cz[k] = rec[2 * k + qz];
      cy[k] = rec[k3J + 2 * k + qy];
      cx[k] = rec[2 * k3J + 2 * k + qx];
    }
    const float2 *tp = tplane + (s_base[3 * i] - sp.z0) * k3ZS + (s_base[3 * i + 1] - sp.y0) * k3SX + (s_base[3 * i + 2] - sp.x0);
    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
    for (int nz = 0; nz < 3; ++nz) {
      float2 g[3][3];
#pragma unroll
      for (int ny = 0; ny < 3; ++ny)
#pragma unroll
        for (int nx = 0; nx < 3; ++nx) g[ny][nx] = tp[2 * nz * k3ZS + 2 * ny * k3SX + 2 * nx];
      float2 row[3];
#pragma unroll
      for (int ny = 0; ny < 3; ++ny) row[ny] = make_float2(0.f, 0.f);
#pragma unroll
      for (int nx = 0; nx < 3; ++nx)
#pragma unroll
        for (int ny = 0; ny < 3; ++ny) cmacf(row[ny], cx[nx], g[ny][nx]);
      float2 plane = make_float2(0.f, 0.f);
#pragma unroll
      for (int ny = 0; ny < 3; ++ny) cmacf(plane, cy[ny], row[ny]);
      cmacf(acc, cz[nz], plane);
    }
#pragma unroll
    for (int off = 4; off < 32; off <<= 1) {
      acc.x += __shfl_xor_sync(0xffffffffu, acc.x, off);
      acc.y += __shfl_xor_sync(0xffffffffu, acc.y, off);
    }
    if (store) out[s_perm[i]] = acc;
  }
}

// -----------------------------------------------------------------------------------------
// adjoint spread: warp w owns the tile z-planes z = w (mod 8)
// -----------------------------------------------------------------------------------------
// ORDERED: store the tile to this sub-problem's scratch slot instead of reducing it into the grid (see
// k_adj_tiled_2d / k_adj_merge_3d): the deterministic mode.
template <bool ORDERED>
__global__ void __launch_bounds__(k3Threads, 2) k_adj_tiled_3d(InterpArgs<float> a, const float2 *__restrict__ kdata,
                                                               float2 *__restrict__ grid,
                                                               const __grid_constant__ CUtensorMap tmap, int use_tma,
                                                               float2 *__restrict__ scratch) {
  constexpr int STAGE_F2 = k3Round * k3NC + k3Round * k3CC;  // coef, val (float2); base ints follow
  constexpr int STAGE_BYTES = STAGE_F2 * 8 + k3Round * 3 * 4;
  static_assert(STAGE_BYTES % 16 == 0, "stage buffers keep 16-byte alignment");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float2 *tile = reinterpret_cast<float2 *>(smem_raw);  // accumulators
  unsigned char *stage0 = smem_raw + k3TileF2 * 8;       // 2 x STAGE_BYTES
  int *s_perm = reinterpret_cast<int *>(stage0 + 2 * STAGE_BYTES);  // 3 x k3Round
  const Sub3 sp = decode3(a);
  if (!sp.valid) return;
  const int Kz = (int)a.K[0], Ky = (int)a.K[1], Kx = (int)a.K[2];
  const int C = (int)a.C;
  const int64_t M = a.M;
  const float2 *pcoef = reinterpret_cast<const float2 *>(a.coef);
  const int rounds = (sp.count + k3Round - 1) / k3Round;

  auto issue_perm = [&](int round) {
    if (round < rounds) {
      const int p0 = round * k3Round, nb = min(k3Round, sp.count - p0);
      int *dst = s_perm + (round % 3) * k3Round;
      for (int e = threadIdx.x; e < nb; e += k3Threads) cp_async4(&dst[e], &a.perm[sp.start + p0 + e]);
    }
  };
  auto issue_data = [&](int round) {
    if (round < rounds) {
      unsigned char *buf = stage0 + (round & 1) * STAGE_BYTES;
      float2 *coef = reinterpret_cast<float2 *>(buf);
      float2 *val = coef + k3Round * k3NC;
      int *sb = reinterpret_cast<int *>(val + k3Round * k3CC);
      const int p0 = round * k3Round, nb = min(k3Round, sp.count - p0), s0 = sp.start + p0;
      const float4 *src = reinterpret_cast<const float4 *>(pcoef + (int64_t)s0 * k3NC);
      float4 *dst = reinterpret_cast<float4 *>(coef);
      for (int e = threadIdx.x; e < nb * (k3NC / 2); e += k3Threads) cp_async16(&dst[e], &src[e]);
      const int *perm = s_perm + (round % 3) * k3Round;
      for (int e = threadIdx.x; e < k3Round * k3CC; e += k3Threads) {
        const int cc = e / k3Round, i = e - cc * k3Round;
        const bool on = sp.c0 + cc < C && i < nb;
        cp_async8(&val[i * k3CC + cc], &kdata[(int64_t)(sp.b * C + (on ? sp.c0 + cc : 0)) * M + (on ? perm[i] : 0)], on);
      }
      for (int e = threadIdx.x; e < nb * 3; e += k3Threads) cp_async4(&sb[e], &a.base[(int64_t)s0 * 3 + e]);
    }
  };

  issue_perm(0);
  issue_perm(1);
  cp_async_commit();
  {
    float4 *t4 = reinterpret_cast<float4 *>(tile);
    for (int e = threadIdx.x; e < k3TileF2 / 2; e += k3Threads) t4[e] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  cp_async_wait_all();
  __syncthreads();
  issue_data(0);
  cp_async_commit();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = lane & 3, qx = (lane >> 2) & 1, qy = (lane >> 3) & 1, h = lane >> 4;
  float2 *tplane = tile + c * k3PS + qy * k3SX + qx;
  for (int round = 0; round < rounds; ++round) {
    cp_async_wait_all();
    __syncthreads();
    issue_data(round + 1);
    issue_perm(round + 2);
    cp_async_commit();
    const unsigned char *buf = stage0 + (round & 1) * STAGE_BYTES;
    const float2 *s_coef = reinterpret_cast<const float2 *>(buf);
    const float2 *s_val = s_coef + k3Round * k3NC;
    const int *s_base = reinterpret_cast<const int *>(s_val + k3Round * k3CC);
    const int nb = min(k3Round, sp.count - round * k3Round);
    for (int i = 0; i < nb; ++i) {
      const int bz = s_base[3 * i] - sp.z0, by = s_base[3 * i + 1] - sp.y0, bx = s_base[3 * i + 2] - sp.x0;
      const int jz = (warp - bz) & (k3Warps - 1);  // the footprint z-plane this warp owns, if any
      if (jz >= k3J) continue;
      const float2 *rec = s_coef + i * k3NC;
      const float2 v = s_val[i * k3CC + c];
      const float2 czv = rec[jz];
      float2 u;  // conj(cz) * v   (cz carries the fftshift phase)
      u.x = fmaf(czv.x, v.x, czv.y * v.y);
      u.y = fmaf(czv.x, v.y, -czv.y * v.x);
      float2 uy[3], cxv[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float2 cyv = rec[k3J + 2 * k + qy];
        uy[k].x = fmaf(cyv.x, u.x, cyv.y * u.y);  // conj(cy) * u
        uy[k].y = fmaf(cyv.x, u.y, -cyv.y * u.x);
        cxv[k] = rec[2 * k3J + 2 * k + qx];
      }
      float2 *tp = tplane + (bz + jz) * k3ZS + by * k3SX + bx;
      // nine 2x2 (y, x) blocks of the 6x6 plane slice, two per warp instruction (halves h = 0, 1)
      float2 t[5];
#pragma unroll
      for (int it = 0; it < 5; ++it) {
        // the half without a ninth block re-reads its own block 7 (a dummy load of another lane's cell would be
        // a read/write hazard inside the warp, flagged by compute-sanitizer racecheck)
        const int blk = 2 * it + h < 9 ? 2 * it + h : 7, ny = blk / 3, nx = blk - ny * 3;
        t[it] = tp[2 * ny * k3SX + 2 * nx];
      }
#pragma unroll
      for (int it = 0; it < 5; ++it) {
        const int blk = 2 * it + h;
        if (blk < 9) {
          const int ny = blk / 3, nx = blk - ny * 3;
          const float2 uyv = ny == 0 ? uy[0] : (ny == 1 ? uy[1] : uy[2]);
          const float2 cx = nx == 0 ? cxv[0] : (nx == 1 ? cxv[1] : cxv[2]);
          cmacf_conj(t[it], cx, uyv);
          tp[2 * ny * k3SX + 2 * nx] = t[it];
        }
      }
      __syncwarp();
    }
  }
  if (ORDERED) {
    __syncthreads();
    const int ncoil = min(k3CC, C - sp.c0);
    const int64_t slot = a.sub_slot[blockIdx.x];
    float4 *dst = reinterpret_cast<float4 *>(scratch + ((slot * gridDim.z + blockIdx.z) * C + sp.c0) * k3PS);
    const float4 *src = reinterpret_cast<const float4 *>(tile);
    for (int e = threadIdx.x; e < ncoil * (k3PS / 2); e += k3Threads) dst[e] = src[e];
    return;
  }
  // merge the tile into the global grid
  if (use_tma && sp.interior) {
    fence_async_proxy();
    __syncthreads();
    if (threadIdx.x == 0) {
      tma_reduce_add_5d(&tmap, 2 * sp.x0, sp.y0, sp.z0, sp.c0, sp.b, tile);
      tma_store_commit_wait();
    }
  } else {
    __syncthreads();
    for (int e = threadIdx.x; e < k3TileF2; e += k3Threads) {
      const float2 v = tile[e];
      if (v.x == 0.f && v.y == 0.f) continue;
      bool on, pad;
      const int64_t gi = tile_global_index(sp, e, C, Kz, Ky, Kx, on, pad);
      if (on && !pad) atomicAdd(&grid[gi], v);
    }
  }
}

// ---- host side ------------------------------------------------------------------------------
// 5-D FP32 view of a coil-major complex64 grid (B, C, Kz, Ky, Kx): dims (2*Kx, Ky, Kz, C, B),
// box (2*14, 14, 13, 4, 1).
static bool make_grid_tmap3(CUtensorMap *map, const void *grid, int64_t B, int64_t C, int64_t Kz, int64_t Ky,
                            int64_t Kx) {
  EncodeTiledFn fn = tensor_map_encoder();
  if (!fn || (Kx & 1) || ((uintptr_t)grid & 15) || Kz < k3SZ || Ky < k3SY || Kx < k3SX) return false;
  const cuuint64_t gdim[5] = {(cuuint64_t)(2 * Kx), (cuuint64_t)Ky, (cuuint64_t)Kz, (cuuint64_t)C, (cuuint64_t)B};
  const cuuint64_t gstride[4] = {(cuuint64_t)(Kx * 8), (cuuint64_t)(Ky * Kx * 8), (cuuint64_t)(Kz * Ky * Kx * 8),
                                 (cuuint64_t)(C * Kz * Ky * Kx * 8)};
  const cuuint32_t box[5] = {2 * k3SX, k3SY, k3SZ, k3CC, 1};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<void *>(grid), gdim, gstride, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static bool tiled3_eligible(const b2n_geom *g, const b2n_points *p, int layout) {
  return g->dtype == B2N_C64 && g->ndim == 3 && layout == B2N_COIL_MAJOR && g->numpoints[0] == k3J &&
         g->numpoints[1] == k3J && g->numpoints[2] == k3J && p->tile[0] == k3Tile && p->tile[1] == k3Tile &&
         p->tile[2] == k3Tile && p->n_points > 0 && p->sub_cap <= k3Cap;
}

// both return 1 when the tiled path does not apply (caller falls back to the generic kernels)
int tiled3_forward(const b2n_geom *g, const b2n_points *p, const void *grid, int64_t B, int64_t C, int layout,
                   void *kdata, cudaStream_t st) {
  if (!tiled3_eligible(g, p, layout)) return 1;
  InterpArgs<float> a;
  int rc = make_args<float>(g, p, B, C, &a);
  if (rc) return rc;
  const size_t smem = sizeof(float2) * (k3TileF2 + k3Cap * k3NC) + sizeof(int) * k3Cap * 4 + 16;
  B2N_SMEM_OPT_IN(k_fwd_tiled_3d, smem);
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  const int use_tma = make_grid_tmap3(&map, grid, a.B, a.C, a.K[0], a.K[1], a.K[2]) ? 1 : 0;
  dim3 gd((unsigned)a.n_sub_max, (unsigned)ceil_div(a.C, k3CC), (unsigned)(a.n_traj == 1 ? a.B : 1));
  k_fwd_tiled_3d<<<gd, k3FwdThreads, smem, st>>>(a, (const float2 *)grid, (float2 *)kdata, map, use_tma);
  B2N_LAUNCH_OK("k_fwd_tiled_3d");
  return 0;
}

int tiled3_adjoint(const b2n_geom *g, const b2n_points *p, const void *kdata, int64_t B, int64_t C, int layout,
                   void *grid, cudaStream_t st) {
  if (!tiled3_eligible(g, p, layout)) return 1;
  InterpArgs<float> a;
  int rc = make_args<float>(g, p, B, C, &a);
  if (rc) return rc;
  const size_t smem = sizeof(float2) * k3TileF2 + 2 * (sizeof(float2) * (k3Round * k3NC + k3Round * k3CC) + sizeof(int) * k3Round * 3) +
                      sizeof(int) * 3 * k3Round;
  B2N_SMEM_OPT_IN(k_adj_tiled_3d<false>, smem);
  B2N_CUDA_OK(cudaMemsetAsync(grid, 0, sizeof(float2) * (size_t)(a.B * a.C * a.Kprod), st));
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  const int use_tma = make_grid_tmap3(&map, grid, a.B, a.C, a.K[0], a.K[1], a.K[2]) ? 1 : 0;
  dim3 gd((unsigned)a.n_sub_max, (unsigned)ceil_div(a.C, k3CC), (unsigned)(a.n_traj == 1 ? a.B : 1));
  k_adj_tiled_3d<false><<<gd, k3Threads, smem, st>>>(a, (const float2 *)kdata, (float2 *)grid, map, use_tma, nullptr);
  B2N_LAUNCH_OK("k_adj_tiled_3d");
  return 0;
}

// -----------------------------------------------------------------------------------------
// deterministic mode: fixed-order merge of the scratch tiles (3-D twin of k_adj_merge_2d): each grid
// cell adds the slots of the <= 3 x 3 x 3 tiles whose 13^3 footprint covers it.  Writes every cell.
// -----------------------------------------------------------------------------------------
constexpr int k3MergeCoils = 8;
__global__ void __launch_bounds__(256) k_adj_merge_3d(InterpArgs<float> a, const float2 *__restrict__ scratch,
                                                      float2 *__restrict__ grid) {
  const int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= a.Kprod) return;
  const int Kz = (int)a.K[0], Ky = (int)a.K[1], Kx = (int)a.K[2], C = (int)a.C;
  const int z = (int)(cell / ((int64_t)Ky * Kx));
  const int yx = (int)(cell - (int64_t)z * Ky * Kx), y = yx / Kx, x = yx - y * Kx;
  const int ntz = a.tiling.nt[0], nty = a.tiling.nt[1], ntx = a.tiling.nt[2];
  const int64_t n_tiles = a.tiling.n_tiles, n_tiles_all = a.n_traj * n_tiles;
  const int n_sub = *a.n_sub;
  const int64_t Bz = a.n_traj == 1 ? a.B : 1;
  const int64_t ncb = (C + k3MergeCoils - 1) / k3MergeCoils, nblk = a.B * ncb;
  constexpr int kFoot = k3Tile + k3J - 1;  // 13 accumulated cells per dimension (rows / columns beyond are padding)
  for (int64_t q = blockIdx.y; q < nblk; q += gridDim.y) {
    const int64_t b = q / ncb, c0 = (q - b * ncb) * k3MergeCoils;
    const int nc = (int)(C - c0 < k3MergeCoils ? C - c0 : k3MergeCoils);
    const int64_t traj = a.n_traj == 1 ? 0 : b, bz = a.n_traj == 1 ? b : 0;
    float2 acc[k3MergeCoils];
#pragma unroll
    for (int k = 0; k < k3MergeCoils; ++k) acc[k] = make_float2(0.f, 0.f);
    for (int kz = 0; kz < min(3, ntz); ++kz) {
      int tz = z / k3Tile - kz;
      if (tz < 0) tz += ntz;
      int rz = z - tz * k3Tile;
      if (rz < 0) rz += Kz;
      if (rz >= kFoot) continue;
      for (int ky = 0; ky < min(3, nty); ++ky) {
        int ty = y / k3Tile - ky;
        if (ty < 0) ty += nty;
        int ry = y - ty * k3Tile;
        if (ry < 0) ry += Ky;
        if (ry >= kFoot) continue;
        for (int kx = 0; kx < min(3, ntx); ++kx) {
          int tx = x / k3Tile - kx;
          if (tx < 0) tx += ntx;
          int rx = x - tx * k3Tile;
          if (rx < 0) rx += Kx;
          if (rx >= kFoot) continue;
          const int64_t t = traj * n_tiles + ((int64_t)tz * nty + ty) * ntx + tx;
          const int s0 = a.tile_sub_start[t], s1 = t + 1 < n_tiles_all ? a.tile_sub_start[t + 1] : n_sub;
          for (int sl = s0; sl < s1; ++sl) {
            const float2 *src = scratch + (((int64_t)sl * Bz + bz) * C + c0) * k3PS + rz * k3ZS + ry * k3SX + rx;
#pragma unroll
            for (int k = 0; k < k3MergeCoils; ++k)
              if (k < nc) {
                const float2 v = src[(int64_t)k * k3PS];
                acc[k].x += v.x;
                acc[k].y += v.y;
              }
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < k3MergeCoils; ++k)
      if (k < nc) grid[(b * C + c0 + k) * a.Kprod + cell] = acc[k];
  }
}

static bool ordered3_eligible(const b2n_geom *g, const b2n_points *p, int layout) {
  constexpr int kFoot = k3Tile + k3J - 1;
  return tiled3_eligible(g, p, layout) && g->grid_size[0] >= kFoot && g->grid_size[1] >= kFoot &&
         g->grid_size[2] >= kFoot && p->sub_slot && p->tile_sub_start;
}

size_t tiled3_adjoint_ordered_bytes(const b2n_geom *g, const b2n_points *p, int64_t B, int64_t C, int layout) {
  if (!ordered3_eligible(g, p, layout)) return 0;
  const int64_t Bz = p->n_traj == 1 ? B : 1;
  return sizeof(float2) * (size_t)p->n_sub_max * (size_t)Bz * (size_t)C * k3PS;
}

int tiled3_adjoint_ordered(const b2n_geom *g, const b2n_points *p, const void *kdata, int64_t B, int64_t C, int layout,
                           void *scratch, size_t scratch_bytes, void *grid, cudaStream_t st) {
  if (!ordered3_eligible(g, p, layout)) return 1;
  if (!scratch || scratch_bytes < tiled3_adjoint_ordered_bytes(g, p, B, C, layout))
    return fail_arg(B2N_E_ARG, "ordered adjoint: scratch too small");
  if ((reinterpret_cast<uintptr_t>(scratch) & 15) != 0)
    return fail_arg(B2N_E_ARG, "ordered adjoint: scratch must be 16-byte aligned");
  InterpArgs<float> a;
  int rc = make_args<float>(g, p, B, C, &a);
  if (rc) return rc;
  const size_t smem = sizeof(float2) * k3TileF2 + 2 * (sizeof(float2) * (k3Round * k3NC + k3Round * k3CC) + sizeof(int) * k3Round * 3) +
                      sizeof(int) * 3 * k3Round;
  B2N_SMEM_OPT_IN(k_adj_tiled_3d<true>, smem);
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  dim3 gd((unsigned)a.n_sub_max, (unsigned)ceil_div(a.C, k3CC), (unsigned)(a.n_traj == 1 ? a.B : 1));
  k_adj_tiled_3d<true><<<gd, k3Threads, smem, st>>>(a, (const float2 *)kdata, (float2 *)grid, map, 0, (float2 *)scratch);
  B2N_LAUNCH_OK("k_adj_tiled_3d<ordered>");
  const int64_t nblk = a.B * ceil_div(a.C, k3MergeCoils);
  dim3 gm((unsigned)ceil_div(a.Kprod, 256), (unsigned)(nblk < 65535 ? nblk : 65535));
  k_adj_merge_3d<<<gm, 256, 0, st>>>(a, (const float2 *)scratch, (float2 *)grid);
  B2N_LAUNCH_OK("k_adj_merge_3d");
  return 0;
}

}  // namespace b2n
