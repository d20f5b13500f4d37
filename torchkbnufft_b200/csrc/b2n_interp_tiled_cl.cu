// b2n_interp_tiled_cl.cu -- tiled spread for CHANNEL-LAST grids (B, Ky, Kx, C), complex64, 2-D, J = 6,
// C a multiple of 16.
//
// With the coils of a grid cell contiguous (16 coils = one 128-byte line) every tile access is a naturally
// aligned 16-byte load / store of TWO coils, a quarter-warp covers exactly one cell (never a bank conflict, no
// plane-stride tricks), and a warp instruction updates 2 rows x 2 columns x 16 coils.  Compared with the
// coil-major kernel (b2n_interp_tiled.cu) a point costs about half the instructions: 3-4 warp visits of
// 3 LDS.128 + 3 STS.128 instead of 6 visits of 3 LDS.64 + 3 STS.64.
//
// Ownership: warp w owns the tile ROW PAIRS k = r >> 1 with k mod 8 == w; a footprint (rows by .. by+5) touches
// 3 pairs (by even) or 4 (by odd, first and last half-used), so each point is visited by 3-4 warps.
// Lanes: (rs = row of the pair, qx = column parity, p = coil pair): cell (2k + rs, bx + qx + 2 nx), coils 2p, 2p+1.
//
// reference loop replaced: torchkbnufft/_nufft/interp.py:689-724 (table_interp_adjoint).
#include "b2n_tiled_common.cuh"

namespace b2n {

namespace cl {
constexpr int kTile = 16, kWarps = 8, kThreads = kWarps * 32, kRound = 32, kJ = 6;
constexpr int kSY = kTile + kJ - 1;      // 21 staged rows
constexpr int kSX = kTile + kJ - 1 + 1;  // 22 staged columns (the coil-major kernels' tile shape, kept for the plan)
constexpr int kCC = 16;                  // coils per CTA = one 128-byte line per cell
constexpr int kCells = kSY * kSX;        // 462 cells of 16 coils
constexpr int kNC = 2 * kJ;
constexpr int kVS = kRound + 1;          // staged samples: coil-major, odd stride (conflict-free both ways)
constexpr int kStage = (kRound * kNC + kVS * kCC + kRound + 1) & ~1;  // float2 slots: weights, samples, base cells
}  // namespace cl

// 4-D FP32 view of a channel-last complex64 grid (B, Ky, Kx, C): dims (2*C, Kx, Ky, B), box (32, 22, 21, 1)
static bool make_grid_tmap_cl(CUtensorMap *map, const void *grid, int64_t B, int64_t C, int64_t Ky, int64_t Kx) {
  EncodeTiledFn fn = tensor_map_encoder();
  if (!fn || ((uintptr_t)grid & 15) || Ky < cl::kSY || Kx < cl::kSX) return false;
  const cuuint64_t gdim[4] = {(cuuint64_t)(2 * C), (cuuint64_t)Kx, (cuuint64_t)Ky, (cuuint64_t)B};
  const cuuint64_t gstride[3] = {(cuuint64_t)(C * 8), (cuuint64_t)(Kx * C * 8), (cuuint64_t)(Ky * Kx * C * 8)};
  const cuuint32_t box[4] = {2 * cl::kCC, cl::kSX, cl::kSY, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void *>(grid), gdim, gstride, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

__global__ void __launch_bounds__(cl::kThreads, 3) k_adj_tiled_cl_2d(InterpArgs<float> a, const float2 *__restrict__ kdata,
                                                                   float2 *__restrict__ grid,
                                                                   const __grid_constant__ CUtensorMap tmap, int use_tma) {
  using namespace cl;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float2 *tile = reinterpret_cast<float2 *>(smem_raw);          // [kSY][kSX][16 coils]
  float2 *stage0 = tile + kCells * kCC;                          // 2 x kStage
  int *s_perm = reinterpret_cast<int *>(stage0 + 2 * kStage);    // 3 x kRound sample indices
  // decode blockIdx -> sub-problem (x), 16-coil chunk (y), batch element (z)
  const int n_sub = *a.n_sub;
  const int tile_all = a.sub_tile[blockIdx.x];
  const int start = a.sub_start[blockIdx.x], count = a.sub_count[blockIdx.x];
  if ((int)blockIdx.x >= n_sub) return;
  const int n_tiles = (int)a.tiling.n_tiles;
  const int traj = tile_all / n_tiles, tid = tile_all - traj * n_tiles;
  const int ty = tid / a.tiling.nt[1], tx = tid - ty * a.tiling.nt[1];
  const int y0 = ty * kTile, x0 = tx * kTile, c0 = blockIdx.y * kCC;
  const int b = a.n_traj == 1 ? (int)blockIdx.z : traj;
  const int Ky = (int)a.K[0], Kx = (int)a.K[1], C = (int)a.C;
  const bool interior = y0 + kSY <= Ky && x0 + kSX <= Kx;
  const int64_t M = a.M;
  const float2 *pcoef = reinterpret_cast<const float2 *>(a.coef);
  const int rounds = (count + kRound - 1) / kRound;

  auto issue_perm = [&](int round) {  // sample indices, two rounds ahead of their use
    if (round < rounds) {
      const int p0 = round * kRound, nb = min(kRound, count - p0);
      int *dst = s_perm + (round % 3) * kRound;
      for (int e = threadIdx.x; e < nb; e += kThreads) cp_async4(&dst[e], &a.perm[start + p0 + e]);
    }
  };
  auto issue_data = [&](int round) {  // weights, base cells and gathered samples, one round ahead
    if (round < rounds) {
      float2 *buf = stage0 + (round & 1) * kStage;
      const int p0 = round * kRound, nb = min(kRound, count - p0), s0 = start + p0;
      const float4 *src = reinterpret_cast<const float4 *>(pcoef + (int64_t)s0 * kNC);
      float4 *dst = reinterpret_cast<float4 *>(buf);
      for (int e = threadIdx.x; e < nb * (kNC / 2); e += kThreads) cp_async16(&dst[e], &src[e]);
      float2 *val = buf + kRound * kNC;
      const int *perm = s_perm + (round % 3) * kRound;
      for (int e = threadIdx.x; e < kRound * kCC; e += kThreads) {
        const int cc = e / kRound, i = e - cc * kRound;  // consecutive threads: consecutive points, one coil
        const bool on = i < nb;
        cp_async8(&val[cc * kVS + i], &kdata[(int64_t)(b * C + c0 + cc) * M + (on ? perm[i] : 0)], on);
      }
      int2 *sb = reinterpret_cast<int2 *>(val + kVS * kCC);
      const int2 *bsrc = reinterpret_cast<const int2 *>(a.base) + s0;
      for (int e = threadIdx.x; e < nb; e += kThreads) cp_async8(&sb[e], &bsrc[e], true);
    }
  };

  issue_perm(0);
  issue_perm(1);
  cp_async_commit();
  {
    float4 *t4 = reinterpret_cast<float4 *>(tile);
    for (int e = threadIdx.x; e < kCells * kCC / 2; e += kThreads) t4[e] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  cp_async_wait_all();
  __syncthreads();
  issue_data(0);
  cp_async_commit();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int p = lane & 7, qx = (lane >> 3) & 1, rs = lane >> 4;
  float4 *tile4 = reinterpret_cast<float4 *>(tile) + p;  // this lane's coil pair; cell stride = 8 float4
  for (int round = 0; round < rounds; ++round) {
    cp_async_wait_all();
    __syncthreads();  // this round's data (and next round's indices) landed; buffers of round-1 are free
    issue_data(round + 1);
    issue_perm(round + 2);
    cp_async_commit();
    const float2 *buf = stage0 + (round & 1) * kStage;
    const float2 *s_coef = buf, *s_val = buf + kRound * kNC;
    const int2 *s_base = reinterpret_cast<const int2 *>(s_val + kVS * kCC);
    const int nb = min(kRound, count - round * kRound);
    for (int i = 0; i < nb; ++i) {
      const int2 bs = s_base[i];
      const int by = bs.x - y0, bx = bs.y - x0;
      const int k0 = by >> 1;                          // first row pair of the footprint
      const int kk = k0 + ((warp - k0) & (kWarps - 1));  // the row pair this warp owns, if it is one of them
      if (kk > ((by + kJ - 1) >> 1)) continue;
      const int r = 2 * kk + rs, jy = r - by;          // this lane's row and footprint row
      const bool row_on = jy >= 0 && jy < kJ;
      const float2 va = s_val[(2 * p) * kVS + i], vb = s_val[(2 * p + 1) * kVS + i];
      const float2 cyv = s_coef[i * kNC + (row_on ? jy : 0)];
      float2 ua, ub;  // conj(cy) * v for the two coils (cy carries the fftshift phase)
      ua.x = fmaf(cyv.x, va.x, cyv.y * va.y);
      ua.y = fmaf(cyv.x, va.y, -cyv.y * va.x);
      ub.x = fmaf(cyv.x, vb.x, cyv.y * vb.y);
      ub.y = fmaf(cyv.x, vb.y, -cyv.y * vb.x);
      float2 cxv[3];
#pragma unroll
      for (int nx = 0; nx < 3; ++nx) cxv[nx] = s_coef[i * kNC + kJ + 2 * nx + qx];
      if (row_on) {
        float4 *trow = tile4 + (r * kSX + bx + qx) * (kCC / 2);
        float4 t[3];
#pragma unroll
        for (int nx = 0; nx < 3; ++nx) t[nx] = trow[2 * nx * (kCC / 2)];
#pragma unroll
        for (int nx = 0; nx < 3; ++nx) {
          float2 ta = make_float2(t[nx].x, t[nx].y), tb = make_float2(t[nx].z, t[nx].w);
          cmacf_conj(ta, cxv[nx], ua);  // += conj(cx) * conj(cy) * v
          cmacf_conj(tb, cxv[nx], ub);
          trow[2 * nx * (kCC / 2)] = make_float4(ta.x, ta.y, tb.x, tb.y);
        }
      }
      __syncwarp();
    }
  }
  // merge the tile into the global grid
  if (use_tma && interior) {
    fence_async_proxy();  // generic-proxy writes to shared memory -> visible to the TMA engine
    __syncthreads();
    if (threadIdx.x == 0) {
      tma_reduce_add_4d(&tmap, 2 * c0, x0, y0, b, tile);
      tma_store_commit_wait();  // shared memory must stay valid until the engine has read it
    }
  } else {
    __syncthreads();
    for (int e = threadIdx.x; e < kCells * kCC; e += kThreads) {
      const int cell = e / kCC, cc = e - cell * kCC;
      const int r = cell / kSX, x = cell - r * kSX;
      if (x >= kSX - 1) continue;  // the 22nd column is never accumulated
      const float2 v = tile[e];
      if (v.x == 0.f && v.y == 0.f) continue;
      int gy = y0 + r, gx = x0 + x;
      gy = gy < Ky ? gy : gy % Ky;
      gx = gx < Kx ? gx : gx % Kx;
      atomicAdd(&grid[(((int64_t)b * Ky + gy) * Kx + gx) * C + c0 + cc], v);
    }
  }
}

// returns 1 when this path does not apply (caller falls back to the generic channel-last kernels)
int tiled_adjoint_cl(const b2n_geom *g, const b2n_points *p, const void *kdata, int64_t B, int64_t C, int layout,
                     void *grid, cudaStream_t st) {
  using namespace cl;
  if (!(g->dtype == B2N_C64 && g->ndim == 2 && layout == B2N_CHANNEL_LAST && g->numpoints[0] == kJ &&
        g->numpoints[1] == kJ && p->tile[0] == kTile && p->tile[1] == kTile && p->n_points > 0 && C % kCC == 0))
    return 1;
  InterpArgs<float> a;
  int rc = make_args<float>(g, p, B, C, &a);
  if (rc) return rc;
  const size_t smem = sizeof(float2) * (kCells * kCC + 2 * kStage) + sizeof(int) * 3 * kRound;
  B2N_SMEM_OPT_IN(k_adj_tiled_cl_2d, smem);
  B2N_CUDA_OK(cudaMemsetAsync(grid, 0, sizeof(float2) * (size_t)(a.B * a.C * a.Kprod), st));
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  const int use_tma = make_grid_tmap_cl(&map, grid, a.B, a.C, a.K[0], a.K[1]) ? 1 : 0;
  dim3 gd((unsigned)a.n_sub_max, (unsigned)(a.C / kCC), (unsigned)(a.n_traj == 1 ? a.B : 1));
  k_adj_tiled_cl_2d<<<gd, kThreads, smem, st>>>(a, (const float2 *)kdata, (float2 *)grid, map, use_tma);
  B2N_LAUNCH_OK("k_adj_tiled_cl_2d");
  return 0;
}

}  // namespace b2n
