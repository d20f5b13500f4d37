// b2n_interp_tiled.cu -- shared-memory tiled gather / spread kernels (complex64, 2-D, J=6).
//
// Measured motivation (profiles/r01_a_*, r01_b_*): the per-point kernels pull every point's
// J^d-cell footprint through L2 (590 MB for BASELINE config 2 against a 52 MB grid) and sit
// on the L2->SM bandwidth; a first tiled version was bound by the latency of dependent
// global loads inside its per-point loop.  This version keeps ALL global traffic of the main
// kernels contiguous and asynchronous:
//   * a CTA owns one sub-problem of the trajectory plan (<= sub_cap consecutive points whose
//     base cell lies in one 16x16 tile) and stages that tile plus its J-1 halo once in shared
//     memory with cp.async, transposed to channel-last [cell][coil] (+1 padding): a warp
//     works on one point with lanes = (cell slot, coil), so every shared-memory access is a
//     run of consecutive coils -> conflict-free;
//   * the per-point plan records (separable weights cy[jy], cx[jx] and the base cell) are
//     contiguous in plan order and are bulk-staged with 16-byte cp.async;
//   * k-space samples cross the kernel boundary in plan order, channel-last ([slot][coil], in
//     a scratch buffer): the forward stores / the adjoint stages one contiguous 128-byte line
//     per point.  Two small transposing kernels convert between that order and the caller's
//     (B, C, M) layout, fully coalesced on both sides, and apply the fftshift phase.
// Adjoint accumulation: each of the 8 warps owns the tile rows r with r mod 8 == warp, so
// the update is a plain shared-memory read-modify-write (shared float atomics are CAS loops
// on sm_100a) with a fixed per-cell order; tiles are merged into the global grid with 8-byte
// L2 reductions (RED.ADD.F32x2).
//
// reference loops replaced: torchkbnufft/_nufft/interp.py:185-203 and :689-724.
#include "b2n_common.cuh"
#include "b2n_interp.cuh"

namespace b2n {

constexpr int kTile = 16;    // must match make_tiling() for ndim == 2
constexpr int kWarps = 8;    // warps per CTA (== row-ownership modulus of the adjoint)
constexpr int kThreads = kWarps * 32;
constexpr int kCap = 128;    // max points per sub-problem the forward kernel stages at once
constexpr int kRound = 32;   // points per staging round of the adjoint kernel

B2N_D float2 cmulf(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
B2N_D void cmacf(float2 &acc, float2 a, float2 b) {
  acc.x += a.x * b.x - a.y * b.y;
  acc.y += a.x * b.y + a.y * b.x;
}
// acc += conj(a) * b
B2N_D void cmacf_conj(float2 &acc, float2 a, float2 b) {
  acc.x += a.x * b.x + a.y * b.y;
  acc.y += a.x * b.y - a.y * b.x;
}

// ---- cp.async (LDGSTS) helpers ----------------------------------------------------------
B2N_D void cp_async8(void *smem_dst, const void *gmem_src, bool valid) {
  const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int src_size = valid ? 8 : 0;  // 0 -> the 8 destination bytes are zero-filled
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(dst), "l"(gmem_src), "r"(src_size) : "memory");
}
B2N_D void cp_async16(void *smem_dst, const void *gmem_src) {
  const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(gmem_src) : "memory");
}
B2N_D void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
B2N_D void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

struct SubProblem {
  int b, c0, y0, x0, start, count;
  int64_t row0;  // scratch row of sorted slot 0 for this batch element
  bool valid;
};

// decode blockIdx -> (sub-problem, coil chunk, batch element)
template <int CC> B2N_D SubProblem decode(const InterpArgs<float> &a) {
  SubProblem sp;
  sp.valid = (int)blockIdx.x < *a.n_sub;
  if (!sp.valid) return sp;
  const int tile_all = a.sub_tile[blockIdx.x];
  const int n_tiles = (int)a.tiling.n_tiles;
  const int traj = tile_all / n_tiles, tid = tile_all - traj * n_tiles;
  const int ty = tid / a.tiling.nt[1], tx = tid - ty * a.tiling.nt[1];
  sp.y0 = ty * kTile;
  sp.x0 = tx * kTile;
  sp.c0 = blockIdx.y * CC;
  sp.b = a.n_traj == 1 ? (int)blockIdx.z : traj;
  sp.row0 = a.n_traj == 1 ? (int64_t)sp.b * a.M : 0;  // batched plans: slots already span all trajectories
  sp.start = a.sub_start[blockIdx.x];
  sp.count = a.sub_count[blockIdx.x];
  return sp;
}

template <int JY, int JX> struct TileShape {
  static constexpr int SY = kTile + JY - 1, SX = kTile + JX - 1;
};
// float2 slots of the staged tile, rounded up so that what follows stays 16-byte aligned
template <int CC, int JY, int JX> constexpr int tile_slots() {
  return ((TileShape<JY, JX>::SY * TileShape<JY, JX>::SX * (CC + 1)) + 1) & ~1;
}

// -----------------------------------------------------------------------------------------
// forward: lanes = (q, c), q = (qy, qx) a QY x QX block of footprint cells, c a coil
// -----------------------------------------------------------------------------------------
template <int CC, int QY, int QX, int JY, int JX>
__global__ void __launch_bounds__(kThreads) k_fwd_tiled_2d(InterpArgs<float> a, const float2 *__restrict__ grid,
                                                           float2 *__restrict__ ysorted) {
  constexpr int SY = TileShape<JY, JX>::SY, SX = TileShape<JY, JX>::SX, CS = CC + 1, NC = JY + JX;
  constexpr int Q = QY * QX, NY = JY / QY, NX = JX / QX;
  static_assert(JY % QY == 0 && JX % QX == 0 && Q * CC <= 32, "bad lane mapping");
  static_assert((NC * 8) % 16 == 0, "plan records must be 16-byte multiples");
  extern __shared__ __align__(16) float2 smem[];
  float2 *tile = smem;                               // [SY*SX][CS]
  float2 *s_coef = tile + tile_slots<CC, JY, JX>();  // [kCap][NC]
  int2 *s_base = (int2 *)(s_coef + kCap * NC);       // [kCap]
  const SubProblem sp = decode<CC>(a);
  if (!sp.valid) return;
  const int Ky = (int)a.K[0], Kx = (int)a.K[1];
  const int C = (int)a.C;

  // ---- stage everything asynchronously: tile (+halo, periodic wrap, transposed), records
  for (int e = threadIdx.x; e < CC * SY * SX; e += kThreads) {
    const int c = e / (SY * SX), rem = e - c * (SY * SX);
    const int r = rem / SX, x = rem - r * SX;
    int gy = sp.y0 + r, gx = sp.x0 + x;
    gy = gy < Ky ? gy : gy % Ky;
    gx = gx < Kx ? gx : gx % Kx;
    const bool on = sp.c0 + c < C;
    cp_async8(&tile[rem * CS + c], &grid[((int64_t)(sp.b * C + (on ? sp.c0 + c : 0)) * Ky + gy) * Kx + gx], on);
  }
  {
    const float4 *src =
        reinterpret_cast<const float4 *>(reinterpret_cast<const float2 *>(a.coef) + (int64_t)sp.start * NC);
    float4 *dst = reinterpret_cast<float4 *>(s_coef);
    for (int e = threadIdx.x; e < sp.count * (NC / 2); e += kThreads) cp_async16(&dst[e], &src[e]);
    const int2 *bsrc = reinterpret_cast<const int2 *>(a.base) + sp.start;
    for (int e = threadIdx.x; e < sp.count; e += kThreads) cp_async8(&s_base[e], &bsrc[e], true);
  }
  cp_async_commit();
  cp_async_wait_all();
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = lane % CC, q = lane / CC;
  const bool lane_on = q < Q;  // lanes beyond the QY x QX block idle on cell (0, 0) and contribute zero
  const int qy = lane_on ? q / QX : 0, qx = lane_on ? q - (q / QX) * QX : 0;
  for (int i = warp; i < sp.count; i += kWarps) {
    const int2 bs = s_base[i];
    const int by = bs.x - sp.y0, bx = bs.y - sp.x0;
    const float2 *rec = s_coef + i * NC;
    // this lane's weights: cy[jy] for jy = ny*QY + qy, cx[jx] for jx = nx*QX + qx
    float2 cy[NY], cx[NX];
#pragma unroll
    for (int ny = 0; ny < NY; ++ny) cy[ny] = rec[ny * QY + qy];
#pragma unroll
    for (int nx = 0; nx < NX; ++nx) cx[nx] = rec[JY + nx * QX + qx];
    float2 acc = make_float2(0.f, 0.f);
    const float2 *t0 = tile + ((by + qy) * SX + bx + qx) * CS + c;
#pragma unroll
    for (int ny = 0; ny < NY; ++ny) {
      float2 row = make_float2(0.f, 0.f);
#pragma unroll
      for (int nx = 0; nx < NX; ++nx) cmacf(row, cx[nx], t0[(ny * QY * SX + nx * QX) * CS]);
      cmacf(acc, cy[ny], row);
    }
    if (!lane_on) acc = make_float2(0.f, 0.f);
#pragma unroll
    for (int off = CC; off < 32; off <<= 1) {
      acc.x += __shfl_xor_sync(0xffffffffu, acc.x, off);
      acc.y += __shfl_xor_sync(0xffffffffu, acc.y, off);
    }
    if (q == 0 && sp.c0 + c < C) ysorted[(sp.row0 + sp.start + i) * C + sp.c0 + c] = acc;
  }
}

// -----------------------------------------------------------------------------------------
// adjoint: warp w owns tile rows r with r % 8 == w; lanes = (qx, c), QX cells along x
// -----------------------------------------------------------------------------------------
template <int CC, int JY, int JX>
__global__ void __launch_bounds__(kThreads) k_adj_tiled_2d(InterpArgs<float> a, const float2 *__restrict__ ysorted,
                                                           float2 *__restrict__ grid) {
  constexpr int SY = TileShape<JY, JX>::SY, SX = TileShape<JY, JX>::SX, CS = CC + 1, NC = JY + JX;
  constexpr int QX = 32 / CC, NX = (JX + QX - 1) / QX;
  constexpr int STAGE = kRound * NC + kRound * CC + kRound;  // float2 slots per staging buffer
  static_assert(JY <= kWarps, "row ownership needs JY <= warps per CTA");
  extern __shared__ __align__(16) float2 smem[];
  float2 *tile = smem;                               // [SY*SX][CS] accumulators
  float2 *stage0 = tile + tile_slots<CC, JY, JX>();  // 2 x { coef[kRound][NC], val[kRound][CC], base[kRound] }
  const SubProblem sp = decode<CC>(a);
  if (!sp.valid) return;
  const int Ky = (int)a.K[0], Kx = (int)a.K[1];
  const int C = (int)a.C;
  const float2 *pcoef = reinterpret_cast<const float2 *>(a.coef);

  auto issue = [&](int round) {
    float2 *buf = stage0 + (round & 1) * STAGE;
    const int p0 = round * kRound, nb = min(kRound, sp.count - p0), s0 = sp.start + p0;
    const float4 *src = reinterpret_cast<const float4 *>(pcoef + (int64_t)s0 * NC);
    float4 *dst = reinterpret_cast<float4 *>(buf);
    for (int e = threadIdx.x; e < nb * (NC / 2); e += kThreads) cp_async16(&dst[e], &src[e]);
    float2 *val = buf + kRound * NC;
    for (int e = threadIdx.x; e < nb * CC; e += kThreads) {
      const int i = e / CC, cc = e - i * CC;
      const bool on = sp.c0 + cc < C;
      cp_async8(&val[e], &ysorted[(sp.row0 + s0 + i) * C + (on ? sp.c0 + cc : 0)], on);
    }
    int2 *sb = reinterpret_cast<int2 *>(val + kRound * CC);
    const int2 *bsrc = reinterpret_cast<const int2 *>(a.base) + s0;
    for (int e = threadIdx.x; e < nb; e += kThreads) cp_async8(&sb[e], &bsrc[e], true);
    cp_async_commit();
  };

  issue(0);
  for (int e = threadIdx.x; e < SY * SX * CS; e += kThreads) tile[e] = make_float2(0.f, 0.f);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = lane % CC, qx = lane / CC;
  const int rounds = (sp.count + kRound - 1) / kRound;
  for (int round = 0; round < rounds; ++round) {
    cp_async_wait_all();
    __syncthreads();  // this round's data landed; everyone is done with the buffer refilled next
    if (round + 1 < rounds) issue(round + 1);
    const float2 *buf = stage0 + (round & 1) * STAGE;
    const float2 *s_coef = buf, *s_val = buf + kRound * NC;
    const int2 *s_base = reinterpret_cast<const int2 *>(s_val + kRound * CC);
    const int nb = min(kRound, sp.count - round * kRound);
    for (int i = 0; i < nb; ++i) {
      const int2 bs = s_base[i];
      const int by = bs.x - sp.y0, bx = bs.y - sp.x0;
      const int jy = (warp - by) & (kWarps - 1);  // the footprint row this warp owns, if any
      if (jy >= JY) continue;
      const float2 v = s_val[i * CC + c];
      const float2 cyv = s_coef[i * NC + jy];
      float2 *trow = tile + ((by + jy) * SX + bx) * CS + c;
      const float2 u = make_float2(cyv.x * v.x + cyv.y * v.y, cyv.x * v.y - cyv.y * v.x);  // conj(cy) * v
      float2 t[NX], cxv[NX];
#pragma unroll
      for (int nx = 0; nx < NX; ++nx) {  // all loads first: independent, latency overlaps
        const int jx = nx * QX + qx;
        const bool on = JX % QX == 0 || jx < JX;
        cxv[nx] = s_coef[i * NC + JY + (on ? jx : 0)];
        t[nx] = trow[(on ? jx : 0) * CS];
      }
#pragma unroll
      for (int nx = 0; nx < NX; ++nx) {
        const int jx = nx * QX + qx;
        if (JX % QX == 0 || jx < JX) {
          cmacf_conj(t[nx], cxv[nx], u);  // += conj(cx) * conj(cy) * v
          trow[jx * CS] = t[nx];
        }
      }
      __syncwarp();
    }
  }
  __syncthreads();
  // merge the tile into the global grid: 8-byte L2 reductions, periodic wrap
  for (int e = threadIdx.x; e < CC * SY * SX; e += kThreads) {
    const int cc = e / (SY * SX), rem = e - cc * (SY * SX);
    if (sp.c0 + cc >= C) break;
    const float2 v = tile[rem * CS + cc];
    if (v.x == 0.f && v.y == 0.f) continue;
    const int r = rem / SX, x = rem - r * SX;
    int gy = sp.y0 + r, gx = sp.x0 + x;
    gy = gy < Ky ? gy : gy % Ky;
    gx = gx < Kx ? gx : gx % Kx;
    atomicAdd(&grid[((int64_t)(sp.b * C + sp.c0 + cc) * Ky + gy) * Kx + gx], v);
  }
}

// -----------------------------------------------------------------------------------------
// (B, C, M) caller order <-> plan order channel-last scratch, through a shared-memory
// transpose: both the global reads and the global writes are coalesced.
// TO_SORTED: ysorted[slot][c] = kdata[c][m] * conj(phase[slot])   (adjoint input)
// else:      kdata[c][m] = ysorted[slot][c] * phase[slot]          (forward output)
// -----------------------------------------------------------------------------------------
template <bool TO_SORTED>
__global__ void __launch_bounds__(256) k_reorder_kdata(InterpArgs<float> a, float2 *__restrict__ kdata,
                                                       float2 *__restrict__ ysorted) {
  constexpr int PM = 64, PC = 16;
  __shared__ float2 buf[PC][PM + 1];
  const int64_t M = a.M;
  const int C = (int)a.C;
  const int b = blockIdx.y;
  const int64_t m0 = (int64_t)blockIdx.x * PM;
  const int64_t traj_off = a.n_traj == 1 ? 0 : (int64_t)b * M;  // inv_perm / slots are global over trajectories
  const int64_t row0 = a.n_traj == 1 ? (int64_t)b * M : 0;
  const float2 *phase = reinterpret_cast<const float2 *>(a.phase);
  for (int c0 = 0; c0 < C; c0 += PC) {
    if (TO_SORTED) {
      for (int e = threadIdx.x; e < PC * PM; e += 256) {
        const int c = e / PM, i = e - c * PM;
        if (c0 + c < C && m0 + i < M) buf[c][i] = kdata[((int64_t)b * C + c0 + c) * M + m0 + i];
      }
      __syncthreads();
      for (int e = threadIdx.x; e < PC * PM; e += 256) {
        const int i = e / PC, c = e - i * PC;
        if (c0 + c < C && m0 + i < M) {
          const int slot = a.inv_perm[traj_off + m0 + i];
          const float2 ph = phase[slot];
          ysorted[(row0 + slot) * C + c0 + c] = cmulf(buf[c][i], make_float2(ph.x, -ph.y));
        }
      }
    } else {
      for (int e = threadIdx.x; e < PC * PM; e += 256) {
        const int i = e / PC, c = e - i * PC;
        if (c0 + c < C && m0 + i < M) {
          const int slot = a.inv_perm[traj_off + m0 + i];
          buf[c][i] = cmulf(ysorted[(row0 + slot) * C + c0 + c], phase[slot]);
        }
      }
      __syncthreads();
      for (int e = threadIdx.x; e < PC * PM; e += 256) {
        const int c = e / PM, i = e - c * PM;
        if (c0 + c < C && m0 + i < M) kdata[((int64_t)b * C + c0 + c) * M + m0 + i] = buf[c][i];
      }
    }
    __syncthreads();
  }
}

// ---- dispatch ---------------------------------------------------------------------------
static bool tiled_eligible(const b2n_geom *g, const b2n_points *p, int layout) {
  return g->dtype == B2N_C64 && g->ndim == 2 && layout == B2N_COIL_MAJOR && g->numpoints[0] == 6 &&
         g->numpoints[1] == 6 && p->tile[0] == kTile && p->tile[1] == kTile && p->n_points > 0 &&
         p->sub_cap <= kCap;
}

size_t tiled_scratch_bytes(const b2n_geom *g, const b2n_points *p, int64_t B, int64_t C) {
  if (!tiled_eligible(g, p, B2N_COIL_MAJOR)) return 0;
  return sizeof(float2) * (size_t)(B * C * p->n_points);
}

template <bool TO_SORTED>
static int launch_reorder(const InterpArgs<float> &a, void *kdata, void *ysorted, cudaStream_t st) {
  dim3 gd((unsigned)ceil_div(a.M, 64), (unsigned)a.B);
  k_reorder_kdata<TO_SORTED><<<gd, 256, 0, st>>>(a, (float2 *)kdata, (float2 *)ysorted);
  B2N_LAUNCH_OK("k_reorder_kdata");
  return 0;
}

template <int CC, int QY, int QX>
static int launch_fwd(const InterpArgs<float> &a, const void *grid, void *ysorted, cudaStream_t st) {
  const size_t smem = sizeof(float2) * (tile_slots<CC, 6, 6>() + kCap * 12) + sizeof(int2) * kCap;
  auto kern = k_fwd_tiled_2d<CC, QY, QX, 6, 6>;
  B2N_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 gd((unsigned)a.n_sub_max, (unsigned)ceil_div(a.C, CC), (unsigned)(a.n_traj == 1 ? a.B : 1));
  kern<<<gd, kThreads, smem, st>>>(a, (const float2 *)grid, (float2 *)ysorted);
  B2N_LAUNCH_OK("k_fwd_tiled_2d");
  return 0;
}

template <int CC>
static int launch_adj(const InterpArgs<float> &a, const void *ysorted, void *grid, cudaStream_t st) {
  const size_t smem = sizeof(float2) * (tile_slots<CC, 6, 6>() + 2 * (kRound * 12 + kRound * CC + kRound));
  auto kern = k_adj_tiled_2d<CC, 6, 6>;
  B2N_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 gd((unsigned)a.n_sub_max, (unsigned)ceil_div(a.C, CC), (unsigned)(a.n_traj == 1 ? a.B : 1));
  kern<<<gd, kThreads, smem, st>>>(a, (const float2 *)ysorted, (float2 *)grid);
  B2N_LAUNCH_OK("k_adj_tiled_2d");
  return 0;
}

// both return 1 when the tiled path does not apply (caller falls back to the generic kernels)
int tiled_forward(const b2n_geom *g, const b2n_points *p, const void *grid, int64_t B, int64_t C, int layout,
                  void *kdata, void *scratch, size_t scratch_bytes, cudaStream_t st) {
  if (!tiled_eligible(g, p, layout) || !scratch || scratch_bytes < tiled_scratch_bytes(g, p, B, C)) return 1;
  InterpArgs<float> a;
  int rc = make_args<float>(g, p, B, C, &a);
  if (rc) return rc;
  if (C > 8) rc = launch_fwd<16, 1, 2>(a, grid, scratch, st);
  else if (C > 4) rc = launch_fwd<8, 2, 2>(a, grid, scratch, st);
  else if (C > 2) rc = launch_fwd<4, 2, 3>(a, grid, scratch, st);
  else if (C > 1) rc = launch_fwd<2, 2, 6>(a, grid, scratch, st);
  else rc = launch_fwd<1, 3, 6>(a, grid, scratch, st);
  if (rc) return rc;
  return launch_reorder<false>(a, kdata, scratch, st);
}

int tiled_adjoint(const b2n_geom *g, const b2n_points *p, const void *kdata, int64_t B, int64_t C, int layout,
                  void *grid, void *scratch, size_t scratch_bytes, cudaStream_t st) {
  if (!tiled_eligible(g, p, layout) || !scratch || scratch_bytes < tiled_scratch_bytes(g, p, B, C)) return 1;
  InterpArgs<float> a;
  int rc = make_args<float>(g, p, B, C, &a);
  if (rc) return rc;
  B2N_CUDA_OK(cudaMemsetAsync(grid, 0, sizeof(float2) * (size_t)(a.B * a.C * a.Kprod), st));
  rc = launch_reorder<true>(a, const_cast<void *>(kdata), scratch, st);
  if (rc) return rc;
  if (C > 8) return launch_adj<16>(a, scratch, grid, st);
  if (C > 4) return launch_adj<8>(a, scratch, grid, st);
  if (C > 2) return launch_adj<4>(a, scratch, grid, st);
  if (C > 1) return launch_adj<2>(a, scratch, grid, st);
  return launch_adj<1>(a, scratch, grid, st);
}

}  // namespace b2n
