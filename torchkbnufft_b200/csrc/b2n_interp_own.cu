// b2n_interp_own.cu -- output-stationary adjoint spread (complex64, 2-D, J = 6): one warp owns an 8 x 8 OUTPUT tile
// of the grid for a chunk of coils and keeps it in REGISTERS.
//
// Measured motivation (profiles/r01_g_ncu_summary.txt, r02_a): the shared-memory spread k_adj_tiled_2d is bound by the
// read-modify-write of the accumulation tile -- 135 shared-memory wavefronts and 448 warp instructions per point, of
// which 72 wavefronts are the RMW of the 36 x 16 cells themselves -- at 69 % data-pipe utilisation; variants that cut the
// operand loads (warp-private tiles, 121 wavefronts / 244 instructions) lose their gain to the few warps the 59 KB
// tiles leave per SM.  Here the accumulators never touch shared memory:
//   * lanes = 8 tile columns x 4 coil groups of CPL coils; registers = 8 tile rows x CPL coils (64 floats at CPL = 4);
//   * the warp walks the tile's visit list (b2n_points.own_*: every point whose 6 x 6 footprint intersects the tile,
//     with the footprint origin (ry, rx) relative to the tile).  Per visit a lane forms u = conj(cx[col - rx]) * v for
//     its coils (weight 0 when its column is outside the footprint) and adds conj(cy[jy]) * u to the rows ry + jy that
//     fall inside the tile.  The row range is data dependent, register indices are not: a 13-way switch on ry
//     selects a fully unrolled block, so every accumulator index is a compile-time constant;
//   * a point is visited by 1, 2 or 4 tiles (2.64 on average), each (point, cell) pair is still updated exactly once;
//   * every tile is written with plain stores exactly once: no zero-initialised grid, no atomics, and the summation
//     order per cell is fixed by the plan -- the result is bit-reproducible, so this kernel serves both the "atomic"
//     and the "sorted" mode of the API;
//   * tiles with more than own_cap visits (the centre of a radial trajectory) are cut into work items; each item
//     stores its partial tile to a scratch slot, the last one to arrive (a counter per (tile, batch, coil chunk))
//     adds the slots in chunk order and stores the tile.  Counters return to zero, the scratch can be kept;
//   * samples, weights and visit records are staged per warp with cp.async, 16 visits per round, double buffered;
//     there is no block-wide barrier anywhere (CTA = one warp).
//
// reference loops replaced: torchkbnufft/_nufft/interp.py:689-724 and accum_tensor_index_add :407-419.
#include "b2n_tiled_common.cuh"

namespace b2n {

constexpr int kOT = 8;    // owner tile edge (must match kOwnTile in b2n_points.cu)
constexpr int kOJ = 6;    // neighbours per dimension
constexpr int kOR = 16;   // visits per staging round
constexpr int kONC = 2 * kOJ;
constexpr int kOW = 6;           // rows updated per visit: a window of kOW consecutive tile rows
constexpr int kOPAD = 24;        // staged weights per visit: [6 zeros][cy 0..5][6 zeros][cx 0..5] (float2 slots)

struct OwnArgs {
  int Ky, Kx, C, ntx;
  int64_t M, Kprod, n_own_tiles;
  int n_traj, n_chunks;  // coil chunks per (tile, batch element)
  const int4 *visits, *items, *tiles;
  const int32_t *counts;
  const float2 *coef;
};

template <int CPL> constexpr size_t own_smem_bytes() {
  return sizeof(int4) * 3 * kOR + sizeof(float2) * 2 * kOR * kOPAD + sizeof(float2) * 2 * kOR * 4 * CPL;
}

// One run of visits [i0, i1) of the staged round whose row window starts at tile row R0.
//
// A visit with footprint row origin ry touches the tile rows [ry, ry + 6) that exist.  Register indices must be
// compile-time constants, and a dispatch on all 13 values of ry costs more in instruction-cache misses and branch
// latency than it saves (measured: profiles/r02_spread_notes.txt).  So every visit updates a WINDOW of six
// consecutive rows [R0, R0 + 6) with R0 = clamp(ry, 0, 2) -- which always covers its rows -- and reads the six row
// weights from a zero-padded copy of cy at the (warp-uniform, data-dependent) offset R0 - ry: rows of the window
// outside the footprint get weight zero.  Three variants of a branch-free body instead of thirteen.
//   acc[r][.] += conj(w) * u,  u = conj(cx[column - rx]) * v  (cx = 0 for columns outside the footprint)
// Packed form (CPL >= 2): accumulators are planar over coil PAIRS -- acc[r][2p] = (re of coil 2p, re of coil 2p+1),
// acc[r][2p+1] the imaginary parts -- so that one FFMA2 (fma.rn.f32x2 with the scalar weight broadcast to both
// halves) updates two coils: 4 FFMA2 per (row, coil pair) instead of 8 FFMA.
template <int CPL> struct OwnOps {  // operands of one visit, as loaded from the staging buffers
  float2 cx;           // x weight of this lane's column (0 outside the footprint)
  float2 w[kOW];       // row weights of the window (0 outside the footprint)
  float4 v[CPL / 2 + 1];  // packed: (re, re', im, im') per coil pair; scalar form: v[0].xy = the sample
};

template <int R0, int CPL>
B2N_D void own_load(OwnOps<CPL> &o, const int4 *__restrict__ ent, const float2 *__restrict__ coef,
                    const float2 *__restrict__ val, int i, int xl, int g) {
  constexpr int CC = 4 * CPL;
  const int2 rel = reinterpret_cast<const int2 *>(ent + i)[1];  // (ry, rx), warp-uniform
  const float2 *row = coef + i * kOPAD;
  const int jx = xl - rel.y;
  const bool on = (unsigned)jx < (unsigned)kOJ;
  o.cx = row[18 + (on ? jx : 0)];
  if (!on) o.cx = make_float2(0.f, 0.f);
  const float2 *wrow = row + (6 + R0 - rel.x);
#pragma unroll
  for (int k = 0; k < kOW; ++k) o.w[k] = wrow[k];
  const float2 *vp = val + i * CC + (g ^ (i & 3)) * CPL;
  if constexpr (CPL >= 2) {
#pragma unroll
    for (int p = 0; p < CPL / 2; ++p) o.v[p] = reinterpret_cast<const float4 *>(vp)[p];
  } else {
    o.v[0] = make_float4(vp[0].x, vp[0].y, 0.f, 0.f);
  }
}

template <int R0, int CPL> B2N_D void own_update(float2 (&acc)[kOT][CPL], const OwnOps<CPL> &o) {
  if constexpr (CPL >= 2) {
    float2 uR[CPL / 2], uI[CPL / 2];  // real / imaginary parts of conj(cx) * v for two coils
    const float2 cxx = make_float2(o.cx.x, o.cx.x), cxy = make_float2(o.cx.y, o.cx.y),
                 ncxy = make_float2(-o.cx.y, -o.cx.y);
#pragma unroll
    for (int p = 0; p < CPL / 2; ++p) {
      const float2 vR = make_float2(o.v[p].x, o.v[p].y), vI = make_float2(o.v[p].z, o.v[p].w);
      uR[p] = __ffma2_rn(cxy, vI, __fmul2_rn(cxx, vR));
      uI[p] = __ffma2_rn(ncxy, vR, __fmul2_rn(cxx, vI));
    }
#pragma unroll
    for (int k = 0; k < kOW; ++k) {
      const float2 wx = make_float2(o.w[k].x, o.w[k].x), wy = make_float2(o.w[k].y, o.w[k].y),
                   nwy = make_float2(-o.w[k].y, -o.w[k].y);
#pragma unroll
      for (int p = 0; p < CPL / 2; ++p) {  // re += w.x uR + w.y uI,  im += w.x uI - w.y uR
        acc[R0 + k][2 * p] = __ffma2_rn(wx, uR[p], acc[R0 + k][2 * p]);
        acc[R0 + k][2 * p + 1] = __ffma2_rn(wx, uI[p], acc[R0 + k][2 * p + 1]);
        acc[R0 + k][2 * p] = __ffma2_rn(wy, uI[p], acc[R0 + k][2 * p]);
        acc[R0 + k][2 * p + 1] = __ffma2_rn(nwy, uR[p], acc[R0 + k][2 * p + 1]);
      }
    }
  } else {
    const float2 v = make_float2(o.v[0].x, o.v[0].y);
    float2 u;  // conj(cx) * v
    u.x = fmaf(o.cx.x, v.x, o.cx.y * v.y);
    u.y = fmaf(o.cx.x, v.y, -o.cx.y * v.x);
#pragma unroll
    for (int k = 0; k < kOW; ++k) cmacf_conj(acc[R0 + k][0], o.w[k], u);
  }
}

// Software-pipelined: the operands of visit i + 1 are fetched from shared memory before the FFMA block of visit i,
// so a warp's two dependent shared-memory latencies per visit overlap its own arithmetic (with 12-16 warps per SM
// there are not enough other warps to hide them).
template <int R0, int CPL, int U>
B2N_D void own_run(float2 (&acc)[kOT][CPL], const int4 *__restrict__ ent, const float2 *__restrict__ coef,
                   const float2 *__restrict__ val, int i0, int i1, int xl, int g) {
  OwnOps<CPL> cur;
  own_load<R0, CPL>(cur, ent, coef, val, i0, xl, g);
#pragma unroll U
  for (int i = i0; i < i1; ++i) {
    OwnOps<CPL> nxt;
    own_load<R0, CPL>(nxt, ent, coef, val, min(i + 1, i1 - 1), xl, g);
    own_update<R0, CPL>(acc, cur);
    cur = nxt;
  }
}

template <int CPL, int MINB, int U>
__global__ void __launch_bounds__(32, MINB) k_adj_own_2d(OwnArgs a, const float2 *__restrict__ kdata, float2 *__restrict__ grid,
                                                   float2 *__restrict__ partials, unsigned *__restrict__ counters,
                                                   int slot_cap) {
  constexpr int CC = 4 * CPL;
  constexpr bool PACKED = CPL >= 2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int4 *s_ent = reinterpret_cast<int4 *>(smem_raw);                 // [3][kOR]
  float2 *s_coef = reinterpret_cast<float2 *>(s_ent + 3 * kOR);     // [2][kOR][kOPAD] zero-padded weights
  float2 *s_val = s_coef + 2 * kOR * kOPAD;                          // [2][kOR][CC], coil groups swizzled by visit
  const int lane = threadIdx.x;
  // {traj * tiles + tile, first visit, visits | chunk index << 12, tile row << 16 | tile column}; the item array has
  // gridDim.x entries, so the count and the item are fetched together (one global-memory latency, not two)
  const int n_items = a.counts[0];
  const int4 item = a.items[blockIdx.x];
  if ((int)blockIdx.x >= n_items) return;
  const int4 tinfo = a.tiles[item.x];  // {visits, first visit, chunks, first partial slot}; used after the loop
  const int y0 = (item.w >> 16) * kOT, x0 = (item.w & 0xffff) * kOT;
  const int b = a.n_traj == 1 ? (int)blockIdx.z : item.x / (int)a.n_own_tiles;
  const int c0 = blockIdx.y * CC;
  const int n = item.z & 0xfff, chunk = item.z >> 12;
  const int rounds = (n + kOR - 1) / kOR;
  const int4 *vis = a.visits + item.y;
  const float2 *kd = kdata + (int64_t)b * a.C * a.M;

  auto issue_ent = [&](int round) {
    if (round < rounds && lane < kOR) {
      const int i = round * kOR + lane;
      // entries past the end repeat the last one: their samples are staged (never used) from valid addresses
      cp_async16(&s_ent[(round % 3) * kOR + lane], &vis[i < n ? i : n - 1]);
    }
  };
  auto issue_data = [&](int round) {
    if (round >= rounds) return;
    const int4 *ent = s_ent + (round % 3) * kOR;
    float2 *coef = s_coef + (round & 1) * kOR * kOPAD;
    float2 *val = s_val + (round & 1) * kOR * CC;
    // weights: 16 visits x 96 bytes -> cy at slots 6..11, cx at slots 18..23 of the visit's padded row
#pragma unroll
    for (int e = lane; e < kOR * (kONC / 2); e += 32) {
      const int i = e / (kONC / 2), part = e - i * (kONC / 2);
      const float4 *src = reinterpret_cast<const float4 *>(a.coef + (int64_t)ent[i].x * kONC) + part;
      cp_async16(coef + i * kOPAD + (part < 3 ? 6 + 2 * part : 12 + 2 * part), src);
    }
    // samples: lanes = 8 visits x 4 coils per instruction (global: consecutive visits are mostly consecutive samples
    // of one coil; shared: the slots of one instruction cover every bank, each at most twice).  Packed form: the two
    // coils of a pair are stored planar, [re, re', im, im'], so that the FFMA2 operands are aligned register pairs.
    const int cl = lane >> 3;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int i = (lane & 7) + 8 * half;
      const int m = ent[i].y;
#pragma unroll
      for (int q = 0; q < CPL; ++q) {
        const int cc = 4 * q + cl, g = cc / CPL, k = cc - g * CPL;
        const bool on = c0 + cc < a.C;
        const float2 *src = &kd[(int64_t)(on ? c0 + cc : 0) * a.M + m];
        float2 *grp = &val[i * CC + (g ^ (i & 3)) * CPL];
        if constexpr (PACKED) {
          float *dst = reinterpret_cast<float *>(grp) + (k >> 1) * 4 + (k & 1);
          cp_async4z(dst, &src->x, on);
          cp_async4z(dst + 2, &src->y, on);
        } else {
          cp_async8(&grp[k], src, on);
        }
      }
    }
  };

  // the plan does not depend on the kernel that produced the samples: fetch the first visit records before the
  // programmatic-dependent-launch wait
  if (n > 0) {
    issue_ent(0);
    issue_ent(1);
  }
  cp_async_commit();
  griddep_wait();
  float2 acc[kOT][CPL];
#pragma unroll
  for (int r = 0; r < kOT; ++r)
#pragma unroll
    for (int k = 0; k < CPL; ++k) acc[r][k] = make_float2(0.f, 0.f);
  const int xl = lane & 7, g = lane >> 3;
  // the zero slots of the padded weight rows (never overwritten by the staging copies): lane = one of the 2 x 16 rows
  {
    float4 *rowp = reinterpret_cast<float4 *>(s_coef + lane * kOPAD);
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    rowp[0] = rowp[1] = rowp[2] = z;  // slots 0..5
    rowp[6] = rowp[7] = rowp[8] = z;  // slots 12..17
  }
  __syncwarp();
  if (n > 0) {
    cp_async_wait_all();
    __syncwarp();
    issue_data(0);
    cp_async_commit();
    for (int round = 0; round < rounds; ++round) {
      cp_async_wait_all();
      __syncwarp();  // this round's data and the next round's records landed; everyone is done with round - 1
      issue_data(round + 1);
      issue_ent(round + 2);
      cp_async_commit();
      const int4 *ent = s_ent + (round % 3) * kOR;
      const float2 *coef = s_coef + (round & 1) * kOR * kOPAD;
      const float2 *val = s_val + (round & 1) * kOR * CC;
      const int nb = min(kOR, n - round * kOR);
      // The visits of a tile are listed by window row, i.e. sorted by ry, so the window origin R0 = clamp(ry, 0, 2)
      // changes at most twice per tile: consecutive visits with the same R0 form a run with a branch-free loop.
      const int my_r0 = min(max(ent[lane < nb ? lane : 0].z, 0), kOT - kOW);
      const int prev_r0 = __shfl_up_sync(0xffffffffu, my_r0, 1);
      const unsigned starts = __ballot_sync(0xffffffffu, lane < nb && (lane == 0 || my_r0 != prev_r0));
      int i0 = 0;
      while (i0 < nb) {
        const unsigned rest = starts >> (i0 + 1);
        const int i1 = rest ? i0 + __ffs(rest) : nb;
        const int r0 = __shfl_sync(0xffffffffu, my_r0, i0);
        if (r0 == 0) own_run<0, CPL, U>(acc, ent, coef, val, i0, i1, xl, g);
        else if (r0 == 1) own_run<1, CPL, U>(acc, ent, coef, val, i0, i1, xl, g);
        else own_run<2, CPL, U>(acc, ent, coef, val, i0, i1, xl, g);
        i0 = i1;
      }
    }
  }

  const int nch = tinfo.z;
  if (nch > 1) {
    // partial tile -> scratch slot; the last item of the tile to arrive adds the slots in chunk order
    const int64_t Bz = gridDim.z, per_slot = (int64_t)Bz * a.n_chunks;
    const int64_t sub = (int64_t)blockIdx.z * a.n_chunks + blockIdx.y;
    if (tinfo.w + nch > slot_cap) __trap();  // the caller's scratch is smaller than the plan needs: fail loudly
    float2 *mine = partials + (((int64_t)(tinfo.w + chunk)) * per_slot + sub) * (kOT * CPL * 32);
#pragma unroll
    for (int r = 0; r < kOT; ++r)
#pragma unroll
      for (int k = 0; k < CPL; ++k) __stcg(&mine[(r * CPL + k) * 32 + lane], acc[r][k]);
    __threadfence();
    __syncwarp();
    unsigned old = 0;
    unsigned *ctr = counters + (int64_t)item.x * per_slot + sub;
    if (lane == 0) old = atomicAdd(ctr, 1u);
    old = __shfl_sync(0xffffffffu, old, 0);
    if (old != (unsigned)(nch - 1)) return;
    if (lane == 0) *ctr = 0u;  // ready for the next launch
    __threadfence();
#pragma unroll
    for (int r = 0; r < kOT; ++r)
#pragma unroll
      for (int k = 0; k < CPL; ++k) acc[r][k] = make_float2(0.f, 0.f);
#pragma unroll 2
    for (int j = 0; j < nch; ++j) {
      const float2 *src = partials + (((int64_t)(tinfo.w + j)) * per_slot + sub) * (kOT * CPL * 32);
#pragma unroll
      for (int r = 0; r < kOT; ++r)
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
          const float2 p = __ldcg(&src[(r * CPL + k) * 32 + lane]);
          acc[r][k].x += p.x;
          acc[r][k].y += p.y;
        }
    }
  }
  // store the tile: 8 lanes (columns) x 8 bytes = one 64-byte segment per (row, coil)
  if (x0 + xl < a.Kx) {
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
      const int c = c0 + g * CPL + k;
      if (c < a.C) {
        float2 *dst = grid + ((int64_t)b * a.C + c) * a.Kprod + (int64_t)y0 * a.Kx + x0 + xl;
#pragma unroll
        for (int r = 0; r < kOT; ++r) {
          float2 out;
          if constexpr (PACKED) {
            out = (k & 1) ? make_float2(acc[r][k - 1].y, acc[r][k].y) : make_float2(acc[r][k].x, acc[r][k + 1].x);
          } else {
            out = acc[r][k];
          }
          if (y0 + r < a.Ky) dst[(int64_t)r * a.Kx] = out;
        }
      }
    }
  }
}

// ---- host side ------------------------------------------------------------------------------
int g_adj_owned = 1;

static bool own_ready(const b2n_geom *g, const b2n_points *p, int layout) {
  return g_adj_owned && g->dtype == B2N_C64 && g->ndim == 2 && layout == B2N_COIL_MAJOR && g->numpoints[0] == kOJ &&
         g->numpoints[1] == kOJ && p->own_tile == kOT && p->own_visits && p->own_items && p->own_tiles &&
         p->own_counts && p->n_points > 0;
}

static int own_cpl(int64_t C) { return C > 8 ? (g_adj_owned == 3 ? 2 : 4) : (C > 4 ? 2 : 1); }  // 3: 8-coil warps (A/B)

// scratch = [arrival counters][partial tiles]; both sized by the plan's upper bounds unless the caller passes the
// partial-slot count it read back from own_counts[1]
size_t own_adjoint_bytes(const b2n_geom *g, const b2n_points *p, int64_t B, int64_t C, int layout, int64_t n_slots,
                         size_t *zero_bytes) {
  if (zero_bytes) *zero_bytes = 0;
  if (!own_ready(g, p, layout)) return 0;
  const int cpl = own_cpl(C);
  const int64_t n_chunks = ceil_div(C, 4 * cpl), Bz = p->n_traj == 1 ? B : 1;
  const int64_t n_tiles_all = (int64_t)p->n_own_tiles[0] * p->n_own_tiles[1] * p->n_traj;
  const size_t ctr = align_up(sizeof(unsigned) * (size_t)(n_tiles_all * Bz * n_chunks), 256);
  const int64_t slots = n_slots > 0 ? n_slots : p->n_own_items_max;
  if (zero_bytes) *zero_bytes = ctr;
  return ctr + sizeof(float2) * (size_t)slots * (size_t)(Bz * n_chunks) * (size_t)(kOT * cpl * 32);
}

template <int CPL, int MINB, int U>
static int launch_own(const OwnArgs &a, const b2n_points *p, const void *kdata, int64_t B, void *scratch, size_t ctr_bytes,
                      int slot_cap, void *grid, cudaStream_t st) {
  auto kern = k_adj_own_2d<CPL, MINB, U>;
  const size_t smem = own_smem_bytes<CPL>();
  dim3 gd((unsigned)p->n_own_items_max, (unsigned)a.n_chunks, (unsigned)(p->n_traj == 1 ? B : 1));
  B2N_CUDA_OK(launch_pdl(kern, gd, dim3(32), smem, st, a, (const float2 *)kdata, (float2 *)grid,
                         (float2 *)((char *)scratch + ctr_bytes), (unsigned *)scratch, slot_cap));
  B2N_LAUNCH_OK("k_adj_own_2d");
  return 0;
}

// returns 1 when the owner-tile path does not apply
int own_adjoint(const b2n_geom *g, const b2n_points *p, const void *kdata, int64_t B, int64_t C, int layout,
                void *scratch, size_t scratch_bytes, void *grid, cudaStream_t st) {
  if (!own_ready(g, p, layout)) return 1;
  if (B < 1 || C < 1) return fail_arg(B2N_E_ARG, "n_batch=%lld n_coils=%lld", (long long)B, (long long)C);
  if (p->n_traj != 1 && p->n_traj != B)
    return fail_arg(B2N_E_ARG, "plan has %lld trajectories but n_batch=%lld", (long long)p->n_traj, (long long)B);
  size_t ctr = 0;
  const size_t one_slot = own_adjoint_bytes(g, p, B, C, layout, 1, &ctr) - 0;
  if (!scratch || scratch_bytes < ctr || (reinterpret_cast<uintptr_t>(scratch) & 15))
    return fail_arg(B2N_E_ARG, "owner-tile adjoint: scratch too small or not 16-byte aligned");
  // partial-sum slots the scratch can hold; the kernel traps if the plan needs more (the caller sized the scratch
  // with a slot count that is not the plan's)
  const size_t slot_bytes = one_slot - ctr;
  const int64_t cap64 = (int64_t)((scratch_bytes - ctr) / slot_bytes);
  const int slot_cap = (int)(cap64 > 0x7fffffff ? 0x7fffffff : cap64);
  const int cpl = own_cpl(C);
  OwnArgs a;
  a.Ky = (int)g->grid_size[0];
  a.Kx = (int)g->grid_size[1];
  a.C = (int)C;
  a.ntx = p->n_own_tiles[1];
  a.M = p->n_points;
  a.Kprod = g->grid_size[0] * g->grid_size[1];
  a.n_own_tiles = (int64_t)p->n_own_tiles[0] * p->n_own_tiles[1];
  a.n_traj = (int)p->n_traj;
  a.n_chunks = (int)ceil_div(C, 4 * cpl);
  a.visits = (const int4 *)p->own_visits;
  a.items = (const int4 *)p->own_items;
  a.tiles = (const int4 *)p->own_tiles;
  a.counts = p->own_counts;
  a.coef = (const float2 *)p->coef;
  // The 16-coil kernel is capped at 128 registers (16 warps per SM), the 8-coil one at 85 (24 warps).
  // B2N_OPT_ADJ_OWNED (A/B): 3 = 8-coil warps for C > 8, 4 = 16-coil kernel capped at 168 registers (12 warps),
  // 5 = inner loops not unrolled (smaller code)
  if (cpl == 4 && g_adj_owned == 4) return launch_own<4, 12, 2>(a, p, kdata, B, scratch, ctr, slot_cap, grid, st);
  if (cpl == 4 && g_adj_owned == 5) return launch_own<4, 16, 1>(a, p, kdata, B, scratch, ctr, slot_cap, grid, st);
  if (cpl == 4) return launch_own<4, 16, 2>(a, p, kdata, B, scratch, ctr, slot_cap, grid, st);
  if (cpl == 2) return launch_own<2, 24, 2>(a, p, kdata, B, scratch, ctr, slot_cap, grid, st);
  return launch_own<1, 32, 2>(a, p, kdata, B, scratch, ctr, slot_cap, grid, st);
}

}  // namespace b2n
