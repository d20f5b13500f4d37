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
// latency than it saves (measured: profiles/r02_spread_notes.txt).  So a visit updates a WINDOW of NR consecutive
// rows [R0, R0 + NR) that covers its rows (seven window classes, own_class) and reads the row weights from a
// zero-padded copy of cy at the (warp-uniform, data-dependent) offset R0 - ry: rows of the window outside the
// footprint get weight zero.  Seven variants of a branch-free body instead of thirteen.
//   acc[r][.] += conj(w) * u,  u = conj(cx[column - rx]) * v  (cx = 0 for columns outside the footprint)
// Packed form (CPL >= 2): accumulators are planar over coil PAIRS -- acc[r][2p] = (re of coil 2p, re of coil 2p+1),
// acc[r][2p+1] the imaginary parts -- so that one FFMA2 (fma.rn.f32x2 with the scalar weight broadcast to both
// halves) updates two coils: 4 FFMA2 per (row, coil pair) instead of 8 FFMA.
template <int CPL, int NR> struct OwnOps {  // operands of one visit, as loaded from the staging buffers
  float2 cx;           // x weight of this lane's column (0 outside the footprint)
  float2 w[NR];        // row weights of the window (0 outside the footprint)
  float4 v[CPL / 2 + 1];  // packed: (re, re', im, im') per coil pair; scalar form: v[0].xy = the sample
};

template <int R0, int NR, int CPL>
B2N_D void own_load(OwnOps<CPL, NR> &o, const int4 *__restrict__ ent, const float2 *__restrict__ coef,
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
  for (int k = 0; k < NR; ++k) o.w[k] = wrow[k];
  const float2 *vp = val + i * CC + (g ^ (i & 3)) * CPL;
  if constexpr (CPL >= 2) {
#pragma unroll
    for (int p = 0; p < CPL / 2; ++p) o.v[p] = reinterpret_cast<const float4 *>(vp)[p];
  } else {
    o.v[0] = make_float4(vp[0].x, vp[0].y, 0.f, 0.f);
  }
}

template <int R0, int NR, int CPL> B2N_D void own_update(float2 (&acc)[kOT][CPL], const OwnOps<CPL, NR> &o) {
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
    for (int k = 0; k < NR; ++k) {
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
    for (int k = 0; k < NR; ++k) cmacf_conj(acc[R0 + k][0], o.w[k], u);
  }
}

// Software-pipelined: the operands of visit i + 1 are fetched from shared memory before the FFMA block of visit i,
// so a warp's two dependent shared-memory latencies per visit overlap its own arithmetic (with 12-16 warps per SM
// there are not enough other warps to hide them).
template <int R0, int NR, int CPL, int U>
B2N_D void own_run(float2 (&acc)[kOT][CPL], const int4 *__restrict__ ent, const float2 *__restrict__ coef,
                   const float2 *__restrict__ val, int i0, int i1, int xl, int g) {
  OwnOps<CPL, NR> cur;
  own_load<R0, NR, CPL>(cur, ent, coef, val, i0, xl, g);
#pragma unroll U
  for (int i = i0; i < i1; ++i) {
    OwnOps<CPL, NR> nxt;
    own_load<R0, NR, CPL>(nxt, ent, coef, val, min(i + 1, i1 - 1), xl, g);
    own_update<R0, NR, CPL>(acc, cur);
    cur = nxt;
  }
}

// Partial tiles of multi-item tiles meet in the scratch (the last item to arrive adds them in chunk order), then the
// tile is stored: 8 lanes (columns) x 8 bytes = one 64-byte segment per (row, coil).
template <int TR, int CPL>
B2N_D void own_finish(const OwnArgs &a, float2 (&acc)[TR][CPL], const int4 tinfo, const int4 item, const int chunk,
                      const int y0, const int x0, const int b, const int c0, const int xl, const int g, const int lane,
                      float2 *__restrict__ grid, float2 *__restrict__ partials, unsigned *__restrict__ counters,
                      const int slot_cap) {
  constexpr bool PACKED = CPL >= 2;
  const int nch = tinfo.z;
  if (nch > 1) {
    // partial tile -> scratch slot; the last item of the tile to arrive adds the slots in chunk order
    const int64_t Bz = gridDim.z, per_slot = (int64_t)Bz * a.n_chunks;
    const int64_t sub = (int64_t)blockIdx.z * a.n_chunks + blockIdx.y;
    if (tinfo.w + nch > slot_cap) __trap();  // the caller's scratch is smaller than the plan needs: fail loudly
    float2 *mine = partials + (((int64_t)(tinfo.w + chunk)) * per_slot + sub) * (TR * CPL * 32);
#pragma unroll
    for (int r = 0; r < TR; ++r)
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
    for (int r = 0; r < TR; ++r)
#pragma unroll
      for (int k = 0; k < CPL; ++k) acc[r][k] = make_float2(0.f, 0.f);
#pragma unroll 2
    for (int j = 0; j < nch; ++j) {
      const float2 *src = partials + (((int64_t)(tinfo.w + j)) * per_slot + sub) * (TR * CPL * 32);
#pragma unroll
      for (int r = 0; r < TR; ++r)
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
        for (int r = 0; r < TR; ++r) {
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

// window class of a footprint row origin ry in [-5, 7]: (first row, rows) =
//   0: (0,2) ry <= -4   1: (0,4) ry in {-3,-2}   2: (0,6) ry in {-1,0}   3: (1,6) ry = 1
//   4: (2,6) ry in {2,3}   5: (4,4) ry in {4,5}   6: (6,2) ry >= 6
// 4.15 rows are updated per visit on average (3.7 carry weight) instead of 6 with one window size.
B2N_D int own_class(int ry) {
  return ry <= -4 ? 0 : (ry <= -2 ? 1 : (ry <= 0 ? 2 : (ry == 1 ? 3 : (ry <= 3 ? 4 : (ry <= 5 ? 5 : 6)))));
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

  const int cl = lane >> 3;
  const float *kdf = reinterpret_cast<const float *>(kd);
  unsigned coil_off[CPL];  // first sample of this lane's q-th coil, in complex elements from kd
#pragma unroll
  for (int q = 0; q < CPL; ++q) coil_off[q] = (unsigned)((c0 + 4 * q + cl < a.C ? c0 + 4 * q + cl : 0) * (int64_t)a.M);

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
    // Addresses: one 64-bit base + 32-bit element offsets (a batch element holds C * M < 2^31 samples).
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int i = (lane & 7) + 8 * half;
      const unsigned m = (unsigned)ent[i].y;
      float *rowf = reinterpret_cast<float *>(val + i * CC);
#pragma unroll
      for (int q = 0; q < CPL; ++q) {
        const int cc = 4 * q + cl, gq = cc / CPL, k = cc - gq * CPL;
        const bool on = c0 + cc < a.C;
        const float *src = kdf + 2 * (size_t)(coil_off[q] + m);
        float *grp = rowf + ((gq ^ (i & 3)) * CPL) * 2;
        if constexpr (PACKED) {
          float *dst = grp + (k >> 1) * 4 + (k & 1);
          cp_async4z(dst, src, on);
          cp_async4z(dst + 2, src + 1, on);
        } else {
          cp_async8(grp + 2 * k, src, on);
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
      // The visits of a tile are listed by window row, i.e. sorted by ry, so the window class changes at most six
      // times per tile: consecutive visits of one class form a run with a branch-free, software-pipelined loop.
      const int my_cls = own_class(ent[lane < nb ? lane : 0].z);
      const int prev_cls = __shfl_up_sync(0xffffffffu, my_cls, 1);
      const unsigned starts = __ballot_sync(0xffffffffu, lane < nb && (lane == 0 || my_cls != prev_cls));
      int i0 = 0;
      while (i0 < nb) {
        const unsigned rest = starts >> (i0 + 1);
        const int i1 = rest ? i0 + __ffs(rest) : nb;
        switch (__shfl_sync(0xffffffffu, my_cls, i0)) {
          case 0: own_run<0, 2, CPL, U>(acc, ent, coef, val, i0, i1, xl, g); break;
          case 1: own_run<0, 4, CPL, U>(acc, ent, coef, val, i0, i1, xl, g); break;
          case 2: own_run<0, 6, CPL, U>(acc, ent, coef, val, i0, i1, xl, g); break;
          case 3: own_run<1, 6, CPL, U>(acc, ent, coef, val, i0, i1, xl, g); break;
          case 4: own_run<2, 6, CPL, U>(acc, ent, coef, val, i0, i1, xl, g); break;
          case 5: own_run<4, 4, CPL, U>(acc, ent, coef, val, i0, i1, xl, g); break;
          default: own_run<6, 2, CPL, U>(acc, ent, coef, val, i0, i1, xl, g); break;
        }
        i0 = i1;
      }
    }
  }

  own_finish<kOT, CPL>(a, acc, tinfo, item, chunk, y0, x0, b, c0, xl, g, lane, grid, partials, counters, slot_cap);
}

// ---- four-row tiles: one branch-free loop ------------------------------------------------------------------------
// The 8 x 8 kernel above pays for its row windows with seven copies of the inner loop: warps of one SM sit in
// different copies and the instruction cache misses ("no instruction" is the top stall reason, 3.6 of 9.9 cycles per
// issue, profiles/r02_spread_notes.txt).  With 4 x 8 tiles every visit updates all four rows: a point is visited by
// 3.66 tiles instead of 2.64, but a visit is 40 FFMA2 in ONE software-pipelined loop, and the staging step already
// writes what the loop reads -- per visit the four window row weights w[k] = cy[k - ry] and the eight column weights
// cx[x - rx], zero outside the footprint (cp.async with a zero source size) -- so the loop has no index arithmetic,
// no predicates and 5 shared-memory loads per visit.
constexpr int kO4R = 4;    // tile rows
constexpr int kO4W = 14;   // float2 slots per staged visit: [w0..w3][cx0..cx7][2 unused]; 112 B keeps 16-byte alignment
                           // and spreads 16 visits over the banks (stride 28 words)

// Sample pre-pass of the four-row kernel.  cp.async pays one shared-memory wavefront per cache line an instruction
// touches, and one staged visit needs the sample of every coil: in the (B, C, M) layout those are 16 different lines
// (measured: 10.6 wavefronts per LDGSTS, half of the kernel's data-pipe work, profiles/r02_spread_notes.txt).  The
// pre-pass transposes the samples once into (B, coil chunk, M, 4 CPL coils), in the form the inner loop reads -- per
// coil pair (re, re', im, im') for CPL >= 2, (re, im) for CPL = 1 -- so that a visit is ONE contiguous 32 CPL byte
// row fetched with 16-byte copies, four (CPL = 4) to sixteen visits per instruction.
template <int CPL>
__global__ void __launch_bounds__(256) k_own_pack(const float2 *__restrict__ kdata, float4 *__restrict__ packed, int C,
                                                  int64_t M, int n_chunks) {
  constexpr int CC = 4 * CPL, MT = 64;
  __shared__ float2 tile[CC][MT + 1];
  const int chunk = blockIdx.y, b = blockIdx.z, t = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * MT;
  griddep_wait();
  {
    const int cm = t & (MT - 1);
#pragma unroll
    for (int cc = t / MT; cc < CC; cc += 256 / MT) {
      const int c = chunk * CC + cc;
      float2 v = make_float2(0.f, 0.f);
      if (c < C && m0 + cm < M) v = __ldg(&kdata[((int64_t)b * C + c) * M + m0 + cm]);
      tile[cc][cm] = v;
    }
  }
  __syncthreads();
  constexpr int PARTS = CC / 2;  // 16-byte parts per sample row
  float4 *out = packed + (((int64_t)b * n_chunks + chunk) * M + m0) * PARTS;
#pragma unroll
  for (int e = t; e < MT * PARTS; e += 256) {
    const int m = e / PARTS, part = e - m * PARTS;
    if (m0 + m < M) {
      const float2 a = tile[2 * part][m], c = tile[2 * part + 1][m];
      out[(int64_t)m * PARTS + part] = CPL >= 2 ? make_float4(a.x, c.x, a.y, c.y) : make_float4(a.x, a.y, c.x, c.y);
    }
  }
}

template <int CPL> constexpr int own4_val_stride() { return 4 * CPL + 2; }  // float2 slots per staged visit (+16 B: banks)
template <int CPL> constexpr size_t own4_smem_bytes() {
  // one spare visit behind the double buffers: the software-pipelined loop loads (and discards) visit nb of a round
  return sizeof(int4) * 3 * kOR + sizeof(float2) * (2 * kOR + 1) * kO4W + sizeof(float2) * (2 * kOR + 1) * own4_val_stride<CPL>();
}

template <int CPL> struct Own4Ops {
  float2 cx;
  float4 w01, w23;
  float4 v[CPL / 2 + 1];
};

// wr: the visit's staged weights (warp-uniform address), cxp = wr + 4 + column, vp: the lane's coil group of the visit
template <int CPL>
B2N_D void own4_load(Own4Ops<CPL> &o, const float2 *__restrict__ wr, const float2 *__restrict__ cxp,
                     const float2 *__restrict__ vp) {
  o.w01 = reinterpret_cast<const float4 *>(wr)[0];
  o.w23 = reinterpret_cast<const float4 *>(wr)[1];
  o.cx = *cxp;
  if constexpr (CPL >= 2) {
#pragma unroll
    for (int p = 0; p < CPL / 2; ++p) o.v[p] = reinterpret_cast<const float4 *>(vp)[p];
  } else {
    o.v[0] = make_float4(vp[0].x, vp[0].y, 0.f, 0.f);
  }
}

template <int CPL> B2N_D void own4_update(float2 (&acc)[kO4R][CPL], const Own4Ops<CPL> &o) {
  const float2 w[kO4R] = {make_float2(o.w01.x, o.w01.y), make_float2(o.w01.z, o.w01.w), make_float2(o.w23.x, o.w23.y),
                          make_float2(o.w23.z, o.w23.w)};
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
    for (int k = 0; k < kO4R; ++k) {
      const float2 wx = make_float2(w[k].x, w[k].x), wy = make_float2(w[k].y, w[k].y), nwy = make_float2(-w[k].y, -w[k].y);
#pragma unroll
      for (int p = 0; p < CPL / 2; ++p) {  // re += w.x uR + w.y uI,  im += w.x uI - w.y uR
        acc[k][2 * p] = __ffma2_rn(wx, uR[p], acc[k][2 * p]);
        acc[k][2 * p + 1] = __ffma2_rn(wx, uI[p], acc[k][2 * p + 1]);
        acc[k][2 * p] = __ffma2_rn(wy, uI[p], acc[k][2 * p]);
        acc[k][2 * p + 1] = __ffma2_rn(nwy, uR[p], acc[k][2 * p + 1]);
      }
    }
  } else {
    const float2 v = make_float2(o.v[0].x, o.v[0].y);
    float2 u;  // conj(cx) * v
    u.x = fmaf(o.cx.x, v.x, o.cx.y * v.y);
    u.y = fmaf(o.cx.x, v.y, -o.cx.y * v.x);
#pragma unroll
    for (int k = 0; k < kO4R; ++k) cmacf_conj(acc[k][0], w[k], u);
  }
}

template <int CPL, int MINB, int U, bool PRE>
__global__ void __launch_bounds__(32, MINB) k_adj_own4_2d(OwnArgs a, const float2 *__restrict__ kdata, float2 *__restrict__ grid,
                                                    float2 *__restrict__ partials, unsigned *__restrict__ counters,
                                                    int slot_cap) {
  constexpr int CC = 4 * CPL, VS = own4_val_stride<CPL>();
  constexpr bool PACKED = CPL >= 2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int4 *s_ent = reinterpret_cast<int4 *>(smem_raw);                 // [3][kOR]
  float2 *s_wts = reinterpret_cast<float2 *>(s_ent + 3 * kOR);      // [2][kOR][kO4W] window weights (+ 1 spare visit)
  float2 *s_val = s_wts + (2 * kOR + 1) * kO4W;                      // [2][kOR][VS] samples, planar per coil pair
  const int lane = threadIdx.x;
  const int n_items = a.counts[0];
  const int4 item = a.items[blockIdx.x];
  if ((int)blockIdx.x >= n_items) return;
  const int4 tinfo = a.tiles[item.x];
  const int y0 = (item.w >> 16) * kO4R, x0 = (item.w & 0xffff) * kOT;
  const int b = a.n_traj == 1 ? (int)blockIdx.z : item.x / (int)a.n_own_tiles;
  const int c0 = blockIdx.y * CC;
  const int n = item.z & 0xfff, chunk = item.z >> 12;
  const int rounds = (n + kOR - 1) / kOR;
  const int4 *vis = a.visits + item.y;

  // staging roles, fixed per lane.  Samples: lane = (visit i8 of 8, coil cl of 4); the q-th copy moves coil 4 q + cl,
  // i.e. coil cl of coil group q.  Weights: lane = (visit i16 of 16, half h).  Addresses are kept cheap on purpose (the
  // staging is ~1/8 of the kernel's instructions): one 64-bit base per lane, a uniform coil stride, immediate offsets.
  const int i8 = lane & 7, cl = lane >> 3, i16 = lane & (kOR - 1), h = lane >> 4;
  const bool all_coils = c0 + CC <= a.C;
  const float2 *kd_lane = kdata + ((int64_t)b * a.C + (c0 + cl < a.C ? c0 + cl : 0)) * a.M;
  const int64_t coil4 = 4 * a.M;  // samples between coil group q and q + 1
  // destination of the lane's copies inside a staging buffer, in floats: visit row + position of coil cl in its group
  const int val_dst = i8 * VS * 2 + (PACKED ? (cl >> 1) * 4 + (cl & 1) : cl * 2);
  const int wts_dst = i16 * kO4W + 6 * h;
  // pre-packed input: first sample row of this (batch element, coil chunk), at the lane's 16-byte part
  const float4 *pk_lane = reinterpret_cast<const float4 *>(kdata) + ((int64_t)b * a.n_chunks + blockIdx.y) * a.M * (2 * CPL) +
                          lane % (2 * CPL);

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
    float2 *wts = s_wts + (round & 1) * kOR * kO4W;
    float *val = reinterpret_cast<float *>(s_val + (round & 1) * kOR * VS);
    {
      // window weights: half 0 writes the four row weights w[k] = cy[k - ry] (slots 0-3) and the column weights
      // cx[x - rx] of columns 0-1 (slots 4-5), half 1 those of columns 2-7 (slots 6-11).  Entries outside the
      // footprint are zero-filled: their copies have a source size of zero and read nothing, so the source address
      // is formed from one base per group with immediate offsets even where it leaves the point's record.
      const int4 e = ent[i16];
      const float2 *rec = a.coef + (int64_t)e.x * kONC;
      const float2 *colp = rec + kOJ + (h ? 2 : -4) - e.w;  // slot 6 h + t (t >= 4 - 4 h) reads colp[t]
      const float2 *rowp = h ? colp : rec - e.z;            // slot t < 4 of half 0 reads rowp[t]
      const int jr = h ? 2 - e.w : -e.z, jc = (h ? 2 : -4) - e.w;
      float2 *dst = wts + wts_dst;
#pragma unroll
      for (int t = 0; t < 4; ++t) cp_async8(dst + t, rowp + t, (unsigned)(jr + t) < (unsigned)kOJ);
#pragma unroll
      for (int t = 4; t < 6; ++t) cp_async8(dst + t, colp + t, (unsigned)(jc + t) < (unsigned)kOJ);
    }
    if constexpr (PRE) {
      // pre-packed samples: kdata is (B, chunk, M, CC coils) in staged form; lane = (visit of 32 / PARTS, 16-byte part)
      constexpr int PARTS = 2 * CPL, VPI = 32 / PARTS;  // parts per visit, visits per instruction
      const int vi = lane / PARTS, part = lane - vi * PARTS;
#pragma unroll
      for (int k = 0; k < kOR / VPI; ++k) {
        const int i = vi + VPI * k;
        cp_async16(val + i * VS * 2 + part * 4, pk_lane + (size_t)(unsigned)ent[i].y * PARTS);
      }
    } else {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const float2 *src = kd_lane + (unsigned)ent[i8 + 8 * half].y;
      float *dst = val + val_dst + half * 8 * VS * 2;
#pragma unroll
      for (int q = 0; q < CPL; ++q) {
        const bool on = all_coils || c0 + 4 * q + cl < a.C;
        const float *sf = reinterpret_cast<const float *>(src);
        if constexpr (PACKED) {
          cp_async4z(dst + q * 8, sf, on);  // coil 4 q + cl = coil cl & 1 of pair 2 q + (cl >> 1): 8 floats per q
          cp_async4z(dst + q * 8 + 2, sf + 1, on);
        } else {
          cp_async8(dst + q * 8, sf, on);
        }
        src += coil4;
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
  float2 acc[kO4R][CPL];
#pragma unroll
  for (int r = 0; r < kO4R; ++r)
#pragma unroll
    for (int k = 0; k < CPL; ++k) acc[r][k] = make_float2(0.f, 0.f);
  const int xl = lane & 7, g = lane >> 3;
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
      const float2 *wr = s_wts + (round & 1) * kOR * kO4W;
      const float2 *cxp = wr + 4 + xl;
      const float2 *vp = s_val + (round & 1) * kOR * VS + g * CPL;
      const int nb = min(kOR, n - round * kOR);
      // software-pipelined: the operands of visit i + 1 are fetched before the FFMA block of visit i (the fetch for
      // i + 1 = nb reads the next buffer or the spare visit and is discarded)
      Own4Ops<CPL> cur;
      own4_load<CPL>(cur, wr, cxp, vp);
#pragma unroll U
      for (int i = 0; i < nb; ++i) {
        wr += kO4W;
        cxp += kO4W;
        vp += VS;
        Own4Ops<CPL> nxt;
        own4_load<CPL>(nxt, wr, cxp, vp);
        own4_update<CPL>(acc, cur);
        cur = nxt;
      }
    }
  }
  own_finish<kO4R, CPL>(a, acc, tinfo, item, chunk, y0, x0, b, c0, xl, g, lane, grid, partials, counters, slot_cap);
}

// ---- host side ------------------------------------------------------------------------------
int g_adj_owned = 1;

static bool own_ready(const b2n_geom *g, const b2n_points *p, int layout) {
  return g_adj_owned && g->dtype == B2N_C64 && g->ndim == 2 && layout == B2N_COIL_MAJOR && g->numpoints[0] == kOJ &&
         g->numpoints[1] == kOJ && (p->own_tile == kOT || p->own_tile == kO4R) && p->own_visits && p->own_items &&
         p->own_tiles && p->own_counts && p->n_points > 0;
}

static int own_cpl(int64_t C) { return C > 8 ? (g_adj_owned == 3 ? 2 : 4) : (C > 4 ? 2 : 1); }  // 3: 8-coil warps (A/B)
// the four-row kernel reads pre-packed samples (k_own_pack) unless B2N_OPT_ADJ_OWNED = 4 (A/B)
static bool own_prepack(const b2n_points *p) { return p->own_tile == kO4R && g_adj_owned != 4; }

// scratch = [arrival counters][pre-packed samples][partial tiles]; the partial tiles are sized by the plan's upper
// bound unless the caller passes the partial-slot count it read back from own_counts[1]
struct OwnScratch {
  size_t ctr, packed, slot, total;  // bytes: counters, packed samples, one partial-tile slot (all its sub-tiles)
};
static OwnScratch own_scratch_layout(const b2n_points *p, int64_t B, int64_t C, int64_t n_slots) {
  OwnScratch o;
  const int cpl = own_cpl(C);
  const int64_t n_chunks = ceil_div(C, 4 * cpl), Bz = p->n_traj == 1 ? B : 1;
  const int64_t n_tiles_all = (int64_t)p->n_own_tiles[0] * p->n_own_tiles[1] * p->n_traj;
  o.ctr = align_up(sizeof(unsigned) * (size_t)(n_tiles_all * Bz * n_chunks), 256);
  o.packed = own_prepack(p) ? align_up(sizeof(float2) * (size_t)(B * n_chunks * 4 * cpl) * (size_t)p->n_points, 256) : 0;
  o.slot = sizeof(float2) * (size_t)(Bz * n_chunks) * (size_t)(p->own_tile * cpl * 32);
  o.total = o.ctr + o.packed + o.slot * (size_t)(n_slots > 0 ? n_slots : p->n_own_items_max);
  return o;
}

size_t own_adjoint_bytes(const b2n_geom *g, const b2n_points *p, int64_t B, int64_t C, int layout, int64_t n_slots,
                         size_t *zero_bytes) {
  if (zero_bytes) *zero_bytes = 0;
  if (!own_ready(g, p, layout) || C * p->n_points >= ((int64_t)1 << 31)) return 0;
  const OwnScratch o = own_scratch_layout(p, B, C, n_slots);
  if (zero_bytes) *zero_bytes = o.ctr;
  return o.total;
}

struct OwnLaunch {
  const b2n_points *p;
  const void *kdata;
  int64_t B, C;
  char *scratch;
  OwnScratch lay;
  int slot_cap;
  void *grid;
  cudaStream_t st;
};

template <int CPL, int MINB, int U>
static int launch_own(const OwnArgs &a, const OwnLaunch &l) {
  auto kern = k_adj_own_2d<CPL, MINB, U>;
  const size_t smem = own_smem_bytes<CPL>();
  dim3 gd((unsigned)l.p->n_own_items_max, (unsigned)a.n_chunks, (unsigned)(l.p->n_traj == 1 ? l.B : 1));
  B2N_CUDA_OK(launch_pdl(kern, gd, dim3(32), smem, l.st, a, (const float2 *)l.kdata, (float2 *)l.grid,
                         (float2 *)(l.scratch + l.lay.ctr + l.lay.packed), (unsigned *)l.scratch, l.slot_cap));
  B2N_LAUNCH_OK("k_adj_own_2d");
  return 0;
}

template <int CPL, int MINB, int U, bool PRE>
static int launch_own4(const OwnArgs &a, const OwnLaunch &l) {
  const void *samples = l.kdata;
  if (PRE) {
    if (!l.lay.packed) return fail_arg(B2N_E_ARG, "owner-tile adjoint: scratch has no room for the packed samples");
    dim3 gp((unsigned)ceil_div(a.M, 64), (unsigned)a.n_chunks, (unsigned)l.B);
    B2N_CUDA_OK(launch_pdl(k_own_pack<CPL>, gp, dim3(256), 0, l.st, (const float2 *)l.kdata,
                           (float4 *)(l.scratch + l.lay.ctr), a.C, a.M, a.n_chunks));
    B2N_LAUNCH_OK("k_own_pack");
    samples = l.scratch + l.lay.ctr;
  }
  auto kern = k_adj_own4_2d<CPL, MINB, U, PRE>;
  const size_t smem = own4_smem_bytes<CPL>();
  dim3 gd((unsigned)l.p->n_own_items_max, (unsigned)a.n_chunks, (unsigned)(l.p->n_traj == 1 ? l.B : 1));
  B2N_CUDA_OK(launch_pdl(kern, gd, dim3(32), smem, l.st, a, (const float2 *)samples, (float2 *)l.grid,
                         (float2 *)(l.scratch + l.lay.ctr + l.lay.packed), (unsigned *)l.scratch, l.slot_cap));
  B2N_LAUNCH_OK("k_adj_own4_2d");
  return 0;
}

// returns 1 when the owner-tile path does not apply
int own_adjoint(const b2n_geom *g, const b2n_points *p, const void *kdata, int64_t B, int64_t C, int layout,
                void *scratch, size_t scratch_bytes, void *grid, cudaStream_t st) {
  if (!own_ready(g, p, layout)) return 1;
  if (B < 1 || C < 1) return fail_arg(B2N_E_ARG, "n_batch=%lld n_coils=%lld", (long long)B, (long long)C);
  if (p->n_traj != 1 && p->n_traj != B)
    return fail_arg(B2N_E_ARG, "plan has %lld trajectories but n_batch=%lld", (long long)p->n_traj, (long long)B);
  if (C * p->n_points >= ((int64_t)1 << 31)) return 1;  // the staging uses 32-bit sample offsets inside a batch element
  OwnLaunch l;
  l.lay = own_scratch_layout(p, B, C, 1);
  if (!scratch || scratch_bytes < l.lay.ctr + l.lay.packed || (reinterpret_cast<uintptr_t>(scratch) & 15))
    return fail_arg(B2N_E_ARG, "owner-tile adjoint: scratch too small or not 16-byte aligned");
  // partial-sum slots the scratch can hold; the kernel traps if the plan needs more (the caller sized the scratch
  // with a slot count that is not the plan's)
  const int64_t cap64 = (int64_t)((scratch_bytes - l.lay.ctr - l.lay.packed) / l.lay.slot);
  l.slot_cap = (int)(cap64 > 0x7fffffff ? 0x7fffffff : cap64);
  l.p = p;
  l.kdata = kdata;
  l.B = B;
  l.C = C;
  l.scratch = (char *)scratch;
  l.grid = grid;
  l.st = st;
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
  if (p->own_tile == kO4R) {
    // four-row tiles: 32 accumulator registers at 16 coils per warp.  B2N_OPT_ADJ_OWNED (A/B): 4 = samples staged
    // straight from (B, C, M) without the pre-pass, 5 / 6 = 24 / 20 warps per SM (80 / 102 registers), 7 = unroll 4
    if (cpl == 4 && g_adj_owned == 4) return launch_own4<4, 16, 2, false>(a, l);
    if (cpl == 4 && g_adj_owned == 5) return launch_own4<4, 24, 2, true>(a, l);
    if (cpl == 4 && g_adj_owned == 6) return launch_own4<4, 20, 2, true>(a, l);
    if (cpl == 4 && g_adj_owned == 7) return launch_own4<4, 16, 4, true>(a, l);
    if (g_adj_owned == 4) {
      if (cpl == 2) return launch_own4<2, 32, 2, false>(a, l);
      return launch_own4<1, 32, 2, false>(a, l);
    }
    if (cpl == 4) return launch_own4<4, 16, 2, true>(a, l);
    if (cpl == 2) return launch_own4<2, 32, 2, true>(a, l);
    return launch_own4<1, 32, 2, true>(a, l);
  }
  // eight-row tiles: the 16-coil kernel is capped at 128 registers (16 warps per SM), the 8-coil one at 85 (24 warps).
  // B2N_OPT_ADJ_OWNED (A/B): 3 = 8-coil warps for C > 8, 4 = 16-coil kernel capped at 168 registers (12 warps),
  // 5 = inner loops not unrolled (smaller code)
  if (cpl == 4 && g_adj_owned == 4) return launch_own<4, 12, 2>(a, l);
  if (cpl == 4 && g_adj_owned == 5) return launch_own<4, 16, 1>(a, l);
  if (cpl == 4) return launch_own<4, 16, 2>(a, l);
  if (cpl == 2) return launch_own<2, 24, 2>(a, l);
  return launch_own<1, 32, 2>(a, l);
}

}  // namespace b2n
