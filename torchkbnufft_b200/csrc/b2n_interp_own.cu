// b2n_interp_own.cu -- output-stationary adjoint spread (complex64, 2-D, J = 6): one warp owns a 4 x 8 OUTPUT tile of
// the grid for a chunk of coils, keeps it in REGISTERS and works with REAL interpolation weights.
//
// How it got here (profiles/r02_spread_notes.txt has the numbers):
//   * the shared-memory spread k_adj_tiled_2d is bound by the read-modify-write of its accumulation tile (135 shared
//     wavefronts per point, 72 of them the RMW itself, 97 us at BASELINE config 2);
//   * accumulators in registers with 8 x 8 tiles need the rows a visit touches as compile-time register indices:
//     seven copies of the inner loop, "no instruction" (instruction cache) the top stall, 81 us;
//   * 4 x 8 tiles update all four rows on every visit -- ONE branch-free loop -- but cp.async of samples in (B, C, M)
//     order costs a shared-memory wavefront per cache line (10.6 per instruction): 80 us; with the samples
//     transposed once by a pre-pass (k_own_pack) a visit is one contiguous row: 71 us, FMA pipe 62 % busy;
//   * FFMA2 issues at half the FFMA rate on sm_100a (profiles/micro/micro_r2.cu: 2.27 vs 1.07 cycles), so the only
//     way further is less arithmetic.  The reference's tables are a real Kaiser-Bessel kernel times a LINEAR phase
//     (b2n_geom.rtable_dev): the phase of neighbour j of a point factors into a per-point part (folded into the
//     samples by the pre-pass), a per-grid-cell part (applied once when the tile is stored) and a sign for footprints
//     that wrap around the grid (folded into the visit's column weights by the plan).  What is left per visit is
//       acc[r][c] += (hy[r] * hx[col]) * v[c]          (real weight, complex sample: 2 FMA instead of 4 + 4 / row)
//     i.e. 4 FMUL + 16 FFMA2 per lane and visit for 4 rows x 4 coils instead of 8 + 32 FFMA2.
//
// Layout of the work:
//   * lanes = 8 tile columns x 4 coil groups of CPL coils; registers = 4 tile rows x CPL coils, planar over coil
//     PAIRS ((re, re'), (im, im')) so that one FFMA2 with a broadcast scalar weight updates two coils;
//   * the warp walks the tile's visit list (b2n_points.own_visits: 64-byte records {hy[4], hx[8], sample index},
//     written by the plan in a fixed order): records are staged with 16-byte cp.async two rounds ahead, samples one
//     round ahead (16 visits per round), no block-wide barrier anywhere (CTA = one warp);
//   * every tile is stored exactly once with plain stores: no zero-initialised grid, no atomics, and the summation
//     order per cell is fixed by the plan -- bit-reproducible, so this kernel serves both the "atomic" and the
//     "sorted" mode of the API;
//   * tiles with more than own_cap visits (the centre of a radial trajectory) are cut into work items; each item
//     stores its partial tile to a scratch slot, the last one to arrive (a counter per (tile, batch, coil chunk))
//     adds the slots in chunk order and stores the tile.  Counters return to zero, the scratch can be kept.
//
// reference loops replaced: torchkbnufft/_nufft/interp.py:689-724 and accum_tensor_index_add :407-419.
#include "b2n_tiled_common.cuh"

namespace b2n {

constexpr int kOTR = 4;   // owner tile rows (kOwnTileRows in b2n_points.cu)
constexpr int kOTC = 8;   // owner tile columns
constexpr int kOJ = 6;    // neighbours per dimension
constexpr int kOR = 16;   // visits per staging round
constexpr int kOVF = 16;  // floats per visit record: hy[4], hx[8], sample index, 3 unused

struct OwnArgs {
  int Ky, Kx, C;
  int64_t M, Kprod, n_own_tiles;
  int n_traj, n_chunks;  // coil chunks per (tile, batch element)
  const float4 *visits;  // 64-byte records
  const int4 *items, *tiles;
  const int32_t *counts;
  const float2 *q;  // per-cell phase factors: [Ky] rows, then [Kx] columns
};

// ---- sample pre-pass ----------------------------------------------------------------------------------------------
// cp.async pays one shared-memory wavefront per cache line an instruction touches, and a staged visit needs the sample
// of every coil: in the (B, C, M) layout those are 16 different lines.  The pre-pass transposes the samples once into
// (B, coil chunk, M, 4 CPL coils) in the form the inner loop reads -- per coil pair (re, re', im, im') for CPL >= 2,
// (re, im) for CPL = 1 -- and multiplies every sample by its point's phase factor (own_fac, looked up through
// inv_perm), so that a visit is ONE contiguous 32 CPL byte row fetched with 16-byte copies.
template <int CPL, bool PLANAR>
__global__ void __launch_bounds__(256) k_own_pack(const float2 *__restrict__ kdata, float4 *__restrict__ packed, int C,
                                                  int64_t M, int n_chunks, int n_traj, const int32_t *__restrict__ inv_perm,
                                                  const float2 *__restrict__ fac) {
  constexpr int CC = 4 * CPL, MT = 64;
  __shared__ float2 tile[CC][MT + 1];
  const int chunk = blockIdx.y, b = blockIdx.z, t = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * MT;
  const int cm = t & (MT - 1);
  // the plan does not depend on the producer of the samples: fetch the factor before the dependent-launch wait
  float2 f = make_float2(0.f, 0.f);
  if (m0 + cm < M) f = fac[inv_perm[(n_traj == 1 ? 0 : (int64_t)b * M) + m0 + cm]];
  griddep_wait();
#pragma unroll
  for (int cc = t / MT; cc < CC; cc += 256 / MT) {
    const int c = chunk * CC + cc;
    float2 v = make_float2(0.f, 0.f);
    if (c < C && m0 + cm < M) v = __ldg(&kdata[((int64_t)b * C + c) * M + m0 + cm]);
    tile[cc][cm] = make_float2(v.x * f.x - v.y * f.y, v.x * f.y + v.y * f.x);
  }
  __syncthreads();
  constexpr int PARTS = CC / 2;  // 16-byte parts per sample row
  float4 *out = packed + (((int64_t)b * n_chunks + chunk) * M + m0) * PARTS;
#pragma unroll
  for (int e = t; e < MT * PARTS; e += 256) {
    const int m = e / PARTS, part = e - m * PARTS;
    if (m0 + m < M) {
      const float2 a = tile[2 * part][m], c = tile[2 * part + 1][m];
      out[(int64_t)m * PARTS + part] = PLANAR ? make_float4(a.x, c.x, a.y, c.y) : make_float4(a.x, a.y, c.x, c.y);
    }
  }
}

// ---- spread -------------------------------------------------------------------------------------------------------
template <int CPL> constexpr int own_val_stride() { return 4 * CPL + 2; }  // float2 slots per staged visit (+16 B: banks)
template <int CPL> constexpr size_t own_smem_bytes() {
  // one spare visit behind the sample buffers: the software-pipelined loop loads (and discards) visit nb of a round
  return sizeof(float) * 3 * kOR * kOVF + sizeof(float2) * (2 * kOR + 1) * own_val_stride<CPL>();
}

template <int CPL> struct OwnOps {
  float4 hy;  // the four row weights of the visit (zero outside the footprint)
  float hx;   // the weight of this lane's column
  float4 v[CPL / 2 + 1];  // packed: (re, re', im, im') per coil pair; CPL = 1: v[0].xy = the sample
};

// rec: the visit's staged record (warp-uniform address), hxp = &rec.hx[column], vp: the lane's coil group of the visit
template <int CPL>
B2N_D void own_load(OwnOps<CPL> &o, const float *__restrict__ rec, const float *__restrict__ hxp,
                    const float2 *__restrict__ vp) {
  o.hy = *reinterpret_cast<const float4 *>(rec);
  o.hx = *hxp;
  if constexpr (CPL >= 2) {
#pragma unroll
    for (int p = 0; p < CPL / 2; ++p) o.v[p] = reinterpret_cast<const float4 *>(vp)[p];
  } else {
    o.v[0] = make_float4(vp[0].x, vp[0].y, 0.f, 0.f);
  }
}

template <int CPL> B2N_D void own_update(float2 (&acc)[kOTR][CPL], const OwnOps<CPL> &o) {
  const float w[kOTR] = {o.hy.x * o.hx, o.hy.y * o.hx, o.hy.z * o.hx, o.hy.w * o.hx};
#pragma unroll
  for (int k = 0; k < kOTR; ++k) {
    if constexpr (CPL >= 2) {
      const float2 ww = make_float2(w[k], w[k]);
#pragma unroll
      for (int p = 0; p < CPL / 2; ++p) {
        acc[k][2 * p] = __ffma2_rn(ww, make_float2(o.v[p].x, o.v[p].y), acc[k][2 * p]);          // real parts
        acc[k][2 * p + 1] = __ffma2_rn(ww, make_float2(o.v[p].z, o.v[p].w), acc[k][2 * p + 1]);  // imaginary parts
      }
    } else {
      acc[k][0].x = fmaf(w[k], o.v[0].x, acc[k][0].x);
      acc[k][0].y = fmaf(w[k], o.v[0].y, acc[k][0].y);
    }
  }
}

template <int CPL, int MINB, int U>
__global__ void __launch_bounds__(32, MINB) k_adj_own_2d(OwnArgs a, const float4 *__restrict__ packed, float2 *__restrict__ grid,
                                                   float2 *__restrict__ partials, unsigned *__restrict__ counters,
                                                   int slot_cap) {
  constexpr int CC = 4 * CPL, VS = own_val_stride<CPL>();
  constexpr bool PACKED = CPL >= 2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float *s_rec = reinterpret_cast<float *>(smem_raw);                 // [3][kOR][kOVF] visit records
  float2 *s_val = reinterpret_cast<float2 *>(s_rec + 3 * kOR * kOVF);  // [2][kOR][VS] samples (+ 1 spare visit)
  const int lane = threadIdx.x;
  // {traj * tiles + tile, first visit, visits | chunk index << 12, tile row << 16 | tile column}; the item array has
  // gridDim.x entries, so the count and the item are fetched together (one global-memory latency, not two)
  const int n_items = a.counts[0];
  const int4 item = a.items[blockIdx.x];
  if ((int)blockIdx.x >= n_items) return;
  const int4 tinfo = a.tiles[item.x];  // {visits, first visit, chunks, first partial slot}; used after the loop
  const int y0 = (item.w >> 16) * kOTR, x0 = (item.w & 0xffff) * kOTC;
  const int b = a.n_traj == 1 ? (int)blockIdx.z : item.x / (int)a.n_own_tiles;
  const int c0 = blockIdx.y * CC;
  const int n = item.z & 0xfff, chunk = item.z >> 12;
  const int rounds = (n + kOR - 1) / kOR;
  const float4 *vis = a.visits + (int64_t)item.y * (kOVF / 4);

  // staging roles, fixed per lane.  Records: a round is kOR * 64 contiguous bytes = two 16-byte copies per lane.
  // Samples: lane = (visit of 32 / PARTS, 16-byte part of the visit's row).
  constexpr int PARTS = 2 * CPL, VPI = 32 / PARTS;  // parts per visit, visits per instruction
  const int vi = lane / PARTS, part = lane - vi * PARTS;
  const float4 *pk_lane = packed + ((int64_t)b * a.n_chunks + blockIdx.y) * a.M * PARTS + part;

  auto issue_rec = [&](int round) {
    if (round >= rounds) return;
    float4 *dst = reinterpret_cast<float4 *>(s_rec + (round % 3) * kOR * kOVF);
    const int base = round * kOR * (kOVF / 4), last = n * (kOVF / 4) - 1;
#pragma unroll
    for (int k = 0; k < kOR * (kOVF / 4) / 32; ++k) {
      const int e = base + lane + 32 * k;
      // parts past the end repeat the last part of the list: their visits are staged (never used) from a valid sample
      cp_async16(dst + lane + 32 * k, vis + (e <= last ? e : last));
    }
  };
  auto issue_val = [&](int round) {
    if (round >= rounds) return;
    const int *rec = reinterpret_cast<const int *>(s_rec + (round % 3) * kOR * kOVF);
    float *val = reinterpret_cast<float *>(s_val + (round & 1) * kOR * VS);
#pragma unroll
    for (int k = 0; k < kOR / VPI; ++k) {
      const int i = vi + VPI * k;
      cp_async16(val + i * VS * 2 + part * 4, pk_lane + (size_t)(unsigned)rec[i * kOVF + 12] * PARTS);
    }
  };

  // the plan does not depend on the kernel that produced the samples: fetch the first visit records before the
  // programmatic-dependent-launch wait
  issue_rec(0);
  issue_rec(1);
  cp_async_commit();
  griddep_wait();
  float2 acc[kOTR][CPL];
#pragma unroll
  for (int r = 0; r < kOTR; ++r)
#pragma unroll
    for (int k = 0; k < CPL; ++k) acc[r][k] = make_float2(0.f, 0.f);
  const int xl = lane & 7, g = lane >> 3;
  if (n > 0) {
    cp_async_wait_all();
    __syncwarp();
    issue_val(0);
    cp_async_commit();
    for (int round = 0; round < rounds; ++round) {
      cp_async_wait_all();
      __syncwarp();  // this round's samples and the next round's records landed; everyone is done with round - 1
      issue_val(round + 1);
      issue_rec(round + 2);
      cp_async_commit();
      const float *rec = s_rec + (round % 3) * kOR * kOVF;
      const float *hxp = rec + 4 + xl;
      const float2 *vp = s_val + (round & 1) * kOR * VS + g * CPL;
      const int nb = min(kOR, n - round * kOR);
      // software-pipelined: the operands of visit i + 1 are fetched before the FFMA block of visit i (the fetch for
      // i + 1 = nb reads the next buffer or the spare visit and is discarded)
      OwnOps<CPL> cur;
      own_load<CPL>(cur, rec, hxp, vp);
#pragma unroll U
      for (int i = 0; i < nb; ++i) {
        rec += kOVF;
        hxp += kOVF;
        vp += VS;
        OwnOps<CPL> nxt;
        own_load<CPL>(nxt, rec, hxp, vp);
        own_update<CPL>(acc, cur);
        cur = nxt;
      }
    }
  }

  const int nch = tinfo.z;
  if (nch > 1) {
    // partial tile -> scratch slot; the last item of the tile to arrive adds the slots in chunk order
    const int64_t Bz = gridDim.z, per_slot = (int64_t)Bz * a.n_chunks;
    const int64_t sub = (int64_t)blockIdx.z * a.n_chunks + blockIdx.y;
    if (tinfo.w + nch > slot_cap) __trap();  // the caller's scratch is smaller than the plan needs: fail loudly
    float2 *mine = partials + (((int64_t)(tinfo.w + chunk)) * per_slot + sub) * (kOTR * CPL * 32);
#pragma unroll
    for (int r = 0; r < kOTR; ++r)
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
    for (int r = 0; r < kOTR; ++r)
#pragma unroll
      for (int k = 0; k < CPL; ++k) acc[r][k] = make_float2(0.f, 0.f);
#pragma unroll 2
    for (int j = 0; j < nch; ++j) {
      const float2 *src = partials + (((int64_t)(tinfo.w + j)) * per_slot + sub) * (kOTR * CPL * 32);
#pragma unroll
      for (int r = 0; r < kOTR; ++r)
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
          const float2 p = __ldcg(&src[(r * CPL + k) * 32 + lane]);
          acc[r][k].x += p.x;
          acc[r][k].y += p.y;
        }
    }
  }
  // store the tile times the per-cell phase factor q_y[row] q_x[column]: 8 lanes (columns) x 8 bytes = one 64-byte
  // segment per (row, coil)
  if (x0 + xl < a.Kx) {
    const float2 qx = a.q[a.Ky + x0 + xl];
    float2 qq[kOTR];
#pragma unroll
    for (int r = 0; r < kOTR; ++r) {
      const float2 qy = a.q[min(y0 + r, a.Ky - 1)];
      qq[r] = make_float2(qy.x * qx.x - qy.y * qx.y, qy.x * qx.y + qy.y * qx.x);
    }
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
      const int c = c0 + g * CPL + k;
      if (c < a.C) {
        float2 *dst = grid + ((int64_t)b * a.C + c) * a.Kprod + (int64_t)y0 * a.Kx + x0 + xl;
#pragma unroll
        for (int r = 0; r < kOTR; ++r) {
          float2 s;
          if constexpr (PACKED) {
            s = (k & 1) ? make_float2(acc[r][k - 1].y, acc[r][k].y) : make_float2(acc[r][k].x, acc[r][k + 1].x);
          } else {
            s = acc[r][k];
          }
          if (y0 + r < a.Ky)
            dst[(int64_t)r * a.Kx] = make_float2(s.x * qq[r].x - s.y * qq[r].y, s.x * qq[r].y + s.y * qq[r].x);
        }
      }
    }
  }
}

// ---- 2-D, one to four coils: lanes = the 32 cells of the tile ------------------------------------------------------
// With few coils the coil-group lanes of k_adj_own_2d idle.  Here a lane owns ONE cell of the 4 x 8 tile and keeps all
// CK coils of the chunk in registers; a visit costs it one weight product and CK FMAs (complex sample x real weight):
//   acc[c] += (hy[row] * hx[col]) * v[c]
// Same visit records, same partial-sum scheme; the sample pre-pass writes rows of CK coils.
template <int CK>
__global__ void __launch_bounds__(256) k_own_pack_small(const float2 *__restrict__ kdata, float2 *__restrict__ packed, int C,
                                                        int64_t M, int n_chunks, int n_traj,
                                                        const int32_t *__restrict__ inv_perm, const float2 *__restrict__ fac) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int chunk = blockIdx.y, b = blockIdx.z;
  if (m >= M) return;
  const float2 f = fac[inv_perm[(n_traj == 1 ? 0 : (int64_t)b * M) + m]];
  griddep_wait();
  float2 v[CK];
#pragma unroll
  for (int k = 0; k < CK; ++k) {
    const int c = chunk * CK + k;
    float2 x = make_float2(0.f, 0.f);
    if (c < C) x = __ldg(&kdata[((int64_t)b * C + c) * M + m]);
    v[k] = make_float2(x.x * f.x - x.y * f.y, x.x * f.y + x.y * f.x);
  }
  float2 *out = packed + (((int64_t)b * n_chunks + chunk) * M + m) * CK;
  if constexpr (CK == 1) {
    out[0] = v[0];
  } else {
#pragma unroll
    for (int p = 0; p < CK / 2; ++p)  // planar per coil pair: (re, re', im, im')
      reinterpret_cast<float4 *>(out)[p] = make_float4(v[2 * p].x, v[2 * p + 1].x, v[2 * p].y, v[2 * p + 1].y);
  }
}

template <int CK> constexpr size_t own_cell_smem_bytes() {
  return sizeof(float) * 3 * kOR * kOVF + sizeof(float2) * (2 * kOR + 1) * CK;
}

template <int CK, int U>
__global__ void __launch_bounds__(32, 32) k_adj_own_2d_cell(OwnArgs a, const float2 *__restrict__ packed, float2 *__restrict__ grid,
                                                      float2 *__restrict__ partials, unsigned *__restrict__ counters,
                                                      int slot_cap) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float *s_rec = reinterpret_cast<float *>(smem_raw);                 // [3][kOR][kOVF] visit records
  float2 *s_val = reinterpret_cast<float2 *>(s_rec + 3 * kOR * kOVF);  // [2][kOR][CK] samples (+ 1 spare visit)
  const int lane = threadIdx.x;
  const int n_items = a.counts[0];
  const int4 item = a.items[blockIdx.x];
  if ((int)blockIdx.x >= n_items) return;
  const int4 tinfo = a.tiles[item.x];
  const int y0 = (item.w >> 16) * kOTR, x0 = (item.w & 0xffff) * kOTC;
  const int b = a.n_traj == 1 ? (int)blockIdx.z : item.x / (int)a.n_own_tiles;
  const int c0 = blockIdx.y * CK;
  const int n = item.z & 0xfff, chunk = item.z >> 12;
  const int rounds = (n + kOR - 1) / kOR;
  const float4 *vis = a.visits + (int64_t)item.y * (kOVF / 4);
  // sample staging: 8-byte (CK = 1) or 16-byte copies; lane = (visit, part)
  constexpr int PARTS = CK == 1 ? 1 : CK / 2;
  const int vi = lane / PARTS, part = lane - vi * PARTS;
  const float2 *pk = packed + ((int64_t)b * a.n_chunks + blockIdx.y) * a.M * CK;

  auto issue_rec = [&](int round) {
    if (round >= rounds) return;
    float4 *dst = reinterpret_cast<float4 *>(s_rec + (round % 3) * kOR * kOVF);
    const int base = round * kOR * (kOVF / 4), last = n * (kOVF / 4) - 1;
#pragma unroll
    for (int k = 0; k < kOR * (kOVF / 4) / 32; ++k) {
      const int e = base + lane + 32 * k;
      cp_async16(dst + lane + 32 * k, vis + (e <= last ? e : last));
    }
  };
  auto issue_val = [&](int round) {
    if (round >= rounds) return;
    const int *rec = reinterpret_cast<const int *>(s_rec + (round % 3) * kOR * kOVF);
    float2 *val = s_val + (round & 1) * kOR * CK;
    if (vi < kOR) {
      const float2 *src = pk + (size_t)(unsigned)rec[vi * kOVF + 12] * CK;
      if constexpr (CK == 1) cp_async8(val + vi, src, true);
      else cp_async16(val + vi * CK + part * 2, src + part * 2);
    }
  };

  issue_rec(0);
  issue_rec(1);
  cp_async_commit();
  griddep_wait();
  float2 acc[CK];
#pragma unroll
  for (int k = 0; k < CK; ++k) acc[k] = make_float2(0.f, 0.f);
  const int xl = lane & 7, yl = lane >> 3;
  if (n > 0) {
    cp_async_wait_all();
    __syncwarp();
    issue_val(0);
    cp_async_commit();
    for (int round = 0; round < rounds; ++round) {
      cp_async_wait_all();
      __syncwarp();
      issue_val(round + 1);
      issue_rec(round + 2);
      cp_async_commit();
      const float *rec = s_rec + (round % 3) * kOR * kOVF;
      const float2 *vp = s_val + (round & 1) * kOR * CK;
      const int nb = min(kOR, n - round * kOR);
#pragma unroll U
      for (int i = 0; i < nb; ++i) {
        const float w = rec[i * kOVF + yl] * rec[i * kOVF + 4 + xl];
        if constexpr (CK == 1) {
          const float2 v = vp[i];
          acc[0].x = fmaf(w, v.x, acc[0].x);
          acc[0].y = fmaf(w, v.y, acc[0].y);
        } else {
          const float2 ww = make_float2(w, w);
#pragma unroll
          for (int p = 0; p < CK / 2; ++p) {
            const float4 v = reinterpret_cast<const float4 *>(vp + i * CK)[p];
            acc[2 * p] = __ffma2_rn(ww, make_float2(v.x, v.y), acc[2 * p]);
            acc[2 * p + 1] = __ffma2_rn(ww, make_float2(v.z, v.w), acc[2 * p + 1]);
          }
        }
      }
    }
  }
  const int nch = tinfo.z;
  if (nch > 1) {
    const int64_t Bz = gridDim.z, per_slot = (int64_t)Bz * a.n_chunks;
    const int64_t sub = (int64_t)blockIdx.z * a.n_chunks + blockIdx.y;
    if (tinfo.w + nch > slot_cap) __trap();
    float2 *mine = partials + (((int64_t)(tinfo.w + chunk)) * per_slot + sub) * (CK * 32);
#pragma unroll
    for (int k = 0; k < CK; ++k) __stcg(&mine[k * 32 + lane], acc[k]);
    __threadfence();
    __syncwarp();
    unsigned old = 0;
    unsigned *ctr = counters + (int64_t)item.x * per_slot + sub;
    if (lane == 0) old = atomicAdd(ctr, 1u);
    old = __shfl_sync(0xffffffffu, old, 0);
    if (old != (unsigned)(nch - 1)) return;
    if (lane == 0) *ctr = 0u;
    __threadfence();
#pragma unroll
    for (int k = 0; k < CK; ++k) acc[k] = make_float2(0.f, 0.f);
    for (int j = 0; j < nch; ++j) {
      const float2 *src = partials + (((int64_t)(tinfo.w + j)) * per_slot + sub) * (CK * 32);
#pragma unroll
      for (int k = 0; k < CK; ++k) {
        const float2 p = __ldcg(&src[k * 32 + lane]);
        acc[k].x += p.x;
        acc[k].y += p.y;
      }
    }
  }
  if (y0 + yl < a.Ky && x0 + xl < a.Kx) {
    const float2 qy = a.q[y0 + yl], qx = a.q[a.Ky + x0 + xl];
    const float2 qq = make_float2(qy.x * qx.x - qy.y * qx.y, qy.x * qx.y + qy.y * qx.x);
#pragma unroll
    for (int k = 0; k < CK; ++k) {
      const int c = c0 + k;
      if (c < a.C) {
        float2 s2;
        if constexpr (CK == 1) s2 = acc[0];
        else s2 = (k & 1) ? make_float2(acc[k - 1].y, acc[k].y) : make_float2(acc[k].x, acc[k + 1].x);
        grid[((int64_t)b * a.C + c) * a.Kprod + (int64_t)(y0 + yl) * a.Kx + x0 + xl] =
            make_float2(s2.x * qq.x - s2.y * qq.y, s2.x * qq.y + s2.y * qq.x);
      }
    }
  }
}

// ---- 3-D: 4 x 4 x 8 output tiles ------------------------------------------------------------------------------------
// Same scheme one dimension up (216 neighbours per point).  Lanes = 8 cells along the last (contiguous) axis x 4 along
// the middle axis; registers = 4 cells along the first axis x CK coils, ALL coils of the chunk in every lane (8 at
// BASELINE config 4), planar over coil pairs.  Per visit a lane forms its weight h1[y] h2[x] once, then
//   acc[z][c] += (h0[z] * h1[y] * h2[x]) * v[c]         5 FMUL + 4 CK FFMA2 (32 at 8 coils)
// with the samples read as warp-uniform 16-byte loads.  A point is visited by 8.2 tiles on average, so a 64-byte record
// per visit as in 2-D would be gigabytes: the visit list holds 16-byte index records and the window weights are formed
// while staging -- copies of 4 bytes with a zero source size outside the footprint, lanes = 2 visits x 16 weight slots
// so that one instruction touches two points' records (two cache lines) only.
constexpr int kO3Z = 4, kO3Y = 4, kO3X = 8;  // tile edges: axis 0 (registers), axis 1 (lane groups), axis 2 (lanes)
constexpr int kO3HW = 24;                    // floats per point record: h0[6], h1[6], h2[6], -h2[6]
constexpr int kO3W = 16;                     // staged weight slots per visit: h0[4], h1[4], h2[8]

struct Own3Args {
  int K[3], nt[3], C;
  int64_t M, Kprod, n_own_tiles;
  int n_traj, n_chunks;
  const int4 *visits, *items, *tiles;
  const int32_t *counts;
  const float *hw;
  const float2 *q;  // per-cell phase factors: [K0], [K1], [K2]
};

template <int CK> constexpr size_t own3_smem_bytes() {
  return sizeof(int4) * 3 * kOR + sizeof(float) * (2 * kOR + 1) * kO3W + sizeof(float4) * (2 * kOR + 1) * (CK / 2);
}

template <int CK> struct Own3Ops {
  float4 h0;
  float h1, h2;
  float4 v[CK / 2];  // (re, re', im, im') per coil pair
};

template <int CK>
B2N_D void own3_load(Own3Ops<CK> &o, const float *__restrict__ w, const float *__restrict__ h1p,
                     const float *__restrict__ h2p, const float4 *__restrict__ vp) {
  o.h0 = *reinterpret_cast<const float4 *>(w);
  o.h1 = *h1p;
  o.h2 = *h2p;
#pragma unroll
  for (int p = 0; p < CK / 2; ++p) o.v[p] = vp[p];
}

template <int CK> B2N_D void own3_update(float2 (&acc)[kO3Z][CK], const Own3Ops<CK> &o) {
  const float wl = o.h1 * o.h2;
  const float w[kO3Z] = {o.h0.x * wl, o.h0.y * wl, o.h0.z * wl, o.h0.w * wl};
#pragma unroll
  for (int z = 0; z < kO3Z; ++z) {
    const float2 ww = make_float2(w[z], w[z]);
#pragma unroll
    for (int p = 0; p < CK / 2; ++p) {
      acc[z][2 * p] = __ffma2_rn(ww, make_float2(o.v[p].x, o.v[p].y), acc[z][2 * p]);
      acc[z][2 * p + 1] = __ffma2_rn(ww, make_float2(o.v[p].z, o.v[p].w), acc[z][2 * p + 1]);
    }
  }
}

template <int CK, int MINB, int U>
__global__ void __launch_bounds__(32, MINB) k_adj_own_3d(Own3Args a, const float4 *__restrict__ packed, float2 *__restrict__ grid,
                                                   float2 *__restrict__ partials, unsigned *__restrict__ counters,
                                                   int slot_cap) {
  constexpr int PARTS = CK / 2;  // 16-byte parts per sample row
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int4 *s_idx = reinterpret_cast<int4 *>(smem_raw);                       // [3][kOR] index records
  float *s_w = reinterpret_cast<float *>(s_idx + 3 * kOR);                // [2][kOR][kO3W] window weights (+ 1 spare)
  float4 *s_val = reinterpret_cast<float4 *>(s_w + (2 * kOR + 1) * kO3W);  // [2][kOR][PARTS] samples (+ 1 spare)
  const int lane = threadIdx.x;
  const int n_items = a.counts[0];
  const int4 item = a.items[blockIdx.x];
  if ((int)blockIdx.x >= n_items) return;
  const int4 tinfo = a.tiles[item.x];
  const int t2 = item.w % a.nt[2], t01 = item.w / a.nt[2], t1 = t01 % a.nt[1], t0 = t01 / a.nt[1];
  const int o0 = t0 * kO3Z, o1 = t1 * kO3Y, o2 = t2 * kO3X;
  const int b = a.n_traj == 1 ? (int)blockIdx.z : item.x / (int)a.n_own_tiles;
  const int c0 = blockIdx.y * CK;
  const int n = item.z & 0xfff, chunk = item.z >> 12;
  const int rounds = (n + kOR - 1) / kOR;
  const int4 *vis = a.visits + item.y;

  // staging roles, fixed per lane.  Weights: lane = (visit of 2, slot of 16); slot class 0 / 1 / 2 = axis.
  const int wv = lane >> 4, slot = lane & 15;
  const int cls = slot < 4 ? 0 : (slot < 8 ? 1 : 2);
  const int local = slot - (cls == 2 ? 8 : 4 * cls);  // cell of the tile along the slot's axis
  // Samples: lane = (visit of 32 / PARTS, 16-byte part).
  constexpr int VPI = 32 / PARTS;
  const int vi = lane / PARTS, part = lane - vi * PARTS;
  const float4 *pk_lane = packed + ((int64_t)b * a.n_chunks + blockIdx.y) * a.M * PARTS + part;

  auto issue_idx = [&](int round) {
    if (round < rounds && lane < kOR) {
      const int i = round * kOR + lane;
      cp_async16(&s_idx[(round % 3) * kOR + lane], &vis[i < n ? i : n - 1]);  // past the end: repeat the last visit
    }
  };
  auto issue_data = [&](int round) {
    if (round >= rounds) return;
    const int4 *idx = s_idx + (round % 3) * kOR;
    float *w = s_w + (round & 1) * kOR * kO3W;
    float4 *val = s_val + (round & 1) * kOR * PARTS;
#pragma unroll
    for (int k = 0; k < kOR / 2; ++k) {
      const int i = 2 * k + wv;
      const int4 e = idx[i];
      const int r = ((e.z >> (8 * cls)) & 0xff) - 16;  // footprint origin along this axis, relative to the tile
      const int j = local - r;
      const bool on = (unsigned)j < (unsigned)kOJ;
      // the record's fourth block is -h2: visits whose footprint wrapped around the grid with a sign flip read it
      const float *src = a.hw + (int64_t)e.x * kO3HW + 6 * cls + (cls == 2 && e.w ? 6 : 0) + (on ? j : 0);
      cp_async4z(w + i * kO3W + slot, src, on);
    }
#pragma unroll
    for (int k = 0; k < kOR / VPI; ++k) {
      const int i = vi + VPI * k;
      cp_async16(val + i * PARTS + part, pk_lane + (size_t)(unsigned)idx[i].y * PARTS);
    }
  };

  issue_idx(0);
  issue_idx(1);
  cp_async_commit();
  griddep_wait();
  float2 acc[kO3Z][CK];
#pragma unroll
  for (int z = 0; z < kO3Z; ++z)
#pragma unroll
    for (int k = 0; k < CK; ++k) acc[z][k] = make_float2(0.f, 0.f);
  const int xl = lane & 7, yl = lane >> 3;
  if (n > 0) {
    cp_async_wait_all();
    __syncwarp();
    issue_data(0);
    cp_async_commit();
    for (int round = 0; round < rounds; ++round) {
      cp_async_wait_all();
      __syncwarp();
      issue_data(round + 1);
      issue_idx(round + 2);
      cp_async_commit();
      const float *w = s_w + (round & 1) * kOR * kO3W;
      const float *h1p = w + 4 + yl, *h2p = w + 8 + xl;
      const float4 *vp = s_val + (round & 1) * kOR * PARTS;
      const int nb = min(kOR, n - round * kOR);
      Own3Ops<CK> cur;
      own3_load<CK>(cur, w, h1p, h2p, vp);
#pragma unroll U
      for (int i = 0; i < nb; ++i) {
        w += kO3W;
        h1p += kO3W;
        h2p += kO3W;
        vp += PARTS;
        Own3Ops<CK> nxt;
        own3_load<CK>(nxt, w, h1p, h2p, vp);
        own3_update<CK>(acc, cur);
        cur = nxt;
      }
    }
  }

  const int nch = tinfo.z;
  if (nch > 1) {
    const int64_t Bz = gridDim.z, per_slot = (int64_t)Bz * a.n_chunks;
    const int64_t sub = (int64_t)blockIdx.z * a.n_chunks + blockIdx.y;
    if (tinfo.w + nch > slot_cap) __trap();
    float2 *mine = partials + (((int64_t)(tinfo.w + chunk)) * per_slot + sub) * (kO3Z * CK * 32);
#pragma unroll
    for (int z = 0; z < kO3Z; ++z)
#pragma unroll
      for (int k = 0; k < CK; ++k) __stcg(&mine[(z * CK + k) * 32 + lane], acc[z][k]);
    __threadfence();
    __syncwarp();
    unsigned old = 0;
    unsigned *ctr = counters + (int64_t)item.x * per_slot + sub;
    if (lane == 0) old = atomicAdd(ctr, 1u);
    old = __shfl_sync(0xffffffffu, old, 0);
    if (old != (unsigned)(nch - 1)) return;
    if (lane == 0) *ctr = 0u;
    __threadfence();
#pragma unroll
    for (int z = 0; z < kO3Z; ++z)
#pragma unroll
      for (int k = 0; k < CK; ++k) acc[z][k] = make_float2(0.f, 0.f);
    for (int j = 0; j < nch; ++j) {
      const float2 *src = partials + (((int64_t)(tinfo.w + j)) * per_slot + sub) * (kO3Z * CK * 32);
#pragma unroll
      for (int z = 0; z < kO3Z; ++z)
#pragma unroll
        for (int k = 0; k < CK; ++k) {
          const float2 p = __ldcg(&src[(z * CK + k) * 32 + lane]);
          acc[z][k].x += p.x;
          acc[z][k].y += p.y;
        }
    }
  }
  // store the tile times the per-cell phase factor: 8 lanes x 8 bytes = one 64-byte segment per (z, y, coil)
  if (o1 + yl < a.K[1] && o2 + xl < a.K[2]) {
    const float2 q1 = a.q[a.K[0] + o1 + yl], q2 = a.q[a.K[0] + a.K[1] + o2 + xl];
    const float2 q12 = make_float2(q1.x * q2.x - q1.y * q2.y, q1.x * q2.y + q1.y * q2.x);
#pragma unroll
    for (int z = 0; z < kO3Z; ++z) {
      if (o0 + z >= a.K[0]) break;
      const float2 q0 = a.q[o0 + z];
      const float2 qq = make_float2(q0.x * q12.x - q0.y * q12.y, q0.x * q12.y + q0.y * q12.x);
      const int64_t cell = ((int64_t)(o0 + z) * a.K[1] + o1 + yl) * a.K[2] + o2 + xl;
#pragma unroll
      for (int k = 0; k < CK; ++k) {
        const int c = c0 + k;
        if (c < a.C) {
          const float2 s2 = (k & 1) ? make_float2(acc[z][k - 1].y, acc[z][k].y) : make_float2(acc[z][k].x, acc[z][k + 1].x);
          grid[((int64_t)b * a.C + c) * a.Kprod + cell] = make_float2(s2.x * qq.x - s2.y * qq.y, s2.x * qq.y + s2.y * qq.x);
        }
      }
    }
  }
}

// ---- exception points ---------------------------------------------------------------------------------------------
// Points whose neighbours do not share one table offset (b2n_points.own_exc: rounding ties next to the k-space
// origin) carry zero weights in the visit lists.  They are spread here with their complex records, output-stationary
// like the main kernel: a CTA owns an output tile, a thread one of its cells, and adds the contributions of the
// tile's exception points (own_xt / own_xv, ascending slot order) to that cell once -- parallel over tiles and
// deterministic.  Launched only while the plan's exception count is not known to be zero.
struct OwnFixArgs {
  int K[3], nt[3], C, n_traj;
  int64_t M, Kprod, n_own_tiles, xcap;
  const int32_t *counts, *perm;
  const int2 *xt, *xv;
  const float2 *coef;
};

// Output tiles looked at by one CTA (most of them have no exception points).  2-D: 32 (config 2: 9 us).  3-D: the
// tiles with exceptions cluster around the k-space origin, where one CTA then works through its whole group tile after
// tile (1.7 ms at config 4 with 32 tiles per CTA; 4 keep the serial part short and the grid
// a quarter of the tile count: 0.65 ms; one tile per CTA, what ships: 0.5 ms).
template <int ND> constexpr int fix_tiles() { return ND == 3 ? 1 : 32; }

template <int ND>
__global__ void __launch_bounds__(128) k_own_fix(OwnFixArgs a, const float2 *__restrict__ kdata, float2 *__restrict__ grid) {
  griddep_wait();
  if (a.counts[3] > a.xcap) __trap();  // more (exception point, tile) pairs than own_xv holds: fail loudly
  const int64_t n_tiles_all = a.n_own_tiles * a.n_traj;
  constexpr int kFixTiles = fix_tiles<ND>();
  const int64_t t_mine = (int64_t)blockIdx.x * kFixTiles + (threadIdx.x & 31);
  const int2 seg_mine = (int)(threadIdx.x & 31) < kFixTiles && t_mine < n_tiles_all ? a.xt[t_mine] : make_int2(0, 0);
  unsigned todo = __ballot_sync(0xffffffffu, seg_mine.y > 0);  // the same in every warp of the CTA
  const int t = threadIdx.x;
  int cl[3] = {0, 0, 0};
  if (ND == 3) {
    cl[0] = t >> 5;
    cl[1] = (t >> 3) & 3;
    cl[2] = t & 7;
  } else {
    cl[0] = t >> 3;
    cl[1] = t & 7;
  }
  while (todo) {
    const int src = __ffs(todo) - 1;
    todo &= todo - 1;
    const int2 seg = make_int2(__shfl_sync(0xffffffffu, seg_mine.x, src), __shfl_sync(0xffffffffu, seg_mine.y, src));
    const int64_t tile_all = (int64_t)blockIdx.x * kFixTiles + src;
    const int64_t traj = tile_all / a.n_own_tiles;
    int64_t tid = tile_all - traj * a.n_own_tiles;
    int o[3] = {0, 0, 0};
    for (int d = ND - 1; d >= 0; --d) {
      o[d] = (int)(tid % a.nt[d]) * (d == ND - 1 ? 8 : 4);
      tid /= a.nt[d];
    }
    int64_t cell = 0;
    bool inside = ND == 3 || t < 32;
    for (int d = 0; d < ND; ++d) {
      inside = inside && o[d] + cl[d] < a.K[d];
      cell = cell * a.K[d] + o[d] + cl[d];
    }
    if (!inside) continue;
    const int b = a.n_traj == 1 ? (int)blockIdx.z : (int)traj;
    for (int c = 0; c < a.C; ++c) {
      float2 acc = make_float2(0.f, 0.f);
      for (int e = 0; e < seg.y; ++e) {
        const int2 xv = a.xv[seg.x + e];
        const float2 *rec = a.coef + (int64_t)xv.x * ND * kOJ;
        float2 w = make_float2(1.f, 0.f);
        bool on = true;
#pragma unroll
        for (int d = 0; d < ND; ++d) {
          const int j = cl[d] - (((xv.y >> (8 * d)) & 0xff) - 16);
          on = on && (unsigned)j < (unsigned)kOJ;
          const float2 cw = rec[d * kOJ + (on ? j : 0)];
          w = make_float2(w.x * cw.x - w.y * cw.y, w.x * cw.y + w.y * cw.x);
        }
        if (on) cmacf_conj(acc, w, kdata[((int64_t)b * a.C + c) * a.M + a.perm[xv.x]]);
      }
      float2 *dst = grid + ((int64_t)b * a.C + c) * a.Kprod + cell;
      float2 gv = *dst;
      gv.x += acc.x;
      gv.y += acc.y;
      *dst = gv;
    }
  }
}

// ---- host side ------------------------------------------------------------------------------
int g_adj_owned = 1;

static bool own_ready(const b2n_geom *g, const b2n_points *p, int layout) {
  if (!g_adj_owned || g->dtype != B2N_C64 || (g->ndim != 2 && g->ndim != 3) || layout != B2N_COIL_MAJOR) return false;
  for (int d = 0; d < g->ndim; ++d)
    if (g->numpoints[d] != kOJ) return false;
  return p->own_tile == kOTR && p->own_visits && p->own_items && p->own_tiles && p->own_counts && p->own_fac &&
         p->own_q && p->own_exc && p->own_xt && p->own_xv && p->own_hw && p->n_points > 0;
}

// coils per warp: 2-D 4 CPL (lanes hold coil groups), 3-D CK (every lane holds all of them)
static int own_chunk_coils(int ndim, int64_t C) {
  if (ndim == 3) return C > 4 ? 8 : 4;
  return C > 8 ? 16 : (C > 4 ? 8 : (C > 2 ? 4 : (int)C));  // <= 4 coils: the lane-per-cell kernel, 1 / 2 / 4 coils
}

// scratch = [arrival counters][pre-packed samples][partial tiles]; the partial tiles are sized by the plan's upper
// bound unless the caller passes the partial-slot count it read back from own_counts[1]
struct OwnScratch {
  size_t ctr, packed, slot, total;  // bytes: counters, packed samples, one partial-tile slot (all its sub-tiles)
};
static OwnScratch own_scratch_layout(const b2n_points *p, int64_t B, int64_t C, int64_t n_slots) {
  OwnScratch o;
  const int cc = own_chunk_coils(p->ndim, C);
  const int64_t n_chunks = ceil_div(C, cc), Bz = p->n_traj == 1 ? B : 1;
  const int64_t n_tiles_all = (int64_t)p->n_own_tiles[0] * p->n_own_tiles[1] * p->n_own_tiles[2] * p->n_traj;
  o.ctr = align_up(sizeof(unsigned) * (size_t)(n_tiles_all * Bz * n_chunks), 256);
  o.packed = align_up(sizeof(float2) * (size_t)(B * n_chunks * cc) * (size_t)p->n_points, 256);
  // accumulators of one warp: 2-D 4 rows x CPL coils x 32 lanes, 3-D 4 planes x CK coils x 32 lanes
  o.slot = sizeof(float2) * (size_t)(Bz * n_chunks) * (size_t)(p->ndim == 3 ? kO3Z * cc * 32 : cc * 32);
  o.total = o.ctr + o.packed + o.slot * (size_t)(n_slots > 0 ? n_slots : p->n_own_items_max);
  return o;
}

size_t own_adjoint_bytes(const b2n_geom *g, const b2n_points *p, int64_t B, int64_t C, int layout, int64_t n_slots,
                         size_t *zero_bytes) {
  if (zero_bytes) *zero_bytes = 0;
  if (!own_ready(g, p, layout)) return 0;
  const OwnScratch o = own_scratch_layout(p, B, C, n_slots);
  if (zero_bytes) *zero_bytes = o.ctr;
  return o.total;
}

struct OwnLaunch {
  const b2n_geom *g;
  const b2n_points *p;
  const void *kdata;
  int64_t B, C;
  int n_chunks;
  char *scratch;
  OwnScratch lay;
  int slot_cap;
  void *grid;
  cudaStream_t st;
};

template <int CPL, bool PLANAR> static int launch_pack(const OwnLaunch &l) {
  dim3 gp((unsigned)ceil_div(l.p->n_points, 64), (unsigned)l.n_chunks, (unsigned)l.B);
  B2N_CUDA_OK(launch_pdl(k_own_pack<CPL, PLANAR>, gp, dim3(256), 0, l.st, (const float2 *)l.kdata,
                         (float4 *)(l.scratch + l.lay.ctr), (int)l.C, l.p->n_points, l.n_chunks, (int)l.p->n_traj,
                         (const int32_t *)l.p->inv_perm, (const float2 *)l.p->own_fac));
  B2N_LAUNCH_OK("k_own_pack");
  return 0;
}

static int launch_fix(const OwnLaunch &l) {
  if (l.p->n_own_exc_max <= 0) return 0;
  OwnFixArgs f;
  f.Kprod = 1;
  f.n_own_tiles = 1;
  for (int d = 0; d < 3; ++d) {
    f.K[d] = d < l.g->ndim ? (int)l.g->grid_size[d] : 1;
    f.nt[d] = d < l.g->ndim ? l.p->n_own_tiles[d] : 1;
    f.Kprod *= f.K[d];
    f.n_own_tiles *= f.nt[d];
  }
  f.C = (int)l.C;
  f.n_traj = (int)l.p->n_traj;
  f.M = l.p->n_points;
  f.xcap = l.p->n_own_xv_max;
  f.counts = l.p->own_counts;
  f.perm = l.p->perm;
  f.xt = (const int2 *)l.p->own_xt;
  f.xv = (const int2 *)l.p->own_xv;
  f.coef = (const float2 *)l.p->coef;
  dim3 gf((unsigned)ceil_div(f.n_own_tiles * l.p->n_traj, l.g->ndim == 3 ? fix_tiles<3>() : fix_tiles<2>()), 1,
          (unsigned)(l.p->n_traj == 1 ? l.B : 1));
  if (l.g->ndim == 3)
    B2N_CUDA_OK(launch_pdl(k_own_fix<3>, gf, dim3(128), 0, l.st, f, (const float2 *)l.kdata, (float2 *)l.grid));
  else
    B2N_CUDA_OK(launch_pdl(k_own_fix<2>, gf, dim3(32), 0, l.st, f, (const float2 *)l.kdata, (float2 *)l.grid));
  B2N_LAUNCH_OK("k_own_fix");
  return 0;
}

static void own_args_2d(OwnArgs &a, const OwnLaunch &l) {
  a.Ky = (int)l.g->grid_size[0];
  a.Kx = (int)l.g->grid_size[1];
  a.C = (int)l.C;
  a.M = l.p->n_points;
  a.Kprod = l.g->grid_size[0] * l.g->grid_size[1];
  a.n_own_tiles = (int64_t)l.p->n_own_tiles[0] * l.p->n_own_tiles[1];
  a.n_traj = (int)l.p->n_traj;
  a.n_chunks = l.n_chunks;
  a.visits = (const float4 *)l.p->own_visits;
  a.items = (const int4 *)l.p->own_items;
  a.tiles = (const int4 *)l.p->own_tiles;
  a.counts = l.p->own_counts;
  a.q = (const float2 *)l.p->own_q;
}

template <int CK> static int launch_own_cell(const OwnLaunch &l) {
  OwnArgs a;
  own_args_2d(a, l);
  dim3 gp((unsigned)ceil_div(a.M, 256), (unsigned)l.n_chunks, (unsigned)l.B);
  B2N_CUDA_OK(launch_pdl(k_own_pack_small<CK>, gp, dim3(256), 0, l.st, (const float2 *)l.kdata,
                         (float2 *)(l.scratch + l.lay.ctr), (int)l.C, a.M, l.n_chunks, a.n_traj,
                         (const int32_t *)l.p->inv_perm, (const float2 *)l.p->own_fac));
  B2N_LAUNCH_OK("k_own_pack_small");
  auto kern = k_adj_own_2d_cell<CK, 4>;
  dim3 gd((unsigned)l.p->n_own_items_max, (unsigned)a.n_chunks, (unsigned)(l.p->n_traj == 1 ? l.B : 1));
  B2N_CUDA_OK(launch_pdl(kern, gd, dim3(32), own_cell_smem_bytes<CK>(), l.st, a, (const float2 *)(l.scratch + l.lay.ctr),
                         (float2 *)l.grid, (float2 *)(l.scratch + l.lay.ctr + l.lay.packed), (unsigned *)l.scratch,
                         l.slot_cap));
  B2N_LAUNCH_OK("k_adj_own_2d_cell");
  return launch_fix(l);
}

template <int CPL, int MINB, int U> static int launch_own(const OwnLaunch &l) {
  OwnArgs a;
  own_args_2d(a, l);
  if (int rc = launch_pack<CPL, (CPL >= 2)>(l)) return rc;
  auto kern = k_adj_own_2d<CPL, MINB, U>;
  dim3 gd((unsigned)l.p->n_own_items_max, (unsigned)a.n_chunks, (unsigned)(l.p->n_traj == 1 ? l.B : 1));
  B2N_CUDA_OK(launch_pdl(kern, gd, dim3(32), own_smem_bytes<CPL>(), l.st, a, (const float4 *)(l.scratch + l.lay.ctr),
                         (float2 *)l.grid, (float2 *)(l.scratch + l.lay.ctr + l.lay.packed), (unsigned *)l.scratch,
                         l.slot_cap));
  B2N_LAUNCH_OK("k_adj_own_2d");
  return launch_fix(l);
}

template <int CK, int MINB, int U> static int launch_own3(const OwnLaunch &l) {
  Own3Args a;
  a.Kprod = 1;
  a.n_own_tiles = 1;
  for (int d = 0; d < 3; ++d) {
    a.K[d] = (int)l.g->grid_size[d];
    a.nt[d] = l.p->n_own_tiles[d];
    a.Kprod *= l.g->grid_size[d];
    a.n_own_tiles *= l.p->n_own_tiles[d];
  }
  a.C = (int)l.C;
  a.M = l.p->n_points;
  a.n_traj = (int)l.p->n_traj;
  a.n_chunks = l.n_chunks;
  a.visits = (const int4 *)l.p->own_visits;
  a.items = (const int4 *)l.p->own_items;
  a.tiles = (const int4 *)l.p->own_tiles;
  a.counts = l.p->own_counts;
  a.hw = l.p->own_hw;
  a.q = (const float2 *)l.p->own_q;
  if (int rc = launch_pack<CK / 4, true>(l)) return rc;
  auto kern = k_adj_own_3d<CK, MINB, U>;
  dim3 gd((unsigned)l.p->n_own_items_max, (unsigned)a.n_chunks, (unsigned)(l.p->n_traj == 1 ? l.B : 1));
  B2N_CUDA_OK(launch_pdl(kern, gd, dim3(32), own3_smem_bytes<CK>(), l.st, a, (const float4 *)(l.scratch + l.lay.ctr),
                         (float2 *)l.grid, (float2 *)(l.scratch + l.lay.ctr + l.lay.packed), (unsigned *)l.scratch,
                         l.slot_cap));
  B2N_LAUNCH_OK("k_adj_own_3d");
  return launch_fix(l);
}

// returns 1 when the owner-tile path does not apply
int own_adjoint(const b2n_geom *g, const b2n_points *p, const void *kdata, int64_t B, int64_t C, int layout,
                void *scratch, size_t scratch_bytes, void *grid, cudaStream_t st) {
  if (!own_ready(g, p, layout)) return 1;
  if (B < 1 || C < 1) return fail_arg(B2N_E_ARG, "n_batch=%lld n_coils=%lld", (long long)B, (long long)C);
  if (p->n_traj != 1 && p->n_traj != B)
    return fail_arg(B2N_E_ARG, "plan has %lld trajectories but n_batch=%lld", (long long)p->n_traj, (long long)B);
  OwnLaunch l;
  l.lay = own_scratch_layout(p, B, C, 1);
  if (!scratch || scratch_bytes < l.lay.ctr + l.lay.packed || (reinterpret_cast<uintptr_t>(scratch) & 15))
    return fail_arg(B2N_E_ARG, "owner-tile adjoint: scratch too small or not 16-byte aligned");
  // partial-sum slots the scratch can hold; the kernel traps if the plan needs more (the caller sized the scratch
  // with a slot count that is not the plan's)
  const int64_t cap64 = (int64_t)((scratch_bytes - l.lay.ctr - l.lay.packed) / l.lay.slot);
  l.slot_cap = (int)(cap64 > 0x7fffffff ? 0x7fffffff : cap64);
  l.g = g;
  l.p = p;
  l.kdata = kdata;
  l.B = B;
  l.C = C;
  const int cc = own_chunk_coils(g->ndim, C);
  l.n_chunks = (int)ceil_div(C, cc);
  l.scratch = (char *)scratch;
  l.grid = grid;
  l.st = st;
  if (g->ndim == 3) {
    // B2N_OPT_ADJ_OWNED (A/B): 4 = 12 resident warps per SM (168 registers) instead of 16, 5 = loop not unrolled
    if (cc == 8 && g_adj_owned == 4) return launch_own3<8, 12, 2>(l);
    if (cc == 8 && g_adj_owned == 5) return launch_own3<8, 16, 1>(l);
    if (cc == 8) return launch_own3<8, 16, 2>(l);
    return launch_own3<4, 24, 2>(l);
  }
  // B2N_OPT_ADJ_OWNED (A/B of the 16-coil kernel): 4 / 5 = 20 / 32 resident warps per SM instead of 24, 6 = unroll 4
  if (cc == 16 && g_adj_owned == 4) return launch_own<4, 20, 2>(l);
  if (cc == 16 && g_adj_owned == 5) return launch_own<4, 32, 2>(l);
  if (cc == 16 && g_adj_owned == 6) return launch_own<4, 24, 4>(l);
  if (cc == 16) return launch_own<4, 24, 2>(l);
  if (cc == 8) return launch_own<2, 32, 2>(l);
  if (cc == 4) return launch_own_cell<4>(l);
  if (cc == 2) return launch_own_cell<2>(l);
  return launch_own_cell<1>(l);
}

}  // namespace b2n
