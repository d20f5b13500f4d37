// b2n_peer.cu -- sum all-reduce of the coil-combined image over NVLink peer memory (coil-sharded SENSE adjoint).
//
// The one collective on the NUFFT path is the sum over coils of the adjoint (reference coupling point:
// torchkbnufft/modules/kbnufft.py:404-405) when the coils are split over GPUs.  The message is small (B*N complex
// values, 0.8 MB at BASELINE config 2), so the all-reduce is bound by launch and link latencies, not by bandwidth
// (21-35 us through NCCL, profiles/r02_bench_*gpu.log).  Here every rank owns a *window* of device memory that all the
// other ranks of the node map through CUDA IPC, and ONE kernel per rank
//   1. pushes its partial image into slot [rank] of every peer's window (plain 16-byte stores over NVLink / NVSwitch),
//   2. polls its own window until the peers' values have replaced the fill pattern, and
//   3. adds the slots in RANK ORDER, so that every rank computes bit-identical sums.
// There are no flags and no fences: the windows are pre-filled with a value that never travels (negative zero, sent
// as +0), and every 4-byte element announces itself by differing from it (Lamport-style; a first version with
// per-chunk flags behind a system-scope release paid an extra link round trip: 20 vs 13 us at 2 GPUs,
// profiles/r02_peer_allreduce_2gpu.log, r02_peer_allreduce_flags_2gpu.log).  Three generations of slots rotate on a call counter that lives in the window
// (device side: the kernel is CUDA-graph capturable and needs no host state): a call reads generation e % 3, peers
// that are one call ahead already write (e + 1) % 3, and (e + 2) % 3 -- last used by call e - 1, which every rank has
// finished reading, or this call could not have been reached -- is reset to the fill pattern for call e + 2.
// Past the latency-bound regime the one-shot exchange moves (world - 1) copies of the message per rank; the TWO-SHOT form
// (k_peer_allreduce_two_shot, same windows and protocol) is a reduce-scatter followed by an all-gather: rank r receives
// everyone's values of shard r only, adds them in rank order, and pushes the reduced shard to everyone -- 2 (world - 1) /
// world copies per rank, one more link latency.  peer_allreduce_launch picks the form from the message size.
#include <string.h>

#include <type_traits>

#include "b2n_peer.cuh"

namespace b2n {

__global__ void k_peer_fill(uint32_t *p, size_t n, uint32_t v) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

// VEC = 4: n a multiple of 4 and 16-byte aligned pointers; VEC = 1: anything
template <int VEC>
__global__ void __launch_bounds__(kPeerThreads) k_peer_allreduce_sum(PeerArgs a, const float *__restrict__ in,
                                                                     float *__restrict__ out, int64_t n) {
  constexpr int PER = kPeerChunk / (kPeerThreads * VEC);  // values (float or float4) per thread and chunk
  using V = typename std::conditional<VEC == 4, float4, float>::type;
  __shared__ uint32_t s_epoch, s_prev;
  griddep_wait();  // the partial image comes from the preceding kernel; the call counter from the preceding call
  unsigned char *mine = a.window[a.rank];
  uint32_t *hdr = reinterpret_cast<uint32_t *>(mine);
  const int64_t n_chunks = (n + kPeerChunk - 1) / kPeerChunk;
  const bool single = n_chunks <= (int64_t)gridDim.x;  // one chunk per CTA: the values stay in registers
  V v[PER];
  bool on[PER];
  auto load = [&](int64_t lo) {
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      const int64_t i = lo + ((int64_t)k * kPeerThreads + threadIdx.x) * VEC;
      on[k] = i < n;
      if (on[k]) {
        v[k] = *reinterpret_cast<const V *>(in + i);
        if constexpr (VEC == 4) v[k] = make_float4(not_fill(v[k].x), not_fill(v[k].y), not_fill(v[k].z), not_fill(v[k].w));
        else v[k] = not_fill(v[k]);
      }
    }
  };
  load((int64_t)blockIdx.x * kPeerChunk);  // in flight while the call counter is read
  if (threadIdx.x == 0) {
    const uint32_t e = *reinterpret_cast<volatile uint32_t *>(hdr) + 1u;
    s_epoch = e;
    s_prev = *reinterpret_cast<volatile uint32_t *>(hdr + 4 + (e + 2u) % 3u);
  }
  __syncthreads();
  const uint32_t epoch = s_epoch, gen = epoch % 3u, gclr = (epoch + 2u) % 3u;
  const size_t gen_off = (size_t)gen * a.world * a.slot_floats;
  // 1. push: chunks blockIdx.x, blockIdx.x + gridDim.x, ...
  for (int64_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
    const int64_t lo = chunk * kPeerChunk;
    if (chunk != (int64_t)blockIdx.x) load(lo);
    for (int q = 1; q < a.world; ++q) {
      const int p = (a.rank + q) % a.world;  // start with the next rank: the peers' ingress is spread evenly
      float *dst = reinterpret_cast<float *>(a.window[p] + a.data_off) + gen_off + (size_t)a.rank * a.slot_floats + lo;
#pragma unroll
      for (int k = 0; k < PER; ++k)
        if (on[k]) *reinterpret_cast<V *>(dst + ((int64_t)k * kPeerThreads + threadIdx.x) * VEC) = v[k];
    }
  }
  // 2. reset the generation call e - 1 used (every rank is done with it) while the peers' values are in flight
  {
    const size_t prev4 = ((size_t)s_prev + 3) / 4;
    const float4 fill = make_float4(__uint_as_float(kPeerFill), __uint_as_float(kPeerFill), __uint_as_float(kPeerFill),
                                    __uint_as_float(kPeerFill));
    for (int r = 0; r < a.world; ++r) {
      if (r == a.rank) continue;
      float4 *dst = reinterpret_cast<float4 *>(reinterpret_cast<float *>(mine + a.data_off) +
                                               ((size_t)gclr * a.world + r) * a.slot_floats);
      for (size_t i = (size_t)blockIdx.x * kPeerThreads + threadIdx.x; i < prev4; i += (size_t)gridDim.x * kPeerThreads)
        dst[i] = fill;
    }
  }
  // 3. wait for the peers' values element by element and add them in rank order (identical sums on every rank)
  const float *slots = reinterpret_cast<const float *>(mine + a.data_off) + gen_off;
  unsigned spins = 0;
  unsigned long long t0 = 0;
  for (int64_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
    const int64_t lo = chunk * kPeerChunk;
    if (!single) load(lo);
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      if (!on[k]) continue;
      const int64_t i = lo + ((int64_t)k * kPeerThreads + threadIdx.x) * VEC;
      V acc;
      for (int r = 0; r < a.world; ++r) {
        V x;
        if (r == a.rank) {
          x = v[k];
        } else {
          const float *src = slots + (size_t)r * a.slot_floats + i;
          if constexpr (VEC == 4) {
            for (x = ld_volatile4(src); !arrived(x); x = ld_volatile4(src)) spin_guard(spins, t0);
          } else {
            for (x = ld_volatile1(src); !arrived(x); x = ld_volatile1(src)) spin_guard(spins, t0);
          }
        }
        if (r == 0) acc = x;
        else if constexpr (VEC == 4) acc = make_float4(acc.x + x.x, acc.y + x.y, acc.z + x.z, acc.w + x.w);
        else acc = acc + x;
      }
      *reinterpret_cast<V *>(out + i) = acc;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {  // the last CTA of the call advances the call counter and records the size of its generation
    __threadfence();
    if (atomicAdd(hdr + 1, 1u) == gridDim.x - 1) {
      hdr[1] = 0;
      hdr[4 + gen] = (uint32_t)n;
      __threadfence();
      *reinterpret_cast<volatile uint32_t *>(hdr) = epoch;
    }
  }
}

// Two-shot form (float4 only).  Shards are runs of `cps` chunks: shard r = chunks [r cps, (r + 1) cps).  Slot [q] of the
// current generation is used twice: floats [0, cps chunks) receive rank q's contribution to MY shard (phase 1), floats
// [cps chunks, 2 cps chunks) receive rank q's reduced shard (phase 2).  A CTA works on the same shard-local chunk j of
// every shard, so that its phase-2 push follows its own phase-1 reduction directly.
__global__ void __launch_bounds__(kPeerThreads) k_peer_allreduce_two_shot(PeerArgs a, const float *__restrict__ in,
                                                                          float *__restrict__ out, int64_t n, int cps) {
  constexpr int PER = kPeerChunk / (kPeerThreads * 4);
  __shared__ uint32_t s_epoch, s_prev;
  griddep_wait();
  unsigned char *mine = a.window[a.rank];
  uint32_t *hdr = reinterpret_cast<uint32_t *>(mine);
  if (threadIdx.x == 0) {
    const uint32_t e = *reinterpret_cast<volatile uint32_t *>(hdr) + 1u;
    s_epoch = e;
    s_prev = *reinterpret_cast<volatile uint32_t *>(hdr + 4 + (e + 2u) % 3u);
  }
  __syncthreads();
  const uint32_t epoch = s_epoch, gen = epoch % 3u, gclr = (epoch + 2u) % 3u;
  const size_t gen_off = (size_t)gen * a.world * a.slot_floats;
  const int64_t n_chunks = (n + kPeerChunk - 1) / kPeerChunk;
  const int64_t half = (int64_t)cps * kPeerChunk;  // floats of a slot's phase-1 part
  float *win = reinterpret_cast<float *>(mine + a.data_off);
  {  // reset the generation the call before last used (all of what it used of every slot)
    const size_t prev4 = ((size_t)s_prev + 3) / 4;
    const float4 fill = make_float4(__uint_as_float(kPeerFill), __uint_as_float(kPeerFill), __uint_as_float(kPeerFill),
                                    __uint_as_float(kPeerFill));
    for (int r = 0; r < a.world; ++r) {
      if (r == a.rank) continue;
      float4 *dst = reinterpret_cast<float4 *>(win + ((size_t)gclr * a.world + r) * a.slot_floats);
      for (size_t i = (size_t)blockIdx.x * kPeerThreads + threadIdx.x; i < prev4; i += (size_t)gridDim.x * kPeerThreads)
        dst[i] = fill;
    }
  }
  unsigned spins = 0;
  unsigned long long t0 = 0;
  for (int64_t j = blockIdx.x; j < cps; j += gridDim.x) {
    const int64_t loc = j * kPeerChunk;  // offset inside a shard / a slot half
    // phase 1: my values of every other rank's shard go to that rank
    for (int q = 1; q < a.world; ++q) {
      const int p = (a.rank + q) % a.world;
      const int64_t g = (int64_t)p * cps + j;
      if (g >= n_chunks) continue;
      float *dst = reinterpret_cast<float *>(a.window[p] + a.data_off) + gen_off + (size_t)a.rank * a.slot_floats + loc;
#pragma unroll
      for (int k = 0; k < PER; ++k) {
        const int64_t off = ((int64_t)k * kPeerThreads + threadIdx.x) * 4, i = g * kPeerChunk + off;
        if (i < n) {
          float4 v = *reinterpret_cast<const float4 *>(in + i);
          v = make_float4(not_fill(v.x), not_fill(v.y), not_fill(v.z), not_fill(v.w));
          *reinterpret_cast<float4 *>(dst + off) = v;
        }
      }
    }
    // my shard: add the arrivals in rank order, publish the result (phase 2)
    const int64_t g_own = (int64_t)a.rank * cps + j;
    if (g_own < n_chunks) {
#pragma unroll
      for (int k = 0; k < PER; ++k) {
        const int64_t off = ((int64_t)k * kPeerThreads + threadIdx.x) * 4, i = g_own * kPeerChunk + off;
        if (i >= n) continue;
        const float4 own = *reinterpret_cast<const float4 *>(in + i);
        float4 acc;
        for (int r = 0; r < a.world; ++r) {
          float4 x = own;
          if (r != a.rank) {
            const float *src = win + gen_off + (size_t)r * a.slot_floats + loc + off;
            for (x = ld_volatile4(src); !arrived(x); x = ld_volatile4(src)) spin_guard(spins, t0);
          }
          acc = r == 0 ? x : make_float4(acc.x + x.x, acc.y + x.y, acc.z + x.z, acc.w + x.w);
        }
        const float4 pub = make_float4(not_fill(acc.x), not_fill(acc.y), not_fill(acc.z), not_fill(acc.w));
        for (int q = 1; q < a.world; ++q) {
          const int p = (a.rank + q) % a.world;
          float *dst = reinterpret_cast<float *>(a.window[p] + a.data_off) + gen_off + (size_t)a.rank * a.slot_floats + half + loc;
          *reinterpret_cast<float4 *>(dst + off) = pub;
        }
        *reinterpret_cast<float4 *>(out + i) = pub;
      }
    }
    // the other ranks' reduced shards
    for (int q = 1; q < a.world; ++q) {
      const int r = (a.rank + q) % a.world;
      const int64_t g = (int64_t)r * cps + j;
      if (g >= n_chunks) continue;
      const float *src = win + gen_off + (size_t)r * a.slot_floats + half + loc;
#pragma unroll
      for (int k = 0; k < PER; ++k) {
        const int64_t off = ((int64_t)k * kPeerThreads + threadIdx.x) * 4, i = g * kPeerChunk + off;
        if (i >= n) continue;
        float4 x;
        for (x = ld_volatile4(src + off); !arrived(x); x = ld_volatile4(src + off)) spin_guard(spins, t0);
        *reinterpret_cast<float4 *>(out + i) = x;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(hdr + 1, 1u) == gridDim.x - 1) {
      hdr[1] = 0;
      hdr[4 + gen] = (uint32_t)(2 * half);  // what this call used of every slot
      __threadfence();
      *reinterpret_cast<volatile uint32_t *>(hdr) = epoch;
    }
  }
}

}  // namespace b2n

using namespace b2n;

extern "C" int b2n_peer_window_bytes(int world, int64_t max_floats, size_t *bytes) {
  if (!bytes || world < 1 || world > B2N_PEER_MAX_RANKS || max_floats < 1 || max_floats >= ((int64_t)1 << 31))
    return fail_arg(B2N_E_ARG, "peer window: world %d (1..%d), max_floats %lld (1..2^31-1)", world, B2N_PEER_MAX_RANKS,
                    (long long)max_floats);
  *bytes = peer_layout(world, max_floats).bytes;
  return 0;
}

extern "C" int b2n_peer_window_create(size_t bytes, void **window_dev, void *handle_out) {
  if (!window_dev || !handle_out || bytes < (size_t)kPeerHeader) return fail_arg(B2N_E_ARG, "peer window: NULL output or no bytes");
  static_assert(sizeof(cudaIpcMemHandle_t) == B2N_PEER_HANDLE_BYTES, "handle size");
  void *p = nullptr;
  B2N_CUDA_OK(cudaMalloc(&p, bytes));
  int rc = check_cuda(cudaMemset(p, 0, kPeerHeader), "cudaMemset(peer window)");
  if (rc == 0) {  // every slot starts out as "nothing has arrived"
    k_peer_fill<<<592, 256>>>(reinterpret_cast<uint32_t *>(static_cast<unsigned char *>(p) + kPeerHeader),
                              (bytes - kPeerHeader) / sizeof(uint32_t), kPeerFill);
    rc = check_cuda(cudaGetLastError(), "k_peer_fill");
  }
  if (rc == 0) rc = check_cuda(cudaDeviceSynchronize(), "cudaDeviceSynchronize(peer window)");
  cudaIpcMemHandle_t h;
  if (rc == 0) rc = check_cuda(cudaIpcGetMemHandle(&h, p), "cudaIpcGetMemHandle");
  if (rc != 0) {
    cudaFree(p);
    return rc;
  }
  memcpy(handle_out, &h, sizeof(h));
  *window_dev = p;
  return 0;
}

extern "C" int b2n_peer_window_open(const void *handle, void **window_dev) {
  if (!handle || !window_dev) return fail_arg(B2N_E_ARG, "peer window: NULL handle");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  B2N_CUDA_OK(cudaIpcOpenMemHandle(window_dev, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

extern "C" int b2n_peer_window_close(void *window_dev) {
  if (!window_dev) return 0;
  B2N_CUDA_OK(cudaIpcCloseMemHandle(window_dev));
  return 0;
}

extern "C" int b2n_peer_window_destroy(void *window_dev) {
  if (!window_dev) return 0;
  B2N_CUDA_OK(cudaFree(window_dev));
  return 0;
}

namespace b2n {

int g_peer_form = 0;  // B2N_OPT_PEER_FORM: 0 = by message size, 1 = one-shot, 2 = two-shot

int peer_args_from_comm(const b2n_peer_comm *comm, PeerArgs *out) {
  if (!comm) return fail_arg(B2N_E_ARG, "peer all-reduce: NULL comm");
  if (comm->world < 1 || comm->world > B2N_PEER_MAX_RANKS || comm->rank < 0 || comm->rank >= comm->world)
    return fail_arg(B2N_E_ARG, "peer all-reduce: rank %d of %d", comm->rank, comm->world);
  if (comm->max_floats < 1 || comm->max_floats >= ((int64_t)1 << 31))
    return fail_arg(B2N_E_ARG, "peer all-reduce: max_floats %lld", (long long)comm->max_floats);
  const PeerLayout l = peer_layout(comm->world, comm->max_floats);
  PeerArgs a;
  a.rank = comm->rank;
  a.world = comm->world;
  a.slot_floats = l.slot_floats;
  a.data_off = l.data_off;
  for (int r = 0; r < B2N_PEER_MAX_RANKS; ++r) {
    a.window[r] = r < comm->world ? static_cast<unsigned char *>(comm->window[r]) : nullptr;
    if (r < comm->world && !a.window[r]) return fail_arg(B2N_E_ARG, "peer all-reduce: window of rank %d is NULL", r);
  }
  *out = a;
  return 0;
}

int peer_allreduce_launch(const b2n_peer_comm *comm, const void *in_dev, void *out_dev, int64_t n_floats, cudaStream_t st) {
  if (!in_dev || !out_dev) return fail_arg(B2N_E_ARG, "peer all-reduce: NULL in/out");
  PeerArgs a;
  const int rc = peer_args_from_comm(comm, &a);
  if (rc) return rc;
  if (n_floats < 1 || n_floats > comm->max_floats)
    return fail_arg(B2N_E_RANGE, "peer all-reduce: %lld floats, window sized for %lld", (long long)n_floats, (long long)comm->max_floats);
  const int64_t chunks = ceil_div(n_floats, kPeerChunk);
  int dev = 0, sms = 0;
  B2N_CUDA_OK(cudaGetDevice(&dev));
  B2N_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const bool vec = !(n_floats & 3) && !(reinterpret_cast<uintptr_t>(in_dev) & 15) && !(reinterpret_cast<uintptr_t>(out_dev) & 15);
  // one-shot moves (S - 1) n, two-shot 2 (S - 1) / S n and pays one more link latency (~6 us): two-shot from the size
  // where the saved bytes outweigh it (never for two ranks); g_peer_form forces a form for tests and A/B runs
  const int64_t S = a.world;
  const double saved_bytes = 4.0 * (double)n_floats * (double)(S - 1) * (double)(S - 2) / (double)S;
  const bool two_shot = vec && S >= 2 && (g_peer_form == 2 || (g_peer_form == 0 && S >= 3 && saved_bytes > 6.0e6));
  if (two_shot) {
    const int64_t cps = ceil_div(chunks, S);
    if (2 * cps * kPeerChunk > a.slot_floats) return fail_arg(B2N_E_RANGE, "peer all-reduce: window too small for the two-shot form");
    const dim3 grid2((unsigned)(cps < 4 * (int64_t)sms ? cps : 4 * (int64_t)sms));
    B2N_CUDA_OK(launch_pdl(k_peer_allreduce_two_shot, grid2, dim3(kPeerThreads), 0, st, a, static_cast<const float *>(in_dev),
                           static_cast<float *>(out_dev), n_floats, (int)cps));
    B2N_LAUNCH_OK("k_peer_allreduce_two_shot");
    return 0;
  }
  const dim3 grid((unsigned)(chunks < 4 * (int64_t)sms ? chunks : 4 * (int64_t)sms));  // <= 4 CTAs per SM: all resident
  // launched with programmatic stream serialization like the FFT passes: the CTAs are resident (and have read their
  // kernel arguments) while the last inverse pass drains, and start pushing the moment it has completed
  if (vec)
    B2N_CUDA_OK(launch_pdl(k_peer_allreduce_sum<4>, grid, dim3(kPeerThreads), 0, st, a, static_cast<const float *>(in_dev),
                           static_cast<float *>(out_dev), n_floats));
  else
    B2N_CUDA_OK(launch_pdl(k_peer_allreduce_sum<1>, grid, dim3(kPeerThreads), 0, st, a, static_cast<const float *>(in_dev),
                           static_cast<float *>(out_dev), n_floats));
  B2N_LAUNCH_OK("k_peer_allreduce_sum");
  return 0;
}

}  // namespace b2n

extern "C" int b2n_peer_allreduce_sum(const b2n_peer_comm *comm, const void *in_dev, void *out_dev, int64_t n_floats,
                                      void *stream) {
  return peer_allreduce_launch(comm, in_dev, out_dev, n_floats, static_cast<cudaStream_t>(stream));
}
