// b2n_peer.cu -- sum all-reduce of the coil-combined image over NVLink peer memory (coil-sharded SENSE adjoint).
//
// The one collective on the NUFFT path is the sum over coils of the adjoint (reference coupling point:
// torchkbnufft/modules/kbnufft.py:404-405) when the coils are split over GPUs.  The message is small (B*N complex
// values, 0.8 MB at BASELINE config 2), so a ring / tree all-reduce is bound by its launch and hop latencies
// (21-35 us through NCCL, profiles/r02_bench_*gpu.log).  Here every rank owns a *window* of device memory that all the
// other ranks of the node map through CUDA IPC; ONE kernel per rank
//   1. pushes its partial image into slot [rank] of every peer's window (plain 16-byte stores over NVLink / NVSwitch),
//   2. publishes one flag per (peer, chunk) with release semantics at system scope,
//   3. waits for the flags of the same chunk from every peer (acquire, system scope) and
//   4. adds the slots in RANK ORDER, so that every rank computes bit-identical sums.
// Chunks (4096 floats, one CTA each) are independent: a CTA starts adding as soon as its own chunk has arrived from
// everyone, the rest of the image is still in flight.  Data slots are double-buffered on the parity of a call counter
// that lives in the window (device side: the kernel is CUDA-graph capturable and needs no host state); a rank can be
// at most one call ahead of its peers because it needs their flags of call k to finish call k, and those are only
// written after the peer has finished reading call k-1.
#include <string.h>

#include <type_traits>

#include "b2n_common.cuh"

namespace b2n {

constexpr int kPeerThreads = 256;
constexpr int kPeerChunk = 4096;   // floats per CTA: 256 threads x 4 x float4
constexpr int kPeerHeader = 128;   // bytes: {calls completed, CTAs of the running call that are done}

struct PeerLayout {
  int64_t n_chunks_max, slot_floats;
  size_t flags_off, data_off, bytes;
};

static PeerLayout peer_layout(int world, int64_t max_floats) {
  PeerLayout l;
  l.n_chunks_max = ceil_div(max_floats, kPeerChunk);
  l.slot_floats = l.n_chunks_max * kPeerChunk;
  l.flags_off = kPeerHeader;
  l.data_off = align_up(l.flags_off + sizeof(uint32_t) * (size_t)world * l.n_chunks_max, 256);
  l.bytes = l.data_off + sizeof(float) * 2 * (size_t)world * l.slot_floats;
  return l;
}

struct PeerArgs {
  int rank, world;
  int64_t n_chunks_max, slot_floats;
  size_t flags_off, data_off;
  unsigned char *window[B2N_PEER_MAX_RANKS];
};

B2N_D uint32_t ld_acquire_sys(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
B2N_D void st_release_sys(uint32_t *p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// VEC = 4: n a multiple of 4 and 16-byte aligned pointers; VEC = 1: anything
template <int VEC>
__global__ void __launch_bounds__(kPeerThreads) k_peer_allreduce_sum(PeerArgs a, const float *__restrict__ in,
                                                                     float *__restrict__ out, int64_t n) {
  constexpr int PER = kPeerChunk / (kPeerThreads * VEC);  // values (float or float4) per thread
  using V = typename std::conditional<VEC == 4, float4, float>::type;
  __shared__ uint32_t s_epoch;
  griddep_wait();  // the partial image comes from the preceding kernel; the call counter from the preceding call
  unsigned char *mine = a.window[a.rank];
  uint32_t *hdr = reinterpret_cast<uint32_t *>(mine);
  if (threadIdx.x == 0) s_epoch = *reinterpret_cast<volatile uint32_t *>(hdr) + 1u;
  __syncthreads();
  const uint32_t epoch = s_epoch;
  const int64_t n_chunks = (n + kPeerChunk - 1) / kPeerChunk;
  // chunks blockIdx.x, blockIdx.x + gridDim.x, ... in the same order on every rank; the grid is small enough to be
  // resident as a whole, so a CTA waiting for a peer never keeps that peer's partner CTA off its GPU
  for (int64_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
  const int64_t lo = chunk * kPeerChunk;
  const size_t slot = ((size_t)(epoch & 1u) * a.world + a.rank) * a.slot_floats + lo;  // where this rank's chunk goes
  V v[PER];
  bool on[PER];
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    const int64_t i = lo + ((int64_t)k * kPeerThreads + threadIdx.x) * VEC;
    on[k] = i < n;
    if (on[k]) v[k] = *reinterpret_cast<const V *>(in + i);
  }
  for (int p = 0; p < a.world; ++p) {
    if (p == a.rank) continue;
    float *dst = reinterpret_cast<float *>(a.window[p] + a.data_off) + slot;
#pragma unroll
    for (int k = 0; k < PER; ++k)
      if (on[k]) *reinterpret_cast<V *>(dst + ((int64_t)k * kPeerThreads + threadIdx.x) * VEC) = v[k];
  }
  __threadfence_system();
  __syncthreads();  // every thread's stores are ordered before the flags
  if ((int)threadIdx.x < a.world && (int)threadIdx.x != a.rank)
    st_release_sys(reinterpret_cast<uint32_t *>(a.window[threadIdx.x] + a.flags_off) + (size_t)a.rank * a.n_chunks_max + chunk,
                   epoch);
  if ((int)threadIdx.x < a.world && (int)threadIdx.x != a.rank) {
    const uint32_t *flag = reinterpret_cast<const uint32_t *>(mine + a.flags_off) + (size_t)threadIdx.x * a.n_chunks_max + chunk;
    // a peer that never issues the matching call is a usage error: trap after ~20 s instead of hanging the device
    unsigned long long t0 = 0;
    for (unsigned spins = 0; (int32_t)(ld_acquire_sys(flag) - epoch) < 0; ++spins) {
      __nanosleep(20);
      if ((spins & 0xFFFFu) == 0xFFFFu) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (!t0) t0 = now;
        else if (now - t0 > 20000000000ull) __trap();
      }
    }
  }
  __syncthreads();
  const float *slots = reinterpret_cast<const float *>(mine + a.data_off) + (size_t)(epoch & 1u) * a.world * a.slot_floats + lo;
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    if (!on[k]) continue;
    const int64_t off = ((int64_t)k * kPeerThreads + threadIdx.x) * VEC;
    V acc;
    for (int r = 0; r < a.world; ++r) {  // rank order on every rank: identical sums everywhere
      const V x = r == a.rank ? v[k] : __ldcg(reinterpret_cast<const V *>(slots + (size_t)r * a.slot_floats + off));
      if (r == 0) acc = x;
      else if constexpr (VEC == 4) acc = make_float4(acc.x + x.x, acc.y + x.y, acc.z + x.z, acc.w + x.w);
      else acc = acc + x;
    }
    *reinterpret_cast<V *>(out + lo + off) = acc;
  }
  }
  __syncthreads();
  if (threadIdx.x == 0) {  // the last CTA of the call advances the call counter
    if (atomicAdd(hdr + 1, 1u) == gridDim.x - 1) {
      hdr[1] = 0;
      __threadfence();
      *reinterpret_cast<volatile uint32_t *>(hdr) = epoch;
    }
  }
}

}  // namespace b2n

using namespace b2n;

extern "C" int b2n_peer_window_bytes(int world, int64_t max_floats, size_t *bytes) {
  if (!bytes || world < 1 || world > B2N_PEER_MAX_RANKS || max_floats < 1)
    return fail_arg(B2N_E_ARG, "peer window: world %d (1..%d), max_floats %lld", world, B2N_PEER_MAX_RANKS, (long long)max_floats);
  *bytes = peer_layout(world, max_floats).bytes;
  return 0;
}

extern "C" int b2n_peer_window_create(size_t bytes, void **window_dev, void *handle_out) {
  if (!window_dev || !handle_out || bytes < (size_t)kPeerHeader) return fail_arg(B2N_E_ARG, "peer window: NULL output or no bytes");
  static_assert(sizeof(cudaIpcMemHandle_t) == B2N_PEER_HANDLE_BYTES, "handle size");
  void *p = nullptr;
  B2N_CUDA_OK(cudaMalloc(&p, bytes));
  int rc = check_cuda(cudaMemset(p, 0, bytes), "cudaMemset(peer window)");
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

extern "C" int b2n_peer_allreduce_sum(const b2n_peer_comm *comm, const void *in_dev, void *out_dev, int64_t n_floats,
                                      void *stream) {
  if (!comm || !in_dev || !out_dev) return fail_arg(B2N_E_ARG, "peer all-reduce: NULL comm/in/out");
  if (comm->world < 1 || comm->world > B2N_PEER_MAX_RANKS || comm->rank < 0 || comm->rank >= comm->world)
    return fail_arg(B2N_E_ARG, "peer all-reduce: rank %d of %d", comm->rank, comm->world);
  if (n_floats < 1 || n_floats > comm->max_floats)
    return fail_arg(B2N_E_RANGE, "peer all-reduce: %lld floats, window sized for %lld", (long long)n_floats, (long long)comm->max_floats);
  const PeerLayout l = peer_layout(comm->world, comm->max_floats);
  PeerArgs a;
  a.rank = comm->rank;
  a.world = comm->world;
  a.n_chunks_max = l.n_chunks_max;
  a.slot_floats = l.slot_floats;
  a.flags_off = l.flags_off;
  a.data_off = l.data_off;
  for (int r = 0; r < B2N_PEER_MAX_RANKS; ++r) {
    a.window[r] = r < comm->world ? static_cast<unsigned char *>(comm->window[r]) : nullptr;
    if (r < comm->world && !a.window[r]) return fail_arg(B2N_E_ARG, "peer all-reduce: window of rank %d is NULL", r);
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t chunks = ceil_div(n_floats, kPeerChunk);
  int dev = 0, sms = 0;
  B2N_CUDA_OK(cudaGetDevice(&dev));
  B2N_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const dim3 grid((unsigned)(chunks < 4 * (int64_t)sms ? chunks : 4 * (int64_t)sms));  // <= 4 CTAs per SM: all resident
  const bool vec = !(n_floats & 3) && !(reinterpret_cast<uintptr_t>(in_dev) & 15) && !(reinterpret_cast<uintptr_t>(out_dev) & 15);
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
