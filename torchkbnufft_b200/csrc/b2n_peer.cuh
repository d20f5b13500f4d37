// b2n_peer.cuh -- the slot protocol of the peer-memory all-reduce (see b2n_peer.cu), shared by the stand-alone kernel
// and by the last inverse FFT pass when it pushes its rows straight into the peers' windows (b2n_fft_fast_kernels.cuh).
#pragma once
#include "b2n_common.cuh"

namespace b2n {

constexpr int kPeerThreads = 256;
constexpr int kPeerChunk = 4096;          // floats per CTA and round: 256 threads x 4 x float4
constexpr int kPeerHeader = 256;          // bytes: {calls completed, CTAs of the running call that are done, -, -, floats of generation 0 / 1 / 2}
constexpr uint32_t kPeerFill = 0x80000000u;  // -0.0f

struct PeerLayout {
  int64_t slot_floats;
  size_t data_off, bytes;
};

static inline PeerLayout peer_layout(int world, int64_t max_floats) {
  PeerLayout l;
  l.slot_floats = (ceil_div(max_floats, kPeerChunk) + 1) * kPeerChunk;  // + 1: the two-shot form rounds its shards up
  l.data_off = kPeerHeader;
  l.bytes = l.data_off + sizeof(float) * 3 * (size_t)world * l.slot_floats;
  return l;
}

struct PeerArgs {
  int rank, world;
  int64_t slot_floats;
  size_t data_off;
  unsigned char *window[B2N_PEER_MAX_RANKS];
};

#ifdef __CUDACC__
B2N_D float not_fill(float x) { return __float_as_uint(x) == kPeerFill ? 0.f : x; }
B2N_D float4 ld_volatile4(const float *p) {
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
B2N_D float ld_volatile1(const float *p) {
  float v;
  asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
B2N_D bool arrived(float4 v) {
  return __float_as_uint(v.x) != kPeerFill && __float_as_uint(v.y) != kPeerFill && __float_as_uint(v.z) != kPeerFill &&
         __float_as_uint(v.w) != kPeerFill;
}
B2N_D bool arrived(float v) { return __float_as_uint(v) != kPeerFill; }
// a peer that never issues the matching call is a usage error: trap after ~20 s instead of hanging the device
B2N_D void spin_guard(unsigned &spins, unsigned long long &t0) {
  if ((++spins & 0xFFFFu) != 0) return;
  unsigned long long now;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
  if (!t0) t0 = now;
  else if (now - t0 > 20000000000ull) __trap();
}
#endif


#ifdef __CUDACC__
B2N_D float2 ld_volatile2(const float *p) {
  float2 v;
  asm volatile("ld.volatile.global.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p) : "memory");
  return v;
}
B2N_D bool arrived(float2 v) { return __float_as_uint(v.x) != kPeerFill && __float_as_uint(v.y) != kPeerFill; }

// One complex value of the coil-combined image through the all-reduce: push to every peer, reset the slot the call
// before last used, wait for the peers' values, add in rank order.  `idx` = complex index in the flattened image.
B2N_D float2 peer_exchange(const PeerArgs &a, uint32_t epoch, int64_t idx, float2 val, unsigned &spins,
                           unsigned long long &t0) {
  const uint32_t gen = epoch % 3u, gclr = (epoch + 2u) % 3u;
  val = make_float2(not_fill(val.x), not_fill(val.y));
  for (int q = 1; q < a.world; ++q) {
    const int p = (a.rank + q) % a.world;
    float *dst = reinterpret_cast<float *>(a.window[p] + a.data_off) + ((size_t)gen * a.world + a.rank) * a.slot_floats;
    *reinterpret_cast<float2 *>(dst + 2 * idx) = val;
  }
  float *mine = reinterpret_cast<float *>(a.window[a.rank] + a.data_off);
  const float2 fill = make_float2(__uint_as_float(kPeerFill), __uint_as_float(kPeerFill));
  for (int r = 0; r < a.world; ++r)
    if (r != a.rank) *reinterpret_cast<float2 *>(mine + ((size_t)gclr * a.world + r) * a.slot_floats + 2 * idx) = fill;
  float2 acc = make_float2(0.f, 0.f);
  for (int r = 0; r < a.world; ++r) {
    float2 x = val;
    if (r != a.rank) {
      const float *src = mine + ((size_t)gen * a.world + r) * a.slot_floats + 2 * idx;
      for (x = ld_volatile2(src); !arrived(x); x = ld_volatile2(src)) spin_guard(spins, t0);
    }
    acc = r == 0 ? x : make_float2(acc.x + x.x, acc.y + x.y);
  }
  return acc;
}
#endif

// host: stand-alone all-reduce launch (b2n_peer.cu) and the kernel-argument form of a communicator
int peer_allreduce_launch(const b2n_peer_comm *comm, const void *in_dev, void *out_dev, int64_t n_floats, cudaStream_t st);
int peer_args_from_comm(const b2n_peer_comm *comm, PeerArgs *out);

}  // namespace b2n
