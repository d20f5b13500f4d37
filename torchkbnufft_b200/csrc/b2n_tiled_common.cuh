// b2n_tiled_common.cuh -- device helpers shared by the 2-D and 3-D tiled interpolation kernels:
// complex FMA forms, cp.async (LDGSTS), mbarrier, TMA tensor loads / reduce-adds, per-CTA tracing.
#pragma once
#include <cuda.h>

#include "b2n_common.cuh"
#include "b2n_interp.cuh"

namespace b2n {

// ---- small device helpers -------------------------------------------------------------
B2N_D void cmacf(float2 &acc, float2 a, float2 b) {  // acc += a * b, 4 FFMA
  acc.x = fmaf(a.x, b.x, acc.x);
  acc.x = fmaf(-a.y, b.y, acc.x);
  acc.y = fmaf(a.x, b.y, acc.y);
  acc.y = fmaf(a.y, b.x, acc.y);
}
B2N_D void cmacf_conj(float2 &acc, float2 a, float2 b) {  // acc += conj(a) * b
  acc.x = fmaf(a.x, b.x, acc.x);
  acc.x = fmaf(a.y, b.y, acc.x);
  acc.y = fmaf(a.x, b.y, acc.y);
  acc.y = fmaf(-a.y, b.x, acc.y);
}

B2N_D unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

B2N_D void cp_async4(void *dst, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
B2N_D void cp_async4z(void *dst, const void *src, bool valid) {
  const int src_size = valid ? 4 : 0;  // 0 -> destination bytes are zero-filled
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(smem_u32(dst)), "l"(src), "r"(src_size) : "memory");
}
B2N_D void cp_async8(void *dst, const void *src, bool valid) {
  const int src_size = valid ? 8 : 0;  // 0 -> destination bytes are zero-filled
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(smem_u32(dst)), "l"(src), "r"(src_size) : "memory");
}
B2N_D void cp_async16(void *dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
// 16-byte copy, zero-filled when !valid (src must still be a mapped address)
B2N_D void cp_async16z(void *dst, const void *src, bool valid) {
  const int src_size = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(dst)), "l"(src), "r"(src_size) : "memory");
}
B2N_D void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
B2N_D void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

B2N_D void mbar_init(uint64_t *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
B2N_D void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
B2N_D void mbar_wait(uint64_t *bar, unsigned parity) {
  unsigned done = 0;
  for (unsigned spins = 0; !done; ++spins) {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (spins > (1u << 22)) __trap();  // a TMA that never lands must not hang the device
  }
}
// TMA: 4-D box (x in floats, y, coil, batch) global -> shared, completion on an mbarrier
B2N_D void tma_load_4d(void *dst, const CUtensorMap *map, int x, int y, int c, int b, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];\n" ::
          "r"(smem_u32(dst)),
      "l"(map), "r"(x), "r"(y), "r"(c), "r"(b), "r"(smem_u32(bar))
      : "memory");
}
// TMA: shared -> global element-wise FP32 add of a 4-D box (out-of-range parts are dropped)
B2N_D void tma_reduce_add_4d(const CUtensorMap *map, int x, int y, int c, int b, const void *src) {
  asm volatile(
      "cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2, %3, %4}], [%5];\n" ::"l"(map),
      "r"(x), "r"(y), "r"(c), "r"(b), "r"(smem_u32(src))
      : "memory");
}
B2N_D void tma_store_commit_wait() {
  asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
}
B2N_D void fence_async_proxy() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

B2N_D long long gtime() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
B2N_D int smid() {
  int v;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(v));
  return v;
}
// per-CTA timeline record (development aid, see b2n_set_trace_buffer)
B2N_D void trace_write(const InterpArgs<float> &a, int points, long long t0, long long t1, long long t2, int flags) {
  if (a.trace && threadIdx.x == 0) {
    const int64_t id = ((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    if (id < a.trace_cap) {
      long long *r = a.trace + id * 6;
      r[0] = smid(); r[1] = points; r[2] = t0; r[3] = t1; r[4] = t2; r[5] = flags;
    }
  }
}


// TMA: 5-D box (x in floats, y, z, coil, batch) global -> shared, completion on an mbarrier
B2N_D void tma_load_5d(void *dst, const CUtensorMap *map, int x, int y, int z, int c, int b, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];\n" ::
          "r"(smem_u32(dst)),
      "l"(map), "r"(x), "r"(y), "r"(z), "r"(c), "r"(b), "r"(smem_u32(bar))
      : "memory");
}
B2N_D void tma_reduce_add_5d(const CUtensorMap *map, int x, int y, int z, int c, int b, const void *src) {
  asm volatile(
      "cp.reduce.async.bulk.tensor.5d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2, %3, %4, %5}], [%6];\n" ::"l"(map),
      "r"(x), "r"(y), "r"(z), "r"(c), "r"(b), "r"(smem_u32(src))
      : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn tensor_map_encoder();  // cuTensorMapEncodeTiled through cudaGetDriverEntryPoint (no libcuda link)

}  // namespace b2n
