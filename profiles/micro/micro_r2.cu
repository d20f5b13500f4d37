// micro_r2.cu -- round-2 design questions for the owner-tile spread, answered on the B200 itself:
//  (1) issue rate of FFMA (3 registers) vs FFMA2 (fma.rn.f32x2 with a scalar broadcast operand) per SM sub-partition;
//  (2) cost of a warp-uniform (broadcast) LDS.64 / LDS.128 and of a 4-address LDS.128 on the shared-memory data pipe;
//  (3) how both scale with the number of resident warps per SM (4 .. 32).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o micro_r2 micro_r2.cu && ./micro_r2
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 4096

__global__ void k_ffma(float *out, float a, float b) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 0.001f + i;
  a += threadIdx.x * 1e-9f;  // register operands, not constant-bank ones
  b += threadIdx.x * 1e-9f;
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_ffma2(float *out, float a, float b) {
  float2 x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = make_float2(threadIdx.x * 0.001f + i, i);
  a += threadIdx.x * 1e-9f;
  b += threadIdx.x * 1e-9f;
  const float2 aa = make_float2(a, a), bb = make_float2(b, b + 1.f);
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = __ffma2_rn(x[i], aa, bb);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i].x + x[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// MODE 0: uniform LDS.64, 1: uniform LDS.128, 2: LDS.128 with 4 distinct addresses (8 lanes each), 3: LDS.64 with 8
// distinct consecutive addresses (4 lanes each), 4: conflict-free LDS.128 (32 distinct addresses)
template <int MODE> __global__ void k_lds(float *out, int stride) {
  extern __shared__ float4 sm[];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = make_float4(i, 1, 2, 3);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  int idx = MODE == 2 ? (lane >> 3) : (MODE == 3 ? (lane & 7) : (MODE == 4 ? lane : 0));
  float s = 0.f;
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (MODE == 0 || MODE == 3) {
        float2 v;
        const unsigned addr = (unsigned)__cvta_generic_to_shared(reinterpret_cast<const float2 *>(sm) + idx + u * 64);
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
        s += v.x + v.y;
      } else {
        float4 v;
        const unsigned addr = (unsigned)__cvta_generic_to_shared(sm + idx + u * 32);
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
        s += (v.x + v.y) + (v.z + v.w);
      }
    }
    idx = (idx + stride) & 127;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F> float time_ms(F f) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  f();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms;
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double ghz = khz * 1e-6;
  printf("%s, %d SMs, %.3f GHz (attribute; cycles below assume it)\n", prop.name, sms, ghz);
  float *out;
  cudaMalloc(&out, sizeof(float) * sms * 1024);
  for (int warps : {4, 8, 16, 32}) {
    const int threads = warps * 32;
    float ms = time_ms([&] { k_ffma<<<sms, threads>>>(out, 1.0001f, 0.5f); });
    const double inst = (double)ITER * 16 * warps;  // warp instructions per SM
    printf("warps/SM %2d  FFMA : %.3f ms  %.2f cycles per warp-instruction per SM sub-partition\n", warps, ms,
           ms * 1e-3 * ghz * 1e9 / (inst / 4));
    ms = time_ms([&] { k_ffma2<<<sms, threads>>>(out, 1.0001f, 0.5f); });
    printf("warps/SM %2d  FFMA2: %.3f ms  %.2f cycles per warp-instruction per SM sub-partition\n", warps, ms,
           ms * 1e-3 * ghz * 1e9 / (inst / 4));
  }
  for (int warps : {4, 8, 16, 32}) {
    const int threads = warps * 32;
    const double inst = (double)ITER * 8 * warps;
    float ms = time_ms([&] { k_lds<0><<<sms, threads, 16384>>>(out, 1); });
    printf("warps/SM %2d  LDS.64 uniform      : %.2f cycles per warp-instruction per SM\n", warps, ms * 1e-3 * ghz * 1e9 / inst);
    ms = time_ms([&] { k_lds<1><<<sms, threads, 16384>>>(out, 1); });
    printf("warps/SM %2d  LDS.128 uniform     : %.2f cycles per warp-instruction per SM\n", warps, ms * 1e-3 * ghz * 1e9 / inst);
    ms = time_ms([&] { k_lds<2><<<sms, threads, 16384>>>(out, 4); });
    printf("warps/SM %2d  LDS.128 4 addresses : %.2f cycles per warp-instruction per SM\n", warps, ms * 1e-3 * ghz * 1e9 / inst);
    ms = time_ms([&] { k_lds<3><<<sms, threads, 16384>>>(out, 8); });
    printf("warps/SM %2d  LDS.64 8 addresses  : %.2f cycles per warp-instruction per SM\n", warps, ms * 1e-3 * ghz * 1e9 / inst);
    ms = time_ms([&] { k_lds<4><<<sms, threads, 16384>>>(out, 32); });
    printf("warps/SM %2d  LDS.128 32 addresses: %.2f cycles per warp-instruction per SM\n", warps, ms * 1e-3 * ghz * 1e9 / inst);
  }
  return 0;
}
