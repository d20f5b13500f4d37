// micro_r1.cu -- round-1 design microbenchmarks (B200): decides the grid layout and
// the adjoint accumulation strategy.  Standalone: nvcc -O3 -arch=sm_100a micro_r1.cu -lcufft
//  (1) cuFFT C2C: coil-major (B,C,Ky,Kx) vs channel-last (B,Ky,Kx,C) batches
//  (2) forward gather prototypes on a cfg2-like problem (golden-angle radial 200x640,
//      640^2 grid, 16 coils): thread-per-(point,coil) coil-major vs warp-per-point channel-last
//  (3) adjoint scatter with L2 reductions: scalar f32 / v2 / v4, coil-major vs channel-last
//  (4) shared-memory accumulation: atomicAdd(float) vs plain read-modify-write
#include <cuda_runtime.h>
#include <cufft.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <numeric>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)
#define CF(x) do { cufftResult r = (x); if (r != CUFFT_SUCCESS) { printf("cuFFT %d at %s:%d\n", (int)r, __FILE__, __LINE__); exit(1);} } while (0)

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; cudaEventElapsedTime(&ms, a, b); return ms; }

// ---------------------------------------------------------------- (1) cuFFT layouts
static void bench_fft(int rank, const int *n, int batch, const char *tag) {
  long long sig = 1; for (int i = 0; i < rank; ++i) sig *= n[i];
  size_t bytes = sizeof(cufftComplex) * sig * batch;
  const int NBUF = (int)std::max<size_t>(2, (300u << 20) / bytes + 1);  // rotate > L2
  std::vector<cufftComplex *> in(NBUF), out(NBUF);
  for (int i = 0; i < NBUF; ++i) { CK(cudaMalloc(&in[i], bytes)); CK(cudaMalloc(&out[i], bytes)); CK(cudaMemset(in[i], 0, bytes)); }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int layout = 0; layout < 2; ++layout) {
    cufftHandle plan;
    int istride = layout ? batch : 1, idist = layout ? 1 : (int)sig;
    CF(cufftPlanMany(&plan, rank, (int *)n, (int *)n, istride, idist, (int *)n, istride, idist, CUFFT_C2C, batch));
    for (int inplace = 0; inplace < 2; ++inplace) {
      for (int w = 0; w < 3; ++w) CF(cufftExecC2C(plan, in[0], inplace ? in[0] : out[0], CUFFT_FORWARD));
      CK(cudaDeviceSynchronize());
      const int reps = 20;
      cudaEventRecord(e0);
      for (int r = 0; r < reps; ++r) { int b = r % NBUF; CF(cufftExecC2C(plan, in[b], inplace ? in[b] : out[b], CUFFT_FORWARD)); }
      cudaEventRecord(e1); CK(cudaDeviceSynchronize());
      float ms = time_ms(e0, e1) / reps;
      printf("fft %-18s %-12s %-8s %8.2f us  %7.1f GB/s (1r+1w)\n", tag, layout ? "channel-last" : "coil-major",
             inplace ? "inplace" : "outplace", ms * 1e3, 2.0 * bytes / ms / 1e6);
    }
    cufftDestroy(plan);
  }
  for (int i = 0; i < NBUF; ++i) { cudaFree(in[i]); cudaFree(out[i]); }
}

// ---------------------------------------------------------------- synthetic plan
struct Plan {
  int M, Ky, Kx;
  int *base;      // [M][2] wrapped base cells, sorted by cell
  int *perm;      // [M]
  float2 *coef;   // [M][12]  cy[6], cx[6]
  float2 *phase;  // [M]
};

static Plan make_plan(int n_spokes, int n_read, int K) {
  int M = n_spokes * n_read;
  std::vector<int> by(M), bx(M), idx(M);
  const double phi = (1 + sqrt(5.0)) / 2;
  for (int s = 0; s < n_spokes; ++s)
    for (int r = 0; r < n_read; ++r) {
      double th = s * M_PI / phi, rad = -M_PI + 2 * M_PI * r / n_read;
      double ty = rad * sin(th) * K / (2 * M_PI), tx = rad * cos(th) * K / (2 * M_PI);
      int b0 = 1 + (int)floor(ty - 3), b1 = 1 + (int)floor(tx - 3);
      by[s * n_read + r] = ((b0 % K) + K) % K; bx[s * n_read + r] = ((b1 % K) + K) % K;
    }
  std::iota(idx.begin(), idx.end(), 0);
  std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return by[a] * K + bx[a] < by[b] * K + bx[b]; });
  std::vector<int> base(2 * M), perm(M); std::vector<float2> coef(12 * M), phase(M);
  for (int s = 0; s < M; ++s) {
    base[2 * s] = by[idx[s]]; base[2 * s + 1] = bx[idx[s]]; perm[s] = idx[s];
    for (int j = 0; j < 12; ++j) coef[12 * s + j] = make_float2(0.1f + 0.01f * j, 0.02f * j);
    phase[s] = make_float2(0.6f, 0.8f);
  }
  Plan p; p.M = M; p.Ky = p.Kx = K;
  CK(cudaMalloc(&p.base, sizeof(int) * 2 * M)); CK(cudaMalloc(&p.perm, sizeof(int) * M));
  CK(cudaMalloc(&p.coef, sizeof(float2) * 12 * M)); CK(cudaMalloc(&p.phase, sizeof(float2) * M));
  CK(cudaMemcpy(p.base, base.data(), sizeof(int) * 2 * M, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(p.perm, perm.data(), sizeof(int) * M, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(p.coef, coef.data(), sizeof(float2) * 12 * M, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(p.phase, phase.data(), sizeof(float2) * M, cudaMemcpyHostToDevice));
  return p;
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ void cmac(float2 &acc, float2 a, float2 b) { acc.x += a.x * b.x - a.y * b.y; acc.y += a.x * b.y + a.y * b.x; }
__device__ __forceinline__ int wrapi(int g, int K) { return g >= K ? g - K : g; }

// ---------------------------------------------------------------- (2) forward gathers
// A: thread per (sorted point, coil), coil-major grid [C][Ky][Kx]
__global__ void fwd_cm_thread(Plan p, int C, const float2 *__restrict__ grid, float2 *__restrict__ out) {
  int s = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y;
  if (s >= p.M) return;
  int by = p.base[2 * s], bx = p.base[2 * s + 1];
  const float2 *rec = p.coef + 12 * s;
  const float2 *g = grid + (size_t)c * p.Ky * p.Kx;
  float2 acc = make_float2(0, 0);
#pragma unroll
  for (int jy = 0; jy < 6; ++jy) {
    int gy = wrapi(by + jy, p.Ky); float2 cy = rec[jy];
#pragma unroll
    for (int jx = 0; jx < 6; ++jx) cmac(acc, cmul(cy, rec[6 + jx]), g[gy * p.Kx + wrapi(bx + jx, p.Kx)]);
  }
  out[(size_t)c * p.M + p.perm[s]] = cmul(acc, p.phase[s]);
}

// B: warp per point, channel-last grid [Ky][Kx][C], C == 16: lane = (half h, coil c)
template <int PTS_PER_WARP>
__global__ void fwd_cl_warp(Plan p, const float2 *__restrict__ grid, float2 *__restrict__ out) {
  const int C = 16;
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  int c = lane & 15, h = lane >> 4;
  for (int i = 0; i < PTS_PER_WARP; ++i) {
    int s = warp * PTS_PER_WARP + i;
    if (s >= p.M) return;
    int by = p.base[2 * s], bx = p.base[2 * s + 1];
    const float2 *rec = p.coef + 12 * s;
    float2 cx0 = rec[6 + h], cx1 = rec[8 + h], cx2 = rec[10 + h];
    int x0 = wrapi(bx + h, p.Kx), x1 = wrapi(bx + 2 + h, p.Kx), x2 = wrapi(bx + 4 + h, p.Kx);
    float2 acc = make_float2(0, 0);
#pragma unroll
    for (int jy = 0; jy < 6; ++jy) {
      const float2 *row = grid + (size_t)wrapi(by + jy, p.Ky) * p.Kx * C + c;
      float2 r = make_float2(0, 0);
      cmac(r, cx0, row[x0 * C]); cmac(r, cx1, row[x1 * C]); cmac(r, cx2, row[x2 * C]);
      cmac(acc, rec[jy], r);
    }
    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 16); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 16);
    if (h == 0) out[(size_t)c * p.M + p.perm[s]] = cmul(acc, p.phase[s]);
  }
}

// ---------------------------------------------------------------- (3) adjoint scatters
__device__ __forceinline__ void red_v2(float2 *addr, float2 v) { atomicAdd(addr, v); }
__device__ __forceinline__ void red_v4(float4 *addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// A: thread per (point, coil), coil-major, float2 reductions (what the generic kernel does)
template <bool SCALAR>
__global__ void adj_cm_thread(Plan p, int C, const float2 *__restrict__ y, float2 *__restrict__ grid) {
  int s = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y;
  if (s >= p.M) return;
  int by = p.base[2 * s], bx = p.base[2 * s + 1];
  const float2 *rec = p.coef + 12 * s;
  float2 *g = grid + (size_t)c * p.Ky * p.Kx;
  float2 val = cmul(y[(size_t)c * p.M + p.perm[s]], p.phase[s]);
#pragma unroll
  for (int jy = 0; jy < 6; ++jy) {
    int gy = wrapi(by + jy, p.Ky); float2 cy = rec[jy];
#pragma unroll
    for (int jx = 0; jx < 6; ++jx) {
      float2 w = cmul(cmul(cy, rec[6 + jx]), val);
      float2 *a = &g[gy * p.Kx + wrapi(bx + jx, p.Kx)];
      if (SCALAR) { atomicAdd(&a->x, w.x); atomicAdd(&a->y, w.y); } else red_v2(a, w);
    }
  }
}

// B: warp per point, channel-last, v2 reductions: each instruction adds 2 full 128 B lines
template <int PTS_PER_WARP>
__global__ void adj_cl_warp_v2(Plan p, const float2 *__restrict__ y, float2 *__restrict__ grid) {
  const int C = 16;
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  int c = lane & 15, h = lane >> 4;
  for (int i = 0; i < PTS_PER_WARP; ++i) {
    int s = warp * PTS_PER_WARP + i;
    if (s >= p.M) return;
    int by = p.base[2 * s], bx = p.base[2 * s + 1];
    const float2 *rec = p.coef + 12 * s;
    float2 val = cmul(y[(size_t)c * p.M + p.perm[s]], p.phase[s]);
    float2 cx0 = cmul(rec[6 + h], val), cx1 = cmul(rec[8 + h], val), cx2 = cmul(rec[10 + h], val);
    int x0 = wrapi(bx + h, p.Kx), x1 = wrapi(bx + 2 + h, p.Kx), x2 = wrapi(bx + 4 + h, p.Kx);
#pragma unroll
    for (int jy = 0; jy < 6; ++jy) {
      float2 *row = grid + (size_t)wrapi(by + jy, p.Ky) * p.Kx * C + c;
      float2 cy = rec[jy];
      red_v2(&row[x0 * C], cmul(cy, cx0)); red_v2(&row[x1 * C], cmul(cy, cx1)); red_v2(&row[x2 * C], cmul(cy, cx2));
    }
  }
}

// C: warp per point, channel-last, v4 reductions: lane = (cell slot q of 4, coil pair cp of 8)
template <int PTS_PER_WARP>
__global__ void adj_cl_warp_v4(Plan p, const float2 *__restrict__ y, float2 *__restrict__ grid) {
  const int C = 16;
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  int cp = lane & 7, q = lane >> 3;
  for (int i = 0; i < PTS_PER_WARP; ++i) {
    int s = warp * PTS_PER_WARP + i;
    if (s >= p.M) return;
    int by = p.base[2 * s], bx = p.base[2 * s + 1];
    const float2 *rec = p.coef + 12 * s;
    float2 ph = p.phase[s];
    float2 v0 = cmul(y[(size_t)(2 * cp) * p.M + p.perm[s]], ph), v1 = cmul(y[(size_t)(2 * cp + 1) * p.M + p.perm[s]], ph);
#pragma unroll
    for (int it = 0; it < 9; ++it) {
      int n = it * 4 + q, jy = n / 6, jx = n - jy * 6;
      float2 w = cmul(rec[jy], rec[6 + jx]);
      float2 a = cmul(w, v0), b = cmul(w, v1);
      float4 *dst = (float4 *)(grid + ((size_t)wrapi(by + jy, p.Ky) * p.Kx + wrapi(bx + jx, p.Kx)) * C + 2 * cp);
      red_v4(dst, make_float4(a.x, a.y, b.x, b.y));
    }
  }
}

// ---------------------------------------------------------------- (4) shared-memory accumulation
// one warp per CTA owns a (T+5)^2 x 16-coil channel-last tile; random points inside the tile
template <bool ATOMIC>
__global__ void smem_accum(int n_points, float2 *__restrict__ sink, unsigned seed) {
  const int T = 16, S = T + 5, C = 16;
  extern __shared__ float2 tile[];  // [S][S][C]
  int lane = threadIdx.x & 31, c = lane & 15, h = lane >> 4;
  for (int i = threadIdx.x; i < S * S * C; i += blockDim.x) tile[i] = make_float2(0, 0);
  __syncthreads();
  unsigned state = seed + blockIdx.x * 7919u;
  for (int pt = 0; pt < n_points; ++pt) {
    state = state * 1664525u + 1013904223u;
    int by = (state >> 8) % T, bx = (state >> 16) % T;
    float2 val = make_float2(1.0f + c, 0.5f);
#pragma unroll
    for (int jy = 0; jy < 6; ++jy)
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        float2 *a = &tile[((by + jy) * S + bx + 2 * k + h) * C + c];
        float2 w = make_float2(val.x * 0.1f * (jy + 1), val.y * 0.2f * (k + 1));
        if (ATOMIC) { atomicAdd(&a->x, w.x); atomicAdd(&a->y, w.y); }
        else { float2 o = *a; o.x += w.x; o.y += w.y; *a = o; }
      }
    __syncwarp();
  }
  __syncthreads();
  float2 acc = make_float2(0, 0);
  for (int i = threadIdx.x; i < S * S * C; i += blockDim.x) { acc.x += tile[i].x; acc.y += tile[i].y; }
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s, %d SMs, L2 %d MB\n", prop.name, prop.multiProcessorCount, prop.l2CacheSize >> 20);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);

  { int n[2] = {640, 640}; bench_fft(2, n, 16, "2D 640^2 x16"); }
  { int n[2] = {512, 512}; bench_fft(2, n, 16, "2D 512^2 x16"); }
  { int n[2] = {768, 768}; bench_fft(2, n, 32, "2D 768^2 x32"); }
  { int n[3] = {256, 256, 256}; bench_fft(3, n, 8, "3D 256^3 x8"); }
  { int n[2] = {640, 640}; bench_fft(2, n, 1, "2D 640^2 x1"); }

  const int C = 16, K = 640;
  Plan p = make_plan(200, 640, K);
  size_t gbytes = sizeof(float2) * (size_t)C * K * K, ybytes = sizeof(float2) * (size_t)C * p.M;
  float2 *grid, *y, *flush; CK(cudaMalloc(&grid, gbytes)); CK(cudaMalloc(&y, ybytes)); CK(cudaMalloc(&flush, 512u << 20));
  CK(cudaMemset(grid, 0, gbytes)); CK(cudaMemset(y, 0, ybytes));
  const int reps = 10;
  auto run = [&](const char *name, auto launch, bool zero_grid) {
    float total = 0;
    for (int r = 0; r < reps + 2; ++r) {
      CK(cudaMemsetAsync(flush, r, 512u << 20));
      if (zero_grid) CK(cudaMemsetAsync(grid, 0, gbytes));
      cudaEventRecord(e0); launch(); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
      if (r >= 2) total += time_ms(e0, e1);
    }
    CK(cudaGetLastError());
    double us = total / reps * 1e3;
    printf("%-34s %9.2f us   %7.2f G coil-pts/s   %7.1f GB/s algorithmic\n", name, us, (double)C * p.M / us / 1e3,
           (gbytes + ybytes + 8.0 * p.M) / us / 1e3);
  };
  int M = p.M;
  run("fwd coil-major thread/(pt,coil)", [&] { fwd_cm_thread<<<dim3((M + 127) / 128, C), 128>>>(p, C, grid, y); }, false);
  run("fwd channel-last warp/pt x1", [&] { fwd_cl_warp<1><<<(M * 32 + 255) / 256, 256>>>(p, grid, y); }, false);
  run("fwd channel-last warp/pt x4", [&] { fwd_cl_warp<4><<<((M + 3) / 4 * 32 + 255) / 256, 256>>>(p, grid, y); }, false);
  run("fwd channel-last warp/pt x16", [&] { fwd_cl_warp<16><<<((M + 15) / 16 * 32 + 255) / 256, 256>>>(p, grid, y); }, false);
  run("adj coil-major thread scalar f32 red", [&] { adj_cm_thread<true><<<dim3((M + 127) / 128, C), 128>>>(p, C, y, grid); }, true);
  run("adj coil-major thread v2 red", [&] { adj_cm_thread<false><<<dim3((M + 127) / 128, C), 128>>>(p, C, y, grid); }, true);
  run("adj channel-last warp v2 red x4", [&] { adj_cl_warp_v2<4><<<((M + 3) / 4 * 32 + 255) / 256, 256>>>(p, y, grid); }, true);
  run("adj channel-last warp v4 red x4", [&] { adj_cl_warp_v4<4><<<((M + 3) / 4 * 32 + 255) / 256, 256>>>(p, y, grid); }, true);
  run("memset grid only (52 MB)", [&] {}, true);

  // shared-memory accumulation: 148*4 single-warp CTAs, 2000 points each
  {
    const int S = 21, n_pts = 2000, ctas = 148 * 4;
    size_t smem = sizeof(float2) * S * S * 16;
    float2 *sink; CK(cudaMalloc(&sink, sizeof(float2) * ctas * 32));
    CK(cudaFuncSetAttribute(smem_accum<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(smem_accum<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int atomic = 0; atomic < 2; ++atomic) {
      for (int w = 0; w < 2; ++w) { if (atomic) smem_accum<true><<<ctas, 32, smem>>>(n_pts, sink, 1); else smem_accum<false><<<ctas, 32, smem>>>(n_pts, sink, 1); }
      CK(cudaDeviceSynchronize());
      cudaEventRecord(e0);
      if (atomic) smem_accum<true><<<ctas, 32, smem>>>(n_pts, sink, 1); else smem_accum<false><<<ctas, 32, smem>>>(n_pts, sink, 1);
      cudaEventRecord(e1); CK(cudaDeviceSynchronize());
      double us = time_ms(e0, e1) * 1e3;
      double updates = (double)ctas * n_pts * 36 * 16;
      printf("smem tile accumulate %-10s %9.2f us  %7.2f G complex updates/s (4 warps/SM, 1 warp per 56 KB tile)\n",
             atomic ? "atomicAdd" : "plain RMW", us, updates / us / 1e3);
    }
  }
  return 0;
}
