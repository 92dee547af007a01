// Micro-benchmark of the MLP inner loop variants (not part of the product):
// a chain of 64x64 Dense+LeakyReLU layers on warp-private row tiles, weights in shared
// memory.  Reports warp-level FMA throughput as a fraction of 128 FMA/clk/SM, and checks
// every variant against a scalar reference.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/mlp_core_bench tools/mlp_core_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); exit(1);} } while (0)

constexpr int K = 64, N = 64, LAYERS = 4;
typedef unsigned long long u64;

__device__ __forceinline__ float leaky(float v) { return v > 0.f ? v : 0.2f * v; }
__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ void ffma2(u64& d, u64 a, u64 b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b)); }
__device__ __forceinline__ int swz(int k) { return (k >> 2) & 7; }

// ---------------------------------------------------------------- variant A ---
// 8 rows x 8 cols per thread, scalar FFMA, 32-row warp tile (the round-1 kernel).
struct VarA {
  static constexpr int ROWS = 32;
  static __device__ void layer(float* act, const float* w, const float* bias, int lane) {
    const int rg = lane >> 3, cg = lane & 7;
    float acc[8][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { float b = bias[(j < 4) ? cg * 4 + j : 32 + cg * 4 + j - 4];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i][j] = b; }
    float a[8], b[8];
    auto lda = [&](const float* p, int ch, float (&o)[8]) { float4 lo = *(const float4*)(p + ch), hi = *(const float4*)(p + (ch ^ 16));
      o[0]=lo.x;o[1]=lo.y;o[2]=lo.z;o[3]=lo.w;o[4]=hi.x;o[5]=hi.y;o[6]=hi.z;o[7]=hi.w; };
    auto ldb = [&](const float* p, float (&o)[8]) { float4 lo = *(const float4*)(p + cg * 4), hi = *(const float4*)(p + 32 + cg * 4);
      o[0]=lo.x;o[1]=lo.y;o[2]=lo.z;o[3]=lo.w;o[4]=hi.x;o[5]=hi.y;o[6]=hi.z;o[7]=hi.w; };
    lda(act, rg << 2, a); ldb(w, b);
#pragma unroll 1
    for (int k0 = 0; k0 < K; k0 += 4) {
      const int c0 = (rg ^ swz(k0)) << 2, c1 = (rg ^ swz(k0 + 4)) << 2;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const int k = k0 + kk; float an[8], bn[8];
        lda(act + (k + 1) * 32, kk == 3 ? c1 : c0, an); ldb(w + (k + 1) * N, bn);
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
#pragma unroll
        for (int i = 0; i < 8; ++i) { a[i] = an[i]; b[i] = bn[i]; }
      }
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = (j < 4) ? cg * 4 + j : 32 + cg * 4 + j - 4; const int ch = (rg ^ swz(c)) << 2;
      *(float4*)(act + c * 32 + ch) = make_float4(leaky(acc[0][j]), leaky(acc[1][j]), leaky(acc[2][j]), leaky(acc[3][j]));
      *(float4*)(act + c * 32 + (ch ^ 16)) = make_float4(leaky(acc[4][j]), leaky(acc[5][j]), leaky(acc[6][j]), leaky(acc[7][j]));
    }
    __syncwarp();
  }
  static __device__ int idx(int k, int row) { return k * 32 + ((((row >> 2) ^ swz(k)) << 2) | (row & 3)); }
};

// 8x8 scalar FFMA with U k-steps per loop iteration (U % 4 == 0, K % U == 0).
template <int U>
struct VarU {
  static constexpr int ROWS = 32;
  static __device__ int idx(int k, int row) { return VarA::idx(k, row); }
  static __device__ void layer(float* act, const float* w, const float* bias, int lane) {
    const int rg = lane >> 3, cg = lane & 7;
    float acc[8][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { float b = bias[(j < 4) ? cg * 4 + j : 32 + cg * 4 + j - 4];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i][j] = b; }
    float a[8], b[8];
    auto lda = [&](const float* p, int ch, float (&o)[8]) { float4 lo = *(const float4*)(p + ch), hi = *(const float4*)(p + (ch ^ 16));
      o[0]=lo.x;o[1]=lo.y;o[2]=lo.z;o[3]=lo.w;o[4]=hi.x;o[5]=hi.y;o[6]=hi.z;o[7]=hi.w; };
    auto ldb = [&](const float* p, float (&o)[8]) { float4 lo = *(const float4*)(p + cg * 4), hi = *(const float4*)(p + 32 + cg * 4);
      o[0]=lo.x;o[1]=lo.y;o[2]=lo.z;o[3]=lo.w;o[4]=hi.x;o[5]=hi.y;o[6]=hi.z;o[7]=hi.w; };
    lda(act, rg << 2, a); ldb(w, b);
#pragma unroll 1
    for (int k0 = 0; k0 < K; k0 += U) {
      const int base = k0 >> 2;
#pragma unroll
      for (int kk = 0; kk < U; ++kk) {
        const int k = k0 + kk; float an[8], bn[8];
        const int ch = (rg ^ ((base + ((kk + 1) >> 2)) & 7)) << 2;
        lda(act + (k + 1) * 32, ch, an); ldb(w + (k + 1) * N, bn);
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
#pragma unroll
        for (int i = 0; i < 8; ++i) { a[i] = an[i]; b[i] = bn[i]; }
      }
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = (j < 4) ? cg * 4 + j : 32 + cg * 4 + j - 4; const int ch = (rg ^ swz(c)) << 2;
      *(float4*)(act + c * 32 + ch) = make_float4(leaky(acc[0][j]), leaky(acc[1][j]), leaky(acc[2][j]), leaky(acc[3][j]));
      *(float4*)(act + c * 32 + (ch ^ 16)) = make_float4(leaky(acc[4][j]), leaky(acc[5][j]), leaky(acc[6][j]), leaky(acc[7][j]));
    }
    __syncwarp();
  }
};

// ---------------------------------------------------------------- variant B ---
// 8x8 per thread, packed FFMA2: accumulator pairs over ROWS (natural from LDS.128 of a),
// weights duplicated into both halves with one mov each.
struct VarB {
  static constexpr int ROWS = 32;
  static __device__ void layer(float* act, const float* w, const float* bias, int lane) {
    const int rg = lane >> 3, cg = lane & 7;
    u64 acc[4][8];  // [row pair][col]
#pragma unroll
    for (int j = 0; j < 8; ++j) { float b = bias[(j < 4) ? cg * 4 + j : 32 + cg * 4 + j - 4];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i][j] = pack2(b, b); }
    u64 a[4]; float b[8];
    auto lda = [&](const float* p, int ch, u64 (&o)[4]) { ulonglong2 lo = *(const ulonglong2*)(p + ch), hi = *(const ulonglong2*)(p + (ch ^ 16));
      o[0]=lo.x;o[1]=lo.y;o[2]=hi.x;o[3]=hi.y; };
    auto ldb = [&](const float* p, float (&o)[8]) { float4 lo = *(const float4*)(p + cg * 4), hi = *(const float4*)(p + 32 + cg * 4);
      o[0]=lo.x;o[1]=lo.y;o[2]=lo.z;o[3]=lo.w;o[4]=hi.x;o[5]=hi.y;o[6]=hi.z;o[7]=hi.w; };
    lda(act, rg << 2, a); ldb(w, b);
#pragma unroll 1
    for (int k0 = 0; k0 < K; k0 += 4) {
      const int c0 = (rg ^ swz(k0)) << 2, c1 = (rg ^ swz(k0 + 4)) << 2;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const int k = k0 + kk; u64 an[4]; float bn[8];
        lda(act + (k + 1) * 32, kk == 3 ? c1 : c0, an); ldb(w + (k + 1) * N, bn);
#pragma unroll
        for (int j = 0; j < 8; ++j) { const u64 bb = pack2(b[j], b[j]);
#pragma unroll
          for (int i = 0; i < 4; ++i) ffma2(acc[i][j], a[i], bb); }
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = an[i];
#pragma unroll
        for (int j = 0; j < 8; ++j) b[j] = bn[j];
      }
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = (j < 4) ? cg * 4 + j : 32 + cg * 4 + j - 4; const int ch = (rg ^ swz(c)) << 2;
      float r[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) unpack2(acc[i][j], r[2 * i], r[2 * i + 1]);
      *(float4*)(act + c * 32 + ch) = make_float4(leaky(r[0]), leaky(r[1]), leaky(r[2]), leaky(r[3]));
      *(float4*)(act + c * 32 + (ch ^ 16)) = make_float4(leaky(r[4]), leaky(r[5]), leaky(r[6]), leaky(r[7]));
    }
    __syncwarp();
  }
  static __device__ int idx(int k, int row) { return VarA::idx(k, row); }
};

// ---------------------------------------------------------------- variant C/D ---
// 16 rows x 8 cols per thread, 64-row warp tile.  act[k][64 rows]; lane = rg*8+cg, rg 0..3
// owns row chunks {rg, rg+4, rg+8, rg+12} (4 rows each); swizzle chunk ^ (swz(k)) on the low 3 bits.
template <bool PACKED>
struct VarCD {
  static constexpr int ROWS = 64;
  static __device__ int idx(int k, int row) { int ch = row >> 2; ch = (ch & 8) | ((ch & 7) ^ swz(k)); return k * 64 + ((ch << 2) | (row & 3)); }
  static __device__ void layer(float* act, const float* w, const float* bias, int lane) {
    const int rg = lane >> 3, cg = lane & 7;
    auto ldb = [&](const float* p, float (&o)[8]) { float4 lo = *(const float4*)(p + cg * 4), hi = *(const float4*)(p + 32 + cg * 4);
      o[0]=lo.x;o[1]=lo.y;o[2]=lo.z;o[3]=lo.w;o[4]=hi.x;o[5]=hi.y;o[6]=hi.z;o[7]=hi.w; };
    if constexpr (PACKED) {
      u64 acc[8][8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { float b = bias[(j < 4) ? cg * 4 + j : 32 + cg * 4 + j - 4];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i][j] = pack2(b, b); }
      u64 a[8]; float b[8];
      auto lda = [&](const float* p, int ch, u64 (&o)[8]) {
        ulonglong2 q0 = *(const ulonglong2*)(p + ch), q1 = *(const ulonglong2*)(p + (ch ^ 16)), q2 = *(const ulonglong2*)(p + 32 + ch), q3 = *(const ulonglong2*)(p + 32 + (ch ^ 16));
        o[0]=q0.x;o[1]=q0.y;o[2]=q1.x;o[3]=q1.y;o[4]=q2.x;o[5]=q2.y;o[6]=q3.x;o[7]=q3.y; };
      lda(act, rg << 2, a); ldb(w, b);
#pragma unroll 1
      for (int k0 = 0; k0 < K; k0 += 2) {
        const int c0 = (rg ^ swz(k0)) << 2, c1 = (rg ^ swz(k0 + 2)) << 2;
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          const int k = k0 + kk; u64 an[8]; float bn[8];
          lda(act + (k + 1) * 64, kk == 1 ? c1 : c0, an); ldb(w + (k + 1) * N, bn);
#pragma unroll
          for (int j = 0; j < 8; ++j) { const u64 bb = pack2(b[j], b[j]);
#pragma unroll
            for (int i = 0; i < 8; ++i) ffma2(acc[i][j], a[i], bb); }
#pragma unroll
          for (int i = 0; i < 8; ++i) { a[i] = an[i]; b[i] = bn[i]; }
        }
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = (j < 4) ? cg * 4 + j : 32 + cg * 4 + j - 4; const int ch = (rg ^ swz(c)) << 2;
        float r[16];
#pragma unroll
        for (int i = 0; i < 8; ++i) unpack2(acc[i][j], r[2 * i], r[2 * i + 1]);
        float* base = act + c * 64;
        *(float4*)(base + ch) = make_float4(leaky(r[0]), leaky(r[1]), leaky(r[2]), leaky(r[3]));
        *(float4*)(base + (ch ^ 16)) = make_float4(leaky(r[4]), leaky(r[5]), leaky(r[6]), leaky(r[7]));
        *(float4*)(base + 32 + ch) = make_float4(leaky(r[8]), leaky(r[9]), leaky(r[10]), leaky(r[11]));
        *(float4*)(base + 32 + (ch ^ 16)) = make_float4(leaky(r[12]), leaky(r[13]), leaky(r[14]), leaky(r[15]));
      }
      __syncwarp();
    } else {
      float acc[16][8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { float b = bias[(j < 4) ? cg * 4 + j : 32 + cg * 4 + j - 4];
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i][j] = b; }
      float a[16], b[8];
      auto lda = [&](const float* p, int ch, float (&o)[16]) {
        float4 q0 = *(const float4*)(p + ch), q1 = *(const float4*)(p + (ch ^ 16)), q2 = *(const float4*)(p + 32 + ch), q3 = *(const float4*)(p + 32 + (ch ^ 16));
        o[0]=q0.x;o[1]=q0.y;o[2]=q0.z;o[3]=q0.w;o[4]=q1.x;o[5]=q1.y;o[6]=q1.z;o[7]=q1.w;o[8]=q2.x;o[9]=q2.y;o[10]=q2.z;o[11]=q2.w;o[12]=q3.x;o[13]=q3.y;o[14]=q3.z;o[15]=q3.w; };
      lda(act, rg << 2, a); ldb(w, b);
#pragma unroll 1
      for (int k0 = 0; k0 < K; k0 += 2) {
        const int c0 = (rg ^ swz(k0)) << 2, c1 = (rg ^ swz(k0 + 2)) << 2;
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          const int k = k0 + kk; float an[16], bn[8];
          lda(act + (k + 1) * 64, kk == 1 ? c1 : c0, an); ldb(w + (k + 1) * N, bn);
#pragma unroll
          for (int i = 0; i < 16; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
#pragma unroll
          for (int i = 0; i < 16; ++i) a[i] = an[i];
#pragma unroll
          for (int j = 0; j < 8; ++j) b[j] = bn[j];
        }
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = (j < 4) ? cg * 4 + j : 32 + cg * 4 + j - 4; const int ch = (rg ^ swz(c)) << 2;
        float* base = act + c * 64;
        *(float4*)(base + ch) = make_float4(leaky(acc[0][j]), leaky(acc[1][j]), leaky(acc[2][j]), leaky(acc[3][j]));
        *(float4*)(base + (ch ^ 16)) = make_float4(leaky(acc[4][j]), leaky(acc[5][j]), leaky(acc[6][j]), leaky(acc[7][j]));
        *(float4*)(base + 32 + ch) = make_float4(leaky(acc[8][j]), leaky(acc[9][j]), leaky(acc[10][j]), leaky(acc[11][j]));
        *(float4*)(base + 32 + (ch ^ 16)) = make_float4(leaky(acc[12][j]), leaky(acc[13][j]), leaky(acc[14][j]), leaky(acc[15][j]));
      }
      __syncwarp();
    }
  }
};

// act row order inside a 64-row tile for VarCD: thread rg owns chunks rg, rg^4.. -> rows
// (ch<<2 | r): chunk list {rg, rg+4, rg+8, rg+12} maps to a[0..3],a[4..7],a[8..11],a[12..15].

template <class V>
__global__ void __launch_bounds__(256, 1) bench_kernel(const float* __restrict__ wg, const float* __restrict__ xin,
                                                       float* __restrict__ out, int iters, int warps_used) {
  extern __shared__ __align__(128) float smem[];
  float* w = smem;  // LAYERS x (K*N + N)
  for (int i = threadIdx.x; i < LAYERS * (K * N + N); i += blockDim.x) w[i] = wg[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp >= warps_used) return;
  float* act = smem + LAYERS * (K * N + N) + 64 + warp * (K + 1) * V::ROWS;
  const int tile = blockIdx.x * warps_used + warp;
  for (int r = lane; r < V::ROWS; r += 32)
    for (int k = 0; k < K; ++k) act[V::idx(k, r)] = xin[((size_t)tile * V::ROWS + r) * K + k];
  __syncwarp();
  for (int it = 0; it < iters; ++it)
    for (int l = 0; l < LAYERS; ++l) V::layer(act, w + l * (K * N + N), w + l * (K * N + N) + K * N, lane);
  for (int r = lane; r < V::ROWS; r += 32)
    for (int k = 0; k < K; ++k) out[((size_t)tile * V::ROWS + r) * K + k] = act[V::idx(k, r)];
}

__global__ void __launch_bounds__(256) peak_ffma(float* out, int iters, float a, float b) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
  float s = 0; for (int i = 0; i < 16; ++i) s += x[i];
  if (s == 123.456f) out[0] = s;
}
__global__ void __launch_bounds__(256) peak_ffma2(float* out, int iters, float a, float b) {
  u64 x[16]; const u64 aa = pack2(a, a), bb = pack2(b, b);
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = pack2(threadIdx.x * 1e-3f + i, i);
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(aa), "l"(bb));
  float s = 0; for (int i = 0; i < 16; ++i) { float lo, hi; unpack2(x[i], lo, hi); s += lo + hi; }
  if (s == 123.456f) out[0] = s;
}

static void cpu_ref(const std::vector<float>& w, std::vector<float>& x, int rows, int iters) {
  std::vector<float> y(K);
  for (int r = 0; r < rows; ++r) {
    float* a = &x[(size_t)r * K];
    for (int it = 0; it < iters; ++it)
      for (int l = 0; l < LAYERS; ++l) {
        const float* W = &w[l * (K * N + N)]; const float* B = W + K * N;
        for (int n = 0; n < N; ++n) { float acc = B[n]; for (int k = 0; k < K; ++k) acc = fmaf(a[k], W[k * N + n], acc); y[n] = acc > 0 ? acc : 0.2f * acc; }
        for (int n = 0; n < N; ++n) a[n] = y[n];
      }
  }
}

template <class V>
static void run(const char* name, int warps, const float* wd, const float* xd, float* od, const std::vector<float>& ref,
                int sms, double clk_ghz) {
  const int smem = (LAYERS * (K * N + N) + 64 + warps * (K + 1) * V::ROWS) * 4;
  CK(cudaFuncSetAttribute(bench_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int rows = sms * warps * V::ROWS;
  // correctness, 2 iterations
  bench_kernel<V><<<sms, warps * 32, smem>>>(wd, xd, od, 2, warps);
  CK(cudaDeviceSynchronize());
  std::vector<float> got((size_t)rows * K);
  CK(cudaMemcpy(got.data(), od, got.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0; int bad = 0;
  for (int i = 0; i < 64 * K; ++i) { double e = fabs(got[i] - ref[i]); if (e > maxerr) maxerr = e; if (got[i] != ref[i]) ++bad; }
  const int iters = 2000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0); bench_kernel<V><<<sms, warps * 32, smem>>>(wd, xd, od, iters, warps); cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  const double fma = (double)rows * iters * LAYERS * K * N;
  const double rate = fma / (best * 1e-3);
  printf("%-28s warps=%d rows/SM=%4d  %.3f ms  %.2f TFLOP/s  %.1f%% of 128 FMA/clk/SM @%.3f GHz  bit-mismatch=%d maxerr=%.2e\n", name, warps,
         warps * V::ROWS, best, 2 * rate / 1e12, 100.0 * rate / (sms * 128.0 * clk_ghz * 1e9), clk_ghz, bad, maxerr);
}

int main() {
  int dev = 0, sms = 0, clk = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev));
  const double clk_ghz = clk * 1e-6;
  printf("SMs=%d clock=%.3f GHz\n", sms, clk_ghz);
  std::vector<float> w(LAYERS * (K * N + N));
  srand(1);
  for (auto& v : w) v = ((rand() % 2001) - 1000) * 1e-3f * 0.25f;
  const int maxrows = sms * 8 * 64;
  std::vector<float> x((size_t)maxrows * K);
  for (auto& v : x) v = ((rand() % 2001) - 1000) * 1e-3f;
  float *wd, *xd, *od;
  CK(cudaMalloc(&wd, w.size() * 4)); CK(cudaMalloc(&xd, x.size() * 4)); CK(cudaMalloc(&od, x.size() * 4));
  CK(cudaMemcpy(wd, w.data(), w.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(xd, x.data(), x.size() * 4, cudaMemcpyHostToDevice));
  std::vector<float> ref(x.begin(), x.begin() + 64 * K);
  cpu_ref(w, ref, 64, 2);
  // raw pipe peaks
  {
    float* o; CK(cudaMalloc(&o, 4)); cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int which = 0; which < 2; ++which) {
      float best = 1e30f;
      for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        if (which == 0) peak_ffma<<<sms * 8, 256>>>(o, 4096, 0.999f, 1e-3f); else peak_ffma2<<<sms * 8, 256>>>(o, 4096, 0.999f, 1e-3f);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
      }
      const double fma = (which ? 2.0 : 1.0) * 16 * 8 * 4096.0 * sms * 8 * 256;
      printf("peak %-6s: %.2f TFLOP/s (%.1f%% of 128 FMA/clk/SM)\n", which ? "FFMA2" : "FFMA", 2 * fma / (best * 1e-3) / 1e12,
             100.0 * fma / (best * 1e-3) / (sms * 128.0 * clk_ghz * 1e9));
    }
  }
  run<VarA>("A 8x8 FFMA", 8, wd, xd, od, ref, sms, clk_ghz);
  run<VarU<4>>("U4 8x8 FFMA", 8, wd, xd, od, ref, sms, clk_ghz);
  run<VarU<8>>("U8 8x8 FFMA", 8, wd, xd, od, ref, sms, clk_ghz);
  run<VarU<16>>("U16 8x8 FFMA", 8, wd, xd, od, ref, sms, clk_ghz);
  run<VarU<32>>("U32 8x8 FFMA", 8, wd, xd, od, ref, sms, clk_ghz);
  run<VarU<64>>("U64 8x8 FFMA", 8, wd, xd, od, ref, sms, clk_ghz);
  run<VarU<16>>("U16 8x8 FFMA", 4, wd, xd, od, ref, sms, clk_ghz);
  run<VarA>("A 8x8 FFMA", 4, wd, xd, od, ref, sms, clk_ghz);
  run<VarB>("B 8x8 FFMA2", 8, wd, xd, od, ref, sms, clk_ghz);
  run<VarB>("B 8x8 FFMA2", 4, wd, xd, od, ref, sms, clk_ghz);
  run<VarCD<false>>("C 16x8 FFMA", 4, wd, xd, od, ref, sms, clk_ghz);
  run<VarCD<false>>("C 16x8 FFMA", 8, wd, xd, od, ref, sms, clk_ghz);
  run<VarCD<true>>("D 16x8 FFMA2", 8, wd, xd, od, ref, sms, clk_ghz);
  run<VarCD<true>>("D 16x8 FFMA2", 4, wd, xd, od, ref, sms, clk_ghz);
  return 0;
}
