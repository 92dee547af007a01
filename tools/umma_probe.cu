// Probe of the tcgen05 building block the tensor-core sampler is made of:
//   D[128 x N] (TMEM, fp32) = A[128 x K] (TMEM, tf32 hi/lo split) * B[K x N] (smem, hi/lo split)
// evaluated as 3xTF32 (hi*hi + lo*hi + hi*lo), A written by the epilogue threads with
// tcgen05.st, B in the no-swizzle K-major canonical layout [K/4][N][4].
// Prints the error of the 1xTF32 and 3xTF32 results against float64 and times a chain
// of dependent layers (ld D -> LeakyReLU -> split -> st A -> MMA) with 1 and 2 warpgroups.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_probe tools/umma_probe.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
               :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                  "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
               : "memory");
}
// round-to-nearest split: hi has 11 significant bits (tf32-exact), lo = v - hi exactly
__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo) {
  hi = (__float_as_uint(v) + 0x1000u) & 0xffffe000u;
  lo = __float_as_uint(v - __uint_as_float(hi));
}
__device__ __forceinline__ uint64_t make_bdesc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint32_t lo = ((saddr >> 4) & 0x3FFFu) | (((lbo >> 4) & 0x3FFFu) << 16);
  uint32_t hi = ((sbo >> 4) & 0x3FFFu) | (1u << 14);
  return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ void mma_tf32_ts(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
               :: "r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  }
}

// smem: B hi [K/4][N][4] | B lo ; mode bit0: swap LBO/SBO, bit1: single pass (hi*hi only)
// One warpgroup (128 threads) per tile; `wgs` warpgroups per CTA, each with its own TMEM slice.
template <int N, int K>
__global__ void __launch_bounds__(256, 1)
probe_kernel(const float* __restrict__ A, const float* __restrict__ W, float* __restrict__ out,
             int mode, int layers, long long* cycles) {
  extern __shared__ __align__(1024) float smem[];
  __shared__ uint64_t mbar[2];
  __shared__ uint32_t tmem_base_s;
  float* Bhi = smem;
  float* Blo = smem + K * N;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wg = warp >> 2, wtid = tid & 127;
  // weights: W[k][n] -> [k/4][n][k%4], split on the fly (the product does this on the host)
  for (int i = tid; i < K * N; i += blockDim.x) {
    const int k = i / N, n = i - k * N;
    uint32_t hi, lo;
    split_tf32(W[i], hi, lo);
    lo = (lo + 0x1000u) & 0xffffe000u;
    const int idx = (k >> 2) * (N * 4) + n * 4 + (k & 3);
    Bhi[idx] = __uint_as_float(hi);
    Blo[idx] = __uint_as_float(lo);
  }
  if (tid == 0) {
    for (int g = 0; g < 2; ++g) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar[g])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> tensor core
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tmem_base_s + (uint32_t)wg * 256u;          // 256 columns per warpgroup
  const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
  const uint32_t tA_hi = tbase, tA_lo = tbase + 64, tD = tbase + 128;
  const uint32_t bar = smem_u32(&mbar[wg]);
  const uint32_t lbo = (mode & 1) ? 128u : (uint32_t)N * 16u;
  const uint32_t sbo = (mode & 1) ? (uint32_t)N * 16u : 128u;
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
  const int row = wg * 128 + wtid;

  // layer 0 input: this thread's row of A
  {
    const float* ar = A + (size_t)row * K;
#pragma unroll
    for (int c = 0; c < K; c += 16) {
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) split_tf32(ar[c + j], hi[j], lo[j]);
      tmem_st16(tA_hi + lane_off + c, hi);
      tmem_st16(tA_lo + lane_off + c, lo);
    }
  }
  uint32_t parity = 0;
  long long t0 = clock64();
  for (int l = 0; l < layers; ++l) {
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    asm volatile("bar.sync %0, 128;" :: "r"(1 + wg) : "memory");
    if (wtid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int ks = 0; ks < K / 8; ++ks) {
        const uint64_t bh = make_bdesc(smem_u32(Bhi) + ks * N * 32, lbo, sbo);
        const uint64_t bl = make_bdesc(smem_u32(Blo) + ks * N * 32, lbo, sbo);
        mma_tf32_ts(tD, tA_hi + ks * 8, bh, idesc, ks > 0);
        if (!(mode & 2)) {
          mma_tf32_ts(tD, tA_lo + ks * 8, bh, idesc, 1);
          mma_tf32_ts(tD, tA_hi + ks * 8, bl, idesc, 1);
        }
      }
      mma_commit(bar);
    }
    mbar_wait(bar, parity);
    parity ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (l + 1 == layers) break;
    // epilogue of a hidden layer: LeakyReLU, split, back into the A slots (needs N == K)
    if constexpr (N == K) {
#pragma unroll
      for (int c = 0; c < N; c += 16) {
        float v[16];
        tmem_ld16(tD + lane_off + c, v);
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float a = fmaxf(v[j], 0.2f * v[j]);
          split_tf32(a, hi[j], lo[j]);
        }
        tmem_st16(tA_hi + lane_off + c, hi);
        tmem_st16(tA_lo + lane_off + c, lo);
      }
    }
  }
  long long t1 = clock64();
  if (tid == 0 && cycles) cycles[blockIdx.x] = t1 - t0;
  float* orow = out + (size_t)row * N;
#pragma unroll
  for (int c = 0; c < N; c += 16) {
    float v[16];
    tmem_ld16(tD + lane_off + c, v);
#pragma unroll
    for (int j = 0; j < 16; ++j) orow[c + j] = v[j];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem_base_s));
}

template <int N, int K>
static void run_case(int mode, int layers, int wgs, bool verbose) {
  const int M = 128 * wgs;
  std::vector<float> A(M * K), W(K * N), out(M * N);
  srand(1234);
  for (auto& a : A) a = (float)rand() / RAND_MAX * 2.f - 1.f;
  for (auto& w : W) w = ((float)rand() / RAND_MAX * 2.f - 1.f) * 0.3f;
  float *dA, *dW, *dO;
  long long* dC;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dW, W.size() * 4)); CK(cudaMalloc(&dO, out.size() * 4));
  CK(cudaMalloc(&dC, 8 * 148));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dO, 0xff, out.size() * 4));
  const int smem = 2 * K * N * 4;
  CK(cudaFuncSetAttribute(probe_kernel<N, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe_kernel<N, K><<<1, 128 * wgs, smem>>>(dA, dW, dO, mode, layers, dC);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(out.data(), dO, out.size() * 4, cudaMemcpyDeviceToHost));
  long long cyc;
  CK(cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost));
  // float64 reference of the layer chain
  std::vector<double> cur(A.begin(), A.end()), nxt(M * N);
  for (int l = 0; l < layers; ++l) {
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        double s = 0;
        for (int k = 0; k < K; ++k) s += cur[m * K + k] * (double)W[k * N + n];
        nxt[m * N + n] = s;
      }
    if (l + 1 < layers) {
      for (int i = 0; i < M * N; ++i) cur[i] = nxt[i] > 0 ? nxt[i] : 0.2 * nxt[i];
    }
  }
  double maxerr = 0, maxref = 0;
  for (int i = 0; i < M * N; ++i) {
    maxerr = fmax(maxerr, fabs((double)out[i] - nxt[i]));
    maxref = fmax(maxref, fabs(nxt[i]));
  }
  printf("N=%d K=%d mode=%d layers=%d wgs=%d: max|err|=%.3e (max|ref|=%.3e, rel %.3e) cycles=%lld (%.1f per layer)\n",
         N, K, mode, layers, wgs, maxerr, maxref, maxerr / maxref, cyc, (double)cyc / layers);
  if (verbose) printf("   out[0..3] = %g %g %g %g   ref = %g %g %g %g\n", out[0], out[1], out[2], out[3], nxt[0], nxt[1], nxt[2], nxt[3]);
  cudaFree(dA); cudaFree(dW); cudaFree(dO); cudaFree(dC);
}

int main() {
  // descriptor conventions: mode 0 = LBO across K chunks / SBO across 8-row groups; mode 1 = swapped
  run_case<64, 64>(2, 1, 1, true);   // 1xTF32, expect rel err ~1e-3
  run_case<64, 64>(3, 1, 1, true);   // swapped strides
  run_case<64, 64>(0, 1, 1, true);   // 3xTF32, expect rel err ~1e-6
  run_case<64, 64>(0, 1, 2, false);
  run_case<32, 64>(0, 1, 1, false);
  run_case<16, 64>(0, 1, 1, false);
  run_case<64, 8>(0, 1, 1, false);
  run_case<64, 64>(0, 5, 1, false);
  run_case<64, 64>(0, 200, 1, false);
  run_case<64, 64>(0, 200, 2, false);
  run_case<64, 64>(2, 200, 2, false);
  return 0;
}
