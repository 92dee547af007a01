// Shared device helpers: error plumbing, Philox4x32-10, fp32 math that mirrors the
// reference's TF/NumPy formulas, the warp-tile MAC used by every MLP layer, and the
// bulk (TMA) shared-memory loader.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "../../include/bgm_b200.h"

namespace bgm {

// ------------------------------------------------------------------ errors --
extern thread_local std::string g_last_error;
int fail(int code, const std::string& msg);
#define BGM_CUDA_OK(expr)                                                          \
  do {                                                                             \
    cudaError_t e__ = (expr);                                                      \
    if (e__ != cudaSuccess)                                                        \
      return ::bgm::fail(BGM_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
  } while (0)

// ------------------------------------------------------------------- math ---
// LeakyReLU(alpha=0.2), networks/base.py:45.
__device__ __forceinline__ float leaky(float v) { return v > 0.f ? v : 0.2f * v; }
// tf.nn.softplus, stable form.
__device__ __forceinline__ float softplus_f(float t) {
  return fmaxf(t, 0.f) + log1pf(expf(-fabsf(t)));
}
__device__ __forceinline__ float sigmoid_f(float t) { return 1.f / (1.f + expf(-t)); }

// ------------------------------------------------------------------ Philox --
// Philox4x32-10 (Salmon et al., SC'11), stateless: every draw is a pure function
// of (seed, global row, iteration, stream kind), so results do not depend on the
// grid shape or on how rows are sharded across GPUs.
__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}
__host__ __device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    uint32_t hi0 = mulhi32(M0, c.x), lo0 = M0 * c.x;
    uint32_t hi1 = mulhi32(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0;
    k.y += W1;
  }
  return c;
}
enum : uint32_t { NOISE_PROPOSAL = 0, NOISE_ACCEPT = 1, NOISE_EFFECT = 2, NOISE_MOMENTUM = 3,
                  NOISE_PREDICT = 4 };
constexpr uint32_t T_INIT = 0xFFFFFFFFu;  // "iteration" of the initial-state draw

__device__ __forceinline__ uint4 noise_block(uint64_t seed, int64_t row, uint32_t t, uint32_t kind,
                                             uint32_t j) {
  uint4 c = make_uint4(t, (uint32_t)row, (uint32_t)((uint64_t)row >> 32), (kind << 24) | j);
  return philox4x32_10(c, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}
// (0,1] with 32 random bits, and [0,1) with 24.
__device__ __forceinline__ float u01_open0(uint32_t x) {
  return fmaf((float)x, 2.3283064365386963e-10f, 1.1641532182693481e-10f);
}
__device__ __forceinline__ float u01_open1(uint32_t x) { return (float)(x >> 8) * 5.9604644775390625e-08f; }
// Box-Muller: two N(0,1) from two words.
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& n0, float& n1) {
  float r = sqrtf(-2.f * logf(u01_open0(a)));
  float s, c;
  sincospif(2.f * u01_open1(b), &s, &c);
  n0 = r * c;
  n1 = r * s;
}
// j-th group of 4 unit normals of (row, t, kind).
__device__ __forceinline__ void normal4(uint64_t seed, int64_t row, uint32_t t, uint32_t kind,
                                        uint32_t j, float (&out)[4]) {
  uint4 r = noise_block(seed, row, t, kind, j);
  box_muller(r.x, r.y, out[0], out[1]);
  box_muller(r.z, r.w, out[2], out[3]);
}
__device__ __forceinline__ float normal1(uint64_t seed, int64_t row, uint32_t t, uint32_t kind,
                                         uint32_t j) {
  uint4 r = noise_block(seed, row, t, kind, j);
  float a, b;
  box_muller(r.x, r.y, a, b);
  return a;
}
__device__ __forceinline__ float uniform1(uint64_t seed, int64_t row, uint32_t t, uint32_t kind) {
  return u01_open1(noise_block(seed, row, t, kind, 0).x);
}

// -------------------------------------------------------------- warp tiles --
// One warp owns a tile of 32 rows and carries it through every layer; nothing is
// shared between warps except the read-only weight image.  Activations live in the
// warp's private shared-memory buffer, TRANSPOSED: act[k][row] (32 floats per k).
// Lane = rg*8 + cg: rg (0..3) picks 8 rows, cg (0..7) picks C columns of an NT-wide
// column tile (NT = 8*C).  Rows/columns are interleaved in groups of 4 so that every
// LDS.128 of a warp touches one contiguous 64 B / 128 B span (conflict-free, the
// rest is broadcast).
constexpr int TILE_ROWS = 32;
constexpr int RPT = 8;  // rows per thread

__device__ __forceinline__ int row_of(int rg, int i) { return (i < 4) ? rg * 4 + i : 16 + rg * 4 + (i - 4); }

template <int C> struct ColMap;
template <> struct ColMap<8> {
  static constexpr int NT = 64;
  static __device__ __forceinline__ int col(int cg, int j) { return (j < 4) ? cg * 4 + j : 32 + cg * 4 + (j - 4); }
};
template <> struct ColMap<4> {
  static constexpr int NT = 32;
  static __device__ __forceinline__ int col(int cg, int j) { return cg * 4 + j; }
};
template <> struct ColMap<1> {
  static constexpr int NT = 8;
  static __device__ __forceinline__ int col(int cg, int) { return cg; }
};

// Activation buffers are XOR-swizzled: element (k, row) lives at
//   k*32 + (((row>>2) ^ ((k>>2)&7)) << 2 | (row&3)),
// i.e. the 16-byte chunk index of the row is XORed with bits 2..4 of k.  A lane's
// loads are unaffected (all 8 lanes of a quarter-warp read the same chunk), while the
// epilogue's STS.128 of 8 lanes (8 different k = output columns, same row chunk) land
// in 8 different bank groups instead of one.
__device__ __forceinline__ int act_swz(int k) { return (k >> 2) & 7; }
__device__ __forceinline__ int act_idx(int k, int row) {
  return k * TILE_ROWS + ((((row >> 2) ^ act_swz(k)) << 2) | (row & 3));
}

// a[0..3] = rows rg*4.., a[4..7] = rows 16+rg*4.. of activation row k; `chunk` is the
// swizzled chunk offset (rg ^ swz(k)) << 2 in floats.
__device__ __forceinline__ void lds_rows(const float* in_k, int chunk, float (&a)[RPT]) {
  float4 lo = *reinterpret_cast<const float4*>(in_k + chunk);
  float4 hi = *reinterpret_cast<const float4*>(in_k + (chunk ^ 16));
  a[0] = lo.x; a[1] = lo.y; a[2] = lo.z; a[3] = lo.w;
  a[4] = hi.x; a[5] = hi.y; a[6] = hi.z; a[7] = hi.w;
}
template <int C>
__device__ __forceinline__ void lds_cols(const float* w, int cg, float (&b)[C]) {
  if constexpr (C == 8) {
    float4 lo = *reinterpret_cast<const float4*>(w + cg * 4);
    float4 hi = *reinterpret_cast<const float4*>(w + 32 + cg * 4);
    b[0] = lo.x; b[1] = lo.y; b[2] = lo.z; b[3] = lo.w;
    b[4] = hi.x; b[5] = hi.y; b[6] = hi.z; b[7] = hi.w;
  } else if constexpr (C == 4) {
    float4 lo = *reinterpret_cast<const float4*>(w + cg * 4);
    b[0] = lo.x; b[1] = lo.y; b[2] = lo.z; b[3] = lo.w;
  } else {
    b[0] = w[cg];
  }
}

// acc[i][j] += sum_k in[k][row_i] * w[k][col_j], k = 0..kp-1 (kp % 4 == 0), strictly
// in k order in fp32 FMA.  Operands for k+1 are fetched while k is being multiplied
// (register double buffer); the fetch after the last k reads one row past the tile,
// which every caller keeps inside the shared-memory allocation, and is discarded.
template <int C>
__device__ __forceinline__ void tile_mac(const float* __restrict__ in, const float* __restrict__ w,
                                         int kp, int rg, int cg, float (&acc)[RPT][C]) {
  constexpr int NT = ColMap<C>::NT;
  float a[RPT], b[C];
  lds_rows(in, rg << 2, a);
  lds_cols<C>(w, cg, b);
#pragma unroll 1
  for (int k0 = 0; k0 < kp; k0 += 4) {
    const int ch_cur = (rg ^ act_swz(k0)) << 2;
    const int ch_nxt = (rg ^ act_swz(k0 + 4)) << 2;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const int k = k0 + kk;
      float an[RPT], bn[C];
      lds_rows(in + (k + 1) * TILE_ROWS, kk == 3 ? ch_nxt : ch_cur, an);
      lds_cols<C>(w + (k + 1) * NT, cg, bn);
#pragma unroll
      for (int i = 0; i < RPT; ++i)
#pragma unroll
        for (int j = 0; j < C; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
#pragma unroll
      for (int i = 0; i < RPT; ++i) a[i] = an[i];
#pragma unroll
      for (int j = 0; j < C; ++j) b[j] = bn[j];
    }
  }
}

// One column tile of one Dense layer.
struct TileOp {
  int w_off;            // float offset of W[kp][NT] inside the weight image
  int b_off;            // float offset of bias[NT]
  short kp;             // reduction length, padded to a multiple of 4
  unsigned char ctype;  // 0: C=8 (NT 64), 1: C=4 (NT 32), 2: C=1 (NT 8)
  unsigned char src;    // 0: input buffer (zin), 1: activation buffer
  unsigned char epi;    // EPI_*
  unsigned char post;   // POST_*: what to assemble after this tile (last tile of a net)
  short c0;             // EPI_SSE: first data column; EPI_OUT: first scratch slot
  short nvalid;         // valid columns in this tile
};
enum : unsigned char { EPI_ACT = 0, EPI_SSE = 1, EPI_OUT = 2, EPI_LIN = 3 };
enum : unsigned char { POST_NONE = 0, POST_G = 1, POST_F = 2, POST_H = 3 };

// ------------------------------------------------------- bulk smem loader ---
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
// Copies `bytes` (multiple of 16, < 1 MiB) from global to shared memory with the
// bulk async-copy engine (TMA, cp.async.bulk), completion on an mbarrier.  Must be
// called by all threads of the CTA; returns when the data is visible to all.
__device__ __forceinline__ void bulk_load_to_smem(void* dst, const void* src, uint32_t bytes,
                                                  uint64_t* bar) {
  const uint32_t bar_a = smem_u32(bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes)
                 : "memory");
    const uint32_t CH = 32768;
    const char* s = reinterpret_cast<const char*>(src);
    uint32_t d = smem_u32(dst);
    for (uint32_t off = 0; off < bytes; off += CH) {
      uint32_t sz = bytes - off < CH ? bytes - off : CH;
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d + off),
          "l"(s + off), "r"(sz), "r"(bar_a)
          : "memory");
    }
  }
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar_a), "r"(0u)
        : "memory");
  }
}

}  // namespace bgm
