// tcgen05 / TMEM primitives (sm_100a inline PTX) shared by the tensor-core kernels.
//
// The building block is  D[128 x N] (TMEM, fp32) (+)= A[128 x K] (TMEM) * B[K x N] (smem)
// with kind::tf32 and fp32 operands split into two tf32-exact halves (hi + lo), evaluated
// as hi*hi + lo*hi + hi*lo ("3xTF32"): the dropped lo*lo term and the tf32 truncation of
// lo are below 2^-21 relative per product, i.e. fp32-level (tools/umma_probe.cu measures
// 6e-7 relative on a 64-term dot product against float64; 3e-4 for a single TF32 pass).
//
// Layouts
//  * A and D: row i of the 128-row tile is TMEM lane i, one 32-bit column per element.
//    The thread that owns row i (warp w%4 == i/32, lane i%32) writes A with tcgen05.st and
//    reads D with tcgen05.ld (shape 32x32b), so a row never leaves its thread.
//  * B: K-major, no swizzle, canonical core matrices of 8 rows x 16 bytes:
//    element (n, k) of B^T at float index (k/4)*(N*4) + n*4 + (k%4), i.e. [K/4][N][4];
//    descriptor LBO = N*16 bytes (next 4 k), SBO = 128 bytes (next 8 n).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bgm {
namespace umma {

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One lane of a converged warp (elect.sync): used with warp-uniform operands so that the
// tcgen05.mma / commit operands stay in uniform registers (a divergent `if (tid == 0)` makes
// the compiler wrap every MMA in an ELECT / R2UR waterfall loop, ~60 cycles per issue).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xffffffff;\n\tselp.u32 %0, 1, 0, px;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---- TMEM allocation (one warp, all 512 columns: the kernels run one CTA per SM) ----
__device__ __forceinline__ void tmem_alloc512(uint32_t* slot_in_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_addr(slot_in_smem)) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc512(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(taddr) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- TMEM <-> registers, 32 lanes x 32 bit x 16 / 32 columns (no wait inside) ----
__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
      "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
      "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}

__device__ __forceinline__ void ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

// ---- fp32 -> (hi, lo), both tf32-exact up to the truncation of lo's last bits ----
// hi = v rounded to nearest at 11 significant bits; lo = v - hi is exact in fp32.
__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo) {
  hi = (__float_as_uint(v) + 0x1000u) & 0xffffe000u;
  lo = __float_as_uint(v - __uint_as_float(hi));
}

// Packed variant (two values per instruction; sm_100 FFMA2), three instructions per PAIR:
//   s  = rn(v + 8192 v)      one rounding of 8193 v: its error is at most half an 11-bit ulp of v
//   hi = s - 8192 v          exact (both are multiples of 2^(e_v - 10)): v rounded to nearest at 11 significant bits
//                            (at 10 for mantissas within 2.4e-4 of 2, where 8193 v crosses a binade: still tf32-exact)
//   lo = v - hi              exact
// (Veltkamp's split with the factor 2^13 + 1 minus one operation: 8192 v is exact, so the c - (c - v) detour that
// protects against the rounding of c - v is not needed.  The kernels are bound by issue slots, and the hi / lo split
// runs once per activation.)  Range: 8193 |v| must not overflow, i.e. |v| < 4.1e34 (beyond that hi is NaN and the
// proposal is rejected); NaN / inf inputs propagate as NaN.
__device__ __forceinline__ void split_tf32_x2(float2 v, float2& hi, float2& lo) {
  const float2 k = make_float2(8192.f, 8192.f), mk = make_float2(-8192.f, -8192.f), m1 = make_float2(-1.f, -1.f);
  const float2 s = __ffma2_rn(v, k, v);
  hi = __ffma2_rn(v, mk, s);
  lo = __ffma2_rn(hi, m1, v);
}

// ---- descriptors ----
// shared-memory matrix descriptor (K-major, SWIZZLE_NONE, version 1): start address, LBO, SBO
// in 16-byte units.  cute::UMMA::SmemDescriptor has the same bit layout.
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  const uint32_t lo = ((saddr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
  const uint32_t hi = ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14);
  return ((uint64_t)hi << 32) | lo;
}
// instruction descriptor, kind::tf32, fp32 accumulate, A and B K-major, M = 128.
__host__ __device__ constexpr uint32_t idesc_tf32_m128(int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
}
// D (+)= A * B, A in TMEM, B through a shared-memory descriptor; issued by ONE thread.
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all MMAs issued so far by this thread arrive on `bar` when they complete
__device__ __forceinline__ void mma_commit(uint32_t bar_saddr) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_saddr) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar_saddr, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_saddr), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar_saddr, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar_saddr), "r"(parity)
        : "memory");
  }
}

// One Dense layer with K = 8*KSTEPS inputs as 3xTF32: per k-step (hi*hi, lo*hi, hi*lo).
// b_hi / b_lo: shared-memory byte addresses of the [K/4][N][4] images.
template <int N, int KSTEPS>
__device__ __forceinline__ void issue_layer(uint32_t tD, uint32_t tA_hi, uint32_t tA_lo, uint32_t b_hi,
                                            uint32_t b_lo) {
  constexpr uint32_t idesc = idesc_tf32_m128(N);
  uint64_t dh = smem_desc(b_hi, N * 16, 128), dl = smem_desc(b_lo, N * 16, 128);
#pragma unroll
  for (int ks = 0; ks < KSTEPS; ++ks) {
    mma_tf32_ts(tD, tA_hi + ks * 8, dh, idesc, ks > 0);
    mma_tf32_ts(tD, tA_lo + ks * 8, dh, idesc, 1);
    mma_tf32_ts(tD, tA_hi + ks * 8, dl, idesc, 1);
    dh += (N * 32) >> 4;   // next 8 k: 2 core-matrix columns of N*16 bytes
    dl += (N * 32) >> 4;
  }
}
template <int N>
__device__ __forceinline__ void issue_layer_k64(uint32_t tD, uint32_t tA_hi, uint32_t tA_lo, uint32_t b_hi,
                                                uint32_t b_lo) {
  issue_layer<N, 8>(tD, tA_hi, tA_lo, b_hi, b_lo);
}

}  // namespace umma
}  // namespace bgm
