// BGM HMC, tensor-core engine (tcgen05 / TMEM): the same contract as hmc_kernel (hmc.cuh) --
// bgm/base.py:665-705 (get_log_posterior) under :709-830 (TFP HamiltonianMonteCarlo +
// SimpleStepSizeAdaptation), same Philox streams, same arguments -- for the standard
// generator shape (every hidden layer 64 wide, at least two of them).
//
// Plan.  A CTA owns a tile of 128 observations = 128 TMEM lanes, worked on by SIXTEEN warps: warps q, q+4, q+8, q+12
// share lane quarter q and split the columns four ways (16 of a hidden layer's 64 outputs, 8 of a head chunk's 32
// features), so that every scheduler has four warps to interleave; the four threads of a row carry the chain
// state (z, momentum, gradient) redundantly and exchange their partial sums (likelihood, first-layer transpose)
// through shared memory, added in a fixed order, so they stay bit-identical.  Every 64-wide product of one gradient evaluation runs on the tensor cores as 3xTF32
// (umma.cuh: hi*hi + lo*hi + hi*lo, fp32-level error):
//   forward   h_l   = LeakyReLU(h_{l-1} W_l + b_l), l = 2..nh             (A = h_{l-1} in TMEM, D 64 columns)
//   heads     [mu | raw]_c = h_nh [Wm_c | Wv_c], 32 features per chunk c   (A = h_nh stays resident)
//             the thread turns its 32 (mu, raw) pairs and its 32 data values into d loss / d mu, d loss / d raw,
//   head bwd  d loss / d h_nh += [dmu | draw]_c [Wm_c | Wv_c]^T            (A = [dmu | draw]_c, accumulating D)
//   backward  d loss / d h_{l-1} = (d loss / d a_l) W_l^T, l = nh..2.
// Only the first layer (z_dim inputs) and its transpose stay on the FMA pipe.  The head passes are software-
// pipelined: MMA stage k carries the forward of chunk k and the backward of chunk k-1.  No weight is resident:
// every operand is a pre-split [hi | lo] image of 32 KB streamed from L2 through a 4-slot shared-memory ring with
// cp.async.bulk on mbarriers (1.28 MB per evaluation at x_dim = 500, shared by the tile's 128 rows), refilled as
// soon as the MMAs that read a slot have committed.
#pragma once
#include "hmc.cuh"
#include "umma.cuh"

namespace bgm {

constexpr int HT_ROWS = 128;
constexpr int HT_TPR = 4;                    // threads per row
constexpr int HT_THREADS = HT_ROWS * HT_TPR;
constexpr int HT_IMG_FLOATS = 2 * 64 * 64;   // [hi | lo] of a K = 64, N = 64 operand in UMMA layout [K/4][N][4]
constexpr int HT_SLOTS = 4;
constexpr uint32_t HT_A_HI = 0, HT_A_LO = 64, HT_D = 128, HT_B_HI = 192, HT_B_LO = 256, HT_DH = 320, HT_D2 = 384;

struct HmcTcProgram {
  int enabled;
  int zd, kin, x_dim, nh;
  int n_chunks;                 // ceil(x_dim / 32)
  int n_img;                    // images per gradient evaluation: (nh - 1) + 2 n_chunks + (nh - 1)
  int n_img_fwd;                // images of the forward-only stream (posterior-predictive draws): (nh - 1) + n_chunks
  float bn_mean[HMC_MAXZ], bn_inv[HMC_MAXZ], bn_beta[HMC_MAXZ];
  // resident small arrays (float offsets): W1 [zd][64], b1 [64], hidden biases [(nh-1)][64], bm, bv [32 n_chunks]
  int off_W1, off_b1, off_bh, off_bm, off_bv, small_floats;
};

struct HtCtx {
  const HmcTcProgram* P;
  const float* stream;          // global image stream of one evaluation, n_img x HT_IMG_FLOATS
  float* ring;                  // shared: HT_SLOTS x HT_IMG_FLOATS
  const float* small;           // shared: resident small arrays
  uint64_t* full;               // [HT_SLOTS]
  uint32_t mma_bar;             // shared address of the MMA-completion barrier
  uint32_t tbase, trow;         // TMEM base / this warp's lane quarter
  int cg;                       // which quarter of the columns this thread owns (0..3)
  int r_in_tile;                // row of the tile (TMEM lane)
  float* xch;                   // shared: [HT_TPR][ZMAX + 1][HT_ROWS] partial sums exchanged between a row's threads
  uint32_t img_it, img_total;   // images consumed so far / in the whole launch (identical in every thread)
  uint32_t n_img;               // images of one pass over `stream` (the stream is periodic)
  uint32_t mma_parity;
  bool issuer_warp;
};

__device__ __forceinline__ void ht_fill(const HtCtx& C, uint32_t img_index_in_launch, int slot) {
  const uint32_t i = img_index_in_launch % C.n_img;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  mbar_expect_tx(C.full + slot, HT_IMG_FLOATS * 4u);
  bulk_g2s(C.ring + slot * HT_IMG_FLOATS, C.stream + (size_t)i * HT_IMG_FLOATS, HT_IMG_FLOATS * 4u, C.full + slot);
}

// D (+)= A * B as 3xTF32, K = 64; acc0: accumulate onto D from the first MMA on.
__device__ __forceinline__ void ht_issue(uint32_t tD, uint32_t tA_hi, uint32_t tA_lo, uint32_t b_img_saddr, uint32_t acc0) {
  constexpr uint32_t idesc = umma::idesc_tf32_m128(64);
  uint64_t dh = umma::smem_desc(b_img_saddr, 64 * 16, 128), dl = umma::smem_desc(b_img_saddr + 64 * 64 * 4, 64 * 16, 128);
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) {
    umma::mma_tf32_ts(tD, tA_hi + ks * 8, dh, idesc, (ks > 0) ? 1u : acc0);
    umma::mma_tf32_ts(tD, tA_lo + ks * 8, dh, idesc, 1);
    umma::mma_tf32_ts(tD, tA_hi + ks * 8, dl, idesc, 1);
    dh += (64 * 32) >> 4;
    dl += (64 * 32) >> 4;
  }
}

// One MMA stage: the tile's A operands are published (tcgen05.st + fence + CTA barrier), the issuer waits for the
// stage's images, issues, commits; everybody waits for the commit; the consumed slots are refilled.
//   n_img = 1: D_target (+)= A_src * image;  n_img = 2: first image -> (DH += B), second -> (D = A).
// The tile's 16 math warps never issue: a 17th CONTROL warp waits for the streamed images, issues the MMAs, commits,
// and refills the ring, following the same barrier schedule (ht_eval_ctrl).  (With the issue on a math warp every
// other warp waited ~500 cycles per head chunk at the next barrier for it: 17 % of the warp time was barrier stall.)
//   math warps:    ht_publish()  ... math ...  ht_wait_mma()
//   control warp:  ht_ctrl_begin(issue) ........ ht_ctrl_end()
__device__ __forceinline__ void ht_publish() {
  umma::wait_st();
  umma::fence_before_sync();
  __syncthreads();
}
__device__ __forceinline__ void ht_wait_mma(HtCtx& C) {
  umma::mbar_wait(C.mma_bar, C.mma_parity);
  C.mma_parity ^= 1u;
  umma::fence_after_sync();
}
template <class Issue>
__device__ __forceinline__ void ht_ctrl_begin(HtCtx& C, int n_img, Issue issue) {
  umma::fence_before_sync();
  __syncthreads();
  if (n_img > 0) {
    if (umma::elect_one()) {
      for (int j = 0; j < n_img; ++j) {
        const uint32_t it = C.img_it + j;
        mbar_wait(C.full + (it % HT_SLOTS), (it / HT_SLOTS) & 1u);
      }
      umma::fence_after_sync();
      issue();
      umma::mma_commit(C.mma_bar);
    }
    __syncwarp();
  }
}
__device__ __forceinline__ void ht_ctrl_end(HtCtx& C, int n_img) {
  if (n_img > 0) {
    ht_wait_mma(C);
    if ((threadIdx.x & 31) == 0) {
      for (int j = 0; j < n_img; ++j) {
        const uint32_t it = C.img_it + j;
        if (it + HT_SLOTS < C.img_total) ht_fill(C, it + HT_SLOTS, it % HT_SLOTS);
      }
    }
    C.img_it += n_img;
  }
}
__device__ __forceinline__ uint32_t ht_slot_addr(const HtCtx& C, uint32_t it) {
  return umma::smem_addr(C.ring + (it % HT_SLOTS) * HT_IMG_FLOATS);
}

// bias + LeakyReLU (sign bits of the thread's 16 columns into `bits`) + split, D -> A
__device__ __forceinline__ void ht_act16(const uint32_t (&r)[16], const float* bias, uint32_t& bits, uint32_t t_hi, uint32_t t_lo) {
  uint32_t hi[16], lo[16];
  bits = 0u;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float v = __uint_as_float(r[i]) + bias[i];
    if (v > 0.f) bits |= 1u << i;
    else v *= 0.2f;
    umma::split_tf32(v, hi[i], lo[i]);
  }
  umma::st16(t_hi, hi);
  umma::st16(t_lo, lo);
}
// gradient through a LeakyReLU whose sign bits are `bits` + split, 16 columns
__device__ __forceinline__ void ht_mask16(const uint32_t (&r)[16], uint32_t bits, uint32_t t_hi, uint32_t t_lo) {
  uint32_t hi[16], lo[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float g = __uint_as_float(r[i]);
    const float v = ((bits >> i) & 1u) ? g : 0.2f * g;
    umma::split_tf32(v, hi[i], lo[i]);
  }
  umma::st16(t_hi, hi);
  umma::st16(t_lo, lo);
}

// One gradient evaluation for the thread's row at z (identical in the row's four threads): returns the likelihood
// loss (0 unless want_lp) and leaves d loss / d z_in (before the BatchNormalization scale) in gz -- both summed over
// the four threads' partials in a fixed order, so the four copies agree bit for bit.
template <int ZMAX>
__device__ __forceinline__ float ht_eval(HtCtx& C, const HmcDev& D, const float (&z)[ZMAX], int lrow, bool want_lp,
                                         float (&gz)[ZMAX]) {
  const HmcTcProgram& P = *C.P;
  const int zd = P.zd, nh = P.nh, NC = P.n_chunks;
  const int cg = C.cg;
  const float* W1 = C.small + P.off_W1 + cg * 16;          // this thread's 16 columns of W1 [zd][64]
  uint32_t sg[HMC_MAXL];
#pragma unroll
  for (int l = 0; l < HMC_MAXL; ++l) sg[l] = 0u;
  const uint32_t tA_hi = C.trow + HT_A_HI + cg * 16, tA_lo = C.trow + HT_A_LO + cg * 16;
  // ---- layer 1 on the FMA pipe: a1 = b1 + BN(z) W1, columns [16 cg, 16 cg + 16) ----
  {
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = C.small[P.off_b1 + cg * 16 + i];
#pragma unroll
    for (int d = 0; d < ZMAX; ++d) {
      if (d < zd) {
        const float zv = (z[d] - P.bn_mean[d]) * P.bn_inv[d] + P.bn_beta[d];
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float4 w = *reinterpret_cast<const float4*>(W1 + d * 64 + i);
          a[i] = fmaf(zv, w.x, a[i]); a[i + 1] = fmaf(zv, w.y, a[i + 1]);
          a[i + 2] = fmaf(zv, w.z, a[i + 2]); a[i + 3] = fmaf(zv, w.w, a[i + 3]);
        }
      }
    }
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float v = a[i];
      if (v > 0.f) sg[0] |= 1u << i;
      else v *= 0.2f;
      umma::split_tf32(v, hi[i], lo[i]);
    }
    umma::st16(tA_hi, hi);
    umma::st16(tA_lo, lo);
  }
  // ---- hidden layers 2..nh ----
#pragma unroll 1
  for (int l = 1; l < nh; ++l) {
    ht_publish();
    ht_wait_mma(C);
    uint32_t r[16], bits;
    umma::ld16(C.trow + HT_D + cg * 16, r);
    umma::wait_ld();
    ht_act16(r, C.small + P.off_bh + (l - 1) * 64 + cg * 16, bits, tA_hi, tA_lo);
#pragma unroll
    for (int q = 0; q < HMC_MAXL; ++q)
      if (q == l) sg[q] = bits;
  }
  // ---- heads over chunks of 32 features; this thread: features 8 cg .. 8 cg + 7 of a chunk.  Software pipeline:
  // while the threads turn chunk k's (mu, raw) (accumulator D_{k&1}) into d loss / d (mu, raw), the tensor cores
  // run the forward of chunk k+1 (into the other accumulator) and the backward of chunk k-1 (from the operand the
  // previous iteration published); the new operand is stored only after those MMAs have committed. ----
  float loss = 0.f;
  const float* xrow = D.a.x_dev + (size_t)lrow * D.a.ldx + cg * 8;
  const float qn = __int_as_float(0x7fc00000);
  // (asm volatile: the loads must be ISSUED here, one chunk ahead, not sunk to their first use)
  auto ldg4 = [](const float* p) -> float4 {
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
  };
  auto load_x = [&](int k, float4& a, float4& b) {
    const int c = k * 32 + cg * 8;
    a = (k < NC && c < D.a.ldx) ? ldg4(xrow + k * 32) : make_float4(qn, qn, qn, qn);
    b = (k < NC && c + 4 < D.a.ldx) ? ldg4(xrow + k * 32 + 4) : make_float4(qn, qn, qn, qn);
  };
  float4 x0, x1;
  load_x(0, x0, x1);
  ht_publish();
  ht_wait_mma(C);
#pragma unroll 1
  for (int k = 0; k < NC; ++k) {
    const int n_img = (k > 0 ? 1 : 0) + (k + 1 < NC ? 1 : 0);
    const uint32_t tD_cur = C.tbase + ((k & 1) ? HT_D2 : HT_D);
    ht_publish();
    float4 xn0, xn1;
    load_x(k + 1, xn0, xn1);
    const float4 bm0 = *reinterpret_cast<const float4*>(C.small + P.off_bm + k * 32 + cg * 8),
                 bm1 = *reinterpret_cast<const float4*>(C.small + P.off_bm + k * 32 + cg * 8 + 4),
                 bv0 = *reinterpret_cast<const float4*>(C.small + P.off_bv + k * 32 + cg * 8),
                 bv1 = *reinterpret_cast<const float4*>(C.small + P.off_bv + k * 32 + cg * 8 + 4);
    const float bm[8] = {bm0.x, bm0.y, bm0.z, bm0.w, bm1.x, bm1.y, bm1.z, bm1.w};
    const float bv[8] = {bv0.x, bv0.y, bv0.z, bv0.w, bv1.x, bv1.y, bv1.z, bv1.w};
    const uint32_t tcur = tD_cur + ((uint32_t)((C.r_in_tile >> 5) * 32) << 16);
    uint32_t rm[8], rr[8];
    umma::ld8(tcur + cg * 8, rm);
    umma::ld8(tcur + 32 + cg * 8, rr);
    umma::wait_ld();
    const float xs[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
    const int n_here = P.x_dim - (k * 32 + cg * 8);      // features of this thread's group that exist
    uint32_t mh[8], ml[8], vh[8], vl[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const bool obs = (i < n_here) && (xs[i] == xs[i]);                               // NaN = missing
      const float mu = __uint_as_float(rm[i]) + bm[i];
      const float raw = __uint_as_float(rr[i]) + bv[i];
      const float e = expf(-fabsf(raw));
      const float s2 = (fmaxf(raw, 0.f) + log1pf(e)) + 1e-6f;                          // softplus + eps
      const float inv = rcp_newton(s2);
      const float d = obs ? xs[i] - mu : 0.f;
      const float r1 = rcp_newton(1.f + e);
      const float sig = raw >= 0.f ? r1 : e * r1;
      if (want_lp && obs) loss += (d * d) * (0.5f * inv) + 0.5f * logf(s2);             // bgm/base.py:683-684
      const float dmu = obs ? -d * inv : 0.f;
      const float draw = obs ? (0.5f * inv - (0.5f * d * d) * (inv * inv)) * sig : 0.f;
      umma::split_tf32(dmu, mh[i], ml[i]);
      umma::split_tf32(draw, vh[i], vl[i]);
    }
    if (n_img > 0) ht_wait_mma(C);   // the backward of chunk k-1 has read the old operand: it may be replaced now
    umma::st8(C.trow + HT_B_HI + cg * 8, mh);
    umma::st8(C.trow + HT_B_LO + cg * 8, ml);
    umma::st8(C.trow + HT_B_HI + 32 + cg * 8, vh);
    umma::st8(C.trow + HT_B_LO + 32 + cg * 8, vl);
    x0 = xn0;
    x1 = xn1;
  }
  ht_publish();
  ht_wait_mma(C);
  // ---- d loss / d h_nh -> through the last LeakyReLU -> A ----
  {
    uint32_t bits = 0u;
#pragma unroll
    for (int q = 0; q < HMC_MAXL; ++q)
      if (q == nh - 1) bits = sg[q];
    uint32_t r[16];
    umma::ld16(C.trow + HT_DH + cg * 16, r);
    umma::wait_ld();
    ht_mask16(r, bits, tA_hi, tA_lo);
  }
  // ---- backward hidden layers nh..2: d loss / d h_{l-1} = (d loss / d a_l) W_l^T ----
  float ga[16];
#pragma unroll 1
  for (int l = nh - 1; l >= 1; --l) {
    ht_publish();
    ht_wait_mma(C);
    uint32_t bits = 0u;
#pragma unroll
    for (int q = 0; q < HMC_MAXL; ++q)
      if (q == l - 1) bits = sg[q];
    uint32_t r[16];
    umma::ld16(C.trow + HT_D + cg * 16, r);
    umma::wait_ld();
    if (l > 1) {
      ht_mask16(r, bits, tA_hi, tA_lo);
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float g = __uint_as_float(r[i]);
        ga[i] = ((bits >> i) & 1u) ? g : 0.2f * g;
      }
    }
  }
  // ---- layer 1 transposed on the FMA pipe: partial d loss / d z_in[d] over this thread's 16 columns; the four
  // partials of a row (and of the likelihood loss) meet in shared memory and are added in a fixed order ----
  float* xc = C.xch + (size_t)cg * (ZMAX + 1) * HT_ROWS + C.r_in_tile;
#pragma unroll
  for (int d = 0; d < ZMAX; ++d) {
    if (d < zd) {
      float sacc = 0.f;
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        const float4 w = *reinterpret_cast<const float4*>(W1 + d * 64 + j);
        sacc = fmaf(ga[j], w.x, sacc); sacc = fmaf(ga[j + 1], w.y, sacc);
        sacc = fmaf(ga[j + 2], w.z, sacc); sacc = fmaf(ga[j + 3], w.w, sacc);
      }
      xc[d * HT_ROWS] = sacc;
    }
  }
  xc[ZMAX * HT_ROWS] = loss;
  __syncthreads();
  const float* x0p = C.xch + C.r_in_tile;
  const int strd = (ZMAX + 1) * HT_ROWS;
#pragma unroll
  for (int d = 0; d < ZMAX; ++d)
    gz[d] = d < zd ? (x0p[d * HT_ROWS] + x0p[strd + d * HT_ROWS]) + (x0p[2 * strd + d * HT_ROWS] + x0p[3 * strd + d * HT_ROWS]) : 0.f;
  return (x0p[ZMAX * HT_ROWS] + x0p[strd + ZMAX * HT_ROWS]) + (x0p[2 * strd + ZMAX * HT_ROWS] + x0p[3 * strd + ZMAX * HT_ROWS]);
}

// The control warp's side of one gradient evaluation: the same barriers as ht_eval, the MMA issue and the ring.
__device__ __forceinline__ void ht_eval_ctrl(HtCtx& C) {
  const HmcTcProgram& P = *C.P;
  const int nh = P.nh, NC = P.n_chunks;
  const uint32_t tA_hi = C.tbase + HT_A_HI, tA_lo = C.tbase + HT_A_LO;
  auto plain = [&]() {                                  // D = A * image
    const uint32_t it = C.img_it;
    ht_ctrl_begin(C, 1, [&]() { ht_issue(C.tbase + HT_D, tA_hi, tA_lo, ht_slot_addr(C, it), 0u); });
    ht_ctrl_end(C, 1);
  };
#pragma unroll 1
  for (int l = 1; l < nh; ++l) plain();                 // forward hidden layers
  plain();                                              // head forward of chunk 0
#pragma unroll 1
  for (int k = 0; k < NC; ++k) {
    const uint32_t it = C.img_it;
    const int n_img = (k > 0 ? 1 : 0) + (k + 1 < NC ? 1 : 0);
    const uint32_t tD_next = C.tbase + ((k & 1) ? HT_D : HT_D2);
    ht_ctrl_begin(C, n_img, [&]() {
      uint32_t j = it;
      if (k > 0) {
        ht_issue(C.tbase + HT_DH, C.tbase + HT_B_HI, C.tbase + HT_B_LO, ht_slot_addr(C, j), k > 1 ? 1u : 0u);
        ++j;
      }
      if (k + 1 < NC) ht_issue(tD_next, tA_hi, tA_lo, ht_slot_addr(C, j), 0u);
    });
    ht_ctrl_end(C, n_img);
  }
  {
    const uint32_t it = C.img_it;
    ht_ctrl_begin(C, 1, [&]() { ht_issue(C.tbase + HT_DH, C.tbase + HT_B_HI, C.tbase + HT_B_LO, ht_slot_addr(C, it), NC > 1 ? 1u : 0u); });
    ht_ctrl_end(C, 1);
  }
#pragma unroll 1
  for (int l = nh - 1; l >= 1; --l) plain();            // backward hidden layers
  __syncthreads();                                      // the math warps' exchange barrier
}

// ---- forward only: posterior-predictive draws x = mu + sqrt(sigma^2) N(0,1) (bgm/base.py:517-521) or the heads ----
// Same Philox keys and arithmetic as the SIMT engine's predict mode (hmc.cuh).  The forward-only stream holds the
// hidden layers and the head-forward images; the control warp issues the forward of chunk k+1 while the math warps
// turn chunk k into draws.
template <int ZMAX>
__device__ __forceinline__ void ht_predict(HtCtx& C, const HmcDev& D, const float (&z)[ZMAX], int r, bool rvalid) {
  const HmcTcProgram& P = *C.P;
  const int zd = P.zd, nh = P.nh, NC = P.n_chunks;
  const int cg = C.cg;
  const float* W1 = C.small + P.off_W1 + cg * 16;
  const uint32_t tA_hi = C.trow + HT_A_HI + cg * 16, tA_lo = C.trow + HT_A_LO + cg * 16;
  {
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = C.small[P.off_b1 + cg * 16 + i];
#pragma unroll
    for (int d = 0; d < ZMAX; ++d) {
      if (d < zd) {
        const float zv = (z[d] - P.bn_mean[d]) * P.bn_inv[d] + P.bn_beta[d];
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float4 w = *reinterpret_cast<const float4*>(W1 + d * 64 + i);
          a[i] = fmaf(zv, w.x, a[i]); a[i + 1] = fmaf(zv, w.y, a[i + 1]);
          a[i + 2] = fmaf(zv, w.z, a[i + 2]); a[i + 3] = fmaf(zv, w.w, a[i + 3]);
        }
      }
    }
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float v = a[i] > 0.f ? a[i] : 0.2f * a[i];
      umma::split_tf32(v, hi[i], lo[i]);
    }
    umma::st16(tA_hi, hi);
    umma::st16(tA_lo, lo);
  }
#pragma unroll 1
  for (int l = 1; l < nh; ++l) {
    ht_publish();
    ht_wait_mma(C);
    uint32_t rr[16], bits;
    umma::ld16(C.trow + HT_D + cg * 16, rr);
    umma::wait_ld();
    ht_act16(rr, C.small + P.off_bh + (l - 1) * 64 + cg * 16, bits, tA_hi, tA_lo);
  }
  ht_publish();
  ht_wait_mma(C);                                   // head forward of chunk 0
  const int s_idx = D.sample0 + r / D.n_per_sample;
  const int64_t grow = D.a.row_offset + (r - (r / D.n_per_sample) * D.n_per_sample);
#pragma unroll 1
  for (int k = 0; k < NC; ++k) {
    const uint32_t tD_cur = C.tbase + ((k & 1) ? HT_D2 : HT_D);
    ht_publish();                                    // everybody is done with the accumulator chunk k+1 goes into
    const int c0 = k * 32 + cg * 8;
    const float4 bm0 = *reinterpret_cast<const float4*>(C.small + P.off_bm + c0), bm1 = *reinterpret_cast<const float4*>(C.small + P.off_bm + c0 + 4),
                 bv0 = *reinterpret_cast<const float4*>(C.small + P.off_bv + c0), bv1 = *reinterpret_cast<const float4*>(C.small + P.off_bv + c0 + 4);
    const float bm[8] = {bm0.x, bm0.y, bm0.z, bm0.w, bm1.x, bm1.y, bm1.z, bm1.w};
    const float bv[8] = {bv0.x, bv0.y, bv0.z, bv0.w, bv1.x, bv1.y, bv1.z, bv1.w};
    const uint32_t tcur = tD_cur + ((uint32_t)((C.r_in_tile >> 5) * 32) << 16);
    uint32_t rm[8], rw[8];
    umma::ld8(tcur + cg * 8, rm);
    umma::ld8(tcur + 32 + cg * 8, rw);
    umma::wait_ld();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = c0 + h * 4;
      float e4[4] = {0.f, 0.f, 0.f, 0.f};
      if (!D.out_var) {
        if (D.noise_x) {
#pragma unroll
          for (int q = 0; q < 4; ++q) e4[q] = (c + q < P.x_dim) ? D.noise_x[(size_t)r * P.x_dim + c + q] : 0.f;
        } else {
          normal4(D.a.seed, grow, (uint32_t)s_idx, NOISE_PREDICT, (uint32_t)(c >> 2), e4);
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int i = h * 4 + q;
        const float mu = __uint_as_float(rm[i]) + bm[i];
        const float s2 = softplus_f(__uint_as_float(rw[i]) + bv[i]) + 1e-6f;
        if (rvalid && c + q < P.x_dim) {
          if (D.out_var) {
            D.out_x[(size_t)r * P.x_dim + c + q] = mu;
            D.out_var[(size_t)r * P.x_dim + c + q] = s2;
          } else {
            D.out_x[(size_t)r * P.x_dim + c + q] = fmaf(e4[q], sqrtf(s2), mu);
          }
        }
      }
    }
    if (k + 1 < NC) ht_wait_mma(C);
  }
}
__device__ __forceinline__ void ht_predict_ctrl(HtCtx& C) {
  const HmcTcProgram& P = *C.P;
  const int nh = P.nh, NC = P.n_chunks;
  const uint32_t tA_hi = C.tbase + HT_A_HI, tA_lo = C.tbase + HT_A_LO;
  auto plain = [&]() {
    const uint32_t it = C.img_it;
    ht_ctrl_begin(C, 1, [&]() { ht_issue(C.tbase + HT_D, tA_hi, tA_lo, ht_slot_addr(C, it), 0u); });
    ht_ctrl_end(C, 1);
  };
#pragma unroll 1
  for (int l = 1; l < nh; ++l) plain();
  plain();
#pragma unroll 1
  for (int k = 0; k < NC; ++k) {
    const uint32_t it = C.img_it;
    const int n_img = k + 1 < NC ? 1 : 0;
    const uint32_t tD_next = C.tbase + ((k & 1) ? HT_D : HT_D2);
    ht_ctrl_begin(C, n_img, [&]() { ht_issue(tD_next, tA_hi, tA_lo, ht_slot_addr(C, it), 0u); });
    ht_ctrl_end(C, n_img);
  }
}

template <int ZMAX>
__global__ void __launch_bounds__(HT_THREADS + 32, 1)
hmc_tc_kernel(const __grid_constant__ HmcTcProgram P, const float* __restrict__ stream, const float* __restrict__ stream_fwd,
              const float* __restrict__ small_g, const __grid_constant__ HmcDev D) {
  extern __shared__ __align__(128) float smem[];
  __shared__ uint64_t full[HT_SLOTS];
  __shared__ uint64_t mma_bar_s;
  __shared__ uint32_t tmem_slot;
  const bgm_hmc_args& A = D.a;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int n_rows = A.n, zd = P.zd;
  const int nblocks = (n_rows + HT_ROWS - 1) / HT_ROWS;
  const int L = A.num_leapfrog;
  const bool need_init = D.mode != HMC_RUN || A.init_mode != 0;
  const int steps = D.mode == HMC_RUN ? A.t_end - A.t_begin : 0;
  const int evals_per_block = (need_init ? 1 : 0) + steps * L;
  int my_blocks = 0;
  for (int b = blockIdx.x; b < nblocks; b += gridDim.x) ++my_blocks;

  HtCtx C;
  C.P = &P;
  const bool predict = D.mode == HMC_PREDICT;
  C.stream = predict ? stream_fwd : stream;
  C.n_img = (uint32_t)(predict ? P.n_img_fwd : P.n_img);
  C.ring = smem;
  float* small_s = smem + HT_SLOTS * HT_IMG_FLOATS;
  C.small = small_s;
  C.xch = small_s + P.small_floats;
  C.cg = warp >> 2;
  C.r_in_tile = (warp & 3) * 32 + lane;
  C.full = full;
  C.mma_bar = umma::smem_addr(&mma_bar_s);
  C.img_it = 0;
  C.img_total = (uint32_t)my_blocks * (uint32_t)evals_per_block * C.n_img;
  C.mma_parity = 0;
  C.issuer_warp = warp == 16;
  if (tid == HT_THREADS) {                     // lane 0 of the control warp owns the ring
    for (int s = 0; s < HT_SLOTS; ++s) mbar_init(full + s, 1);
    umma::mbar_init(C.mma_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int s = 0; s < HT_SLOTS; ++s)
      if ((uint32_t)s < C.img_total) ht_fill(C, (uint32_t)s, s);
  }
  for (int i = tid; i < P.small_floats; i += HT_THREADS + 32) small_s[i] = small_g[i];
  if (warp == 0) umma::tmem_alloc512(&tmem_slot);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  C.tbase = __shfl_sync(0xffffffffu, tmem_slot, 0);
  C.trow = C.tbase + ((uint32_t)((warp & 3) * 32) << 16);
  const float eps = (D.mode == HMC_RUN && A.step_dev) ? *A.step_dev : 0.f;

  if (C.issuer_warp) {                         // control warp: one ht_eval_ctrl per evaluation of the math warps
    for (int b = blockIdx.x; b < nblocks; b += gridDim.x)
      for (int e = 0; e < evals_per_block; ++e) {
        if (predict) ht_predict_ctrl(C);
        else ht_eval_ctrl(C);
      }
  } else if (predict) {
    for (int b = blockIdx.x; b < nblocks; b += gridDim.x) {
      const int row = b * HT_ROWS + C.r_in_tile;
      const bool rvalid = row < n_rows;
      const int lrow = rvalid ? row : n_rows - 1;
      float z[ZMAX];
#pragma unroll
      for (int d = 0; d < ZMAX; ++d) z[d] = d < zd ? D.z_in[(size_t)lrow * zd + d] : 0.f;
      ht_predict<ZMAX>(C, D, z, lrow, rvalid);
    }
  } else
  for (int b = blockIdx.x; b < nblocks; b += gridDim.x) {
    const int row = b * HT_ROWS + C.r_in_tile;
    const bool owner = C.cg == 0;               // one of the row's four threads owns the outputs
    const bool valid = row < n_rows && owner;
    const bool in_range = row < n_rows;
    const int lrow = in_range ? row : n_rows - 1;
    const int64_t grow = A.row_offset + lrow;
    float z[ZMAX], p[ZMAX], g[ZMAX], gz[ZMAX];
#pragma unroll
    for (int d = 0; d < ZMAX; ++d) { z[d] = 0.f; p[d] = 0.f; g[d] = 0.f; }

    // log posterior and gradient from an evaluation's outputs (bgm/base.py:702-704)
    auto finish = [&](float loss) -> float {
      float prior = 0.f;
#pragma unroll
      for (int d = 0; d < ZMAX; ++d)
        if (d < zd) {
          prior = fmaf(z[d], z[d], prior);
          g[d] = -(gz[d] * P.bn_inv[d] + z[d]);
        }
      return -(0.5f * prior + loss);
    };

    if (D.mode != HMC_RUN) {   // EVAL: log p and gradient at z_in
#pragma unroll
      for (int d = 0; d < ZMAX; ++d)
        if (d < zd) z[d] = D.z_in[(size_t)lrow * zd + d];
      const float loss = ht_eval<ZMAX>(C, D, z, lrow, true, gz);
      const float lp = finish(loss);
      if (valid) {
        A.lp_state_dev[row] = lp;
        if (D.out_grad)
#pragma unroll
          for (int d = 0; d < ZMAX; ++d)
            if (d < zd) D.out_grad[(size_t)row * zd + d] = g[d];
      }
      continue;
    }

    // ---- chain state (bgm/base.py:778) ----
    if (A.init_mode == 2) {
#pragma unroll
      for (int gi = 0; gi < ZMAX / 4; ++gi) {
        if (gi * 4 < zd) {
          float e[4];
          normal4(A.seed, grow, T_INIT, NOISE_MOMENTUM, gi, e);
#pragma unroll
          for (int q = 0; q < 4; ++q) z[gi * 4 + q] = e[q];
        }
      }
    } else {
#pragma unroll
      for (int d = 0; d < ZMAX; ++d)
        if (d < zd) z[d] = A.z_state_dev[(size_t)lrow * zd + d];
    }
    float lp_cur;
    if (need_init) {
      const float loss = ht_eval<ZMAX>(C, D, z, lrow, true, gz);
      lp_cur = finish(loss);
      if (valid) {
#pragma unroll
        for (int d = 0; d < ZMAX; ++d)
          if (d < zd) {
            A.z_state_dev[(size_t)row * zd + d] = z[d];
            A.g_state_dev[(size_t)row * zd + d] = g[d];
          }
      }
    } else {
      lp_cur = A.lp_state_dev[lrow];
#pragma unroll
      for (int d = 0; d < ZMAX; ++d)
        if (d < zd) g[d] = A.g_state_dev[(size_t)lrow * zd + d];
    }

    // ---- HMC steps (TFP 0.18 HamiltonianMonteCarlo, unit mass) ----
#pragma unroll 1
    for (int t = A.t_begin; t < A.t_end; ++t) {
      float ke0 = 0.f;
#pragma unroll
      for (int gi = 0; gi < ZMAX / 4; ++gi) {
        if (gi * 4 < zd) {
          float e[4];
          if (A.mom_dev) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              e[q] = (gi * 4 + q < zd) ? A.mom_dev[((size_t)t * A.n + lrow) * zd + gi * 4 + q] : 0.f;
          } else {
            normal4(A.seed, grow, (uint32_t)t, NOISE_MOMENTUM, gi, e);
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int d = gi * 4 + q;
            if (d < zd) {
              ke0 = fmaf(e[q], e[q], ke0);
              p[d] = e[q] + (0.5f * eps) * g[d];
            }
          }
        }
      }
      ke0 *= 0.5f;
      float lp_new = lp_cur;
#pragma unroll 1
      for (int l = 0; l < L; ++l) {
#pragma unroll
        for (int d = 0; d < ZMAX; ++d)
          if (d < zd) z[d] += eps * p[d];
        const bool last = l == L - 1;
        const float loss = ht_eval<ZMAX>(C, D, z, lrow, last, gz);
        const float lp = finish(loss);
        if (last) lp_new = lp;
#pragma unroll
        for (int d = 0; d < ZMAX; ++d)
          if (d < zd) p[d] += eps * g[d];
      }
      float ke1 = 0.f;
#pragma unroll
      for (int d = 0; d < ZMAX; ++d)
        if (d < zd) {
          const float pp = p[d] - (0.5f * eps) * g[d];
          ke1 = fmaf(pp, pp, ke1);
        }
      ke1 *= 0.5f;
      float log_accept = (lp_new - lp_cur) + (ke0 - ke1);
      if (!(fabsf(log_accept) <= 3.0e38f)) log_accept = -INFINITY;   // NaN / inf energy: reject
      float logu;
      if (A.logu_dev) logu = A.logu_dev[(size_t)t * A.n + lrow];
      else logu = logf(u01_open1(noise_block(A.seed, grow, (uint32_t)t, NOISE_ACCEPT, 0).x));
      const bool acc = logu < log_accept;
      if (acc) {
        lp_cur = lp_new;
        if (valid)
#pragma unroll
          for (int d = 0; d < ZMAX; ++d)
            if (d < zd) {
              A.z_state_dev[(size_t)row * zd + d] = z[d];
              A.g_state_dev[(size_t)row * zd + d] = g[d];
            }
      } else {
#pragma unroll
        for (int d = 0; d < ZMAX; ++d)
          if (d < zd) {
            z[d] = A.z_state_dev[(size_t)lrow * zd + d];
            g[d] = A.g_state_dev[(size_t)lrow * zd + d];
          }
      }
      if (A.accept_mask_dev && valid) A.accept_mask_dev[(size_t)t * A.n + row] = acc ? 1 : 0;
      if (A.log_accept_dev && valid) A.log_accept_dev[(size_t)t * A.n + row] = log_accept;
      if (A.accept_count_dev) {
        const unsigned bal = __ballot_sync(0xffffffffu, acc && valid);
        if (lane == 0 && bal) atomicAdd(A.accept_count_dev + t, __popc(bal));
      }
      if (A.accept_stat_dev) {   // SimpleStepSizeAdaptation: mean over ALL chains of exp(min(log_accept, 0))
        float a = valid ? expf(fminf(log_accept, 0.f)) : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) atomicAdd(A.accept_stat_dev + t, (double)a);
      }
      if (t >= A.burn_in && A.out_samples_dev && valid) {
        float* dst = A.out_samples_dev + ((size_t)(t - A.burn_in) * A.n + row) * zd;
#pragma unroll
        for (int d = 0; d < ZMAX; ++d)
          if (d < zd) dst[d] = z[d];
      }
    }
    if (valid) A.lp_state_dev[row] = lp_cur;
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc512(C.tbase);
}

}  // namespace bgm
