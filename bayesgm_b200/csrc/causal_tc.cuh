// CausalBGM random-walk MH sampler, tensor-core engine (tcgen05 / TMEM).
//
// Same contract as causal_mh_kernel (causal.cuh): causalbgm/base.py:765-817
// (get_log_posterior) inside :820-904 (metropolis_hastings_sampler), same Philox streams,
// same scheduling, same arguments -- a different execution plan for the standard net shape
// (g: zd -> 64 x (L-1) -> v_dim+1 with the covariate likelihood projected to 64 columns,
// f / h: in -> 64 -> 32 -> 8 -> 2):
//
//  * a warpgroup (4 warps, 128 threads) owns a 128-row tile; THREAD = ROW = TMEM LANE.  The
//    chain state, the proposal, the per-row losses and the accept decision never leave the
//    thread; there is no shuffle and no shared-memory activation buffer.
//  * the 64x64 layers of g_net (hidden layers 2.. and the projected output layer) run on the
//    tensor cores as 3xTF32 (umma.cuh): the thread writes its activation row, split into
//    hi/lo, into TMEM (tcgen05.st), one elected thread issues 24 tcgen05.mma (A from TMEM,
//    weights from the shared-memory image), the accumulator row comes back with tcgen05.ld
//    for bias + LeakyReLU + split.  fp32-level error.
//  * f_net / h_net: their 64->32 layers and (as one block-diagonal 64->16 product) their 32->8
//    layers run the same way; only the first layers (a handful of inputs) and the 8->2 heads
//    stay on the fp32 FMA pipe, one row per thread with warp-uniform (broadcast) weight loads,
//    scheduled while the thread's MMAs are in flight.  Per iteration: 3 + (g.L-1) MMA stages.
//  * two warpgroups per CTA (one CTA per SM, 512 TMEM columns = 2 x [A_hi 64 | A_lo 64 | D 64 | t 64])
//    work on different tiles, so one tile's epilogue overlaps the other's MMAs.
#pragma once
#include "causal.cuh"
#include "umma.cuh"

namespace bgm {

constexpr int TC_ROWS = 128;
constexpr int TC_MAX_MMA = 8;

// TMEM columns of one warpgroup (256 of the CTA's 512)
constexpr uint32_t TC_A_HI = 0, TC_A_LO = 64, TC_D = 128, TC_P = 192;

struct TcProgram {
  int enabled;
  int zd, p, binary;
  float s2v, s2x, s2y;
  int n_mma;                                  // g_net layers on the tensor cores (g.L - 1)
  int w_hi[TC_MAX_MMA], w_lo[TC_MAX_MMA];     // g: [16][64][4] tf32 images (hi / lo), float offsets
  int gb[TC_MAX_MMA];                         // bias added after g MMA layer m (m < n_mma-1)
  int gW1, gb1;                               // g first layer [zd][64], bias[64]
  int wsig, bsig;                             // sigma_v head: column v_dim of g's last layer
  // f / h: in -> 64 -> 32 -> 8 -> 2
  int fW1, fb1, hW1, hb1;                     // first layers: the rows of [z.., x] each net uses, in order: [popc(mask)][64]
  unsigned long long fmask, hmask;            // which entries of the input vector [z.., x] those rows multiply
  int f2_hi, f2_lo, h2_hi, h2_lo;             // second layers, [16][32][4] images
  int fb2, hb2;
  int w3_hi, w3_lo, b3;                       // third layers of f and h as one block-diagonal
                                              // [f_h2 | h_h2] (64) -> [f 8 | h 8]: [16][16][4]
  int fW4, fb4, hW4, hb4;                     // [8][2], [2]
  int f3_hi, f3_lo, fb3;                      // effect kernel: f third layer alone, [8][16][4] (cols 8.. zero)
  int image_floats;
  // causal_mh_tc16_kernel<ZMAX, true> only (its own image): the first layers of f and h as ONE tensor-core product
  // [z.., x, 0.., 1] (l1_k = 8 inputs, the ones column last) -> [f_h1 (64) | h_h1 (64)], biases in the row of the ones column,
  // structural zeros where a net does not read an input: [l1_k/4][128][4] images.  l1_k == 0: not packed.
  int l1_k, l1_hi, l1_lo;
};

__device__ __forceinline__ void wg_sync(int wg) {
  if (wg == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
  else asm volatile("bar.sync 2, 128;" ::: "memory");
}
// Publishing a stage: only the warp that issues the MMAs has to WAIT for the other warps' TMEM
// writes; the others only announce theirs (bar.arrive) and go on to whatever does not depend on
// the MMAs.  A warp cannot run a whole stage ahead: its next step waits on the MMA mbarrier, which
// completes only after the issuer has passed this barrier.
__device__ __forceinline__ void wg_publish(int wg, bool issuer_warp) {
  if (issuer_warp) {
    if (wg == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
    else asm volatile("bar.sync 2, 128;" ::: "memory");
  } else {
    if (wg == 0) asm volatile("bar.arrive 1, 128;" ::: "memory");
    else asm volatile("bar.arrive 2, 128;" ::: "memory");
  }
}
__device__ __forceinline__ void tile_publish256(int slot, bool issuer_warp) {
  if (issuer_warp) {
    if (slot == 0) asm volatile("bar.sync 1, 256;" ::: "memory");
    else asm volatile("bar.sync 2, 256;" ::: "memory");
  } else {
    if (slot == 0) asm volatile("bar.arrive 1, 256;" ::: "memory");
    else asm volatile("bar.arrive 2, 256;" ::: "memory");
  }
}
__device__ __forceinline__ float leaky_mx(float v) { return fmaxf(v, 0.2f * v); }  // == leaky(v)
__device__ __forceinline__ float2 leaky_x2(float2 v) {
  const float2 m = __fmul2_rn(v, make_float2(0.2f, 0.2f));
  return make_float2(fmaxf(v.x, m.x), fmaxf(v.y, m.y));
}

// First layer of a net on the FMA pipe for the thread's row: out = LeakyReLU(b + in W), 64 wide,
// rows of W (inputs) taken in ascending order like the SIMT engine; `mask` skips the rows
// whose weights are structurally zero for this net.
template <int KINMAX>
__device__ __forceinline__ void first_layer(const float* __restrict__ W, const float* __restrict__ b,
                                            unsigned long long mask, int nin, const float (&in)[KINMAX],
                                            float (&out)[64]) {
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    const float4 bb = *reinterpret_cast<const float4*>(b + q * 4);
    out[q * 4 + 0] = bb.x; out[q * 4 + 1] = bb.y; out[q * 4 + 2] = bb.z; out[q * 4 + 3] = bb.w;
  }
  int wr = 0;   // packed row of W
#pragma unroll
  for (int d = 0; d < KINMAX; ++d) {
    if (d < nin && ((mask >> d) & 1ull)) {
      const float zv = in[d];
      const float* Wd = W + wr * 64;
      ++wr;
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const float4 ww = *reinterpret_cast<const float4*>(Wd + q * 4);
        out[q * 4 + 0] = fmaf(zv, ww.x, out[q * 4 + 0]);
        out[q * 4 + 1] = fmaf(zv, ww.y, out[q * 4 + 1]);
        out[q * 4 + 2] = fmaf(zv, ww.z, out[q * 4 + 2]);
        out[q * 4 + 3] = fmaf(zv, ww.w, out[q * 4 + 3]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 64; ++j) out[j] = leaky_mx(out[j]);
}

// activation row (fp32, 64 wide) -> hi / lo halves in the A slots of the thread's TMEM lane
__device__ __forceinline__ void split_store64(const float (&a)[64], uint32_t tA_hi, uint32_t tA_lo) {
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    uint32_t hi[32], lo[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) umma::split_tf32(a[c * 32 + j], hi[j], lo[j]);
    umma::st32(tA_hi + c * 32, hi);
    umma::st32(tA_lo + c * 32, lo);
  }
}

// r (32 raw accumulator columns) -> LeakyReLU(r + bias) as floats
__device__ __forceinline__ void bias_act32(const uint32_t (&r)[32], const float* __restrict__ bias, float* out) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 b = *reinterpret_cast<const float4*>(bias + q * 4);
    out[q * 4 + 0] = leaky_mx(__uint_as_float(r[q * 4 + 0]) + b.x);
    out[q * 4 + 1] = leaky_mx(__uint_as_float(r[q * 4 + 1]) + b.y);
    out[q * 4 + 2] = leaky_mx(__uint_as_float(r[q * 4 + 2]) + b.z);
    out[q * 4 + 3] = leaky_mx(__uint_as_float(r[q * 4 + 3]) + b.w);
  }
}

// bias + LeakyReLU + hi/lo split of 32 accumulator columns, back into the A slots.
template <bool SIG>
__device__ __forceinline__ void act_block32(uint32_t (&r)[32], const float* __restrict__ bias,
                                            const float* __restrict__ wsig, float& sig, uint32_t tA_hi,
                                            uint32_t tA_lo) {
  uint32_t lo[32];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 b = *reinterpret_cast<const float4*>(bias + q * 4);
    const float bb[4] = {b.x, b.y, b.z, b.w};
    float ws[4] = {0.f, 0.f, 0.f, 0.f};
    if constexpr (SIG) {
      const float4 s4 = *reinterpret_cast<const float4*>(wsig + q * 4);
      ws[0] = s4.x; ws[1] = s4.y; ws[2] = s4.z; ws[3] = s4.w;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float a = leaky_mx(__uint_as_float(r[q * 4 + i]) + bb[i]);
      if constexpr (SIG) sig = fmaf(a, ws[i], sig);
      umma::split_tf32(a, r[q * 4 + i], lo[q * 4 + i]);
    }
  }
  umma::st32(tA_hi, r);
  umma::st32(tA_lo, lo);
}

template <int ZMAX>
__global__ void __launch_bounds__(256, 1)
causal_mh_tc_kernel(const __grid_constant__ TcProgram P, const float* __restrict__ image,
                    const __grid_constant__ MhDev D) {
  constexpr int KINMAX = ZMAX + 1;
  extern __shared__ __align__(128) float smem[];
  __shared__ uint64_t bar_img;
  __shared__ uint64_t bar_mma[2];
  __shared__ uint32_t tmem_slot;
  __shared__ long long unit_s[2];
  bulk_load_to_smem(smem, image, (uint32_t)P.image_floats * 4u, &bar_img);
  const float* wimg = smem;
  const uint32_t wimg_s = umma::smem_addr(smem);

  const bgm_mh_args& A = D.a;
  // warp-uniform by construction (shfl), so that role branches and MMA operands are uniform
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int wg = warp >> 2, wtid = tid & 127;
  const bool issuer_warp = (warp & 3) == 3;   // highest warp id of the group: the arbiter favours it
  if (tid == 0) {
    umma::mbar_init(umma::smem_addr(&bar_mma[0]), 1);
    umma::mbar_init(umma::smem_addr(&bar_mma[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) umma::tmem_alloc512(&tmem_slot);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem0 = __shfl_sync(0xffffffffu, tmem_slot, 0);
  const uint32_t tbase = tmem0 + (uint32_t)wg * 256u;                  // MMA operand addresses (lane 0)
  const uint32_t trow = tbase + ((uint32_t)((warp & 3) * 32) << 16);   // this warp's 32 lanes
  const uint32_t bar = umma::smem_addr(&bar_mma[wg]);
  uint32_t parity = 0;

  const int n = A.n, zd = P.zd;
  const int ntiles = (n + TC_ROWS - 1) / TC_ROWS;
  const double q_sd = A.q_sd_dev ? *A.q_sd_dev : 1.0;
  const bool need_init = !(A.init_mode == 0 && D.mode == 0);
  const int t_first = need_init ? A.t_begin - 1 : A.t_begin;
  const int t_last = D.mode == 1 ? A.t_begin : A.t_end;
  int* sched = A.sched_dev;
  const int n_iter = t_last - t_first;
  const int nchunks = D.nchunks;
  const int chunk_len = (n_iter + nchunks - 1) / nchunks;
  const long long total_units = (long long)ntiles * nchunks;
  const int n_mma = P.n_mma;

  // One MMA stage: publish the A rows written by the 128 threads; one thread issues + commits.
  auto publish = [&]() {
    umma::wait_st();
    umma::fence_before_sync();
    wg_publish(wg, issuer_warp);
  };
  auto stage_wait = [&]() {
    umma::mbar_wait(bar, parity);
    parity ^= 1u;
    umma::fence_after_sync();
  };

  for (;;) {
    if (wtid == 0) {
      const long long u = atomicAdd(reinterpret_cast<unsigned int*>(sched), 1u);
      if (u < total_units) {
        const int chunk = (int)(u / ntiles);
        const int tile = (int)(u - (long long)chunk * ntiles);
        if (chunk > 0) {
          const volatile int* flag = sched + 1 + tile;
          while (*flag < chunk) __nanosleep(200);
          __threadfence();
        }
      }
      unit_s[wg] = u;
    }
    wg_sync(wg);
    const long long u = unit_s[wg];
    if (u >= total_units) break;
    const int chunk = (int)(u / ntiles);
    const int tile = (int)(u - (long long)chunk * ntiles);
    const int ta = t_first + chunk * chunk_len;
    const int tb = min(ta + chunk_len, t_last);
    const int row = tile * TC_ROWS + wtid;
    const bool valid = row < n;
    const int lrow = valid ? row : n - 1;
    const float x_l = A.x_dev[lrow], y_l = A.y_dev[lrow];
    const float r0_l = A.r0_dev[lrow];
    const int64_t grow = A.row_offset + lrow;
    // the row's projected covariates t (64 floats) stay in the thread's TMEM lane (columns
    // TC_P..) for the whole unit: one uncoalesced read per ~16 iterations instead of one each
    {
      const float4* tg = reinterpret_cast<const float4*>(A.vproj_dev + (size_t)lrow * A.ldvproj);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t tr[32];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 v4 = __ldg(tg + c * 8 + q);
          tr[q * 4 + 0] = __float_as_uint(v4.x); tr[q * 4 + 1] = __float_as_uint(v4.y);
          tr[q * 4 + 2] = __float_as_uint(v4.z); tr[q * 4 + 3] = __float_as_uint(v4.w);
        }
        umma::st32(trow + TC_P + c * 32, tr);
      }
      umma::wait_st();
    }
    float zc[ZMAX];
    // ---- initial state (:842) ----
    if (chunk == 0 && A.init_mode == 2) {
#pragma unroll
      for (int g = 0; g < ZMAX / 4; ++g) {
        if (g * 4 < zd) {
          float e[4];
          normal4(A.seed, grow, T_INIT, NOISE_PROPOSAL, g, e);
#pragma unroll
          for (int q = 0; q < 4; ++q) zc[g * 4 + q] = e[q];
        }
      }
    } else {
#pragma unroll
      for (int d = 0; d < ZMAX; ++d) zc[d] = (d < zd) ? __ldcg(A.z_state_dev + (size_t)lrow * zd + d) : 0.f;
    }
    float lp_cur = (need_init && chunk == 0) ? 0.f : __ldcg(A.lp_state_dev + lrow);
    // unit normals of the current iteration's proposal (Philox mode): drawn here for the first
    // iteration of the unit, afterwards under the previous iteration's MMA waits
    float en[ZMAX];
    float u_acc = 0.f;
#pragma unroll
    for (int d = 0; d < ZMAX; ++d) en[d] = 0.f;
    if (!A.eps_dev && ta < tb) {
#pragma unroll
      for (int g = 0; g < ZMAX / 4; ++g) {
        if (g * 4 < zd) {
          float e[4];
          normal4(A.seed, grow, (uint32_t)ta, NOISE_PROPOSAL, g, e);
#pragma unroll
          for (int q = 0; q < 4; ++q) en[g * 4 + q] = e[q];
        }
      }
    }
    // ---- iterations (:860-898) ----
#pragma unroll 1
    for (int t = ta; t < tb; ++t) {
      const bool init_pass = t < A.t_begin;
      float in[KINMAX];   // [proposal z' (zd), x]
#pragma unroll
      for (int d = 0; d < KINMAX; ++d) in[d] = 0.f;
      if (init_pass) {
#pragma unroll
        for (int d = 0; d < ZMAX; ++d)
          if (d < zd) in[d] = zc[d];
      } else if (A.eps_dev) {
        // z' = z + float32(q_sd * eps) (:862): float64 scale, one rounding, then the fp32 add
        const float* e = A.eps_dev + ((size_t)t * n + lrow) * zd;
#pragma unroll
        for (int d = 0; d < ZMAX; ++d)
          if (d < zd) in[d] = __fadd_rn(zc[d], (float)(q_sd * (double)e[d]));
      } else {
#pragma unroll
        for (int d = 0; d < ZMAX; ++d)
          if (d < zd) in[d] = __fadd_rn(zc[d], (float)(q_sd * (double)en[d]));
      }
      // prior on the proposal (:812), before x joins the input vector
      float prior = 0.f;
#pragma unroll
      for (int d = 0; d < ZMAX; ++d)
        if (d < zd) prior = fmaf(in[d], in[d], prior);
      prior *= 0.5f;
#pragma unroll
      for (int d = 0; d < KINMAX; ++d)
        if (d == zd) in[d] = x_l;

      // ---- f_net layer 1 (FMA pipe) -> A; stage F2: f layer 2 on the tensor cores ----
      {
        float a1[64];
        first_layer<KINMAX>(wimg + P.fW1, wimg + P.fb1, P.fmask, zd + 1, in, a1);
        split_store64(a1, trow + TC_A_HI, trow + TC_A_LO);
      }
      publish();
      if (issuer_warp) {
        if (umma::elect_one()) {
          umma::fence_after_sync();
          umma::issue_layer_k64<32>(tbase + TC_D, tbase + TC_A_HI, tbase + TC_A_LO, wimg_s + 4u * (uint32_t)P.f2_hi,
                                    wimg_s + 4u * (uint32_t)P.f2_lo);
          umma::mma_commit(bar);
        }
        __syncwarp();
      }
      float h2[64];   // [f_h2 (32) | h_h2 (32)], post-activation
      {
        // h_net layer 1 while f's MMAs run
        float a1[64];
        first_layer<KINMAX>(wimg + P.hW1, wimg + P.hb1, P.hmask, zd + 1, in, a1);
        stage_wait();
        uint32_t r[32];
        umma::ld32(trow + TC_D, r);
        umma::wait_ld();
        bias_act32(r, wimg + P.fb2, h2);
        split_store64(a1, trow + TC_A_HI, trow + TC_A_LO);
      }
      publish();
      if (issuer_warp) {
        if (umma::elect_one()) {
          umma::fence_after_sync();
          umma::issue_layer_k64<32>(tbase + TC_D, tbase + TC_A_HI, tbase + TC_A_LO, wimg_s + 4u * (uint32_t)P.h2_hi,
                                    wimg_s + 4u * (uint32_t)P.h2_lo);
          umma::mma_commit(bar);
        }
        __syncwarp();
      }
      stage_wait();
      {
        uint32_t r[32];
        umma::ld32(trow + TC_D, r);
        umma::wait_ld();
        bias_act32(r, wimg + P.hb2, h2 + 32);
        split_store64(h2, trow + TC_A_HI, trow + TC_A_LO);
      }
      publish();
      if (issuer_warp) {
        if (umma::elect_one()) {
          umma::fence_after_sync();
          umma::issue_layer_k64<16>(tbase + TC_D, tbase + TC_A_HI, tbase + TC_A_LO, wimg_s + 4u * (uint32_t)P.w3_hi,
                                    wimg_s + 4u * (uint32_t)P.w3_lo);
          umma::mma_commit(bar);
        }
        __syncwarp();
      }
      float g1[64];   // g_net layer 1 while the layer-3 MMAs run
      first_layer<KINMAX>(wimg + P.gW1, wimg + P.gb1, ~0ull, zd, in, g1);
      stage_wait();
      float loss_py, loss_px;
      {
        uint32_t r[16];
        umma::ld16(trow + TC_D, r);
        umma::wait_ld();
        float h3[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 b = *reinterpret_cast<const float4*>(wimg + P.b3 + q * 4);
          h3[q * 4 + 0] = leaky_mx(__uint_as_float(r[q * 4 + 0]) + b.x);
          h3[q * 4 + 1] = leaky_mx(__uint_as_float(r[q * 4 + 1]) + b.y);
          h3[q * 4 + 2] = leaky_mx(__uint_as_float(r[q * 4 + 2]) + b.z);
          h3[q * 4 + 3] = leaky_mx(__uint_as_float(r[q * 4 + 3]) + b.w);
        }
        float mu_y = wimg[P.fb4], raw_y = wimg[P.fb4 + 1], mu_x = wimg[P.hb4], raw_x = wimg[P.hb4 + 1];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float2 wf = *reinterpret_cast<const float2*>(wimg + P.fW4 + k * 2);
          const float2 wh = *reinterpret_cast<const float2*>(wimg + P.hW4 + k * 2);
          mu_y = fmaf(h3[k], wf.x, mu_y);
          raw_y = fmaf(h3[k], wf.y, raw_y);
          mu_x = fmaf(h3[8 + k], wh.x, mu_x);
          raw_x = fmaf(h3[8 + k], wh.y, raw_x);
        }
        // outcome model (:809-810), treatment model (:803-807)
        const float s2y = P.s2y >= 0.f ? P.s2y : softplus_f(raw_y) + 1e-6f;
        const float dy = y_l - mu_y;
        loss_py = (dy * dy) / (2.f * s2y) + logf(s2y) / 2.f;
        if (P.binary) {
          loss_px = fmaxf(mu_x, 0.f) - mu_x * x_l + log1pf(expf(-fabsf(mu_x)));
        } else {
          const float s2x = P.s2x >= 0.f ? P.s2x : softplus_f(raw_x) + 1e-6f;
          const float dx = x_l - mu_x;
          loss_px = (dx * dx) / (2.f * s2x) + logf(s2x) / 2.f;
        }
      }
      // ---- g_net: layers 2.. and the projected output layer on the tensor cores ----
      split_store64(g1, trow + TC_A_HI, trow + TC_A_LO);
      float sig = wimg[P.bsig];
#pragma unroll 1
      for (int m = 0; m < n_mma; ++m) {
        if (m > 0) {
          // epilogue of MMA layer m-1 (a hidden layer of g_net)
          uint32_t r0[32], r1[32];
          umma::ld32(trow + TC_D, r0);
          umma::ld32(trow + TC_D + 32, r1);
          umma::wait_ld();
          const float* bias = wimg + P.gb[m - 1];
          if (m == n_mma - 1) {
            act_block32<true>(r0, bias, wimg + P.wsig, sig, trow + TC_A_HI, trow + TC_A_LO);
            act_block32<true>(r1, bias + 32, wimg + P.wsig + 32, sig, trow + TC_A_HI + 32, trow + TC_A_LO + 32);
          } else {
            act_block32<false>(r0, bias, nullptr, sig, trow + TC_A_HI, trow + TC_A_LO);
            act_block32<false>(r1, bias + 32, nullptr, sig, trow + TC_A_HI + 32, trow + TC_A_LO + 32);
          }
        }
        publish();
        if (issuer_warp) {
          if (umma::elect_one()) {
            umma::fence_after_sync();
            umma::issue_layer_k64<64>(tbase + TC_D, tbase + TC_A_HI, tbase + TC_A_LO,
                                      wimg_s + 4u * (uint32_t)P.w_hi[m], wimg_s + 4u * (uint32_t)P.w_lo[m]);
            umma::mma_commit(bar);
          }
          __syncwarp();
        }
        // independent work under the MMA waits: noise of the next iteration, this iteration's uniform
        if (!A.eps_dev) {
          if (m == 0 && t + 1 < tb) {
#pragma unroll
            for (int g = 0; g < ZMAX / 4; ++g) {
              if (g * 4 < zd) {
                float e[4];
                normal4(A.seed, grow, (uint32_t)(t + 1), NOISE_PROPOSAL, g, e);
#pragma unroll
                for (int q = 0; q < 4; ++q) en[g * 4 + q] = e[q];
              }
            }
          }
          if (m == 1 && !init_pass) u_acc = uniform1(A.seed, grow, (uint32_t)t, NOISE_ACCEPT);
        }
        stage_wait();
      }
      // ---- covariate likelihood in the row space of g's last layer (:800-801) ----
      float sse = 0.f;
      {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t r[32], tt[32];
          umma::ld32(trow + TC_D + c * 32, r);
          umma::ld32(trow + TC_P + c * 32, tt);
          umma::wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float dd = __uint_as_float(r[j]) - __uint_as_float(tt[j]);
            sse = fmaf(dd, dd, sse);
          }
        }
      }
      const float s2v = P.s2v >= 0.f ? P.s2v : softplus_f(sig) + 1e-6f;
      const float loss_pv = (sse + r0_l) / (2.f * s2v) + ((float)P.p * logf(s2v)) / 2.f;
      const float lp_prop = -(((loss_pv + loss_px) + loss_py) + prior);              // :814-816
      if (init_pass) {
        lp_cur = lp_prop;
        continue;
      }
      // accept: u < exp(min(lp' - lp, 0))  (:868-870); a NaN ratio never accepts
      const float dlp = lp_prop - lp_cur;
      const float ratio = (dlp < 0.f) ? expf(dlp) : ((dlp >= 0.f) ? 1.f : __int_as_float(0x7fc00000));
      bool acc;
      if (A.u_dev) acc = A.u_dev[(size_t)t * n + lrow] < (double)ratio;
      else acc = u_acc < ratio;
      if (acc) {                                                                     // :871
#pragma unroll
        for (int d = 0; d < ZMAX; ++d)
          if (d < zd) zc[d] = in[d];
        lp_cur = lp_prop;
      }
      if (A.accept_mask_dev && valid) A.accept_mask_dev[(size_t)t * n + row] = acc ? 1 : 0;
      if (A.lp_trace_dev && valid) A.lp_trace_dev[(size_t)t * n + row] = lp_prop;
      if (A.accept_count_dev) {
        const unsigned b = __ballot_sync(0xffffffffu, acc && valid);
        if (lane == 0 && b) atomicAdd(A.accept_count_dev + t, __popc(b));
      }
      if (t >= A.burn_in && A.out_samples_dev && valid) {                            // :895-896
        float* dst = A.out_samples_dev + ((size_t)(t - A.burn_in) * n + row) * zd;
#pragma unroll
        for (int d = 0; d < ZMAX; ++d)
          if (d < zd) dst[d] = zc[d];
      }
    }
    // ---- save state, publish the chunk ----
    if (valid) {
      if (D.mode == 0) {
#pragma unroll
        for (int d = 0; d < ZMAX; ++d)
          if (d < zd) A.z_state_dev[(size_t)row * zd + d] = zc[d];
      }
      A.lp_state_dev[row] = lp_cur;
    }
    __threadfence();
    wg_sync(wg);
    if (wtid == 0) *reinterpret_cast<volatile int*>(sched + 1 + tile) = chunk + 1;
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc512(tmem0);
}

// ---------------------------------------------------------------------------------------
// The same sampler with EIGHT warps per 128-row tile (16 warps, two tiles per CTA): warps q and
// q+4 of a tile share TMEM lane quarter q, i.e. two threads per row, and split every layer's
// columns in halves (c = 0 / 1).  Both threads carry the identical chain state and make the
// identical accept decision; the per-row scalars that each computes for its half (f-net loss by
// c = 0, h-net loss by c = 1, the halves of the sigma_v dot product and of the covariate SSE) are
// exchanged through shared memory once per iteration.  Per stage the epilogue latency halves and
// each SM sub-partition has 4 warps to interleave instead of 2.
__device__ __forceinline__ void tile_sync256(int slot) {
  if (slot == 0) asm volatile("bar.sync 1, 256;" ::: "memory");
  else asm volatile("bar.sync 2, 256;" ::: "memory");
}

// 32 outputs [c0, c0+32) of a first layer: out = LeakyReLU(b + in W)
template <int KINMAX>
__device__ __forceinline__ void first_layer32(const float* __restrict__ W, const float* __restrict__ b,
                                              unsigned long long mask, int nin, const float (&in)[KINMAX],
                                              int c0, float (&out)[32]) {
  float2 o2[16];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 bb = *reinterpret_cast<const float4*>(b + c0 + q * 4);
    o2[q * 2 + 0] = make_float2(bb.x, bb.y);
    o2[q * 2 + 1] = make_float2(bb.z, bb.w);
  }
  int wr = 0;   // packed row of W
#pragma unroll
  for (int d = 0; d < KINMAX; ++d) {
    if (d < nin && ((mask >> d) & 1ull)) {
      const float2 zv = make_float2(in[d], in[d]);
      const float* Wd = W + wr * 64 + c0;
      ++wr;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 ww = *reinterpret_cast<const float4*>(Wd + q * 4);
        o2[q * 2 + 0] = __ffma2_rn(zv, make_float2(ww.x, ww.y), o2[q * 2 + 0]);
        o2[q * 2 + 1] = __ffma2_rn(zv, make_float2(ww.z, ww.w), o2[q * 2 + 1]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float2 a = leaky_x2(o2[j]);
    out[2 * j] = a.x;
    out[2 * j + 1] = a.y;
  }
}
__device__ __forceinline__ void split_store32(const float (&a)[32], uint32_t tA_hi, uint32_t tA_lo) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {   // 16 columns at a time: the 16-warp kernels have 128 registers per thread
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
      float2 h2, l2;
      umma::split_tf32_x2(make_float2(a[h * 16 + j], a[h * 16 + j + 1]), h2, l2);
      hi[j] = __float_as_uint(h2.x); hi[j + 1] = __float_as_uint(h2.y);
      lo[j] = __float_as_uint(l2.x); lo[j + 1] = __float_as_uint(l2.y);
    }
    umma::st16(tA_hi + h * 16, hi);
    umma::st16(tA_lo + h * 16, lo);
  }
}
// bias + LeakyReLU (+ sigma_v partial dot) + hi/lo split of 16 accumulator columns, back into the A slots
template <bool SIG>
__device__ __forceinline__ void act_block16(uint32_t (&r)[16], const float* __restrict__ bias,
                                            const float* __restrict__ wsig, float& sig, uint32_t tA_hi,
                                            uint32_t tA_lo) {
  uint32_t lo[16];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 b = *reinterpret_cast<const float4*>(bias + i * 4);
    float4 ws = make_float4(0.f, 0.f, 0.f, 0.f);
    if constexpr (SIG) ws = *reinterpret_cast<const float4*>(wsig + i * 4);
    const float2 a0 = leaky_x2(__fadd2_rn(make_float2(__uint_as_float(r[i * 4 + 0]), __uint_as_float(r[i * 4 + 1])),
                                          make_float2(b.x, b.y)));
    const float2 a1 = leaky_x2(__fadd2_rn(make_float2(__uint_as_float(r[i * 4 + 2]), __uint_as_float(r[i * 4 + 3])),
                                          make_float2(b.z, b.w)));
    if constexpr (SIG) {
      sig = fmaf(a0.x, ws.x, sig); sig = fmaf(a0.y, ws.y, sig);
      sig = fmaf(a1.x, ws.z, sig); sig = fmaf(a1.y, ws.w, sig);
    }
    float2 h0, l0, h1, l1;
    umma::split_tf32_x2(a0, h0, l0);
    umma::split_tf32_x2(a1, h1, l1);
    r[i * 4 + 0] = __float_as_uint(h0.x); r[i * 4 + 1] = __float_as_uint(h0.y);
    r[i * 4 + 2] = __float_as_uint(h1.x); r[i * 4 + 3] = __float_as_uint(h1.y);
    lo[i * 4 + 0] = __float_as_uint(l0.x); lo[i * 4 + 1] = __float_as_uint(l0.y);
    lo[i * 4 + 2] = __float_as_uint(l1.x); lo[i * 4 + 3] = __float_as_uint(l1.y);
  }
  umma::st16(tA_hi, r);
  umma::st16(tA_lo, lo);
}
// 16 accumulator columns -> LeakyReLU(r + bias) -> hi / lo into the A slots
__device__ __forceinline__ void act_store16(uint32_t (&r)[16], const float* __restrict__ bias, uint32_t tA_hi,
                                            uint32_t tA_lo) {
  float dummy = 0.f;
  act_block16<false>(r, bias, nullptr, dummy, tA_hi, tA_lo);
}

constexpr int TC16_XCH_FLOATS = 2 * 2 * TC_ROWS * 4;   // [slot][half][row][4] exchange buffer (x2: partials, noise)

template <int ZMAX, bool L1>
__global__ void __launch_bounds__(512, 1)
causal_mh_tc16_kernel(const __grid_constant__ TcProgram P, const float* __restrict__ image,
                      const __grid_constant__ MhDev D) {
  constexpr int KINMAX = ZMAX + 1;
  extern __shared__ __align__(128) float smem[];
  __shared__ uint64_t bar_img;
  __shared__ uint64_t bar_mma[2];
  __shared__ uint32_t tmem_slot;
  __shared__ long long unit_s[2];
  bulk_load_to_smem(smem, image, (uint32_t)P.image_floats * 4u, &bar_img);
  const float* wimg = smem;
  float* xch = smem + P.image_floats;
  const uint32_t wimg_s = umma::smem_addr(smem);

  const bgm_mh_args& A = D.a;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int slot = warp >> 3;          // which of the CTA's two tiles
  const int q = warp & 3;              // TMEM lane quarter
  const int c = (warp >> 2) & 1;       // column half
  const bool issuer_warp = (warp & 7) == 7;   // highest warp id of the tile: the arbiter favours it
  const bool leader = issuer_warp && lane == 0;
  const int r_in_tile = q * 32 + lane;
  if (tid == 0) {
    umma::mbar_init(umma::smem_addr(&bar_mma[0]), 1);
    umma::mbar_init(umma::smem_addr(&bar_mma[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) umma::tmem_alloc512(&tmem_slot);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem0 = __shfl_sync(0xffffffffu, tmem_slot, 0);
  const uint32_t tbase = tmem0 + (uint32_t)slot * 256u;
  const uint32_t trow = tbase + ((uint32_t)(q * 32) << 16);
  const uint32_t bar = umma::smem_addr(&bar_mma[slot]);
  uint32_t parity = 0;
  float* my_x = xch + ((slot * 2 + c) * TC_ROWS + r_in_tile) * 4;
  const float* other_x = xch + ((slot * 2 + (c ^ 1)) * TC_ROWS + r_in_tile) * 4;
  // zd <= 8: the row's two threads each draw one group of 4 proposal normals (group c) and
  // pass it on through shared memory instead of both drawing both
  constexpr bool SHARE_NOISE = ZMAX == 8;
  float* xn = xch + TC16_XCH_FLOATS;   // [slot][group][row][4]
  float* my_n = xn + ((slot * 2 + c) * TC_ROWS + r_in_tile) * 4;
  const float* n0 = xn + ((slot * 2 + 0) * TC_ROWS + r_in_tile) * 4;
  const float* n1 = xn + ((slot * 2 + 1) * TC_ROWS + r_in_tile) * 4;
  // L1 (the image is 2.3 KB larger): compact noise buffer, group 0 as [slot][row][4], then only the zd - 4 values of
  // group 1 that are used as [slot][row][zd - 4]
  const int n1s = P.zd > 4 ? P.zd - 4 : 0;
  if constexpr (L1) {
    n0 = xn + (slot * TC_ROWS + r_in_tile) * 4;
    n1 = xn + 2 * TC_ROWS * 4 + (slot * TC_ROWS + r_in_tile) * n1s;
    my_n = const_cast<float*>(c == 0 ? n0 : n1);
  }

  const int n = A.n, zd = P.zd;
  const int ntiles = (n + TC_ROWS - 1) / TC_ROWS;
  const double q_sd = A.q_sd_dev ? *A.q_sd_dev : 1.0;
  const bool need_init = !(A.init_mode == 0 && D.mode == 0);
  const int t_first = need_init ? A.t_begin - 1 : A.t_begin;
  const int t_last = D.mode == 1 ? A.t_begin : A.t_end;
  int* sched = A.sched_dev;
  const int n_iter = t_last - t_first;
  const int nchunks = D.nchunks;
  const int chunk_len = (n_iter + nchunks - 1) / nchunks;
  const long long total_units = (long long)ntiles * nchunks;
  const int n_mma = P.n_mma;

  auto publish = [&]() {
    umma::wait_st();
    umma::fence_before_sync();
    tile_publish256(slot, issuer_warp);
  };
  auto stage_wait = [&]() {
    umma::mbar_wait(bar, parity);
    parity ^= 1u;
    umma::fence_after_sync();
  };

  for (;;) {
    if (leader) {
      const long long u = atomicAdd(reinterpret_cast<unsigned int*>(sched), 1u);
      if (u < total_units) {
        const int chunk = (int)(u / ntiles);
        const int tile = (int)(u - (long long)chunk * ntiles);
        if (chunk > 0) {
          const volatile int* flag = sched + 1 + tile;
          while (*flag < chunk) __nanosleep(200);
          __threadfence();
        }
      }
      unit_s[slot] = u;
    }
    tile_sync256(slot);
    const long long u = unit_s[slot];
    if (u >= total_units) break;
    const int chunk = (int)(u / ntiles);
    const int tile = (int)(u - (long long)chunk * ntiles);
    const int ta = t_first + chunk * chunk_len;
    const int tb = min(ta + chunk_len, t_last);
    const int row = tile * TC_ROWS + r_in_tile;
    const bool valid = row < n;
    const bool writer = valid && c == 0;   // one of the row's two threads owns the outputs
    const int lrow = valid ? row : n - 1;
    const float x_l = A.x_dev[lrow], y_l = A.y_dev[lrow];
    const float r0_l = A.r0_dev[lrow];
    const int64_t grow = A.row_offset + lrow;
    // this thread's half of the row's projected covariates stays in its TMEM lane for the unit
    {
      const float4* tg = reinterpret_cast<const float4*>(A.vproj_dev + (size_t)lrow * A.ldvproj) + c * 8;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t tr[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 v4 = __ldg(tg + h * 4 + i);
          tr[i * 4 + 0] = __float_as_uint(v4.x); tr[i * 4 + 1] = __float_as_uint(v4.y);
          tr[i * 4 + 2] = __float_as_uint(v4.z); tr[i * 4 + 3] = __float_as_uint(v4.w);
        }
        umma::st16(trow + TC_P + c * 32 + h * 16, tr);
      }
      umma::wait_st();
    }
    float zc[ZMAX];
    // ---- initial state (:842) ----
    if (chunk == 0 && A.init_mode == 2) {
#pragma unroll
      for (int g = 0; g < ZMAX / 4; ++g) {
        if (g * 4 < zd) {
          float e[4];
          normal4(A.seed, grow, T_INIT, NOISE_PROPOSAL, g, e);
#pragma unroll
          for (int i = 0; i < 4; ++i) zc[g * 4 + i] = e[i];
        }
      }
    } else {
#pragma unroll
      for (int d = 0; d < ZMAX; ++d) zc[d] = (d < zd) ? __ldcg(A.z_state_dev + (size_t)lrow * zd + d) : 0.f;
    }
    float lp_cur = (need_init && chunk == 0) ? 0.f : __ldcg(A.lp_state_dev + lrow);
    float en[ZMAX];
    float u_acc = 0.f;
#pragma unroll
    for (int d = 0; d < ZMAX; ++d) en[d] = 0.f;
    if (!A.eps_dev && ta < tb) {
#pragma unroll
      for (int g = 0; g < ZMAX / 4; ++g) {
        if (g * 4 < zd) {
          float e[4];
          normal4(A.seed, grow, (uint32_t)ta, NOISE_PROPOSAL, g, e);
#pragma unroll
          for (int i = 0; i < 4; ++i) en[g * 4 + i] = e[i];
        }
      }
    }
    // ---- iterations (:860-898) ----
#pragma unroll 1
    for (int t = ta; t < tb; ++t) {
      const bool init_pass = t < A.t_begin;
      float in[KINMAX];   // [proposal z' (zd), x]
#pragma unroll
      for (int d = 0; d < KINMAX; ++d) in[d] = 0.f;
      if (init_pass) {
#pragma unroll
        for (int d = 0; d < ZMAX; ++d)
          if (d < zd) in[d] = zc[d];
      } else if (A.eps_dev) {
        const float* e = A.eps_dev + ((size_t)t * n + lrow) * zd;
#pragma unroll
        for (int d = 0; d < ZMAX; ++d)
          if (d < zd) in[d] = __fadd_rn(zc[d], (float)(q_sd * (double)e[d]));
      } else {
        if (SHARE_NOISE && t > ta) {
          const float4 a4 = *reinterpret_cast<const float4*>(n0);
          en[0] = a4.x; en[1] = a4.y; en[2] = a4.z; en[3] = a4.w;
          if constexpr (L1) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (ZMAX > 4 && i < n1s) en[4 + i] = n1[i];
          } else {
            const float4 b4 = *reinterpret_cast<const float4*>(n1);
            if (ZMAX > 4) { en[4] = b4.x; en[5] = b4.y; en[6] = b4.z; en[7] = b4.w; }
          }
        }
#pragma unroll
        for (int d = 0; d < ZMAX; ++d)
          if (d < zd) in[d] = __fadd_rn(zc[d], (float)(q_sd * (double)en[d]));
      }
      float prior = 0.f;                                                              // :812
#pragma unroll
      for (int d = 0; d < ZMAX; ++d)
        if (d < zd) prior = fmaf(in[d], in[d], prior);
      prior *= 0.5f;
#pragma unroll
      for (int d = 0; d < KINMAX; ++d)
        if (d == zd) in[d] = x_l;

      float a1[32];   // h layer 1 (my 32 columns), kept in registers until f's second-layer MMAs have read A
      if constexpr (L1) {
        // ---- f and h layer 1 on the tensor cores: [z', x, 0.., 1] (8 inputs) -> [f_h1 | h_h1] (TMEM columns 64..191) ----
        {
          uint32_t w8[8];
#pragma unroll
          for (int j = 0; j < 7; ++j) {   // in[] is [z' (zd), x, 0..]: zd + 1 <= 7 entries in use
            uint32_t hi, lo;
            umma::split_tf32(in[j], hi, lo);
            w8[j] = c == 0 ? hi : lo;
          }
          w8[7] = c == 0 ? 0x3f800000u : 0u;   // the ones column (row 7 of the images holds the biases)
          umma::st8(trow + TC_A_HI + c * 16, w8);   // c = 0: hi at columns 0..7, c = 1: lo at columns 16..23
        }
        publish();
        if (issuer_warp) {
          if (umma::elect_one()) {
            umma::fence_after_sync();
            umma::issue_layer<128, 1>(tbase + TC_A_LO, tbase + TC_A_HI, tbase + TC_A_HI + 16,
                                      wimg_s + 4u * (uint32_t)P.l1_hi, wimg_s + 4u * (uint32_t)P.l1_lo);
            umma::mma_commit(bar);
          }
          __syncwarp();
        }
        stage_wait();
#pragma unroll
        for (int h = 0; h < 2; ++h) {   // f_h1: LeakyReLU + split, back into the A slots (same columns)
          uint32_t rf[16], lo[16];
          umma::ld16(trow + TC_A_LO + c * 32 + h * 16, rf);
          umma::wait_ld();
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            float2 h2, l2;
            umma::split_tf32_x2(leaky_x2(make_float2(__uint_as_float(rf[j]), __uint_as_float(rf[j + 1]))), h2, l2);
            rf[j] = __float_as_uint(h2.x); rf[j + 1] = __float_as_uint(h2.y);
            lo[j] = __float_as_uint(l2.x); lo[j + 1] = __float_as_uint(l2.y);
          }
          umma::st16(trow + TC_A_HI + c * 32 + h * 16, rf);
          umma::st16(trow + TC_A_LO + c * 32 + h * 16, lo);
        }
        {   // h_h1 (before activation) leaves the D slot: f's second-layer product is about to overwrite it
          uint32_t rh0[16], rh1[16];
          umma::ld16(trow + TC_D + c * 32, rh0);
          umma::ld16(trow + TC_D + c * 32 + 16, rh1);
          umma::wait_ld();
#pragma unroll
          for (int j = 0; j < 16; ++j) { a1[j] = __uint_as_float(rh0[j]); a1[16 + j] = __uint_as_float(rh1[j]); }
        }
      } else {
        // ---- f layer 1 (my 32 columns) -> A; stage F2 ----
        float f1[32];
        first_layer32<KINMAX>(wimg + P.fW1, wimg + P.fb1, P.fmask, zd + 1, in, c * 32, f1);
        split_store32(f1, trow + TC_A_HI + c * 32, trow + TC_A_LO + c * 32);
      }
      publish();
      if (issuer_warp) {
        if (umma::elect_one()) {
          umma::fence_after_sync();
          umma::issue_layer_k64<32>(tbase + TC_D, tbase + TC_A_HI, tbase + TC_A_LO, wimg_s + 4u * (uint32_t)P.f2_hi,
                                    wimg_s + 4u * (uint32_t)P.f2_lo);
          umma::mma_commit(bar);
        }
        __syncwarp();
      }
      {
        if constexpr (L1) {   // h layer 1 came out of the same product: only its LeakyReLU is left, while f's MMAs run
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float2 a = leaky_x2(make_float2(a1[j], a1[j + 1]));
            a1[j] = a.x; a1[j + 1] = a.y;
          }
        } else {              // h layer 1 while f's MMAs run
          first_layer32<KINMAX>(wimg + P.hW1, wimg + P.hb1, P.hmask, zd + 1, in, c * 32, a1);
        }
        stage_wait();
        // f_h2 columns [16c, 16c+16) -> A3 columns [16c ..) ; A3 = [f_h2 (32) | h_h2 (32)]
        uint32_t r[16];
        umma::ld16(trow + TC_D + c * 16, r);
        umma::wait_ld();
        // A3 lives in TC_P? no: TC_P holds t.  A3 is written after h's MMAs (A slots busy), keep f_h2 in registers.
        float fh2[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 b = *reinterpret_cast<const float4*>(wimg + P.fb2 + c * 16 + i * 4);
          const float2 a0 = leaky_x2(__fadd2_rn(make_float2(__uint_as_float(r[i * 4 + 0]), __uint_as_float(r[i * 4 + 1])),
                                                make_float2(b.x, b.y)));
          const float2 a1p = leaky_x2(__fadd2_rn(make_float2(__uint_as_float(r[i * 4 + 2]), __uint_as_float(r[i * 4 + 3])),
                                                 make_float2(b.z, b.w)));
          fh2[i * 4 + 0] = a0.x; fh2[i * 4 + 1] = a0.y; fh2[i * 4 + 2] = a1p.x; fh2[i * 4 + 3] = a1p.y;
        }
        split_store32(a1, trow + TC_A_HI + c * 32, trow + TC_A_LO + c * 32);
        publish();
        if (issuer_warp) {
          if (umma::elect_one()) {
            umma::fence_after_sync();
            umma::issue_layer_k64<32>(tbase + TC_D, tbase + TC_A_HI, tbase + TC_A_LO,
                                      wimg_s + 4u * (uint32_t)P.h2_hi, wimg_s + 4u * (uint32_t)P.h2_lo);
            umma::mma_commit(bar);
          }
          __syncwarp();
        }
        stage_wait();
        {
          uint32_t rh[16];
          umma::ld16(trow + TC_D + c * 16, rh);
          umma::wait_ld();
          // f_h2 -> A3 columns [16c..], h_h2 -> A3 columns [32 + 16c..]
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            float2 h2, l2;
            umma::split_tf32_x2(make_float2(fh2[j], fh2[j + 1]), h2, l2);
            hi[j] = __float_as_uint(h2.x); hi[j + 1] = __float_as_uint(h2.y);
            lo[j] = __float_as_uint(l2.x); lo[j + 1] = __float_as_uint(l2.y);
          }
          umma::st16(trow + TC_A_HI + c * 16, hi);
          umma::st16(trow + TC_A_LO + c * 16, lo);
          act_store16(rh, wimg + P.hb2 + c * 16, trow + TC_A_HI + 32 + c * 16, trow + TC_A_LO + 32 + c * 16);
        }
        publish();
        if (issuer_warp) {
          if (umma::elect_one()) {
            umma::fence_after_sync();
            umma::issue_layer_k64<16>(tbase + TC_D, tbase + TC_A_HI, tbase + TC_A_LO,
                                      wimg_s + 4u * (uint32_t)P.w3_hi, wimg_s + 4u * (uint32_t)P.w3_lo);
            umma::mma_commit(bar);
          }
          __syncwarp();
        }
        float g1[32];   // g layer 1 while the layer-3 MMAs run
        first_layer32<KINMAX>(wimg + P.gW1, wimg + P.gb1, ~0ull, zd, in, c * 32, g1);
        stage_wait();
        // ---- heads: c = 0 finishes f_net (outcome model :809-810), c = 1 h_net (treatment model :803-807) ----
        // The 8 -> 2 head is taken now (g's first product overwrites the D slot); the loss itself (softplus, log,
        // division: ~200 instructions of dependent math) waits for one of g's MMA stages, where the thread is idle.
        float my_loss = 0.f, head_mu, head_raw;
        {
          uint32_t r8[8];
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                       : "=r"(r8[0]), "=r"(r8[1]), "=r"(r8[2]), "=r"(r8[3]), "=r"(r8[4]), "=r"(r8[5]), "=r"(r8[6]),
                         "=r"(r8[7])
                       : "r"(trow + TC_D + c * 8));
          umma::wait_ld();
          const float* W4 = wimg + (c == 0 ? P.fW4 : P.hW4);
          const float* b4 = wimg + (c == 0 ? P.fb4 : P.hb4);
          float mu = b4[0], raw = b4[1];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float a = leaky_mx(__uint_as_float(r8[k]) + wimg[P.b3 + c * 8 + k]);
            const float2 w2 = *reinterpret_cast<const float2*>(W4 + k * 2);
            mu = fmaf(a, w2.x, mu);
            raw = fmaf(a, w2.y, raw);
          }
          head_mu = mu;
          head_raw = raw;
        }
        auto head_loss = [&]() {
          const float mu = head_mu, raw = head_raw;
          if (c == 0) {
            const float s2y = P.s2y >= 0.f ? P.s2y : softplus_f(raw) + 1e-6f;
            const float dy = y_l - mu;
            my_loss = (dy * dy) / (2.f * s2y) + logf(s2y) / 2.f;
          } else if (P.binary) {
            my_loss = fmaxf(mu, 0.f) - mu * x_l + log1pf(expf(-fabsf(mu)));
          } else {
            const float s2x = P.s2x >= 0.f ? P.s2x : softplus_f(raw) + 1e-6f;
            const float dx = x_l - mu;
            my_loss = (dx * dx) / (2.f * s2x) + logf(s2x) / 2.f;
          }
        };
        // ---- g_net on the tensor cores ----
        split_store32(g1, trow + TC_A_HI + c * 32, trow + TC_A_LO + c * 32);
        float sig = 0.f;
        auto g_issue = [&](int m) {
          publish();
          if (issuer_warp) {
            if (umma::elect_one()) {
              umma::fence_after_sync();
              umma::issue_layer_k64<64>(tbase + TC_D, tbase + TC_A_HI, tbase + TC_A_LO,
                                        wimg_s + 4u * (uint32_t)P.w_hi[m], wimg_s + 4u * (uint32_t)P.w_lo[m]);
              umma::mma_commit(bar);
            }
            __syncwarp();
          }
        };
        // first g product: nothing to read back yet.  Under its wait (the stage with the fewest live registers):
        // the noise of the next iteration and the loss of the f / h head.
        g_issue(0);
        if (!A.eps_dev && t + 1 < tb) {
          if constexpr (SHARE_NOISE) {
            if (c * 4 < zd) {
              float e[4];
              normal4(A.seed, grow, (uint32_t)(t + 1), NOISE_PROPOSAL, (uint32_t)c, e);
              if (!L1 || c == 0) {
                *reinterpret_cast<float4*>(my_n) = make_float4(e[0], e[1], e[2], e[3]);
              } else {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  if (i < n1s) my_n[i] = e[i];
              }
            }
          } else {
#pragma unroll
            for (int g = 0; g < ZMAX / 4; ++g) {
              if (g * 4 < zd) {
                float e[4];
                normal4(A.seed, grow, (uint32_t)(t + 1), NOISE_PROPOSAL, g, e);
#pragma unroll
                for (int i = 0; i < 4; ++i) en[g * 4 + i] = e[i];
              }
            }
          }
        }
        head_loss();
        stage_wait();
#pragma unroll 1
        for (int m = 1; m < n_mma; ++m) {
          {
            uint32_t ra[16], rb[16];
            umma::ld16(trow + TC_D + c * 32, ra);
            umma::ld16(trow + TC_D + c * 32 + 16, rb);
            umma::wait_ld();
            const float* bias = wimg + P.gb[m - 1] + c * 32;
            const uint32_t ah = trow + TC_A_HI + c * 32, al = trow + TC_A_LO + c * 32;
            if (m == n_mma - 1) {
              act_block16<true>(ra, bias, wimg + P.wsig + c * 32, sig, ah, al);
              act_block16<true>(rb, bias + 16, wimg + P.wsig + c * 32 + 16, sig, ah + 16, al + 16);
            } else {
              act_block16<false>(ra, bias, nullptr, sig, ah, al);
              act_block16<false>(rb, bias + 16, nullptr, sig, ah + 16, al + 16);
            }
          }
          g_issue(m);
          // the accept uniform under the second wait: drawn by the c == 0 thread, handed over with the partials below
          if (m == 1 && !A.eps_dev && !init_pass && c == 0) u_acc = uniform1(A.seed, grow, (uint32_t)t, NOISE_ACCEPT);
          stage_wait();
        }
        // ---- my half of the covariate SSE (:800), then the row's two threads swap their partials ----
        float sse = 0.f;
        {
          float2 sse2 = make_float2(0.f, 0.f);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t r[16], tt[16];
            umma::ld16(trow + TC_D + c * 32 + h * 16, r);
            umma::ld16(trow + TC_P + c * 32 + h * 16, tt);
            umma::wait_ld();
#pragma unroll
            for (int j = 0; j < 16; j += 2) {   // two running sums (even / odd columns), packed
              const float2 dd = __ffma2_rn(make_float2(__uint_as_float(tt[j]), __uint_as_float(tt[j + 1])), make_float2(-1.f, -1.f),
                                           make_float2(__uint_as_float(r[j]), __uint_as_float(r[j + 1])));
              sse2 = __ffma2_rn(dd, dd, sse2);
            }
          }
          sse = sse2.x + sse2.y;
        }
        *reinterpret_cast<float4*>(my_x) = make_float4(my_loss, sig, sse, u_acc);
        tile_sync256(slot);
        const float4 ox = *reinterpret_cast<const float4*>(other_x);
        if (c == 1) u_acc = ox.w;
        const float loss_py = c == 0 ? my_loss : ox.x;
        const float loss_px = c == 0 ? ox.x : my_loss;
        const float sig0 = c == 0 ? sig : ox.y, sig1 = c == 0 ? ox.y : sig;
        const float sse0 = c == 0 ? sse : ox.z, sse1 = c == 0 ? ox.z : sse;
        const float sig_all = (wimg[P.bsig] + sig0) + sig1;
        const float sse_all = sse0 + sse1;
        const float s2v = P.s2v >= 0.f ? P.s2v : softplus_f(sig_all) + 1e-6f;
        const float loss_pv = (sse_all + r0_l) / (2.f * s2v) + ((float)P.p * logf(s2v)) / 2.f;
        const float lp_prop = -(((loss_pv + loss_px) + loss_py) + prior);              // :814-816
        if (init_pass) {
          lp_cur = lp_prop;
          continue;
        }
        // accept: u < exp(min(lp' - lp, 0))  (:868-870); a NaN ratio never accepts
        const float dlp = lp_prop - lp_cur;
        const float ratio = (dlp < 0.f) ? expf(dlp) : ((dlp >= 0.f) ? 1.f : __int_as_float(0x7fc00000));
        bool acc;
        if (A.u_dev) acc = A.u_dev[(size_t)t * n + lrow] < (double)ratio;
        else acc = u_acc < ratio;
        if (acc) {                                                                     // :871
#pragma unroll
          for (int d = 0; d < ZMAX; ++d)
            if (d < zd) zc[d] = in[d];
          lp_cur = lp_prop;
        }
        if (A.accept_mask_dev && writer) A.accept_mask_dev[(size_t)t * n + row] = acc ? 1 : 0;
        if (A.lp_trace_dev && writer) A.lp_trace_dev[(size_t)t * n + row] = lp_prop;
        if (A.accept_count_dev && c == 0) {
          const unsigned b = __ballot_sync(0xffffffffu, acc && valid);
          if (lane == 0 && b) atomicAdd(A.accept_count_dev + t, __popc(b));
        }
        if (t >= A.burn_in && A.out_samples_dev && writer) {                           // :895-896
          float* dst = A.out_samples_dev + ((size_t)(t - A.burn_in) * n + row) * zd;
#pragma unroll
          for (int d = 0; d < ZMAX; ++d)
            if (d < zd) dst[d] = zc[d];
        }
      }
    }
    // ---- save state, publish the chunk ----
    if (writer) {
      if (D.mode == 0) {
#pragma unroll
        for (int d = 0; d < ZMAX; ++d)
          if (d < zd) A.z_state_dev[(size_t)row * zd + d] = zc[d];
      }
      A.lp_state_dev[row] = lp_cur;
    }
    __threadfence();
    tile_sync256(slot);
    if (leader) *reinterpret_cast<volatile int*>(sched + 1 + tile) = chunk + 1;
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc512(tmem0);
}

// ---------------------------------------------------------------------------------------
// Effect kernel, tensor-core engine: f_net on kept states for every dose
// (infer_from_latent_posterior, causalbgm/base.py:671-763; same contract and the same Philox
// stream as causal_effect_kernel).  A tile is 128 (kept state, row) pairs = 128 TMEM lanes,
// worked on by EIGHT warps: warps q and q+4 of the tile share lane quarter q and split the
// columns (first-layer outputs 32 + 32, second-layer outputs 16 + 16) and the doses of the
// 8 -> 2 head (even / odd), so that each SM sub-partition has 4 warps to interleave.  Two
// tiles per CTA (16 warps, 512 TMEM columns).  The z part of the first layer is computed
// once per tile; per dose the thread adds x*w_x, and the 64->32 and 32->8 layers run on the
// tensor cores, software-pipelined over doses: MMA stage j carries layer 2 of dose j and
// layer 3 of dose j-1.
constexpr uint32_t EF_A2_HI = 0, EF_A2_LO = 64, EF_D2 = 128, EF_A3_HI = 160, EF_A3_LO = 192, EF_D3 = 224;

template <int ZMAX>
__global__ void __launch_bounds__(512, 1)
causal_effect_tc_kernel(const __grid_constant__ TcProgram P, const float* __restrict__ image,
                        const __grid_constant__ EffectDev E) {
  extern __shared__ __align__(128) float smem[];
  __shared__ uint64_t bar_img;
  __shared__ uint64_t bar_mma[2];
  __shared__ uint32_t tmem_slot;
  bulk_load_to_smem(smem, image, (uint32_t)P.image_floats * 4u, &bar_img);
  const float* wimg = smem;
  const uint32_t wimg_s = umma::smem_addr(smem);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int slot = warp >> 3;          // which of the CTA's two tiles
  const int q = warp & 3;              // TMEM lane quarter
  const int c = (warp >> 2) & 1;       // column half / dose parity
  const bool issuer_warp = (warp & 7) == 7;   // highest warp id of the tile: the arbiter favours it
  if (tid == 0) {
    umma::mbar_init(umma::smem_addr(&bar_mma[0]), 1);
    umma::mbar_init(umma::smem_addr(&bar_mma[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) umma::tmem_alloc512(&tmem_slot);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem0 = __shfl_sync(0xffffffffu, tmem_slot, 0);
  const uint32_t tbase = tmem0 + (uint32_t)slot * 256u;
  const uint32_t trow = tbase + ((uint32_t)(q * 32) << 16);
  const uint32_t bar = umma::smem_addr(&bar_mma[slot]);
  uint32_t parity = 0;

  const int n = E.n, zd = P.zd, n_x = E.n_x;
  const int tiles_per_s = (n + TC_ROWS - 1) / TC_ROWS;
  const long long ntiles = (long long)tiles_per_s * E.n_keep;
  // first-layer weights of the treatment input: the last packed row of f's first layer
  const float* wx = wimg + P.fW1 + (__popcll(P.fmask) - 1) * 64 + c * 32;
  for (long long tile = (long long)blockIdx.x * 2 + slot; tile < ntiles; tile += (long long)gridDim.x * 2) {
    const int s = (int)(tile / tiles_per_s);
    const int row = (int)(tile - (long long)s * tiles_per_s) * TC_ROWS + q * 32 + lane;
    const bool valid = row < n;
    const int lrow = valid ? row : n - 1;
    const int64_t grow = E.row_offset + lrow;
    // z part of f's first layer, this thread's 32 columns: base = b1 + [z0, z1] W1
    float base[32];
    {
      const float* zs = E.z_samples + ((size_t)s * n + lrow) * zd;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 bb = *reinterpret_cast<const float4*>(wimg + P.fb1 + c * 32 + i * 4);
        base[i * 4 + 0] = bb.x; base[i * 4 + 1] = bb.y; base[i * 4 + 2] = bb.z; base[i * 4 + 3] = bb.w;
      }
      int wr = 0;
#pragma unroll
      for (int d = 0; d < ZMAX; ++d) {
        if (d < zd && ((P.fmask >> d) & 1ull)) {
          const float zv = zs[d];
          const float* Wd = wimg + P.fW1 + wr * 64 + c * 32;
          ++wr;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 ww = *reinterpret_cast<const float4*>(Wd + i * 4);
            base[i * 4 + 0] = fmaf(zv, ww.x, base[i * 4 + 0]);
            base[i * 4 + 1] = fmaf(zv, ww.y, base[i * 4 + 1]);
            base[i * 4 + 2] = fmaf(zv, ww.z, base[i * 4 + 2]);
            base[i * 4 + 3] = fmaf(zv, ww.w, base[i * 4 + 3]);
          }
        }
      }
    }
    float y_prev = 0.f;
    float nz[4] = {0.f, 0.f, 0.f, 0.f};
    int nz_group = -1;
#pragma unroll 1
    for (int j = 0; j < n_x + 2; ++j) {
      if (j < n_x) {
        // layer 1 of dose j, columns [32c, 32c+32) -> A2
        const float xv = E.x_values ? E.x_values[j] : (j == 0 ? 1.f : 0.f);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t hi[16], lo[16];
          const float2 xv2 = make_float2(xv, xv);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 ww = *reinterpret_cast<const float4*>(wx + h * 16 + i * 4);
            const float2 a0 = leaky_x2(__ffma2_rn(xv2, make_float2(ww.x, ww.y),
                                                  make_float2(base[h * 16 + i * 4 + 0], base[h * 16 + i * 4 + 1])));
            const float2 a1 = leaky_x2(__ffma2_rn(xv2, make_float2(ww.z, ww.w),
                                                  make_float2(base[h * 16 + i * 4 + 2], base[h * 16 + i * 4 + 3])));
            float2 h0, l0, h1, l1;
            umma::split_tf32_x2(a0, h0, l0);
            umma::split_tf32_x2(a1, h1, l1);
            hi[i * 4 + 0] = __float_as_uint(h0.x); hi[i * 4 + 1] = __float_as_uint(h0.y);
            hi[i * 4 + 2] = __float_as_uint(h1.x); hi[i * 4 + 3] = __float_as_uint(h1.y);
            lo[i * 4 + 0] = __float_as_uint(l0.x); lo[i * 4 + 1] = __float_as_uint(l0.y);
            lo[i * 4 + 2] = __float_as_uint(l1.x); lo[i * 4 + 3] = __float_as_uint(l1.y);
          }
          umma::st16(trow + EF_A2_HI + c * 32 + h * 16, hi);
          umma::st16(trow + EF_A2_LO + c * 32 + h * 16, lo);
        }
      }
      const int jd = j - 2;                                 // dose whose layer-3 output is ready
      // this warp finishes dose jd (binary: the c == 1 warps, which write y(1) - y(0), do both)
      const bool head = jd >= 0 && ((jd & 1) == c || (P.binary && jd == 0));
      uint32_t r3[16];
      if (head) umma::ld16(trow + EF_D3, r3);
      if (j >= 1 && j <= n_x) {
        // layer-2 output of dose j-1, columns [16c, 16c+16) -> bias, LeakyReLU, split -> A3
        uint32_t r[16];
        umma::ld16(trow + EF_D2 + c * 16, r);
        umma::wait_ld();
        act_store16(r, wimg + P.fb2 + c * 16, trow + EF_A3_HI + c * 16, trow + EF_A3_LO + c * 16);
      } else {
        umma::wait_ld();
      }
      if (j <= n_x) {
        umma::wait_st();
        umma::fence_before_sync();
        tile_publish256(slot, issuer_warp);
        if (issuer_warp) {
          if (umma::elect_one()) {
            umma::fence_after_sync();
            if (j < n_x)
              umma::issue_layer<32, 8>(tbase + EF_D2, tbase + EF_A2_HI, tbase + EF_A2_LO,
                                       wimg_s + 4u * (uint32_t)P.f2_hi, wimg_s + 4u * (uint32_t)P.f2_lo);
            if (j >= 1)
              umma::issue_layer<16, 4>(tbase + EF_D3, tbase + EF_A3_HI, tbase + EF_A3_LO,
                                       wimg_s + 4u * (uint32_t)P.f3_hi, wimg_s + 4u * (uint32_t)P.f3_lo);
            umma::mma_commit(bar);
          }
          __syncwarp();
        }
      }
      if (head) {
        // dose jd: layer 4 (8 -> 2), optional y ~ N(mu, sigma^2) (:703-708), reduction
        float mu = wimg[P.fb4], raw = wimg[P.fb4 + 1];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float a = leaky_mx(__uint_as_float(r3[k]) + wimg[P.fb3 + k]);
          const float2 wf = *reinterpret_cast<const float2*>(wimg + P.fW4 + k * 2);
          mu = fmaf(a, wf.x, mu);
          raw = fmaf(a, wf.y, raw);
        }
        float y = mu;
        if (E.heads) {            // heads mode: (mu_y, raw sigma head) per (state, dose), nothing else
          if (valid && (jd & 1) == c)
            reinterpret_cast<float2*>(E.heads)[((size_t)s * n + row) * n_x + jd] = make_float2(mu, raw);
        } else {
        if (E.sample_y) {
          const float s2 = P.s2y >= 0.f ? P.s2y : softplus_f(raw) + 1e-6f;
          float e;
          if (E.noise) {
            e = E.noise[((size_t)jd * E.n_keep + s) * n + lrow];
          } else {
            // doses 4g..4g+3 share one Philox block: component jd & 3 of normal4(.., g)
            if ((jd >> 2) != nz_group) {
              nz_group = jd >> 2;
              normal4(E.seed, grow, (uint32_t)s, NOISE_EFFECT, (uint32_t)nz_group, nz);
            }
            const int k4 = jd & 3;
            e = k4 == 0 ? nz[0] : (k4 == 1 ? nz[1] : (k4 == 2 ? nz[2] : nz[3]));
          }
          y = fmaf(sqrtf(s2), e, y);
        }
        if (P.binary) {                                                            // :731
          if (jd == 0) y_prev = y;
          else if (valid) E.ite[(size_t)s * n + row] = y_prev - y;
        } else {                                                                   // :759
          float part = valid ? y : 0.f;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
          if (lane == 0) atomicAdd(E.adrf_sum + (size_t)jd * E.n_keep + s, (double)part);
        }
        }
      }
      if (j <= n_x) {
        umma::mbar_wait(bar, parity);
        parity ^= 1u;
        umma::fence_after_sync();
      }
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc512(tmem0);
}

}  // namespace bgm
