// CausalBGM posterior path on BAYESIAN networks (`use_bnn=True`): networks/bnn.py:4-38
// (`BayesianFullyConnectedNet` = input BatchNormalization on BATCH statistics + tfp DenseFlipout
// layers) under causalbgm/base.py:765-817 (get_log_posterior), :820-904 (MH sampler) and :671-763
// (infer_from_latent_posterior).
//
// What changes against the deterministic engines (causal.cuh / causal_tc.cuh):
//  * every network CALL draws a fresh kernel perturbation dW = sigma * eps (shared by the rows of the
//    batch) and fresh per-row sign vectors:  y = x loc + ((x o s_in) dW) o s_out + b  -- twice the
//    multiply-adds, weights that change per call (no resident image, no QR projection);
//  * the input BatchNormalization uses the statistics of the CALL's batch, so every evaluation is
//    preceded by a reduction over all rows of the slice: the chains are coupled, the sampler is one
//    launch per iteration (stream order is the grid-wide barrier) and the current state's
//    log-posterior is re-evaluated with fresh noise every iteration like :866 (no caching).
//
// Execution plan: THREAD = ROW, 256 rows per CTA.  A layer is processed in chunks of 32 output
// columns: the CTA stages loc[K][32] and dW[K][32] = sigma * N(0,1) (Philox, generated in place --
// every CTA regenerates the same dW from the same counters) in shared memory, then each thread
// runs its row: activations of the previous layer in a shared-memory column act[k][tid]
// (conflict-free), weights as warp-uniform LDS.128 broadcasts, 32 accumulators in registers; the
// perturbation product first, sign flip, then the loc product into the same accumulators.
// The batch statistics of iteration t+1 (of its proposal z + q_sd*eps and of the state) are
// accumulated at the end of launch t as per-CTA partial sums and folded, in a fixed order, by the
// prologue of launch t+1: deterministic for a given n.
//
// Noise streams (Philox4x32-10, restated in oracle/bnn.py `PhiloxFlipout`):
//   eps   element (k,c) of (net, layer): normal ((k*N4+c) & 3) of normal4(seed, row = slice<<44 |
//         net<<40 | layer<<36 | (k*N4+c)>>2, t = call, NOISE_BNN_W, 0), N4 = ceil4(N)
//   signs bit i of the blocks noise_block(seed, global row, t = call, NOISE_BNN_SIGN,
//         j = net<<8 | layer<<4 | blk): s_in[k] = bit k, s_out[c] = bit K+c, set = -1
//   call  = 2t (proposal) / 2t+1 (current state) in the sampler; s*n_x + j in the effect kernel.
#pragma once
#include "bnn_noise.cuh"

namespace bgm {
namespace bnn {

constexpr int BNN_THREADS = 256;
constexpr int BNN_MAXL = 8;
constexpr int BNN_MAXK = 64;

struct BnnLayer {
  int K, N, N32;                       // in, out, out padded to a multiple of 32
  int loc_off, scale_off, bias_off;    // float offsets into the device image: loc[K][N32], sigma[K][N32], bias[N32]
};
struct BnnNet {
  int L, kin;
  int bn_off;                          // gamma[kin] | beta[kin]
  BnnLayer layer[BNN_MAXL];
};
struct BnnProgram {
  int zd, d0, d1, d2, p, binary;
  float s2v, s2x, s2y;                 // fixed variances (sigma^2) or < 0: learned softplus head
  BnnNet g, f, h;
};

constexpr int ACT_FLOATS = BNN_MAXK * BNN_THREADS;          // one activation buffer [64][rows per CTA], sized for 256
constexpr int W_FLOATS = BNN_MAXK * 32;                     // one weight chunk [64][32]

__device__ __forceinline__ float flip(float a, uint32_t bit) {
  return __uint_as_float(__float_as_uint(a) ^ (bit << 31));
}

// CTA-wide: loc and dW = sigma * eps of output columns [32c, 32c+32) of a layer -> Wl / Wd [K][32]
__device__ __forceinline__ void stage_chunk(const BnnLayer& Ly, const float* __restrict__ image, float* Wl, float* Wd,
                                            uint64_t seed, int slice, int net_id, int l, uint32_t call, int c, int tid) {
  const int N4 = (Ly.N + 3) & ~3;
  const int64_t base = ((int64_t)slice << 44) | ((int64_t)net_id << 40) | ((int64_t)l << 36);
  for (int i = tid; i < Ly.K * 8; i += (int)blockDim.x) {
    const int k = i >> 3, q = i & 7;
    const int col = c * 32 + q * 4;
    const float4 lo = __ldg(reinterpret_cast<const float4*>(image + Ly.loc_off + (size_t)k * Ly.N32 + col));
    float4 dw = make_float4(0.f, 0.f, 0.f, 0.f);
    if (col < N4) {
      const float4 sc = __ldg(reinterpret_cast<const float4*>(image + Ly.scale_off + (size_t)k * Ly.N32 + col));
      float e[4];
      normal4(seed, base | (int64_t)((k * N4 + col) >> 2), call, NOISE_BNN_W, 0, e);
      dw = make_float4(sc.x * e[0], sc.y * e[1], sc.z * e[2], sc.w * e[3]);
    }
    *reinterpret_cast<float4*>(Wl + k * 32 + q * 4) = lo;
    *reinterpret_cast<float4*>(Wd + k * 32 + q * 4) = dw;
  }
}

// One chunk of W (32 or 8) output columns for the thread's row.
template <int W>
__device__ __forceinline__ void chunk_mac(const float* __restrict__ in, const float* __restrict__ Wl,
                                          const float* __restrict__ Wd, int K, uint64_t sin, uint32_t sout,
                                          const float* __restrict__ bias, int tid, int nth, float (&acc)[W]) {
#pragma unroll
  for (int j = 0; j < W; ++j) acc[j] = 0.f;
#pragma unroll 4
  for (int k = 0; k < K; ++k) {                       // ((x o s_in) dW)
    const float a = flip(in[k * nth + tid], (uint32_t)((sin >> k) & 1ull));
    const float4* w = reinterpret_cast<const float4*>(Wd + k * 32);
#pragma unroll
    for (int q = 0; q < W / 4; ++q) {
      const float4 w4 = w[q];
      acc[q * 4 + 0] = fmaf(a, w4.x, acc[q * 4 + 0]);
      acc[q * 4 + 1] = fmaf(a, w4.y, acc[q * 4 + 1]);
      acc[q * 4 + 2] = fmaf(a, w4.z, acc[q * 4 + 2]);
      acc[q * 4 + 3] = fmaf(a, w4.w, acc[q * 4 + 3]);
    }
  }
#pragma unroll
  for (int j = 0; j < W; ++j) acc[j] = flip(acc[j], (sout >> j) & 1u);   // o s_out
#pragma unroll 4
  for (int k = 0; k < K; ++k) {                       // + x loc
    const float a = in[k * nth + tid];
    const float4* w = reinterpret_cast<const float4*>(Wl + k * 32);
#pragma unroll
    for (int q = 0; q < W / 4; ++q) {
      const float4 w4 = w[q];
      acc[q * 4 + 0] = fmaf(a, w4.x, acc[q * 4 + 0]);
      acc[q * 4 + 1] = fmaf(a, w4.y, acc[q * 4 + 1]);
      acc[q * 4 + 2] = fmaf(a, w4.z, acc[q * 4 + 2]);
      acc[q * 4 + 3] = fmaf(a, w4.w, acc[q * 4 + 3]);
    }
  }
#pragma unroll
  for (int j = 0; j < W; ++j) acc[j] += __ldg(bias + j);
}

struct NetCtx {
  const float* image;
  float *actA, *actB, *Wl, *Wd;      // shared memory
  uint64_t seed;
  int slice;
  int64_t grow;                      // global row of this thread (sign streams)
  int tid;
  int nth;                           // rows (threads) per CTA = stride of the activation columns
};

// Forward pass of one Bayesian net for the thread's row.  The (already batch-normalised) input must
// be in X.actA[k][tid], k < net.kin.  Hidden layers ping-pong between actA / actB; the final layer's
// accumulators are handed chunk by chunk to fin(c, acc, W) (W = 32 or 8 valid-width accumulators).
template <class Fin>
__device__ __forceinline__ void net_forward(const BnnNet& net, int net_id, const NetCtx& X, uint32_t call, Fin&& fin) {
  float* in = X.actA;
  float* out = X.actB;
  for (int l = 0; l < net.L; ++l) {
    const BnnLayer& Ly = net.layer[l];
    const bool last = l == net.L - 1;
    const uint4 b0 = noise_block(X.seed, X.grow, call, NOISE_BNN_SIGN, ((uint32_t)net_id << 8) | ((uint32_t)l << 4));
    const uint64_t sin = ((uint64_t)b0.y << 32) | (uint64_t)b0.x;
    const int nchunk = Ly.N32 >> 5;
    for (int c = 0; c < nchunk; ++c) {
      __syncthreads();                                  // the previous chunk's readers are done with Wl / Wd
      stage_chunk(Ly, X.image, X.Wl, X.Wd, X.seed, X.slice, net_id, l, call, c, X.tid);
      __syncthreads();
      const uint32_t sout = sign_bits32(X.seed, X.grow, call, net_id, l, Ly.K + c * 32);
      const float* bias = X.image + Ly.bias_off + c * 32;
      if (Ly.N <= 8) {                                  // narrow layers (32 -> 8, 8 -> 2): 8 accumulators
        float acc[8];
        chunk_mac<8>(in, X.Wl, X.Wd, Ly.K, sin, sout, bias, X.tid, X.nth, acc);
        if (!last) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (j < Ly.N) out[j * X.nth + X.tid] = leaky(acc[j]);
        } else {
          fin(c, acc, 8);
        }
      } else {
        float acc[32];
        chunk_mac<32>(in, X.Wl, X.Wd, Ly.K, sin, sout, bias, X.tid, X.nth, acc);
        if (!last) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (c * 32 + j < Ly.N) out[(c * 32 + j) * X.nth + X.tid] = leaky(acc[j]);
        } else {
          fin(c, acc, 32);
        }
      }
    }
    float* t = in; in = out; out = t;
  }
}

// Batch statistics handed to an evaluation: mean / 1/sqrt(var + 1e-3) of every z column and of x.
struct ColStats {
  const float* zmean;   // [zd]
  const float* zinv;    // [zd]
  float xmean, xinv;
};

__device__ __forceinline__ float bn_apply(float v, float mean, float inv, float gamma, float beta) {
  return (v - mean) * inv * gamma + beta;
}

// -log p terms of causalbgm/base.py:800-812 for one row; z in registers.
template <int ZMAX>
__device__ __forceinline__ float eval_logpost(const BnnProgram& P, const NetCtx& X, const float (&z)[ZMAX], float x_l,
                                              float y_l, const float* __restrict__ vrow, int ldv, const ColStats& S,
                                              uint32_t call) {
  const float* img = X.image;
  const int zd = P.zd, d0 = P.d0, d1 = P.d1, d2 = P.d2;
  float prior = 0.f;
#pragma unroll
  for (int d = 0; d < ZMAX; ++d)
    if (d < zd) prior = fmaf(z[d], z[d], prior);
  prior *= 0.5f;
  // ---- g_net(z): covariate model (:779-784, :800) ----
  {
    const float* gm = img + P.g.bn_off;
#pragma unroll
    for (int d = 0; d < ZMAX; ++d)
      if (d < zd) X.actA[d * X.nth + X.tid] = bn_apply(z[d], S.zmean[d], S.zinv[d], gm[d], gm[zd + d]);
  }
  float sse = 0.f, raw_v = 0.f;
  net_forward(P.g, NET_G, X, call, [&](int c, const float* acc, int W) {
    const int c0 = c * 32;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      if (q * 4 < W) {
        const int col = c0 + q * 4;
        float4 v4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (col < ldv) v4 = __ldg(reinterpret_cast<const float4*>(vrow + col));
        const float vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int cc = col + i;
          if (cc < P.p) {
            const float dd = vv[i] - acc[q * 4 + i];
            sse = fmaf(dd, dd, sse);
          } else if (cc == P.p) {
            raw_v = acc[q * 4 + i];
          }
        }
      }
    }
  });
  const float s2v = P.s2v >= 0.f ? P.s2v : softplus_f(raw_v) + 1e-6f;
  const float loss_pv = sse / (2.f * s2v) + ((float)P.p * logf(s2v)) / 2.f;
  // ---- h_net([z0, z2]): treatment model (:786-791, :803-807) ----
  {
    const float* hm = img + P.h.bn_off;
    const int kin = P.h.kin;
#pragma unroll
    for (int d = 0; d < ZMAX; ++d) {
      int idx = -1;
      if (d < d0) idx = d;
      else if (d >= d0 + d1 && d < d0 + d1 + d2) idx = d - d1;
      if (idx >= 0 && d < zd) X.actA[idx * X.nth + X.tid] = bn_apply(z[d], S.zmean[d], S.zinv[d], hm[idx], hm[kin + idx]);
    }
  }
  float mu_x = 0.f, raw_x = 0.f;
  net_forward(P.h, NET_H, X, call, [&](int, const float* acc, int) { mu_x = acc[0]; raw_x = acc[1]; });
  float loss_px;
  if (P.binary) {
    loss_px = fmaxf(mu_x, 0.f) - mu_x * x_l + log1pf(expf(-fabsf(mu_x)));
  } else {
    const float s2x = P.s2x >= 0.f ? P.s2x : softplus_f(raw_x) + 1e-6f;
    const float dx = x_l - mu_x;
    loss_px = (dx * dx) / (2.f * s2x) + logf(s2x) / 2.f;
  }
  // ---- f_net([z0, z1, x]): outcome model (:793-798, :809-810) ----
  {
    const float* fm = img + P.f.bn_off;
    const int kin = P.f.kin;
#pragma unroll
    for (int d = 0; d < ZMAX; ++d)
      if (d < d0 + d1 && d < zd) X.actA[d * X.nth + X.tid] = bn_apply(z[d], S.zmean[d], S.zinv[d], fm[d], fm[kin + d]);
    X.actA[(d0 + d1) * X.nth + X.tid] = bn_apply(x_l, S.xmean, S.xinv, fm[d0 + d1], fm[kin + d0 + d1]);
  }
  float mu_y = 0.f, raw_y = 0.f;
  net_forward(P.f, NET_F, X, call, [&](int, const float* acc, int) { mu_y = acc[0]; raw_y = acc[1]; });
  const float s2y = P.s2y >= 0.f ? P.s2y : softplus_f(raw_y) + 1e-6f;
  const float dy = y_l - mu_y;
  const float loss_py = (dy * dy) / (2.f * s2y) + logf(s2y) / 2.f;
  return -(((loss_pv + loss_px) + loss_py) + prior);                     // :814-816
}

// ---------------------------------------------------------------------------------------
// Per-CTA partial sums of the batch statistics: NP = 4*ZMAX + 2 values per CTA,
//   [0,ZMAX) sum z'   [ZMAX,2ZMAX) sum z'^2   [2ZMAX,3ZMAX) sum z   [3ZMAX,4ZMAX) sum z^2   then sum x, sum x^2
template <int ZMAX>
__device__ __forceinline__ void stats_partial(const float (&zp)[ZMAX], const float (&zc)[ZMAX], float x_l, bool valid,
                                              int zd, float* red /* smem [8][NP] */, double* part /* global [NP] */) {
  constexpr int NP = 4 * ZMAX + 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  auto put = [&](int i, float v) {
    v = valid ? v : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp * NP + i] = v;
  };
#pragma unroll
  for (int d = 0; d < ZMAX; ++d) {
    if (d < zd) {
      put(d, zp[d]);
      put(ZMAX + d, zp[d] * zp[d]);
      put(2 * ZMAX + d, zc[d]);
      put(3 * ZMAX + d, zc[d] * zc[d]);
    }
  }
  put(4 * ZMAX, x_l);
  put(4 * ZMAX + 1, x_l * x_l);
  __syncthreads();
  for (int i = threadIdx.x; i < NP; i += (int)blockDim.x) {
    const int d = i % ZMAX;
    double s = 0.0;
    if (i >= 4 * ZMAX || d < zd)
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += (double)red[w * NP + i];
    part[i] = s;
  }
  __syncthreads();
}

struct BnnMhDev {
  bgm_mh_args a;
  int mode;            // 0: one MH iteration t; 1: log-posterior of z_in (call id = call0); 2: statistics only
  int t;               // iteration of this launch
  int slice;
  uint32_t call0;      // mode 1
  const float* z_in;   // mode 1: (n, zd)
  float* out_lp;       // mode 1: (n)
  double* part;        // [2][ncta][NP] partial sums; launch t reads parity t&1, writes (t+1)&1
  float* lp_cur_trace; // (T, n) or NULL
  float* dw;           // plan 2: [2 parity][2 evaluations][image floats] kernel perturbations of the iteration
};

template <int ZMAX>
__device__ __forceinline__ void propose(const bgm_mh_args& A, int t, int lrow, int64_t grow, int zd, double q_sd,
                                        const float (&zc)[ZMAX], float (&zp)[ZMAX]) {
  if (A.eps_dev) {
    const float* e = A.eps_dev + ((size_t)t * A.n + lrow) * zd;
#pragma unroll
    for (int d = 0; d < ZMAX; ++d) zp[d] = d < zd ? __fadd_rn(zc[d], (float)(q_sd * (double)e[d])) : 0.f;
  } else {
#pragma unroll
    for (int g = 0; g < ZMAX / 4; ++g) {
      if (g * 4 < zd) {
        float e[4];
        normal4(A.seed, grow, (uint32_t)t, NOISE_PROPOSAL, g, e);
#pragma unroll
        for (int q = 0; q < 4; ++q) zp[g * 4 + q] = (g * 4 + q < zd) ? __fadd_rn(zc[g * 4 + q], (float)(q_sd * (double)e[q])) : 0.f;
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) zp[g * 4 + q] = 0.f;
      }
    }
  }
}

template <int ZMAX>
__global__ void __launch_bounds__(BNN_THREADS, 1)
bnn_mh_kernel(const __grid_constant__ BnnProgram P, const float* __restrict__ image, const __grid_constant__ BnnMhDev D) {
  constexpr int NP = 4 * ZMAX + 2;
  extern __shared__ __align__(16) float smem[];
  float* actA = smem;
  float* actB = actA + ACT_FLOATS;
  float* Wl = actB + ACT_FLOATS;
  float* Wd = Wl + W_FLOATS;
  float* red = Wd + W_FLOATS;                       // [8][NP]
  float* st = red + 8 * NP;                         // zmean_p | zinv_p | zmean_c | zinv_c (ZMAX each) | xmean, xinv
  const bgm_mh_args& A = D.a;
  const int tid = threadIdx.x;
  const int n = A.n, zd = P.zd;
  const int ncta = gridDim.x;
  const int nth = (int)blockDim.x;     // rows per CTA: chosen by the host so that one wave of CTAs covers the slice
  const int row = blockIdx.x * nth + tid;
  const bool valid = row < n;
  const int lrow = valid ? row : n - 1;
  const int64_t grow = A.row_offset + lrow;
  const int t = D.t;
  const double q_sd = A.q_sd_dev ? *A.q_sd_dev : 1.0;
  const float x_l = A.x_dev[lrow], y_l = A.y_dev[lrow];

  float zc[ZMAX], zp[ZMAX];
  if (D.mode == 2) {
    // statistics of iteration t from the stored state (or the initial draw, :842)
    if (A.init_mode == 2) {
#pragma unroll
      for (int g = 0; g < ZMAX / 4; ++g) {
        float e[4] = {0.f, 0.f, 0.f, 0.f};
        if (g * 4 < zd) normal4(A.seed, grow, T_INIT, NOISE_PROPOSAL, g, e);
#pragma unroll
        for (int q = 0; q < 4; ++q) zc[g * 4 + q] = (g * 4 + q < zd) ? e[q] : 0.f;
      }
      if (valid)
#pragma unroll
        for (int d = 0; d < ZMAX; ++d)
          if (d < zd) A.z_state_dev[(size_t)row * zd + d] = zc[d];
    } else {
      const float* src = D.z_in ? D.z_in : A.z_state_dev;
#pragma unroll
      for (int d = 0; d < ZMAX; ++d) zc[d] = d < zd ? src[(size_t)lrow * zd + d] : 0.f;
    }
    if (D.z_in) {
#pragma unroll
      for (int d = 0; d < ZMAX; ++d) zp[d] = zc[d];
    } else {
      propose<ZMAX>(A, t, lrow, grow, zd, q_sd, zc, zp);
    }
    stats_partial<ZMAX>(zp, zc, x_l, valid, zd, red, D.part + ((size_t)(t & 1) * ncta + blockIdx.x) * NP);
    return;
  }

  // ---- prologue: fold the per-CTA partial sums of this iteration (fixed order) ----
  {
    const double* part = D.part + (size_t)(t & 1) * ncta * NP;
    double* sums = reinterpret_cast<double*>(actA);   // scratch before the activations are written
    for (int i = tid; i < NP; i += nth) {
      double s = 0.0;
      for (int b = 0; b < ncta; ++b) s += part[(size_t)b * NP + i];
      sums[i] = s;
    }
    __syncthreads();
    const double inv_n = 1.0 / (double)n;
    for (int i = tid; i < 2 * ZMAX; i += nth) {
      const int which = i / ZMAX, d = i % ZMAX;           // 0: proposal, 1: current state
      if (d < zd) {
        const double m = sums[which * 2 * ZMAX + d] * inv_n;
        const double var = fmax(sums[which * 2 * ZMAX + ZMAX + d] * inv_n - m * m, 0.0);
        st[which * 2 * ZMAX + d] = (float)m;
        st[which * 2 * ZMAX + ZMAX + d] = 1.f / sqrtf((float)var + 1e-3f);
      }
    }
    if (tid == 0) {
      const double m = sums[4 * ZMAX] * inv_n;
      const double var = fmax(sums[4 * ZMAX + 1] * inv_n - m * m, 0.0);
      st[4 * ZMAX] = (float)m;
      st[4 * ZMAX + 1] = 1.f / sqrtf((float)var + 1e-3f);
    }
    __syncthreads();
  }
  NetCtx X;
  X.image = image; X.actA = actA; X.actB = actB; X.Wl = Wl; X.Wd = Wd;
  X.seed = A.seed; X.slice = D.slice; X.grow = grow; X.tid = tid; X.nth = nth;
  const float* vrow = A.v_dev + (size_t)lrow * A.ldv;
  ColStats Sp{st, st + ZMAX, st[4 * ZMAX], st[4 * ZMAX + 1]};
  ColStats Sc{st + 2 * ZMAX, st + 3 * ZMAX, st[4 * ZMAX], st[4 * ZMAX + 1]};

  if (D.mode == 1) {
#pragma unroll
    for (int d = 0; d < ZMAX; ++d) zc[d] = d < zd ? D.z_in[(size_t)lrow * zd + d] : 0.f;
    const float lp = eval_logpost<ZMAX>(P, X, zc, x_l, y_l, vrow, A.ldv, Sc, D.call0);
    if (valid) D.out_lp[row] = lp;
    return;
  }

#pragma unroll
  for (int d = 0; d < ZMAX; ++d) zc[d] = d < zd ? A.z_state_dev[(size_t)lrow * zd + d] : 0.f;
  propose<ZMAX>(A, t, lrow, grow, zd, q_sd, zc, zp);
  const float lp_prop = eval_logpost<ZMAX>(P, X, zp, x_l, y_l, vrow, A.ldv, Sp, 2u * (uint32_t)t);       // :865
  const float lp_cur = eval_logpost<ZMAX>(P, X, zc, x_l, y_l, vrow, A.ldv, Sc, 2u * (uint32_t)t + 1u);   // :866
  const float dlp = lp_prop - lp_cur;
  const float ratio = (dlp < 0.f) ? expf(dlp) : ((dlp >= 0.f) ? 1.f : __int_as_float(0x7fc00000));      // :868
  bool acc;
  if (A.u_dev) acc = A.u_dev[(size_t)t * n + lrow] < (double)ratio;                                      // :870
  else acc = uniform1(A.seed, grow, (uint32_t)t, NOISE_ACCEPT) < ratio;
  if (acc) {
#pragma unroll
    for (int d = 0; d < ZMAX; ++d) zc[d] = zp[d];                                                        // :871
  }
  if (valid) {
    if (acc)
#pragma unroll
      for (int d = 0; d < ZMAX; ++d)
        if (d < zd) A.z_state_dev[(size_t)row * zd + d] = zc[d];
    if (A.accept_mask_dev) A.accept_mask_dev[(size_t)t * n + row] = acc ? 1 : 0;
    if (A.lp_trace_dev) A.lp_trace_dev[(size_t)t * n + row] = lp_prop;
    if (D.lp_cur_trace) D.lp_cur_trace[(size_t)t * n + row] = lp_cur;
    if (t >= A.burn_in && A.out_samples_dev) {                                                           // :895-896
      float* dst = A.out_samples_dev + ((size_t)(t - A.burn_in) * n + row) * zd;
#pragma unroll
      for (int d = 0; d < ZMAX; ++d)
        if (d < zd) dst[d] = zc[d];
    }
  }
  if (A.accept_count_dev) {
    const unsigned b = __ballot_sync(0xffffffffu, acc && valid);
    if ((tid & 31) == 0 && b) atomicAdd(A.accept_count_dev + t, __popc(b));
  }
  // ---- statistics of iteration t+1 (its proposal and its current state) ----
  if (t + 1 < A.t_end) {
    __syncthreads();
    propose<ZMAX>(A, t + 1, lrow, grow, zd, q_sd, zc, zp);
    stats_partial<ZMAX>(zp, zc, x_l, valid, zd, red, D.part + ((size_t)((t + 1) & 1) * ncta + blockIdx.x) * NP);
  }
}

// ---------------------------------------------------------------------------------------
// infer_from_latent_posterior (:671-763): per kept state s the batch statistics of its z0 / z1
// columns over the n rows ...
__global__ void __launch_bounds__(256)
bnn_sample_stats_kernel(const float* __restrict__ zs, int n_keep, int n, int zd, int ncols, float* __restrict__ stats) {
  // one CTA per kept state; stats[s][0..ncols) = mean, [ncols..2 ncols) = 1/sqrt(var + 1e-3)
  __shared__ double red[2][8];
  const int s = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* base = zs + (size_t)s * n * zd;
  for (int d = 0; d < ncols; ++d) {
    double a = 0.0, b = 0.0;
    for (int r = threadIdx.x; r < n; r += 256) {
      const double v = (double)base[(size_t)r * zd + d];
      a += v;
      b += v * v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (lane == 0) { red[0][warp] = a; red[1][warp] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double sa = 0.0, sb = 0.0;
      for (int w = 0; w < 8; ++w) { sa += red[0][w]; sb += red[1][w]; }
      const double m = sa / n, var = fmax(sb / n - m * m, 0.0);
      stats[(size_t)s * 2 * ncols + d] = (float)m;
      stats[(size_t)s * 2 * ncols + ncols + d] = 1.f / sqrtf((float)var + 1e-3f);
    }
    __syncthreads();
  }
}

struct BnnEffectDev {
  const float* zs;          // (n_keep, n, zd)
  const float* stats;       // (n_keep, 2*(d0+d1))
  const float* x_values;    // (n_x) or NULL (binary: doses 1, 0)
  const float* noise;       // optional injected N(0,1): (n_x, n_keep, n)
  double* adrf_sum;         // (n_x, n_keep)
  float* ite;               // (n_keep, n)
  int n_keep, n, n_x, sample_y;
  uint64_t seed;
  int64_t row_offset;
};

// ... and one f_net CALL per (kept state, dose): grid = (row blocks, n_keep); a CTA runs its 256 rows
// of state s through every dose.  The tiled dose column is constant over the batch: batch variance 0,
// normalised value 0, so the net sees beta for it -- the reference's behaviour restated
// (oracle/bnn.py).
__global__ void __launch_bounds__(BNN_THREADS, 1)
bnn_effect_kernel(const __grid_constant__ BnnProgram P, const float* __restrict__ image,
                  const __grid_constant__ BnnEffectDev E) {
  extern __shared__ __align__(16) float smem[];
  float* actA = smem;
  float* actB = actA + ACT_FLOATS;
  float* Wl = actB + ACT_FLOATS;
  float* Wd = Wl + W_FLOATS;
  float* red = Wd + W_FLOATS;     // [8]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int s = blockIdx.y;
  const int n = E.n;
  const int row = blockIdx.x * BNN_THREADS + tid;
  const bool valid = row < n;
  const int lrow = valid ? row : n - 1;
  const int64_t grow = E.row_offset + lrow;
  const int d0 = P.d0, d1 = P.d1, kin = P.f.kin, nc = d0 + d1;
  NetCtx X;
  X.image = image; X.actA = actA; X.actB = actB; X.Wl = Wl; X.Wd = Wd;
  X.seed = E.seed; X.slice = 0; X.grow = grow; X.tid = tid; X.nth = BNN_THREADS;
  const float* z = E.zs + ((size_t)s * n + lrow) * P.zd;
  const float* stt = E.stats + (size_t)s * 2 * nc;
  const float* fm = image + P.f.bn_off;
  float zin[BNN_MAXK];            // only the first nc entries are used (nc <= 32 in practice)
#pragma unroll
  for (int d = 0; d < 32; ++d) zin[d] = d < nc ? bn_apply(z[d], stt[d], stt[nc + d], fm[d], fm[kin + d]) : 0.f;
  const float xin = fm[kin + nc];                      // BN of a constant column = beta
  float y_prev = 0.f;
  float nz4[4] = {0.f, 0.f, 0.f, 0.f};
  for (int j = 0; j < E.n_x; ++j) {
    __syncthreads();
#pragma unroll
    for (int d = 0; d < 32; ++d)
      if (d < nc) actA[d * BNN_THREADS + tid] = zin[d];
    actA[nc * BNN_THREADS + tid] = xin;
    float mu = 0.f, raw = 0.f;
    net_forward(P.f, NET_F, X, (uint32_t)(s * E.n_x + j), [&](int, const float* acc, int) { mu = acc[0]; raw = acc[1]; });
    float y = mu;
    if (E.sample_y) {                                                              // :704, :725, :753
      const float s2 = P.s2y >= 0.f ? P.s2y : softplus_f(raw) + 1e-6f;
      float e;
      if (E.noise) {
        e = E.noise[((size_t)j * E.n_keep + s) * n + lrow];
      } else {
        if ((j & 3) == 0) normal4(E.seed, grow, (uint32_t)s, NOISE_EFFECT, (uint32_t)(j >> 2), nz4);
        const int k4 = j & 3;
        e = k4 == 0 ? nz4[0] : (k4 == 1 ? nz4[1] : (k4 == 2 ? nz4[2] : nz4[3]));
      }
      y = fmaf(sqrtf(s2), e, y);
    }
    if (P.binary) {                                                                // :731
      if (j == 0) y_prev = y;
      else if (valid) E.ite[(size_t)s * n + row] = y_prev - y;
    } else {                                                                       // :759
      float part = valid ? y : 0.f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      __syncthreads();
      if (lane == 0) red[warp] = part;
      __syncthreads();
      if (tid == 0) {
        double tot = 0.0;
        for (int w = 0; w < BNN_THREADS / 32; ++w) tot += (double)red[w];
        atomicAdd(E.adrf_sum + (size_t)j * E.n_keep + s, tot);
      }
    }
  }
}

// Test hook: the Flipout noise of one (net, layer, call) as the kernels draw it.
__global__ void bnn_noise_kernel(uint64_t seed, int slice, int net_id, int l, uint32_t call, int K, int N,
                                 int64_t row_offset, int rows, float* eps, signed char* s_in, signed char* s_out) {
  const int N4 = (N + 3) & ~3;
  const int64_t base = ((int64_t)slice << 44) | ((int64_t)net_id << 40) | ((int64_t)l << 36);
  const int gid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
  if (eps)
    for (int g = gid; g < K * N4 / 4; g += gsz) {
      float e[4];
      normal4(seed, base | (int64_t)g, call, NOISE_BNN_W, 0, e);
      const int k = (g * 4) / N4, c = (g * 4) % N4;
      for (int i = 0; i < 4; ++i)
        if (c + i < N) eps[(size_t)k * N + c + i] = e[i];
    }
  for (int r = gid; r < rows; r += gsz) {
    const int64_t grow = row_offset + r;
    if (s_in)
      for (int k0 = 0; k0 < K; k0 += 32) {
        const uint32_t b = sign_bits32(seed, grow, call, net_id, l, k0);
        for (int i = 0; i < 32 && k0 + i < K; ++i) s_in[(size_t)r * K + k0 + i] = ((b >> i) & 1u) ? -1 : 1;
      }
    if (s_out)
      for (int c0 = 0; c0 < N; c0 += 32) {
        const uint32_t b = sign_bits32(seed, grow, call, net_id, l, K + c0);
        for (int i = 0; i < 32 && c0 + i < N; ++i) s_out[(size_t)r * N + c0 + i] = ((b >> i) & 1u) ? -1 : 1;
      }
  }
}

}  // namespace bnn
}  // namespace bgm
