// bnn_mh2_kernel: the Bayesian-net sampler of bnn.cuh (same contract, same noise streams, same arithmetic per
// accumulator) with a different execution plan, chosen from the ncu capture of plan 1
// (profiles/r02j_bnn_mh_kernel_*: fp32 pipe 22 % busy with 5-8 warps per SM; stalls: LDS latency 17 %, barrier 16 %,
// instruction fetch 13 %):
//  * TWO threads per row (lanes l and l^16 of a warp): each owns 16 of the 32 columns of a chunk and half of the
//    latent dimensions, so an SM holds the same rows with twice the warps (and two CTAs per SM: 16 warps);
//    per-row scalars cross with one shuffle;
//  * the three nets of an evaluation are ONE table-driven loop over weight chunks (no per-net / per-evaluation
//    copies of the unrolled inner loops: a quarter of the SASS);
//  * weight chunks are double-buffered: a thread stages its share of chunk i+1 right after computing chunk i,
//    one __syncthreads per chunk instead of two.
#pragma once
#include "bnn.cuh"

namespace bgm {
namespace bnn {

constexpr int B2_MAXCH = 48;
enum : unsigned char { CH_LAYER_FIRST = 1, CH_FINAL = 2, CH_NARROW = 4, CH_NET_FIRST = 8, CH_LAYER_LAST = 16 };

struct BnnChunk {
  int loc_off, scale_off, bias_off;   // of the layer (column offset 32 c added in the kernel)
  short K, N, N32, N4;
  unsigned char c, net, layer, flags;
};
struct BnnProgram2 {
  BnnProgram P;
  int nchunks;
  BnnChunk ch[B2_MAXCH];
};

// CTA-wide: loc and dW = sigma * eps of one chunk -> Wl / Wd [K][32]
__device__ __forceinline__ void stage_chunk2(const BnnChunk& C, const float* __restrict__ image, float* Wl, float* Wd,
                                             uint64_t seed, int slice, uint32_t call, int tid, int nth) {
  const int64_t base = ((int64_t)slice << 44) | ((int64_t)C.net << 40) | ((int64_t)C.layer << 36);
  for (int i = tid; i < C.K * 8; i += nth) {
    const int k = i >> 3, q = i & 7;
    const int col = C.c * 32 + q * 4;
    const float4 lo = __ldg(reinterpret_cast<const float4*>(image + C.loc_off + (size_t)k * C.N32 + col));
    float4 dw = make_float4(0.f, 0.f, 0.f, 0.f);
    if (col < C.N4) {
      const float4 sc = __ldg(reinterpret_cast<const float4*>(image + C.scale_off + (size_t)k * C.N32 + col));
      float e[4];
      normal4(seed, base | (int64_t)((k * C.N4 + col) >> 2), call, NOISE_BNN_W, 0, e);
      dw = make_float4(sc.x * e[0], sc.y * e[1], sc.z * e[2], sc.w * e[3]);
    }
    *reinterpret_cast<float4*>(Wl + k * 32 + q * 4) = lo;
    *reinterpret_cast<float4*>(Wd + k * 32 + q * 4) = dw;
  }
}

// W (16 or 8) output columns starting at column `c0` of the staged chunk, for the thread's row
template <int W>
__device__ __forceinline__ void chunk_mac2(const float* __restrict__ in, const float* __restrict__ Wl,
                                           const float* __restrict__ Wd, int K, uint64_t sin, uint32_t sout, int c0,
                                           const float* __restrict__ bias, int rl, int nrows, float (&acc)[W]) {
#pragma unroll
  for (int j = 0; j < W; ++j) acc[j] = 0.f;
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    const float a = flip(in[k * nrows + rl], (uint32_t)((sin >> k) & 1ull));
    const float4* w = reinterpret_cast<const float4*>(Wd + k * 32 + c0);
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
  for (int j = 0; j < W; ++j) acc[j] = flip(acc[j], (sout >> (c0 + j)) & 1u);
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    const float a = in[k * nrows + rl];
    const float4* w = reinterpret_cast<const float4*>(Wl + k * 32 + c0);
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
  for (int j = 0; j < W; ++j) acc[j] += __ldg(bias + c0 + j);
}

// this half's ZH latent dimensions of the proposal
template <int ZH>
__device__ __forceinline__ void propose2(const bgm_mh_args& A, int t, int lrow, int64_t grow, int zd, int dbase, double q_sd,
                                         const float (&zc)[ZH], float (&zp)[ZH]) {
  if (A.eps_dev) {
    const float* e = A.eps_dev + ((size_t)t * A.n + lrow) * zd;
#pragma unroll
    for (int k = 0; k < ZH; ++k) zp[k] = (dbase + k < zd) ? __fadd_rn(zc[k], (float)(q_sd * (double)e[dbase + k])) : 0.f;
  } else {
#pragma unroll
    for (int g = 0; g < ZH / 4; ++g) {
      float e[4] = {0.f, 0.f, 0.f, 0.f};
      if (dbase + g * 4 < zd) normal4(A.seed, grow, (uint32_t)t, NOISE_PROPOSAL, (uint32_t)(dbase / 4 + g), e);
#pragma unroll
      for (int q = 0; q < 4; ++q)
        zp[g * 4 + q] = (dbase + g * 4 + q < zd) ? __fadd_rn(zc[g * 4 + q], (float)(q_sd * (double)e[q])) : 0.f;
    }
  }
}

// per-CTA partial sums of the batch statistics (layout as stats_partial in bnn.cuh), two threads per row
template <int ZH>
__device__ __forceinline__ void stats_partial2(const float (&zp)[ZH], const float (&zc)[ZH], float x_l, bool valid, int zd,
                                               int dbase, float* red, double* part) {
  constexpr int ZMAX = 2 * ZH, NP = 4 * ZMAX + 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  auto put = [&](int i, float v) {          // sum over the 16 rows of this half-warp
    v = valid ? v : 0.f;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((lane & 15) == 0) red[warp * NP + i] = v;
  };
#pragma unroll
  for (int k = 0; k < ZH; ++k) {
    const int d = dbase + k;             // d < ZMAX always; entries with d >= zd are never read
    put(d, zp[k]);
    put(ZMAX + d, zp[k] * zp[k]);
    put(2 * ZMAX + d, zc[k]);
    put(3 * ZMAX + d, zc[k] * zc[k]);
  }
  if (dbase == 0) {                      // x: the lower half-warp only
    float a = valid ? x_l : 0.f, b = valid ? x_l * x_l : 0.f;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) { a += __shfl_xor_sync(0x0000ffffu, a, o); b += __shfl_xor_sync(0x0000ffffu, b, o); }
    if (lane == 0) { red[warp * NP + 4 * ZMAX] = a; red[warp * NP + 4 * ZMAX + 1] = b; }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NP; i += (int)blockDim.x) {
    const int d = i % ZMAX;
    double s = 0.0;
    if (i >= 4 * ZMAX || d < zd)
      for (int w = 0; w < nwarps; ++w) s += (double)red[w * NP + i];
    part[i] = s;
  }
  __syncthreads();
}

template <int ZH>
__global__ void __launch_bounds__(256, 2)
bnn_mh2_kernel(const __grid_constant__ BnnProgram2 Q, const float* __restrict__ image, const __grid_constant__ BnnMhDev D) {
  constexpr int ZMAX = 2 * ZH, NP = 4 * ZMAX + 2;
  const BnnProgram& P = Q.P;
  extern __shared__ __align__(16) float smem[];
  const int nth = (int)blockDim.x, nrows = nth >> 1;
  float* act0 = smem;                               // [64][nrows]
  float* act1 = act0 + BNN_MAXK * nrows;
  float* Wbuf = act1 + BNN_MAXK * nrows;            // [2][Wl | Wd][64 * 32]
  float* red = Wbuf + 4 * W_FLOATS;                 // [8][NP]
  float* st = red + 8 * NP;                         // zmean_p | zinv_p | zmean_c | zinv_c | xmean, xinv
  const bgm_mh_args& A = D.a;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int half = lane >> 4;
  const int rl = warp * 16 + (lane & 15);
  const int n = A.n, zd = P.zd;
  const int ncta = gridDim.x;
  const int row = blockIdx.x * nrows + rl;
  const bool valid = row < n;
  const int lrow = valid ? row : n - 1;
  const int64_t grow = A.row_offset + lrow;
  const int t = D.t;
  const int dbase = half * ZH;
  const double q_sd = A.q_sd_dev ? *A.q_sd_dev : 1.0;
  const float x_l = A.x_dev[lrow], y_l = A.y_dev[lrow];
  float zc[ZH], zp[ZH];

  if (D.mode == 2) {
    if (A.init_mode == 2) {
#pragma unroll
      for (int g = 0; g < ZH / 4; ++g) {
        float e[4] = {0.f, 0.f, 0.f, 0.f};
        if (dbase + g * 4 < zd) normal4(A.seed, grow, T_INIT, NOISE_PROPOSAL, (uint32_t)(dbase / 4 + g), e);
#pragma unroll
        for (int q = 0; q < 4; ++q) zc[g * 4 + q] = (dbase + g * 4 + q < zd) ? e[q] : 0.f;
      }
      if (valid)
#pragma unroll
        for (int k = 0; k < ZH; ++k)
          if (dbase + k < zd) A.z_state_dev[(size_t)row * zd + dbase + k] = zc[k];
    } else {
      const float* src = D.z_in ? D.z_in : A.z_state_dev;
#pragma unroll
      for (int k = 0; k < ZH; ++k) zc[k] = (dbase + k < zd) ? src[(size_t)lrow * zd + dbase + k] : 0.f;
    }
    if (D.z_in) {
#pragma unroll
      for (int k = 0; k < ZH; ++k) zp[k] = zc[k];
    } else {
      propose2<ZH>(A, t, lrow, grow, zd, dbase, q_sd, zc, zp);
    }
    stats_partial2<ZH>(zp, zc, x_l, valid, zd, dbase, red, D.part + ((size_t)(t & 1) * ncta + blockIdx.x) * NP);
    return;
  }

  // ---- prologue: fold the per-CTA partial sums of this iteration (fixed order) ----
  {
    const double* part = D.part + (size_t)(t & 1) * ncta * NP;
    double* sums = reinterpret_cast<double*>(act0);
    for (int i = tid; i < NP; i += nth) {
      double s = 0.0;
      for (int b = 0; b < ncta; ++b) s += part[(size_t)b * NP + i];
      sums[i] = s;
    }
    __syncthreads();
    const double inv_n = 1.0 / (double)n;
    for (int i = tid; i < 2 * ZMAX; i += nth) {
      const int which = i / ZMAX, d = i % ZMAX;
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
  const float xmean = st[4 * ZMAX], xinv = st[4 * ZMAX + 1];
  const float* vrow = A.v_dev + (size_t)lrow * A.ldv;
  const int d0 = P.d0, d1 = P.d1, d2 = P.d2;
  const int n_eval = D.mode == 1 ? 1 : 2;

  if (D.mode == 1) {
#pragma unroll
    for (int k = 0; k < ZH; ++k) { zc[k] = (dbase + k < zd) ? D.z_in[(size_t)lrow * zd + dbase + k] : 0.f; zp[k] = zc[k]; }
  } else {
#pragma unroll
    for (int k = 0; k < ZH; ++k) zc[k] = (dbase + k < zd) ? A.z_state_dev[(size_t)lrow * zd + dbase + k] : 0.f;
    propose2<ZH>(A, t, lrow, grow, zd, dbase, q_sd, zc, zp);
  }

  // writes this half's share of the (batch-normalised) input of net `net` for evaluation e into act0
  auto write_inputs = [&](int net, int e) {
    const bool cur = (D.mode == 1) || e == 1;
    const float* zm = st + (cur ? 2 * ZMAX : 0);
    const float* zi = zm + ZMAX;
    const BnnNet& N = net == NET_G ? P.g : (net == NET_H ? P.h : P.f);
    const float* gm = image + N.bn_off;
    const int kin = N.kin;
#pragma unroll
    for (int k = 0; k < ZH; ++k) {
      const int d = dbase + k;
      const float zv = cur ? zc[k] : zp[k];
      int idx = -1;
      if (net == NET_G) idx = d;
      else if (net == NET_H) { if (d < d0) idx = d; else if (d >= d0 + d1 && d < d0 + d1 + d2) idx = d - d1; }
      else { if (d < d0 + d1) idx = d; }
      if (d < zd && idx >= 0) act0[idx * nrows + rl] = bn_apply(zv, zm[d], zi[d], gm[idx], gm[kin + idx]);
    }
    if (net == NET_F && half == 0) act0[(d0 + d1) * nrows + rl] = bn_apply(x_l, xmean, xinv, gm[d0 + d1], gm[kin + d0 + d1]);
  };
  auto call_of = [&](int e) -> uint32_t { return D.mode == 1 ? D.call0 : 2u * (uint32_t)t + (uint32_t)e; };

  const int nch = Q.nchunks, total = nch * n_eval;
  write_inputs(Q.ch[0].net, 0);
  stage_chunk2(Q.ch[0], image, Wbuf, Wbuf + W_FLOATS, A.seed, D.slice, call_of(0), tid, nth);
  __syncthreads();
  float lp0 = 0.f, lp1 = 0.f;
  float* in = act0;
  float* out = act1;
  float sse = 0.f, raw_v = 0.f, loss_pv = 0.f, loss_px = 0.f, loss_py = 0.f;
  uint64_t sin = 0ull;
#pragma unroll 1
  for (int gi = 0; gi < total; ++gi) {
    const int e = gi >= nch ? 1 : 0;
    const int i = gi - e * nch;
    const BnnChunk& C = Q.ch[i];
    const uint32_t call = call_of(e);
    const float* Wl = Wbuf + (gi & 1) * 2 * W_FLOATS;
    const float* Wd = Wl + W_FLOATS;
    if (C.flags & CH_NET_FIRST) { in = act0; out = act1; sse = 0.f; raw_v = 0.f; }
    if (C.flags & CH_LAYER_FIRST) {
      const uint4 b0 = noise_block(A.seed, grow, call, NOISE_BNN_SIGN, ((uint32_t)C.net << 8) | ((uint32_t)C.layer << 4));
      sin = ((uint64_t)b0.y << 32) | (uint64_t)b0.x;
    }
    const uint32_t sout = sign_bits32(A.seed, grow, call, C.net, C.layer, C.K + C.c * 32);
    const float* bias = image + C.bias_off + C.c * 32;
    if (C.flags & CH_NARROW) {
      float acc[8];
      chunk_mac2<8>(in, Wl, Wd, C.K, sin, sout, 0, bias, rl, nrows, acc);
      if (!(C.flags & CH_FINAL)) {
        if (half == 0)
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (j < C.N) out[j * nrows + rl] = leaky(acc[j]);
      } else if (C.net == NET_H) {
        const float mu_x = acc[0], raw_x = acc[1];
        if (P.binary) {
          loss_px = fmaxf(mu_x, 0.f) - mu_x * x_l + log1pf(expf(-fabsf(mu_x)));
        } else {
          const float s2x = P.s2x >= 0.f ? P.s2x : softplus_f(raw_x) + 1e-6f;
          const float dx = x_l - mu_x;
          loss_px = (dx * dx) / (2.f * s2x) + logf(s2x) / 2.f;
        }
      } else {
        const float mu_y = acc[0], raw_y = acc[1];
        const float s2y = P.s2y >= 0.f ? P.s2y : softplus_f(raw_y) + 1e-6f;
        const float dy = y_l - mu_y;
        loss_py = (dy * dy) / (2.f * s2y) + logf(s2y) / 2.f;
      }
    } else {
      float acc[16];
      const int c0 = half * 16;
      chunk_mac2<16>(in, Wl, Wd, C.K, sin, sout, c0, bias, rl, nrows, acc);
      const int col0 = C.c * 32 + c0;
      if (!(C.flags & CH_FINAL)) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (col0 + j < C.N) out[(col0 + j) * nrows + rl] = leaky(acc[j]);
      } else {                                      // g_net's output layer: squared error against the row's covariates
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int col = col0 + q * 4;
          float4 v4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (col < A.ldv) v4 = __ldg(reinterpret_cast<const float4*>(vrow + col));
          const float vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int cc = col + u;
            if (cc < P.p) {
              const float dd = vv[u] - acc[q * 4 + u];
              sse = fmaf(dd, dd, sse);
            } else if (cc == P.p) {
              raw_v = acc[q * 4 + u];
            }
          }
        }
        if (C.flags & CH_LAYER_LAST) {
          const float sse_t = sse + __shfl_xor_sync(0xffffffffu, sse, 16);
          const float raw_t = raw_v + __shfl_xor_sync(0xffffffffu, raw_v, 16);
          const float s2v = P.s2v >= 0.f ? P.s2v : softplus_f(raw_t) + 1e-6f;
          loss_pv = sse_t / (2.f * s2v) + ((float)P.p * logf(s2v)) / 2.f;
        }
      }
    }
    if ((C.flags & CH_LAYER_LAST) && !(C.flags & CH_FINAL)) { float* tmp = in; in = out; out = tmp; }
    // end of an evaluation: assemble the log-posterior (:814-816)
    if (i == nch - 1) {
      const bool cur = (D.mode == 1) || e == 1;
      float pr = 0.f;
#pragma unroll
      for (int k = 0; k < ZH; ++k) {
        const float zv = cur ? zc[k] : zp[k];
        if (dbase + k < zd) pr = fmaf(zv, zv, pr);
      }
      pr += __shfl_xor_sync(0xffffffffu, pr, 16);
      const float lpv = -(((loss_pv + loss_px) + loss_py) + 0.5f * pr);
      if (e == 0) lp0 = lpv; else lp1 = lpv;
    }
    // next chunk: its net's inputs (the activation buffers are free once a net's final layer has been consumed),
    // its weights into the other buffer
    if (gi + 1 < total) {
      const int e2 = (gi + 1) >= nch ? 1 : 0;
      const int i2 = gi + 1 - e2 * nch;
      const BnnChunk& C2 = Q.ch[i2];
      if (C2.flags & CH_NET_FIRST) {
        __syncwarp();
        write_inputs(C2.net, e2);
      }
      float* Wl2 = Wbuf + ((gi + 1) & 1) * 2 * W_FLOATS;
      stage_chunk2(C2, image, Wl2, Wl2 + W_FLOATS, A.seed, D.slice, call_of(e2), tid, nth);
    }
    __syncthreads();
  }

  if (D.mode == 1) {
    if (valid && half == 0) D.out_lp[row] = lp0;
    return;
  }
  const float lp_prop = lp0, lp_cur = lp1;
  const float dlp = lp_prop - lp_cur;
  const float ratio = (dlp < 0.f) ? expf(dlp) : ((dlp >= 0.f) ? 1.f : __int_as_float(0x7fc00000));      // :868
  bool acc_;
  if (A.u_dev) acc_ = A.u_dev[(size_t)t * n + lrow] < (double)ratio;                                     // :870
  else acc_ = uniform1(A.seed, grow, (uint32_t)t, NOISE_ACCEPT) < ratio;
  if (acc_) {
#pragma unroll
    for (int k = 0; k < ZH; ++k) zc[k] = zp[k];                                                          // :871
  }
  if (valid) {
    if (acc_)
#pragma unroll
      for (int k = 0; k < ZH; ++k)
        if (dbase + k < zd) A.z_state_dev[(size_t)row * zd + dbase + k] = zc[k];
    if (half == 0) {
      if (A.accept_mask_dev) A.accept_mask_dev[(size_t)t * n + row] = acc_ ? 1 : 0;
      if (A.lp_trace_dev) A.lp_trace_dev[(size_t)t * n + row] = lp_prop;
      if (D.lp_cur_trace) D.lp_cur_trace[(size_t)t * n + row] = lp_cur;
    }
    if (t >= A.burn_in && A.out_samples_dev) {                                                           // :895-896
      float* dst = A.out_samples_dev + ((size_t)(t - A.burn_in) * n + row) * zd;
#pragma unroll
      for (int k = 0; k < ZH; ++k)
        if (dbase + k < zd) dst[dbase + k] = zc[k];
    }
  }
  if (A.accept_count_dev) {
    const unsigned b = __ballot_sync(0xffffffffu, acc_ && valid && half == 0);
    if (lane == 0 && b) atomicAdd(A.accept_count_dev + t, __popc(b));
  }
  if (t + 1 < A.t_end) {
    propose2<ZH>(A, t + 1, lrow, grow, zd, dbase, q_sd, zc, zp);
    stats_partial2<ZH>(zp, zc, x_l, valid, zd, dbase, red, D.part + ((size_t)((t + 1) & 1) * ncta + blockIdx.x) * NP);
  }
}

}  // namespace bnn
}  // namespace bgm
