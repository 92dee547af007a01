// CausalBGM EGM training steps (causalbgm/base.py:305-377) and Keras Adam, sm_100a.
//
// A mini-batch is at most 32 rows, every layer at most a few hundred units: one step is
// ~25 MFLOP spread over ~150 tiny dependent GEMMs -- latency-bound by construction
// (SURVEY 2.2 K5/K6).  The reference pays one TensorFlow op dispatch per GEMM; here ONE
// CTA runs the whole forward / backward (and, for the discriminator's gradient penalty,
// the hand-derived double backward through Dense -> BatchNorm(batch stats) -> tanh) in a
// single launch, activations in shared memory as [feature][row] columns, saved
// activations in an L2-resident tape, and a second multi-CTA kernel applies Adam.
// Between the two sits the only collective of the data-parallel path (gradient
// all-reduce over the flat gradient buffer).
//
// Work mapping: a thread owns an output FEATURE (column) and half (or all) of the 32
// batch rows in registers; the other operand is broadcast from shared memory with
// LDS.128.  Column reductions over the batch (BatchNorm statistics, bias gradients) are
// therefore thread-local.
#pragma once
#include "common.cuh"

namespace bgm {
namespace tr {

constexpr int NTH = 128;     // threads per CTA
constexpr int LD = 36;       // padded length of one feature column (32 rows + 4: conflict-free LDS/STS.128)
constexpr int MAXL = 8;      // Dense layers per net
constexpr float BN_EPS = 1e-3f;

struct Net {                 // a Dense stack inside a flat parameter buffer
  int L;
  int dims[MAXL + 1];
  int w_off[MAXL], b_off[MAXL];   // float offsets of kernel[in][out] / bias[out]
};
struct Disc {                // Dense -> BN -> tanh blocks + output Dense
  int L;                     // hidden blocks
  int dims[MAXL + 1];        // [in, units..., 1]
  int w_off[MAXL], b_off[MAXL], g_off[MAXL], be_off[MAXL];   // g/be only for l < L
  int n_params;
};

struct GenArgs {
  Net g, e, f, h;
  Disc dz;
  int z_dims[4];
  int zd, p, binary;
  float use_z_rec;
  int bs;
  const float* theta;        // gen group: [g | e | f | h]
  const float* theta_d;      // disc group
  float* grad;               // gen group gradient (written)
  float* tape;               // global scratch
  int tape_floats;
  const float *z, *v, *x, *y;   // batch: (bs,zd) (bs,p) (bs) (bs), row-major
  float* losses;             // [6]: e_loss_adv, l2_v, l2_z, l2_x, l2_y, g_e_loss (:377)
  int wm;                    // widest matrix (features) -> shared-memory carve-up
};

struct DiscArgs {
  Net e;
  Disc dz;
  int zd, p, bs;
  const float* theta;        // gen group (e_net is read, not trained here; Net offsets are absolute)
  const float* theta_d;
  float* grad_d;             // disc group gradient (written)
  const float *z, *v;
  float epsilon;             // the tf.random.uniform([]) draw (:307)
  float gp_weight;           // 10 (:323)
  float* losses;             // [2]: dz_loss, d_loss
  int wm;
  const float* zenc_in;      // optional (bs, zd): z_ = e_net(v) computed elsewhere (Bayesian e_net, layered engine)
  const float* eps_dev;      // optional: epsilon read from device memory (replayed CUDA graphs)
  int stage;                 // != 0: 2 * dz.n_params extra floats of shared memory hold the discriminator's parameters
                             // and its gradient accumulators for the whole kernel (thread = feature loops otherwise pay an
                             // L2 round trip per weight and a global read-modify-write per gradient)
};

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float leaky02(float v) { return v > 0.f ? v : 0.2f * v; }

// ---- batch loads: row-major global (bs, dim) -> shared [dim][LD], rows >= bs zero ----
__device__ void load_cols(const float* __restrict__ src, int ld, int col0, int ncols, int bs, float* dst) {
  for (int i = threadIdx.x; i < ncols * 32; i += NTH) {
    const int r = i / ncols, c = i - r * ncols;
    dst[c * LD + r] = r < bs ? src[(size_t)r * ld + col0 + c] : 0.f;
  }
}
__device__ void copy_mat(const float* __restrict__ src, float* __restrict__ dst, int nfeat) {
  for (int i = threadIdx.x; i < nfeat * (LD / 4); i += NTH) st4(dst + i * 4, ld4(src + i * 4));
}

// ---- Dense forward: ys[n] = act(b[n] + sum_k xs[k] W[k][n]) ; optional tape copy ----
// thread = (column n, half of the rows)
__device__ void dense_fwd(const float* __restrict__ W, const float* __restrict__ b, int K, int N,
                          const float* xs, float* ys, bool act, float* tape) {
  const int half = threadIdx.x >> 6, cl = threadIdx.x & 63;
  for (int n0 = 0; n0 < N; n0 += 64) {
    const int n = n0 + cl;
    if (n < N) {
      float acc[16];
      const float bn = b[n];
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = bn;
      const float* xr = xs + half * 16;
#pragma unroll 4
      for (int k = 0; k < K; ++k) {
        const float w = __ldg(W + (size_t)k * N + n);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 x4 = ld4(xr + k * LD + q * 4);
          acc[q * 4 + 0] = fmaf(x4.x, w, acc[q * 4 + 0]);
          acc[q * 4 + 1] = fmaf(x4.y, w, acc[q * 4 + 1]);
          acc[q * 4 + 2] = fmaf(x4.z, w, acc[q * 4 + 2]);
          acc[q * 4 + 3] = fmaf(x4.w, w, acc[q * 4 + 3]);
        }
      }
      if (act) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = leaky02(acc[i]);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 o = make_float4(acc[q * 4], acc[q * 4 + 1], acc[q * 4 + 2], acc[q * 4 + 3]);
        st4(ys + n * LD + half * 16 + q * 4, o);
        if (tape) st4(tape + n * LD + half * 16 + q * 4, o);
      }
    }
  }
  __syncthreads();
}

// ---- Dense backward, inputs: gx[k] = (sum_n gs[n] W[k][n]) * leaky'(xprev[k]) ----
// xprev: post-activation of the producing layer (sign == sign of its pre-activation), or
// NULL for a linear input.  accumulate: gx += ...
__device__ void dense_bwd_x(const float* __restrict__ W, int K, int N, const float* gs, float* gx,
                            const float* xprev, bool accumulate) {
  const int half = threadIdx.x >> 6, cl = threadIdx.x & 63;
  for (int k0 = 0; k0 < K; k0 += 64) {
    const int k = k0 + cl;
    if (k < K) {
      float acc[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = 0.f;
      const float* wr = W + (size_t)k * N;
      const float* gr = gs + half * 16;
#pragma unroll 4
      for (int n = 0; n < N; ++n) {
        const float w = __ldg(wr + n);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 g4 = ld4(gr + n * LD + q * 4);
          acc[q * 4 + 0] = fmaf(g4.x, w, acc[q * 4 + 0]);
          acc[q * 4 + 1] = fmaf(g4.y, w, acc[q * 4 + 1]);
          acc[q * 4 + 2] = fmaf(g4.z, w, acc[q * 4 + 2]);
          acc[q * 4 + 3] = fmaf(g4.w, w, acc[q * 4 + 3]);
        }
      }
      float* o = gx + k * LD + half * 16;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float4 r = make_float4(acc[q * 4], acc[q * 4 + 1], acc[q * 4 + 2], acc[q * 4 + 3]);
        if (xprev) {
          const float4 xp = ld4(xprev + k * LD + half * 16 + q * 4);
          r.x *= xp.x > 0.f ? 1.f : 0.2f;
          r.y *= xp.y > 0.f ? 1.f : 0.2f;
          r.z *= xp.z > 0.f ? 1.f : 0.2f;
          r.w *= xp.w > 0.f ? 1.f : 0.2f;
        }
        if (accumulate) {
          const float4 old = ld4(o + q * 4);
          r.x += old.x; r.y += old.y; r.z += old.z; r.w += old.w;
        }
        st4(o + q * 4, r);
      }
    }
  }
  __syncthreads();
}

// ---- Dense backward, parameters: dW[k][n] (+)= sum_r xs[k][r] gs[n][r]; db[n] (+)= sum_r gs[n][r]
// thread = (column n, half of the k range)
__device__ void dense_bwd_w(float* __restrict__ dW, float* __restrict__ db, int K, int N, const float* xs,
                            const float* gs, bool accumulate, float scale = 1.f) {
  const int khalf = threadIdx.x >> 6, cl = threadIdx.x & 63;
  const int kb = khalf == 0 ? 0 : (K + 1) / 2, ke = khalf == 0 ? (K + 1) / 2 : K;
  for (int n0 = 0; n0 < N; n0 += 64) {
    const int n = n0 + cl;
    if (n < N) {
      float g[32];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 g4 = ld4(gs + n * LD + q * 4);
        g[q * 4] = g4.x; g[q * 4 + 1] = g4.y; g[q * 4 + 2] = g4.z; g[q * 4 + 3] = g4.w;
      }
      if (khalf == 0) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) s += g[i];
        s *= scale;
        db[n] = accumulate ? db[n] + s : s;
      }
#pragma unroll 2
      for (int k = kb; k < ke; ++k) {
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int q = 0; q < 8; q += 2) {
          const float4 a = ld4(xs + k * LD + q * 4), c = ld4(xs + k * LD + q * 4 + 4);
          s0 = fmaf(a.x, g[q * 4], s0); s0 = fmaf(a.y, g[q * 4 + 1], s0);
          s0 = fmaf(a.z, g[q * 4 + 2], s0); s0 = fmaf(a.w, g[q * 4 + 3], s0);
          s1 = fmaf(c.x, g[q * 4 + 4], s1); s1 = fmaf(c.y, g[q * 4 + 5], s1);
          s1 = fmaf(c.z, g[q * 4 + 6], s1); s1 = fmaf(c.w, g[q * 4 + 7], s1);
        }
        const float s = (s0 + s1) * scale;
        float* o = dW + (size_t)k * N + n;
        *o = accumulate ? *o + s : s;
      }
    }
  }
  __syncthreads();
}

// ---- BaseFullyConnectedNet forward (networks/base.py:30-51) ----
// xs holds the input; buffers ping-pong between bufA / bufB; every layer's output is also
// written to the tape (tape_base + cumulative offset).  Returns the buffer with the output.
__device__ float* mlp_forward(const Net& net, const float* theta, const float* xs, float* bufA, float* bufB,
                              float* tape) {
  const float* in = xs;
  float* out = bufA;
  int toff = 0;
  for (int l = 0; l < net.L; ++l) {
    out = (in == bufA) ? bufB : bufA;
    dense_fwd(theta + net.w_off[l], theta + net.b_off[l], net.dims[l], net.dims[l + 1], in, out,
              l < net.L - 1, tape ? tape + toff : nullptr);
    toff += net.dims[l + 1] * LD;
    in = out;
  }
  return out;
}
__device__ __host__ inline int net_tape_floats(const Net& net) {
  int t = 0;
  for (int l = 0; l < net.L; ++l) t += net.dims[l + 1] * LD;
  return t;
}

// ---- backward through a Dense stack ----
// gs: gradient w.r.t. the net's (linear) output, in shared memory; x0: the net's input
// (shared or global matrix [dims[0]][LD]); tape: the forward's saved outputs.  Writes the
// parameter gradients (accumulate: add to what is there) and, if gin != NULL, the gradient
// w.r.t. the input.  ga/gb: ping-pong matrices wide enough for the hidden layers (gb may
// be the seed buffer gs if that one is wide enough); xb: scratch for the layer inputs.
__device__ void mlp_backward(const Net& net, const float* theta, float* grad, const float* x0, const float* tape,
                             float* gs, float* ga, float* gb, float* xb, float* gin, bool accumulate) {
  int toff[MAXL];
  int t = 0;
  for (int l = 0; l < net.L; ++l) { toff[l] = t; t += net.dims[l + 1] * LD; }
  float* g = gs;
  for (int l = net.L - 1; l >= 0; --l) {
    const int K = net.dims[l], N = net.dims[l + 1];
    const float* xprev = l == 0 ? x0 : tape + toff[l - 1];
    copy_mat(xprev, xb, K);
    __syncthreads();
    dense_bwd_w(grad + net.w_off[l], grad + net.b_off[l], K, N, xb, g, accumulate);
    if (l > 0 || gin) {
      float* o = l == 0 ? gin : (g == ga ? gb : ga);
      dense_bwd_x(theta + net.w_off[l], K, N, g, o, l == 0 ? nullptr : xb, false);
      g = o;
    }
  }
}

// =========================== Discriminator ====================================
// Shared-memory state of one discriminator pass (all hidden blocks).
struct DiscBufs {
  float* X[MAXL + 1];   // X[0] input, X[l] = tanh output of block l          [d_l][LD]
  float* N[MAXL];       // normalised pre-activations of block l (index l-1)
  float* H[MAXL];       // gradient w.r.t. the pre-activation A_l (first-order backward)
  float* U[MAXL + 1];   // gradient w.r.t. X_l (first-order backward)
  float* Q[MAXL];       // V * gamma
  float* s;             // [sum d_l] 1/sqrt(var+eps) per feature
  float* m2;            // [sum d_l] mean_r(Q*N) per feature
  int foff[MAXL + 1];   // feature offset of block l in s / m2
};

__device__ __forceinline__ float rowmask(int r, int bs) { return r < bs ? 1.f : 0.f; }

// forward of all blocks; returns nothing, out[r] written to `outv` (shared, 32 floats)
__device__ void disc_forward(const Disc& dz, const float* th, const DiscBufs& B, int bs, float* outv) {
  const float inv_bs = 1.f / (float)bs;
  for (int l = 1; l <= dz.L; ++l) {
    const int K = dz.dims[l - 1], Nf = dz.dims[l];
    const float* W = th + dz.w_off[l - 1];
    for (int j = threadIdx.x; j < Nf; j += NTH) {
      float a[32];
      const float bj = th[dz.b_off[l - 1] + j];
#pragma unroll
      for (int i = 0; i < 32; ++i) a[i] = bj;
      for (int k = 0; k < K; ++k) {
        const float w = W[k * Nf + j];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 x4 = ld4(B.X[l - 1] + k * LD + q * 4);
          a[q * 4] = fmaf(x4.x, w, a[q * 4]); a[q * 4 + 1] = fmaf(x4.y, w, a[q * 4 + 1]);
          a[q * 4 + 2] = fmaf(x4.z, w, a[q * 4 + 2]); a[q * 4 + 3] = fmaf(x4.w, w, a[q * 4 + 3]);
        }
      }
      float mu = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) mu += i < bs ? a[i] : 0.f;
      mu *= inv_bs;
      float var = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) { const float d = i < bs ? a[i] - mu : 0.f; var = fmaf(d, d, var); }
      var *= inv_bs;
      const float s = 1.f / sqrtf(var + BN_EPS);
      const float gam = th[dz.g_off[l - 1] + j], bet = th[dz.be_off[l - 1] + j];
      B.s[B.foff[l] + j] = s;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float nn[4], xx[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int i = q * 4 + c;
          nn[c] = i < bs ? (a[i] - mu) * s : 0.f;
          xx[c] = i < bs ? tanhf(fmaf(gam, nn[c], bet)) : 0.f;
        }
        st4(B.N[l - 1] + j * LD + q * 4, make_float4(nn[0], nn[1], nn[2], nn[3]));
        st4(B.X[l] + j * LD + q * 4, make_float4(xx[0], xx[1], xx[2], xx[3]));
      }
    }
    __syncthreads();
  }
  // output Dense (units -> 1)
  const int K = dz.dims[dz.L];
  if (threadIdx.x < 32) {
    const int r = threadIdx.x;
    float o = th[dz.b_off[dz.L]];
    for (int k = 0; k < K; ++k) o = fmaf(B.X[dz.L][k * LD + r], th[dz.w_off[dz.L] + k], o);
    outv[r] = r < bs ? o : 0.f;
  }
  __syncthreads();
}

// First-order backward with d(loss)/d(out[r]) = seed for r < bs.
//   gacc != NULL: accumulate scale * parameter gradients into gacc (disc group layout)
//   keeps H, U, Q, m2 in B for the double backward.  U[0] = d(loss)/d(input).
__device__ void disc_backward(const Disc& dz, const float* th, const DiscBufs& B, int bs, float seed,
                              float* gacc, float scale, const float* seedv = nullptr) {
  // seedv != NULL: per-row seeds d(loss)/d(out[r]) (shared memory, 32 floats) instead of `seed`
  const float inv_bs = 1.f / (float)bs;
  const int L = dz.L;
  {  // output layer
    const int K = dz.dims[L];
    for (int j = threadIdx.x; j < K; j += NTH) {
      const float w = th[dz.w_off[L] + j];
      float sx = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 x4 = ld4(B.X[L] + j * LD + q * 4);
        float sd[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) sd[c] = rowmask(q * 4 + c, bs) * (seedv ? seedv[q * 4 + c] : seed);
        st4(B.U[L] + j * LD + q * 4, make_float4(sd[0] * w, sd[1] * w, sd[2] * w, sd[3] * w));
        sx += sd[0] * x4.x + sd[1] * x4.y + sd[2] * x4.z + sd[3] * x4.w;
      }
      if (gacc) gacc[dz.w_off[L] + j] += scale * sx;
    }
    if (gacc && threadIdx.x == 0) {
      float sb = 0.f;
      for (int r = 0; r < bs; ++r) sb += seedv ? seedv[r] : seed;
      gacc[dz.b_off[L]] += scale * sb;
    }
    __syncthreads();
  }
  for (int l = L; l >= 1; --l) {
    const int K = dz.dims[l - 1], Nf = dz.dims[l];
    for (int j = threadIdx.x; j < Nf; j += NTH) {
      const float gam = th[dz.g_off[l - 1] + j], s = B.s[B.foff[l] + j];
      float q_[32];
      float dgam = 0.f, dbet = 0.f, m1 = 0.f, m2 = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 u4 = ld4(B.U[l] + j * LD + q * 4), x4 = ld4(B.X[l] + j * LD + q * 4),
                     n4 = ld4(B.N[l - 1] + j * LD + q * 4);
        const float uu[4] = {u4.x, u4.y, u4.z, u4.w}, xx[4] = {x4.x, x4.y, x4.z, x4.w},
                    nn[4] = {n4.x, n4.y, n4.z, n4.w};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float v = uu[c] * (1.f - xx[c] * xx[c]);
          dgam = fmaf(v, nn[c], dgam);
          dbet += v;
          const float qq = v * gam;
          q_[q * 4 + c] = qq;
          m1 += qq;
          m2 = fmaf(qq, nn[c], m2);
        }
      }
      m1 *= inv_bs;
      m2 *= inv_bs;
      B.m2[B.foff[l] + j] = m2;
      float dbias = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 n4 = ld4(B.N[l - 1] + j * LD + q * 4);
        const float nn[4] = {n4.x, n4.y, n4.z, n4.w};
        float hh[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int i = q * 4 + c;
          hh[c] = i < bs ? s * (q_[i] - m1 - nn[c] * m2) : 0.f;
          dbias += hh[c];
        }
        st4(B.H[l - 1] + j * LD + q * 4, make_float4(hh[0], hh[1], hh[2], hh[3]));
        if (B.Q[l - 1]) st4(B.Q[l - 1] + j * LD + q * 4, make_float4(q_[q * 4], q_[q * 4 + 1], q_[q * 4 + 2], q_[q * 4 + 3]));
      }
      if (gacc) {
        gacc[dz.g_off[l - 1] + j] += scale * dgam;
        gacc[dz.be_off[l - 1] + j] += scale * dbet;
        gacc[dz.b_off[l - 1] + j] += scale * dbias;
        // dW[i][j] = sum_r X_{l-1}[i][r] H[j][r]
        float h[32];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 n4 = ld4(B.N[l - 1] + j * LD + q * 4);
          const float nn[4] = {n4.x, n4.y, n4.z, n4.w};
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int i = q * 4 + c;
            h[i] = i < bs ? s * (q_[i] - m1 - nn[c] * m2) : 0.f;
          }
        }
        for (int i = 0; i < K; ++i) {
          float d = 0.f;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 x4 = ld4(B.X[l - 1] + i * LD + q * 4);
            d = fmaf(x4.x, h[q * 4], d); d = fmaf(x4.y, h[q * 4 + 1], d);
            d = fmaf(x4.z, h[q * 4 + 2], d); d = fmaf(x4.w, h[q * 4 + 3], d);
          }
          gacc[dz.w_off[l - 1] + i * Nf + j] += scale * d;
        }
      }
    }
    __syncthreads();
    // U_{l-1}[i] = sum_j H[j] W[i][j]
    const float* W = th + dz.w_off[l - 1];
    for (int i = threadIdx.x; i < K; i += NTH) {
      float u[32];
#pragma unroll
      for (int r = 0; r < 32; ++r) u[r] = 0.f;
      for (int j = 0; j < Nf; ++j) {
        const float w = W[i * Nf + j];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 h4 = ld4(B.H[l - 1] + j * LD + q * 4);
          u[q * 4] = fmaf(h4.x, w, u[q * 4]); u[q * 4 + 1] = fmaf(h4.y, w, u[q * 4 + 1]);
          u[q * 4 + 2] = fmaf(h4.z, w, u[q * 4 + 2]); u[q * 4 + 3] = fmaf(h4.w, w, u[q * 4 + 3]);
        }
      }
#pragma unroll
      for (int q = 0; q < 8; ++q)
        st4(B.U[l - 1] + i * LD + q * 4, make_float4(u[q * 4], u[q * 4 + 1], u[q * 4 + 2], u[q * 4 + 3]));
    }
    __syncthreads();
  }
}

// Double backward: given Ubar0 = d P / d U_0 (P: a function of the input gradient U_0 of
// a first-order backward with seed 1), accumulate scale * dP/dtheta into gacc.
// Scratch matrices (each [maxd][LD]): UB, HB (H-bar / A-bar), XB, NB per block are carved
// from `scr`; sbar per feature in `sbar`.
__device__ void disc_double_backward(const Disc& dz, const float* th, const DiscBufs& B, int bs, float* UB0,
                                     float* scr, float* sbar, float* gacc, float scale, int maxd) {
  const float inv_bs = 1.f / (float)bs;
  const int L = dz.L;
  // per-block scratch: XB[l], NB[l] (kept for the second sweep); UB ping-pong
  float* XB[MAXL];
  float* NB[MAXL];
  float* p = scr;
  for (int l = 1; l <= L; ++l) { XB[l - 1] = p; p += dz.dims[l] * LD; NB[l - 1] = p; p += dz.dims[l] * LD; }
  float* UBa = p; p += maxd * LD;
  float* AB = p; p += maxd * LD;       // A-bar of the block above (second sweep)
  float* ABn = p; p += maxd * LD;
  const float* UBin = UB0;
  // ---- sweep 1: reverse of the first-order backward, blocks 1..L ----
  for (int l = 1; l <= L; ++l) {
    const int K = dz.dims[l - 1], Nf = dz.dims[l];
    const float* W = th + dz.w_off[l - 1];
    float* UBout = (UBin == UBa) ? AB : UBa;    // AB is free during sweep 1
    for (int j = threadIdx.x; j < Nf; j += NTH) {
      const float gam = th[dz.g_off[l - 1] + j], s = B.s[B.foff[l] + j], m2 = B.m2[B.foff[l] + j];
      // Hbar[r] = sum_i Ubar_{l-1}[i][r] W[i][j];  Wbar[i][j] += sum_r Ubar_{l-1}[i][r] H[j][r]
      float hb[32], h[32];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 h4 = ld4(B.H[l - 1] + j * LD + q * 4);
        h[q * 4] = h4.x; h[q * 4 + 1] = h4.y; h[q * 4 + 2] = h4.z; h[q * 4 + 3] = h4.w;
        hb[q * 4] = hb[q * 4 + 1] = hb[q * 4 + 2] = hb[q * 4 + 3] = 0.f;
      }
      for (int i = 0; i < K; ++i) {
        const float w = W[i * Nf + j];
        float d = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 u4 = ld4(UBin + i * LD + q * 4);
          hb[q * 4] = fmaf(u4.x, w, hb[q * 4]); hb[q * 4 + 1] = fmaf(u4.y, w, hb[q * 4 + 1]);
          hb[q * 4 + 2] = fmaf(u4.z, w, hb[q * 4 + 2]); hb[q * 4 + 3] = fmaf(u4.w, w, hb[q * 4 + 3]);
          d = fmaf(u4.x, h[q * 4], d); d = fmaf(u4.y, h[q * 4 + 1], d);
          d = fmaf(u4.z, h[q * 4 + 2], d); d = fmaf(u4.w, h[q * 4 + 3], d);
        }
        gacc[dz.w_off[l - 1] + i * Nf + j] += scale * d;
      }
      // column-local part
      float sb = 0.f, m1b = 0.f, m2b = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 n4 = ld4(B.N[l - 1] + j * LD + q * 4);
        const float nn[4] = {n4.x, n4.y, n4.z, n4.w};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int i = q * 4 + c;
          const float hbr = i < bs ? hb[i] : 0.f;
          const float cc = h[i] / s;                 // C = H / s
          sb = fmaf(hbr, cc, sb);
          const float cb = hbr * s;                  // C-bar
          m1b -= cb;
          m2b = fmaf(-cb, nn[c], m2b);
        }
      }
      sbar[B.foff[l] + j] = sb;
      float gamb = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 n4 = ld4(B.N[l - 1] + j * LD + q * 4), q4 = ld4(B.Q[l - 1] + j * LD + q * 4),
                     u4 = ld4(B.U[l] + j * LD + q * 4), x4 = ld4(B.X[l] + j * LD + q * 4);
        const float nn[4] = {n4.x, n4.y, n4.z, n4.w}, qq[4] = {q4.x, q4.y, q4.z, q4.w},
                    uu[4] = {u4.x, u4.y, u4.z, u4.w}, xx[4] = {x4.x, x4.y, x4.z, x4.w};
        float ub[4], xb[4], nb[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int i = q * 4 + c;
          const float msk = rowmask(i, bs);
          const float cb = msk * hb[i] * s;
          const float qb = msk * (cb + m1b * inv_bs + m2b * inv_bs * nn[c]);       // Q-bar
          nb[c] = msk * (-cb * m2 + m2b * inv_bs * qq[c]);                         // N-bar (part 1)
          const float t = 1.f - xx[c] * xx[c];
          const float v = uu[c] * t;                                               // V = U*T
          gamb = fmaf(qb, v, gamb);
          const float vb = qb * gam;                                               // V-bar
          ub[c] = vb * t;                                                          // U_l-bar
          xb[c] = -2.f * xx[c] * (vb * uu[c]);                                     // X_l-bar via T
        }
        st4(UBout + j * LD + q * 4, make_float4(ub[0], ub[1], ub[2], ub[3]));
        st4(XB[l - 1] + j * LD + q * 4, make_float4(xb[0], xb[1], xb[2], xb[3]));
        st4(NB[l - 1] + j * LD + q * 4, make_float4(nb[0], nb[1], nb[2], nb[3]));
      }
      gacc[dz.g_off[l - 1] + j] += scale * gamb;
    }
    __syncthreads();
    UBin = UBout;
  }
  // U_L = seed * w_out^T (seed 1): w_out-bar[j] += sum_r Ubar_L[j][r]
  for (int j = threadIdx.x; j < dz.dims[L]; j += NTH) {
    float sacc = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 u4 = ld4(UBin + j * LD + q * 4);
      sacc += u4.x + u4.y + u4.z + u4.w;
    }
    gacc[dz.w_off[L] + j] += scale * sacc;
  }
  __syncthreads();
  // ---- sweep 2: reverse of the forward pass, blocks L..1 ----
  float* ABcur = AB;    // A-bar of block l+1
  float* ABnew = ABn;
  for (int l = L; l >= 1; --l) {
    const int K = dz.dims[l - 1], Nf = dz.dims[l];
    const float* Wup = l < L ? th + dz.w_off[l] : nullptr;   // W_{l+1}: [Nf][dims[l+1]]
    const int Nup = l < L ? dz.dims[l + 1] : 0;
    for (int j = threadIdx.x; j < Nf; j += NTH) {
      const float gam = th[dz.g_off[l - 1] + j], s = B.s[B.foff[l] + j];
      float xb[32];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 x4 = ld4(XB[l - 1] + j * LD + q * 4);
        xb[q * 4] = x4.x; xb[q * 4 + 1] = x4.y; xb[q * 4 + 2] = x4.z; xb[q * 4 + 3] = x4.w;
      }
      for (int k = 0; k < Nup; ++k) {      // X_l feeds A_{l+1} = X_l W_{l+1} + b
        const float w = Wup[j * Nup + k];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 a4 = ld4(ABcur + k * LD + q * 4);
          xb[q * 4] = fmaf(a4.x, w, xb[q * 4]); xb[q * 4 + 1] = fmaf(a4.y, w, xb[q * 4 + 1]);
          xb[q * 4 + 2] = fmaf(a4.z, w, xb[q * 4 + 2]); xb[q * 4 + 3] = fmaf(a4.w, w, xb[q * 4 + 3]);
        }
      }
      float gamb = 0.f, betb = 0.f, stot = sbar[B.foff[l] + j];
      float nb[32];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 x4 = ld4(B.X[l] + j * LD + q * 4), n4 = ld4(B.N[l - 1] + j * LD + q * 4),
                     nb4 = ld4(NB[l - 1] + j * LD + q * 4);
        const float xx[4] = {x4.x, x4.y, x4.z, x4.w}, nn[4] = {n4.x, n4.y, n4.z, n4.w},
                    nbp[4] = {nb4.x, nb4.y, nb4.z, nb4.w};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int i = q * 4 + c;
          const float yb = rowmask(i, bs) * xb[i] * (1.f - xx[c] * xx[c]);          // Y-bar
          gamb = fmaf(yb, nn[c], gamb);
          betb += yb;
          nb[i] = nbp[c] + yb * gam;                                               // N-bar total
          stot = fmaf(nb[i], nn[c] / s, stot);                                     // + sum N-bar * D, D = N/s
        }
      }
      const float varb = -0.5f * stot * s * s * s;                                 // d/d var of s=(var+eps)^-1/2
      float db_[32];
      float mean_db = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 n4 = ld4(B.N[l - 1] + j * LD + q * 4);
        const float nn[4] = {n4.x, n4.y, n4.z, n4.w};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int i = q * 4 + c;
          const float d = nn[c] / s;
          db_[i] = rowmask(i, bs) * (nb[i] * s + varb * 2.f * inv_bs * d);
          mean_db += db_[i];
        }
      }
      mean_db *= inv_bs;
      float bb = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        db_[i] = i < bs ? db_[i] - mean_db : 0.f;                                  // A-bar
        bb += db_[i];
      }
#pragma unroll
      for (int q = 0; q < 8; ++q)
        st4(ABnew + j * LD + q * 4, make_float4(db_[q * 4], db_[q * 4 + 1], db_[q * 4 + 2], db_[q * 4 + 3]));
      gacc[dz.g_off[l - 1] + j] += scale * gamb;
      gacc[dz.be_off[l - 1] + j] += scale * betb;
      gacc[dz.b_off[l - 1] + j] += scale * bb;
      for (int i = 0; i < K; ++i) {      // W_l-bar[i][j] += sum_r X_{l-1}[i][r] A-bar[j][r]
        float d = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 x4 = ld4(B.X[l - 1] + i * LD + q * 4);
          d = fmaf(x4.x, db_[q * 4], d); d = fmaf(x4.y, db_[q * 4 + 1], d);
          d = fmaf(x4.z, db_[q * 4 + 2], d); d = fmaf(x4.w, db_[q * 4 + 3], d);
        }
        gacc[dz.w_off[l - 1] + i * dz.dims[l] + j] += scale * d;
      }
    }
    __syncthreads();
    float* tmp = ABcur; ABcur = ABnew; ABnew = tmp;
  }
}

__device__ __host__ inline int disc_feat_total(const Disc& dz) {
  int t = 0;
  for (int l = 0; l <= dz.L; ++l) t += dz.dims[l];
  return t;
}
__device__ __host__ inline int disc_maxd(const Disc& dz) {
  int m = 0;
  for (int l = 0; l <= dz.L; ++l) m = dz.dims[l] > m ? dz.dims[l] : m;
  return m;
}
// floats of shared memory the discriminator machinery needs.  with_double: keep what the
// double backward reads (Q) and its scratch; ext_x0: the input matrix X[0] lives elsewhere.
__device__ __host__ inline int disc_smem_floats(const Disc& dz, bool with_double, bool ext_x0 = false) {
  const int ft = disc_feat_total(dz), md = disc_maxd(dz);
  int hid = 0;
  for (int l = 1; l <= dz.L; ++l) hid += dz.dims[l];
  const int mats = (ext_x0 ? hid : ft) + ft + 2 * hid + (with_double ? hid + 2 * hid + 3 * md : 0);
  return mats * LD + 3 * ft + 64 + 16;
}

// Parameters are read from global memory (L1-cached) and gradients accumulate straight into
// the global gradient buffer (each parameter is owned by one thread per phase; phases are
// separated by __syncthreads), so neither takes shared memory.
__device__ void disc_carve(const Disc& dz, float* base, bool with_double, DiscBufs& B, float*& scr, float*& sbar,
                           float*& outv, float* x0_ext = nullptr) {
  float* p = base;
  int f = 0;
  for (int l = 0; l <= dz.L; ++l) {
    if (l == 0 && x0_ext) B.X[0] = x0_ext;
    else { B.X[l] = p; p += dz.dims[l] * LD; }
    B.foff[l] = f;
    f += dz.dims[l];
  }
  for (int l = 0; l <= dz.L; ++l) { B.U[l] = p; p += dz.dims[l] * LD; }
  for (int l = 1; l <= dz.L; ++l) {
    B.N[l - 1] = p; p += dz.dims[l] * LD;
    B.H[l - 1] = p; p += dz.dims[l] * LD;
    if (with_double) { B.Q[l - 1] = p; p += dz.dims[l] * LD; }
    else B.Q[l - 1] = nullptr;
  }
  int hid = 0;
  for (int l = 1; l <= dz.L; ++l) hid += dz.dims[l];
  scr = p; p += with_double ? (2 * hid + 3 * disc_maxd(dz)) * LD : 0;
  B.s = p; p += f;
  B.m2 = p; p += f;
  sbar = p; p += f;
  outv = p; p += 64;
}

// ================================ kernels ======================================
// train_disc_step gradients (causalbgm/base.py:305-323).
__global__ void __launch_bounds__(NTH, 1) disc_grad_kernel(const __grid_constant__ DiscArgs A) {
  extern __shared__ __align__(16) float sm[];
  const Disc& dz = A.dz;
  const int bs = A.bs;
  const float epsilon = A.eps_dev ? *A.eps_dev : A.epsilon;
  float* bufA = sm;
  float* bufB = bufA + A.wm * LD;
  float* zmat = bufB + A.wm * LD;           // z      [zd][LD]
  float* zenc = zmat + A.zd * LD;           // e(v)   [zd][LD]
  float* zhat = zenc + A.zd * LD;
  DiscBufs B;
  float *scr, *sbar, *outv;
  disc_carve(dz, zhat + A.zd * LD, true, B, scr, sbar, outv);
  const float* th_s = A.theta_d;
  float* gacc = A.grad_d;
  if (A.stage) {
    float* th_sm = outv + 64;
    gacc = th_sm + dz.n_params;
    for (int i = threadIdx.x; i < dz.n_params; i += NTH) th_sm[i] = A.theta_d[i];
    th_s = th_sm;
  }
  for (int i = threadIdx.x; i < dz.n_params; i += NTH) gacc[i] = 0.f;
  if (A.zenc_in) load_cols(A.zenc_in, A.zd, 0, A.zd, bs, bufA);
  else load_cols(A.v, A.p, 0, A.p, bs, bufA);
  load_cols(A.z, A.zd, 0, A.zd, bs, zmat);
  __syncthreads();
  // z_ = e_net(v)   (:310; no gradient flows to e_net here)
  float* ze = A.zenc_in ? bufA : mlp_forward(A.e, A.theta, bufA, bufA, bufB, nullptr);
  const float inv_bs = 1.f / (float)bs;
  for (int i = threadIdx.x; i < A.zd * 32; i += NTH) {
    const int d = i >> 5, r = i & 31;
    const float zz = zmat[d * LD + r], ze_ = r < bs ? ze[d * LD + r] : 0.f;
    zenc[d * LD + r] = ze_;
    zhat[d * LD + r] = r < bs ? zz * epsilon + ze_ * (1.f - epsilon) : 0.f;      // :311
  }
  __syncthreads();
  float mean_d = 0.f, mean_d_ = 0.f;
  // ---- D(z): -mean ----
  copy_mat(zmat, B.X[0], A.zd);
  __syncthreads();
  disc_forward(dz, th_s, B, bs, outv);
  for (int r = 0; r < bs; ++r) mean_d += outv[r];
  mean_d *= inv_bs;
  __syncthreads();
  disc_backward(dz, th_s, B, bs, -inv_bs, gacc, 1.f);
  // ---- D(z_): +mean ----
  copy_mat(zenc, B.X[0], A.zd);
  __syncthreads();
  disc_forward(dz, th_s, B, bs, outv);
  for (int r = 0; r < bs; ++r) mean_d_ += outv[r];
  mean_d_ *= inv_bs;
  __syncthreads();
  disc_backward(dz, th_s, B, bs, inv_bs, gacc, 1.f);
  // ---- gradient penalty on z_hat (:319-321) ----
  copy_mat(zhat, B.X[0], A.zd);
  __syncthreads();
  disc_forward(dz, th_s, B, bs, outv);
  disc_backward(dz, th_s, B, bs, 1.f, nullptr, 0.f);       // U[0] = d sum(D(z_hat)) / d z_hat
  float* UB0 = bufA;                                       // [zd][LD]  (bufA is free now)
  float* normv = outv + 32;
  if (threadIdx.x < 32) {
    const int r = threadIdx.x;
    float s2 = 0.f;
    for (int d = 0; d < A.zd; ++d) { const float g = B.U[0][d * LD + r]; s2 = fmaf(g, g, s2); }
    normv[r] = sqrtf(s2);
  }
  __syncthreads();
  float gp = 0.f;
  for (int r = 0; r < bs; ++r) { const float d = normv[r] - 1.f; gp = fmaf(d, d, gp); }
  gp *= inv_bs;
  for (int i = threadIdx.x; i < A.zd * 32; i += NTH) {
    const int d = i >> 5, r = i & 31;
    const float nr = normv[r];
    // d gp / d G[r][d] = (2/bs) (|G_r| - 1) G[r][d] / |G_r|
    UB0[d * LD + r] = (r < bs && nr > 0.f) ? 2.f * inv_bs * (nr - 1.f) * B.U[0][d * LD + r] / nr : 0.f;
  }
  __syncthreads();
  disc_double_backward(dz, th_s, B, bs, UB0, scr, sbar, gacc, A.gp_weight, disc_maxd(dz));
  __syncthreads();
  if (A.stage)
    for (int i = threadIdx.x; i < dz.n_params; i += NTH) A.grad_d[i] = gacc[i];
  if (threadIdx.x == 0) {
    const float dz_loss = -mean_d + mean_d_;               // :316
    A.losses[0] = dz_loss;
    A.losses[1] = dz_loss + A.gp_weight * gp;              // :323
  }
}

// train_gen_step gradients (causalbgm/base.py:332-370).
__global__ void __launch_bounds__(NTH, 1) gen_grad_kernel(const __grid_constant__ GenArgs A) {
  extern __shared__ __align__(16) float sm[];
  const int bs = A.bs, p = A.p, zd = A.zd;
  const float inv_bs = 1.f / (float)bs;
  float* bufA = sm;
  float* bufB = bufA + A.wm * LD;
  float* bufC = bufB + A.wm * LD;
  float* bufD = bufC + A.wm * LD;
  float* zmat = bufD + A.wm * LD;          // z   [zd][LD]
  float* gz = zmat + zd * LD;              // d loss / d z_  [zd][LD]
  float* xy = gz + zd * LD;                // x, y columns   [2][LD]
  float* fin = xy + 2 * LD;                // f / h inputs   [zd+1][LD]
  float* red = fin + (zd + 1) * LD;        // [16] loss accumulators
  DiscBufs B;
  float *scr, *sbar, *outv;
  disc_carve(A.dz, red + 16, false, B, scr, sbar, outv);
  const float* th_s = A.theta_d;
  if (threadIdx.x < 16) red[threadIdx.x] = 0.f;
  // tape layout
  float* T_g1 = A.tape;
  float* T_e1 = T_g1 + net_tape_floats(A.g);
  float* T_e2 = T_e1 + net_tape_floats(A.e);
  float* T_g2 = T_e2 + net_tape_floats(A.e);
  float* T_f = T_g2 + net_tape_floats(A.g);
  float* T_h = T_f + net_tape_floats(A.f);
  float* T_v = T_h + net_tape_floats(A.h);           // v  [p][LD]
  float* T_zenc = T_v + p * LD;                      // z_ [zd][LD]
  const float* th = A.theta;
  auto last_off = [](const Net& n) { int t = 0; for (int l = 0; l < n.L - 1; ++l) t += n.dims[l + 1] * LD; return t; };
  auto block_sum = [&](float v, int slot) {   // adds the CTA-wide sum of v to red[slot]
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(red + slot, v);
  };

  load_cols(A.z, zd, 0, zd, bs, zmat);
  load_cols(A.x, 1, 0, 1, bs, xy);
  load_cols(A.y, 1, 0, 1, bs, xy + LD);
  __syncthreads();
  // F1: g(z)  -> v_ (first p outputs), raw sigma head (:336-337)
  float* g1 = mlp_forward(A.g, th, zmat, bufA, bufB, T_g1);
  {
    float s = 0.f;
    if (threadIdx.x < bs) { const float t = g1[p * LD + threadIdx.x]; s = t * t; }
    block_sum(s, 5);
  }
  // F2: z_ = e(v)  (:338)
  load_cols(A.v, p, 0, p, bs, bufC);
  __syncthreads();
  copy_mat(bufC, T_v, p);
  float* ze = mlp_forward(A.e, th, bufC, bufC, bufD, T_e1);
  copy_mat(ze, T_zenc, zd);
  __syncthreads();
  // F3: z__ = e(v_)  (:344)  -- v_ lives in the tape of F1 (rows [0,p) of its last layer)
  const float* v1 = T_g1 + last_off(A.g);
  copy_mat(v1, bufA, p);
  __syncthreads();
  float* z2 = mlp_forward(A.e, th, bufA, bufA, bufB, T_e2);
  // l2_loss_z (:350) and its seed into bufC rows [0,zd): d/d z__ = use_z_rec * 2 (z__ - z) / (bs*zd)
  {
    float s = 0.f;
    for (int i = threadIdx.x; i < zd * 32; i += NTH) {
      const int d = i >> 5, r = i & 31;
      const float df = r < bs ? z2[d * LD + r] - zmat[d * LD + r] : 0.f;
      s = fmaf(df, df, s);
      bufC[d * LD + r] = A.use_z_rec * 2.f * df / (float)(bs * zd);
    }
    block_sum(s, 2);
    __syncthreads();
  }
  // B2: back through e (pass F3, input v_) -> e grads (first pass) and d/d v_ in bufD
  mlp_backward(A.e, th, A.grad, v1, T_e2, bufC, bufA, bufC, bufB, bufD, false);
  // B3: seed of g pass F1: [d/d v_ (p rows), 0.001 * 2 s_g / bs] ; back through g (input z)
  if (threadIdx.x < 32) {
    const int r = threadIdx.x;
    bufD[p * LD + r] = r < bs ? 0.001f * 2.f * T_g1[last_off(A.g) + p * LD + r] * inv_bs : 0.f;
  }
  __syncthreads();
  mlp_backward(A.g, th, A.grad, zmat, T_g1, bufD, bufA, bufD, bufB, nullptr, false);
  // F4: v__ = g(z_)[:, :p]  (:345)
  copy_mat(T_zenc, bufC, zd);
  __syncthreads();
  float* g2 = mlp_forward(A.g, th, bufC, bufA, bufB, T_g2);
  {  // l2_loss_v (:349) and its seed (in place): 2 (v__ - v) / (bs*p); sigma column unused -> 0
    float s = 0.f;
    for (int i = threadIdx.x; i < (p + 1) * 32; i += NTH) {
      const int c = i >> 5, r = i & 31;
      float df = 0.f;
      if (c < p && r < bs) df = g2[c * LD + r] - T_v[c * LD + r];
      s = fmaf(df, df, s);
      g2[c * LD + r] = 2.f * df / (float)(bs * p);
    }
    block_sum(s, 1);
    __syncthreads();
  }
  // B1: back through g (pass F4, input z_) -> g grads (accumulate) and d/d z_ (part A) in gz
  {
    float* other = g2 == bufA ? bufB : bufA;
    mlp_backward(A.g, th, A.grad, T_zenc, T_g2, g2, other, g2, bufC, gz, true);
  }
  // F5/B4: -mean D(z_) (:347,:352) and its gradient w.r.t. z_ (dz_net is not trained here)
  copy_mat(T_zenc, B.X[0], zd);
  __syncthreads();
  disc_forward(A.dz, th_s, B, bs, outv);
  {
    float s = 0.f;
    if (threadIdx.x < bs) s = outv[threadIdx.x];
    block_sum(s, 0);
  }
  disc_backward(A.dz, th_s, B, bs, -inv_bs, nullptr, 0.f);
  for (int i = threadIdx.x; i < zd * (LD / 4); i += NTH) {
    float4 a = ld4(gz + i * 4);
    const float4 b = ld4(B.U[0] + i * 4);
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    st4(gz + i * 4, a);
  }
  __syncthreads();
  const int d0 = A.z_dims[0], d1 = A.z_dims[1], d2 = A.z_dims[2];
  // F6/B5: f([z0, z1, x]) (:354-355, :366)
  {
    const int kin = d0 + d1 + 1;
    copy_mat(T_zenc, fin, d0 + d1);
    copy_mat(xy, fin + (d0 + d1) * LD, 1);
    __syncthreads();
    float* fo = mlp_forward(A.f, th, fin, bufA, bufB, T_f);
    float sy = 0.f, ss = 0.f;
    if (threadIdx.x < 32) {
      const int r = threadIdx.x;
      const float dy = r < bs ? fo[r] - xy[LD + r] : 0.f;
      const float sg = r < bs ? fo[LD + r] : 0.f;
      sy = dy * dy;
      ss = sg * sg;
      fo[r] = 2.f * dy * inv_bs;                       // d l2_loss_y / d y_
      fo[LD + r] = 0.001f * 2.f * sg * inv_bs;         // d 0.001*sigma_square_loss / d raw head
    }
    block_sum(sy, 4);
    block_sum(ss, 5);
    __syncthreads();
    float* other = fo == bufA ? bufB : bufA;
    mlp_backward(A.f, th, A.grad, fin, T_f, fo, other, fo, bufC, bufD, false);
    for (int i = threadIdx.x; i < (d0 + d1) * 32; i += NTH) gz[(i >> 5) * LD + (i & 31)] += bufD[(i >> 5) * LD + (i & 31)];
    __syncthreads();
    (void)kin;
  }
  // F7/B6: h([z0, z2]) (:357-365)
  {
    copy_mat(T_zenc, fin, d0);
    copy_mat(T_zenc + (d0 + d1) * LD, fin + d0 * LD, d2);
    __syncthreads();
    float* ho = mlp_forward(A.h, th, fin, bufA, bufB, T_h);
    float sx = 0.f, ss = 0.f;
    if (threadIdx.x < 32) {
      const int r = threadIdx.x;
      const float lg = ho[r], xv = xy[r];
      const float sg = r < bs ? ho[LD + r] : 0.f;
      float seed;
      if (A.binary) {   // sigmoid cross-entropy with logits (:361)
        sx = r < bs ? fmaxf(lg, 0.f) - lg * xv + log1pf(expf(-fabsf(lg))) : 0.f;
        seed = r < bs ? (sigmoid_f(lg) - xv) * inv_bs : 0.f;
      } else {
        const float dx = r < bs ? lg - xv : 0.f;
        sx = dx * dx;
        seed = 2.f * dx * inv_bs;
      }
      ss = sg * sg;
      ho[r] = seed;
      ho[LD + r] = 0.001f * 2.f * sg * inv_bs;
    }
    block_sum(sx, 3);
    block_sum(ss, 5);
    __syncthreads();
    float* other = ho == bufA ? bufB : bufA;
    mlp_backward(A.h, th, A.grad, fin, T_h, ho, other, ho, bufC, bufD, false);
    for (int i = threadIdx.x; i < d0 * 32; i += NTH) gz[(i >> 5) * LD + (i & 31)] += bufD[(i >> 5) * LD + (i & 31)];
    for (int i = threadIdx.x; i < d2 * 32; i += NTH)
      gz[(d0 + d1 + (i >> 5)) * LD + (i & 31)] += bufD[(d0 + (i >> 5)) * LD + (i & 31)];
    __syncthreads();
  }
  // B7: back through e (pass F2, input v) with the summed d/d z_ -> e grads (accumulate)
  mlp_backward(A.e, th, A.grad, T_v, T_e1, gz, bufA, bufC, bufB, nullptr, true);
  if (threadIdx.x == 0) {
    const float e_adv = -red[0] * inv_bs;
    const float l2_v = red[1] / (float)(bs * p);
    const float l2_z = red[2] / (float)(bs * zd);
    const float l2_x = red[3] * inv_bs;
    const float l2_y = red[4] * inv_bs;
    const float sig = red[5] * inv_bs;
    A.losses[0] = e_adv; A.losses[1] = l2_v; A.losses[2] = l2_z; A.losses[3] = l2_x; A.losses[4] = l2_y;
    A.losses[5] = e_adv + (l2_v + A.use_z_rec * l2_z) + (l2_x + l2_y) + 0.001f * sig;     // :367
  }
}

// Keras Adam (TF 2.10 optimizer_v2, dense): t is the step AFTER increment.
__global__ void adam_kernel(float* __restrict__ theta, const float* __restrict__ grad, float* __restrict__ m,
                            float* __restrict__ v, int n, float lr_t, float b1, float b2, float eps,
                            float grad_scale) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float g = grad[i] * grad_scale;
    const float mi = m[i] + (g - m[i]) * (1.f - b1);
    const float vi = v[i] + (g * g - v[i]) * (1.f - b2);
    m[i] = mi;
    v[i] = vi;
    theta[i] -= lr_t * mi / (sqrtf(vi) + eps);
  }
}

__global__ void gather_rows_kernel(const float* __restrict__ src, int ld, const int* __restrict__ idx, int bs,
                                   int dim, float* __restrict__ dst) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < bs * dim; i += gridDim.x * blockDim.x) {
    const int r = i / dim, c = i - r * dim;
    dst[i] = src[(size_t)idx[r] * ld + c];
  }
}

}  // namespace tr
}  // namespace bgm

// ======================= iterative phase of CausalBGM.fit (N2) ====================
namespace bgm {
namespace tr {

struct IterArgs {
  Net g, f, h;
  int z_dims[4];
  int zd, p, binary, bs;
  float s2v, s2x, s2y;        // fixed variances (sigma_* keys), < 0: learned softplus head
  const float* theta;
  float* grad;                // gen-group layout; g, f, h ranges are written
  float* tape;
  const float* zt;            // latent table (n, zd)
  const float *x, *y, *v;     // full data (n), (n), (n, p)
  const int* idx;             // (bs) rows of this mini-batch
  float* losses;              // NET step: [6] loss_v, mse_v, loss_x, mse_x, loss_y, mse_y ; LATENT step: [1]
  float* gz_out;              // LATENT step: (bs, zd) d loss_postrior_z / d z rows
  int wm;
};

// gather rows idx[r] of a row-major (n, dim) array into a [dim][LD] column matrix
__device__ void gather_cols(const float* __restrict__ src, int ld, const int* __restrict__ idx, int ncols, int bs,
                            float* dst) {
  for (int i = threadIdx.x; i < ncols * 32; i += NTH) {
    const int r = i / ncols, c = i - r * ncols;
    dst[c * LD + r] = r < bs ? src[(size_t)idx[r] * ld + c] : 0.f;
  }
}

// Gaussian negative log-likelihood head (causalbgm/base.py:166-169 and the like):
// out rows [0,nd) = mu, row nd = raw variance head; tgt [nd][LD].  Writes the seeds
// d mean_r(loss_r) / d out in place; returns (thread 0 .. all threads via shared) nothing --
// partial sums go to red[slot] (loss) and red[slot+1] (sum of squared errors).
__device__ void nll_seed(float* out, const float* tgt, int nd, int bs, float s2_fixed, float* red, int slot,
                         float* rowbuf) {
  const float inv_bs = 1.f / (float)bs;
  if (threadIdx.x < 32) {
    const int r = threadIdx.x;
    float sse = 0.f;
    if (r < bs)
      for (int c = 0; c < nd; ++c) { const float d = tgt[c * LD + r] - out[c * LD + r]; sse = fmaf(d, d, sse); }
    const float raw = out[nd * LD + r];
    const float s2 = s2_fixed >= 0.f ? s2_fixed : softplus_f(raw) + 1e-6f;
    float loss = r < bs ? sse / (2.f * s2) + (float)nd * logf(s2) / 2.f : 0.f;
    rowbuf[r] = r < bs ? 1.f / s2 : 0.f;                         // 1/s2 per row for the mu seeds
    const float draw = (r < bs && s2_fixed < 0.f)
                           ? (-sse / (2.f * s2 * s2) + (float)nd / (2.f * s2)) * sigmoid_f(raw) * inv_bs : 0.f;
    out[nd * LD + r] = draw;
    float sl = loss, ss = sse;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sl += __shfl_xor_sync(0xffffffffu, sl, o);
      ss += __shfl_xor_sync(0xffffffffu, ss, o);
    }
    if (r == 0) { red[slot] = sl; red[slot + 1] = ss; }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nd * 32; i += NTH) {
    const int c = i >> 5, r = i & 31;
    out[c * LD + r] = r < bs ? -(tgt[c * LD + r] - out[c * LD + r]) * rowbuf[r] * inv_bs : 0.f;
  }
  __syncthreads();
}
// sigmoid cross-entropy head (:196): out row 0 = logit; row 1 (unused head) gets a zero seed
__device__ void bce_seed(float* out, const float* tgt, int bs, float* red, int slot) {
  if (threadIdx.x < 32) {
    const int r = threadIdx.x;
    const float lg = out[r], xv = tgt[r];
    float l = r < bs ? fmaxf(lg, 0.f) - lg * xv + log1pf(expf(-fabsf(lg))) : 0.f;
    out[r] = r < bs ? (sigmoid_f(lg) - xv) / (float)bs : 0.f;
    out[LD + r] = 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
    if (r == 0) { red[slot] = l; red[slot + 1] = l; }
  }
  __syncthreads();
}

// update_g_net / update_h_net / update_f_net gradients (:156-243), one launch.
// mode 0: parameter gradients of the three nets.  mode 1: update_latent_variable_sgd
// (:246-295): gradient of mean_r(loss_pv + loss_px + loss_py + |z|^2/2) w.r.t. the batch rows of z.
template <int MODE>
__global__ void __launch_bounds__(NTH, 1) iter_grad_kernel(const __grid_constant__ IterArgs A) {
  extern __shared__ __align__(16) float sm[];
  const int bs = A.bs, p = A.p, zd = A.zd;
  const float inv_bs = 1.f / (float)bs;
  float* bufA = sm;
  float* bufB = bufA + A.wm * LD;
  float* bufC = bufB + A.wm * LD;
  float* bufD = bufC + A.wm * LD;
  float* zmat = bufD + A.wm * LD;       // [zd][LD]
  float* gz = zmat + zd * LD;           // [zd][LD]
  float* xy = gz + zd * LD;             // [2][LD]
  float* fin = xy + 2 * LD;             // [zd+1][LD]
  float* red = fin + (zd + 1) * LD;     // [16]
  float* rowbuf = red + 16;             // [32]
  const float* th = A.theta;
  float* T_g = A.tape;
  float* T_f = T_g + net_tape_floats(A.g);
  float* T_h = T_f + net_tape_floats(A.f);
  float* T_v = T_h + net_tape_floats(A.h);
  const int d0 = A.z_dims[0], d1 = A.z_dims[1], d2 = A.z_dims[2];
  gather_cols(A.zt, zd, A.idx, zd, bs, zmat);
  gather_cols(A.x, 1, A.idx, 1, bs, xy);
  gather_cols(A.y, 1, A.idx, 1, bs, xy + LD);
  gather_cols(A.v, p, A.idx, p, bs, bufC);
  if (threadIdx.x < 16) red[threadIdx.x] = 0.f;
  __syncthreads();
  copy_mat(bufC, T_v, p);
  for (int i = threadIdx.x; i < zd * (LD / 4); i += NTH) st4(gz + i * 4, make_float4(0.f, 0.f, 0.f, 0.f));
  // ---- g: covariates ----
  {
    float* go = mlp_forward(A.g, th, zmat, bufA, bufB, T_g);
    nll_seed(go, bufC, p, bs, A.s2v, red, 0, rowbuf);
    float* other = go == bufA ? bufB : bufA;
    if (MODE == 0) mlp_backward(A.g, th, A.grad, zmat, T_g, go, other, go, bufD, nullptr, false);
    else {
      // input gradient only: reuse mlp_backward's structure without the parameter part
      int toff[MAXL]; int t = 0;
      for (int l = 0; l < A.g.L; ++l) { toff[l] = t; t += A.g.dims[l + 1] * LD; }
      float* g = go;
      for (int l = A.g.L - 1; l >= 0; --l) {
        float* o = l == 0 ? bufD : (g == other ? go : other);
        const float* xprev = l == 0 ? nullptr : T_g + toff[l - 1];
        if (xprev) { copy_mat(xprev, bufC, A.g.dims[l]); __syncthreads(); }
        dense_bwd_x(th + A.g.w_off[l], A.g.dims[l], A.g.dims[l + 1], g, o, xprev ? bufC : nullptr, false);
        g = o;
      }
      for (int i = threadIdx.x; i < zd * 32; i += NTH) gz[(i >> 5) * LD + (i & 31)] += bufD[(i >> 5) * LD + (i & 31)];
      __syncthreads();
    }
  }
  auto input_grad = [&](const Net& n, const float* tape, float* seed, float* other, float* out) {
    int toff[MAXL]; int t = 0;
    for (int l = 0; l < n.L; ++l) { toff[l] = t; t += n.dims[l + 1] * LD; }
    float* g = seed;
    for (int l = n.L - 1; l >= 0; --l) {
      float* o = l == 0 ? out : (g == other ? seed : other);
      const float* xprev = l == 0 ? nullptr : tape + toff[l - 1];
      if (xprev) { copy_mat(xprev, bufC, n.dims[l]); __syncthreads(); }
      dense_bwd_x(th + n.w_off[l], n.dims[l], n.dims[l + 1], g, o, xprev ? bufC : nullptr, false);
      g = o;
    }
  };
  // ---- h: treatment ([z0, z2]) ----
  {
    copy_mat(zmat, fin, d0);
    copy_mat(zmat + (d0 + d1) * LD, fin + d0 * LD, d2);
    __syncthreads();
    float* ho = mlp_forward(A.h, th, fin, bufA, bufB, T_h);
    if (A.binary) bce_seed(ho, xy, bs, red, 2);
    else nll_seed(ho, xy, 1, bs, A.s2x, red, 2, rowbuf);
    float* other = ho == bufA ? bufB : bufA;
    if (MODE == 0) mlp_backward(A.h, th, A.grad, fin, T_h, ho, other, ho, bufC, nullptr, false);
    else {
      input_grad(A.h, T_h, ho, other, bufD);
      for (int i = threadIdx.x; i < d0 * 32; i += NTH) gz[(i >> 5) * LD + (i & 31)] += bufD[(i >> 5) * LD + (i & 31)];
      for (int i = threadIdx.x; i < d2 * 32; i += NTH)
        gz[(d0 + d1 + (i >> 5)) * LD + (i & 31)] += bufD[(d0 + (i >> 5)) * LD + (i & 31)];
      __syncthreads();
    }
  }
  // ---- f: outcome ([z0, z1, x]) ----
  {
    copy_mat(zmat, fin, d0 + d1);
    copy_mat(xy, fin + (d0 + d1) * LD, 1);
    __syncthreads();
    float* fo = mlp_forward(A.f, th, fin, bufA, bufB, T_f);
    nll_seed(fo, xy + LD, 1, bs, A.s2y, red, 4, rowbuf);
    float* other = fo == bufA ? bufB : bufA;
    if (MODE == 0) mlp_backward(A.f, th, A.grad, fin, T_f, fo, other, fo, bufC, nullptr, false);
    else {
      input_grad(A.f, T_f, fo, other, bufD);
      for (int i = threadIdx.x; i < (d0 + d1) * 32; i += NTH) gz[(i >> 5) * LD + (i & 31)] += bufD[(i >> 5) * LD + (i & 31)];
      __syncthreads();
    }
  }
  if (MODE == 0) {
    if (threadIdx.x == 0) {
      A.losses[0] = red[0] * inv_bs; A.losses[1] = red[1] / (float)(bs * p);
      A.losses[2] = red[2] * inv_bs; A.losses[3] = red[3] * inv_bs;
      A.losses[4] = red[4] * inv_bs; A.losses[5] = red[5] * inv_bs;
    }
  } else {
    // prior |z|^2/2 (:288-289): gradient z / bs ; write the batch-row gradients
    float pr = 0.f;
    for (int i = threadIdx.x; i < zd * 32; i += NTH) {
      const int d = i >> 5, r = i & 31;
      if (r < bs) {
        const float z = zmat[d * LD + r];
        pr = fmaf(z, z, pr);
        A.gz_out[(size_t)r * zd + d] = gz[d * LD + r] + z * inv_bs;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pr += __shfl_xor_sync(0xffffffffu, pr, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(red + 6, pr);
    __syncthreads();
    if (threadIdx.x == 0) A.losses[0] = (red[0] + red[2] + red[4] + 0.5f * red[6]) * inv_bs;   // :291
  }
}

// Keras Adam on a tf.gather'ed variable (TF 2.10 optimizer_v2 _resource_apply_sparse, SURVEY
// A.4): m and v of the WHOLE table decay, the batch rows receive (1-beta) g / (1-beta) g^2,
// and every row moves by lr_t * m / (sqrt(v) + eps).  slot[row] = position of the row in
// this mini-batch or -1 (set by latent_mark_kernel, reset here).  HBM-bound: 6 floats of
// traffic per table element.
__global__ void latent_mark_kernel(const int* __restrict__ idx, int bs, int* __restrict__ slot) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < bs) slot[idx[r]] = r;
}
// The one HBM-bound kernel of the path: 6 floats of traffic per table element (z, m, v read and written) plus the
// row's slot.  Thread = 4 consecutive elements of the flat (n * zd) arrays (16-byte loads / stores); a quad lies in
// one or two rows (up to four when zd < 4) and almost never in a batch row, so the rows' slots are tested first and
// the gradient gather (with its per-element division) only runs for the bs quads that need it.  lr_dev: optional
// device-resident learning rate (replayed CUDA graphs).
template <typename IndexT>
__global__ void __launch_bounds__(256) latent_adam_sweep_kernel(float* __restrict__ z, float* __restrict__ m, float* __restrict__ v,
                                                                const int* __restrict__ slot, const float* __restrict__ gz,
                                                                long long n, int zd, float lr_t, float b1, float b2, float eps,
                                                                const float* __restrict__ lr_dev) {
  if (lr_dev) lr_t = *lr_dev;
  const long long total = n * zd;
  const long long nq = total >> 2;
  const float c1 = 1.f - b1, c2 = 1.f - b2;
  for (long long q = blockIdx.x * 256ll + threadIdx.x; q < nq; q += (long long)gridDim.x * 256ll) {
    const long long i = q << 2;
    // the three streams first: the slot test must not sit between their issue and their use
    const float4 m4 = *reinterpret_cast<const float4*>(m + i), v4 = *reinterpret_cast<const float4*>(v + i),
                 z4 = *reinterpret_cast<const float4*>(z + i);
    const IndexT r0 = (IndexT)i / (IndexT)zd, r3 = (IndexT)(i + 3) / (IndexT)zd;
    int smax = __ldg(slot + r0);
    if (r3 != r0) {
      smax = max(smax, __ldg(slot + r3));
      for (IndexT r = r0 + 1; r < r3; ++r) smax = max(smax, __ldg(slot + r));
    }
    const bool hit = smax >= 0;
    float mm[4] = {m4.x * b1, m4.y * b1, m4.z * b1, m4.w * b1};
    float vv[4] = {v4.x * b2, v4.y * b2, v4.z * b2, v4.w * b2};
    float zz[4] = {z4.x, z4.y, z4.z, z4.w};
    if (hit) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const IndexT row = (IndexT)(i + e) / (IndexT)zd;
        const int sl = slot[row];
        if (sl >= 0) {
          const float g = gz[(size_t)sl * zd + (int)((i + e) - (long long)row * zd)];
          mm[e] += c1 * g;
          vv[e] += c2 * g * g;
        }
      }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) zz[e] -= lr_t * mm[e] / (sqrtf(vv[e]) + eps);
    *reinterpret_cast<float4*>(m + i) = make_float4(mm[0], mm[1], mm[2], mm[3]);
    *reinterpret_cast<float4*>(v + i) = make_float4(vv[0], vv[1], vv[2], vv[3]);
    *reinterpret_cast<float4*>(z + i) = make_float4(zz[0], zz[1], zz[2], zz[3]);
  }
  // tail (total not a multiple of 4)
  if (blockIdx.x == 0 && threadIdx.x < (total & 3)) {
    const long long i = (nq << 2) + threadIdx.x;
    const long long row = i / zd;
    const int sl = slot[row];
    float mi = m[i] * b1, vi = v[i] * b2;
    if (sl >= 0) {
      const float g = gz[(size_t)sl * zd + (int)(i - row * zd)];
      mi += c1 * g;
      vi += c2 * g * g;
    }
    m[i] = mi;
    v[i] = vi;
    z[i] -= lr_t * mi / (sqrtf(vi) + eps);
  }
}
// launch helper shared by the fused and the layered trainers
static inline void launch_latent_adam_sweep(float* z, float* m, float* v, const int* slot, const float* gz, long long n, int zd,
                                            float lr_t, float b1, float b2, float eps, const float* lr_dev, int sm_count,
                                            cudaStream_t st) {
  const long long nq = (n * zd) >> 2;
  const int grid = (int)std::max<long long>(1, std::min<long long>((nq + 255) / 256, (long long)sm_count * 16));
  if (n * zd < (1ll << 31))
    latent_adam_sweep_kernel<unsigned int><<<grid, 256, 0, st>>>(z, m, v, slot, gz, n, zd, lr_t, b1, b2, eps, lr_dev);
  else
    latent_adam_sweep_kernel<long long><<<grid, 256, 0, st>>>(z, m, v, slot, gz, n, zd, lr_t, b1, b2, eps, lr_dev);
}
__global__ void latent_unmark_kernel(const int* __restrict__ idx, int bs, int* __restrict__ slot) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < bs) slot[idx[r]] = -1;
}

// evaluate (:534-556): sums of squared errors of the three models over all rows, and
// e_net / any Dense stack applied to many rows.  One CTA per 32 rows.
struct EvalArgs {
  Net g, f, h, e;
  int z_dims[4];
  int zd, p, binary, n;
  const float* theta;
  const float* zt;            // (n, zd) latent table, or NULL: z = e_net(v)
  const float *x, *y, *v;
  double* sums;               // [3] sum (v-v^)^2, sum (x-x^)^2, sum (y-y^)^2
  float* z_out;               // optional (n, zd): the z that was used
  int wm;
};
__global__ void __launch_bounds__(NTH) eval_kernel(const __grid_constant__ EvalArgs A) {
  extern __shared__ __align__(16) float sm[];
  const int p = A.p, zd = A.zd;
  float* bufA = sm;
  float* bufB = bufA + A.wm * LD;
  float* bufC = bufB + A.wm * LD;
  float* zmat = bufC + A.wm * LD;
  float* xy = zmat + zd * LD;
  float* fin = xy + 2 * LD;
  __shared__ float red[3];
  const int d0 = A.z_dims[0], d1 = A.z_dims[1], d2 = A.z_dims[2];
  for (int blk = blockIdx.x; blk * 32 < A.n; blk += gridDim.x) {
    const int r0 = blk * 32, bs = min(32, A.n - r0);
    if (threadIdx.x < 3) red[threadIdx.x] = 0.f;
    load_cols(A.v + (size_t)r0 * p, p, 0, p, bs, bufC);
    load_cols(A.x + r0, 1, 0, 1, bs, xy);
    load_cols(A.y + r0, 1, 0, 1, bs, xy + LD);
    __syncthreads();
    if (A.zt) {
      load_cols(A.zt + (size_t)r0 * zd, zd, 0, zd, bs, zmat);
      __syncthreads();
    } else {
      float* ze = mlp_forward(A.e, A.theta, bufC, bufA, bufB, nullptr);
      copy_mat(ze, zmat, zd);
      __syncthreads();
    }
    if (A.z_out)
      for (int i = threadIdx.x; i < zd * 32; i += NTH)
        if ((i & 31) < bs) A.z_out[(size_t)(r0 + (i & 31)) * zd + (i >> 5)] = zmat[(i >> 5) * LD + (i & 31)];
    auto acc = [&](float v, int slot) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((threadIdx.x & 31) == 0) atomicAdd(red + slot, v);
    };
    {
      float* go = mlp_forward(A.g, A.theta, zmat, bufA, bufB, nullptr);
      float s = 0.f;
      for (int i = threadIdx.x; i < p * 32; i += NTH) {
        const int c = i >> 5, r = i & 31;
        if (r < bs) { const float d = bufC[c * LD + r] - go[c * LD + r]; s = fmaf(d, d, s); }
      }
      acc(s, 0);
    }
    {
      copy_mat(zmat, fin, d0);
      copy_mat(zmat + (d0 + d1) * LD, fin + d0 * LD, d2);
      __syncthreads();
      float* ho = mlp_forward(A.h, A.theta, fin, bufA, bufB, nullptr);
      float s = 0.f;
      if (threadIdx.x < bs) {
        float xh = ho[threadIdx.x];
        if (A.binary) xh = sigmoid_f(xh);                         // :549-550
        const float d = xy[threadIdx.x] - xh;
        s = d * d;
      }
      acc(s, 1);
    }
    {
      copy_mat(zmat, fin, d0 + d1);
      copy_mat(xy, fin + (d0 + d1) * LD, 1);
      __syncthreads();
      float* fo = mlp_forward(A.f, A.theta, fin, bufA, bufB, nullptr);
      float s = 0.f;
      if (threadIdx.x < bs) { const float d = xy[LD + threadIdx.x] - fo[threadIdx.x]; s = d * d; }
      acc(s, 2);
    }
    __syncthreads();
    if (threadIdx.x < 3) atomicAdd(A.sums + threadIdx.x, (double)red[threadIdx.x]);
    __syncthreads();
  }
}

}  // namespace tr
}  // namespace bgm
