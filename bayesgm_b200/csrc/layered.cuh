// Layered training engine: the training steps of CausalBGM on nets of ANY width and batch size, and on
// BAYESIAN nets (networks/bnn.py: input BatchNormalization on batch statistics + DenseFlipout), as a
// sequence of small generic kernels -- one per layer-level operation, forward and hand-derived
// backward -- instead of one fused single-CTA kernel (train.cuh, which stays the fast path for
// deterministic nets at batch <= 32).  Every kernel is a plain grid over output elements; the host
// side (layered_api.cuh) strings them together per training step.
//
// DenseFlipout (tfp.layers.DenseFlipout, restated in oracle/bnn.py):
//   forward   y = a loc + ((a o s_in) dW) o s_out + b,  dW = sigma o eps, sigma = feps + softplus(rho)
//   backward  d a   = dy loc^T + ((dy o s_out) dW^T) o s_in
//             d loc = a^T dy ;  d b = sum_rows dy
//             d rho = ((a o s_in)^T (dy o s_out)) o eps o sigmoid(rho)
// Noise: the Philox streams of bnn.cuh (same keys), written once per (net call, layer) into the
// workspace as dW (K,N) and int8 sign matrices, so forward, input-backward and parameter-backward
// read the same draw.
#pragma once
#include "bnn_noise.cuh"

namespace bgm {
namespace lt {

constexpr float FEPS = 1.1920928955078125e-07f;

__device__ __forceinline__ float softplus_l(float t) { return fmaxf(t, 0.f) + log1pf(expf(-fabsf(t))); }
__device__ __forceinline__ float sigmoid_l(float t) { return 1.f / (1.f + expf(-t)); }

// dW[k][c] = (FEPS + softplus(rho[k][c])) * eps(k, c);  s_in (B,K), s_out (B,N) as +-1 int8.
// Scalars that change from step to step, kept in device memory so that a captured CUDA graph of a step can be
// replayed unchanged (layered_api.cuh, run_step): the noise call counter, the WGAN-GP interpolation draw and the
// bias-corrected Adam learning rates.
struct StepScalars {
  uint32_t ctr;
  float eps;
  float lr[4];
};
__global__ void set_scalars_kernel(StepScalars* dst, const StepScalars v) { *dst = v; }

// ctr_dev != NULL: the call id is 16 * ctr_dev[0] + call (call = the k of the step's net call).
__global__ void flipout_noise_kernel(const float* __restrict__ rho, int K, int N, uint64_t seed, int slice, int net_id,
                                     int l, uint32_t call, int64_t row0, int B, float* __restrict__ dW,
                                     signed char* __restrict__ s_in, signed char* __restrict__ s_out,
                                     const uint32_t* __restrict__ ctr_dev) {
  using namespace bgm::bnn;
  if (ctr_dev) call += 16u * ctr_dev[0];
  const int N4 = (N + 3) & ~3;
  const int64_t base = ((int64_t)slice << 44) | ((int64_t)net_id << 40) | ((int64_t)l << 36);
  const int gid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
  for (int g = gid; g < K * N4 / 4; g += gsz) {
    float e[4];
    normal4(seed, base | (int64_t)g, call, NOISE_BNN_W, 0, e);
    const int k = (g * 4) / N4, c = (g * 4) % N4;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (c + i < N) dW[(size_t)k * N + c + i] = (FEPS + softplus_l(rho[(size_t)k * N + c + i])) * e[i];
  }
  // one thread per (row, 32-bit word of the row's sign stream)
  const int words = (K + N + 31) / 32;
  for (int i = gid; i < B * words; i += gsz) {
    const int r = i / words, w = i - r * words;
    const uint32_t bits = sign_bits32(seed, row0 + r, call, net_id, l, w * 32);
    for (int j = 0; j < 32; ++j) {
      const int pos = w * 32 + j;
      const signed char s = ((bits >> j) & 1u) ? -1 : 1;
      if (pos < K) s_in[(size_t)r * K + pos] = s;
      else if (pos < K + N) s_out[(size_t)r * N + (pos - K)] = s;
    }
  }
}

// Dense / DenseFlipout forward and input gradient:
//   out[b][c] = act( sum_k A[b][k] W[k][c] + s_out[b][c] * sum_k A[b][k] s_in[b][k] dW[k][c] + bias[c] )
//   dA[b][k] (+)= mask * ( sum_c dY[b][c] W[k][c] + s_in[b][k] sum_c dY[b][c] s_out[b][c] dW[k][c] ),
//   mask = LeakyReLU'(A_post[b][k]) when A_post != NULL (the layer input was a LeakyReLU output);
//   act: 0 none, 1 LeakyReLU(0.2); dW == NULL: plain Dense.
// A 32 x 32 output tile per CTA, reduction in chunks of 32 staged through shared memory; 256 threads, thread (ty, tx)
// owns rows 4 ty .. 4 ty + 3 of column tx.  TRANS = false: forward (reduce over k, weight element W[r][m]); TRANS = true:
// input gradient (reduce over c, weight element W[m][r]).  (One thread per output element walking the whole reduction
// was latency-bound at the path's batch sizes: 24 us per launch at batch 32 against ~4 us here.)
struct DenseTileArgs {
  const float* X; int ldx;           // (B, R) rows to reduce over
  const signed char* sX;             // (B, R) signs applied to X for the perturbation product (Flipout) or NULL
  const float* W; const float* dW;   // (K, N) kernel and perturbation (NULL: plain Dense)
  const signed char* sO;             // (B, M) signs applied to the perturbation product
  const float* bias;                 // forward only
  const float* A_post; int lda;      // input gradient only: LeakyReLU mask source
  float* O; int ldo;
  int B, K, N;
  int act, accumulate;
};

template <bool TRANS>
__global__ void __launch_bounds__(256) dense_tile_kernel(const DenseTileArgs a) {
  __shared__ __align__(16) float Xs[32][36];
  __shared__ __align__(16) float Xp[32][36];
  __shared__ float Ws[32][33];
  __shared__ float Ds[32][33];
  const int R = TRANS ? a.N : a.K, M = TRANS ? a.K : a.N;
  const int m0 = blockIdx.x * 32, b0 = blockIdx.y * 32;
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  const bool flip = a.dW != nullptr;
  float acc[4] = {0.f, 0.f, 0.f, 0.f}, accp[4] = {0.f, 0.f, 0.f, 0.f};
  for (int r0 = 0; r0 < R; r0 += 32) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + i * 256, row = e >> 5, col = e & 31;
      const int b = b0 + row, r = r0 + col;
      const bool ok = b < a.B && r < R;
      const float x = ok ? a.X[(size_t)b * a.ldx + r] : 0.f;
      Xs[row][col] = x;
      if (flip) Xp[row][col] = ok ? x * (float)a.sX[(size_t)b * R + r] : 0.f;
      // weights: e -> (row-of-W-tile, contiguous index)
      float w = 0.f, d = 0.f;
      if (!TRANS) {
        const int rr = r0 + row, m = m0 + col;
        if (rr < R && m < M) {
          w = a.W[(size_t)rr * a.N + m];
          if (flip) d = a.dW[(size_t)rr * a.N + m];
        }
        Ws[row][col] = w;
        if (flip) Ds[row][col] = d;
      } else {
        const int m = m0 + row, rr = r0 + col;
        if (m < M && rr < R) {
          w = a.W[(size_t)m * a.N + rr];
          if (flip) d = a.dW[(size_t)m * a.N + rr];
        }
        Ws[col][row] = w;
        if (flip) Ds[col][row] = d;
      }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 32; r += 4) {
      float4 xv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) xv[i] = *reinterpret_cast<const float4*>(&Xs[ty * 4 + i][r]);
      const float w0 = Ws[r][tx], w1 = Ws[r + 1][tx], w2 = Ws[r + 2][tx], w3 = Ws[r + 3][tx];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[i] = fmaf(xv[i].x, w0, acc[i]);
        acc[i] = fmaf(xv[i].y, w1, acc[i]);
        acc[i] = fmaf(xv[i].z, w2, acc[i]);
        acc[i] = fmaf(xv[i].w, w3, acc[i]);
      }
      if (flip) {
#pragma unroll
        for (int i = 0; i < 4; ++i) xv[i] = *reinterpret_cast<const float4*>(&Xp[ty * 4 + i][r]);
        const float d0 = Ds[r][tx], d1 = Ds[r + 1][tx], d2 = Ds[r + 2][tx], d3 = Ds[r + 3][tx];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          accp[i] = fmaf(xv[i].x, d0, accp[i]);
          accp[i] = fmaf(xv[i].y, d1, accp[i]);
          accp[i] = fmaf(xv[i].z, d2, accp[i]);
          accp[i] = fmaf(xv[i].w, d3, accp[i]);
        }
      }
    }
    __syncthreads();
  }
  const int m = m0 + tx;
  if (m >= M) return;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int b = b0 + ty * 4 + i;
    if (b >= a.B) continue;
    float v = acc[i];
    if (flip) v += (float)a.sO[(size_t)b * M + m] * accp[i];
    if (!TRANS) {
      v += a.bias[m];
      if (a.act == 1) v = v > 0.f ? v : 0.2f * v;
      a.O[(size_t)b * a.ldo + m] = v;
    } else {
      if (a.A_post) v *= (a.A_post[(size_t)b * a.lda + m] > 0.f ? 1.f : 0.2f);
      float* dst = a.O + (size_t)b * a.ldo + m;
      *dst = a.accumulate ? *dst + v : v;
    }
  }
}

// g_W[k][c] += sum_b A[b][k] dY[b][c];  g_rho[k][c] += (sum_b A s_in s_out dY) * eps * sigmoid(rho);
// g_b[c] += sum_b dY[b][c] (threads of row k == 0).  scale multiplies everything (1 here; kept for DP).
__global__ void dense_bwd_param_kernel(const float* __restrict__ A, int lda, const float* __restrict__ dY, int ldy,
                                       const float* __restrict__ rho, const float* __restrict__ dW,
                                       const signed char* __restrict__ s_in, const signed char* __restrict__ s_out, int B,
                                       int K, int N, float* __restrict__ gW, float* __restrict__ grho,
                                       float* __restrict__ gb) {
  const int total = K * N;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int k = i / N, c = i - k * N;
    float acc = 0.f, accp = 0.f, accb = 0.f;
#pragma unroll 8
    for (int b = 0; b < B; ++b) {
      const float a = A[(size_t)b * lda + k], d = dY[(size_t)b * ldy + c];
      acc = fmaf(a, d, acc);
      if (dW) accp = fmaf(a * (float)s_in[(size_t)b * K + k], d * (float)s_out[(size_t)b * N + c], accp);
      accb += d;
    }
    gW[i] += acc;
    if (dW) {
      const float r = rho[i];
      const float sigma = FEPS + softplus_l(r);
      grho[i] += accp * (dW[i] / sigma) * sigmoid_l(r);
    }
    if (k == 0) gb[c] += accb;
  }
}

// KL(N(loc, sigma) || N(0,1)) summed over a kernel, weight w: adds w*KL to *loss and its gradient.
__global__ void kl_grad_kernel(const float* __restrict__ loc, const float* __restrict__ rho, int n, float w,
                               float* __restrict__ gloc, float* __restrict__ grho, float* __restrict__ loss) {
  float part = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float m = loc[i], r = rho[i];
    const float s = FEPS + softplus_l(r);
    part += -logf(s) + 0.5f * (s * s + m * m) - 0.5f;
    gloc[i] += w * m;
    grho[i] += w * (s - 1.f / s) * sigmoid_l(r);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if ((threadIdx.x & 31) == 0 && loss) atomicAdd(loss, w * part);
}

// Column statistics of X (B, K): mean[k], inv[k] = 1/sqrt(var + 1e-3) (biased variance).  One block per column.
__global__ void col_stats_kernel(const float* __restrict__ X, int ldx, int B, int K, float* __restrict__ mean,
                                 float* __restrict__ inv) {
  __shared__ double ra[8], rb[8];
  const int k = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double a = 0.0, b = 0.0;
  for (int r = threadIdx.x; r < B; r += blockDim.x) {
    const double v = (double)X[(size_t)r * ldx + k];
    a += v;
    b += v * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if (lane == 0) { ra[warp] = a; rb[warp] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double sa = 0.0, sb = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { sa += ra[w]; sb += rb[w]; }
    const double m = sa / B, var = fmax(sb / B - m * m, 0.0);
    mean[k] = (float)m;
    inv[k] = 1.f / sqrtf((float)var + 1e-3f);
  }
}

// xhat = (x - mean) * inv ; out = act(gamma * xhat + beta), act: 0 none, 2 tanh.  const_col >= 0: that column is
// constant over the batch (a tiled dose): xhat = 0 exactly.
__global__ void bn_fwd_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ mean,
                              const float* __restrict__ inv, const float* __restrict__ gamma,
                              const float* __restrict__ beta, int B, int K, float* __restrict__ xhat,
                              float* __restrict__ out, int act, int const_col) {
  const long long total = (long long)B * K;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / K), k = (int)(i - (long long)b * K);
    float xh = (X[(size_t)b * ldx + k] - mean[k]) * inv[k];
    if (k == const_col) xh = 0.f;
    if (xhat) xhat[i] = xh;
    float y = xh * gamma[k] + beta[k];
    if (act == 2) y = tanhf(y);
    out[i] = y;
  }
}

// BatchNorm backward, part 1: per column  s1[k] = sum_b dy,  s2[k] = sum_b dy * xhat, with
// dy = dOut * act'(out) (tanh: 1 - out^2).  Adds s2 to g_gamma and s1 to g_beta when given.
__global__ void bn_bwd_sums_kernel(const float* __restrict__ dOut, const float* __restrict__ out,
                                   const float* __restrict__ xhat, int B, int K, int act, float* __restrict__ s1,
                                   float* __restrict__ s2, float* __restrict__ ggamma, float* __restrict__ gbeta) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < K; k += gridDim.x * blockDim.x) {
    float a = 0.f, c = 0.f;
    for (int b = 0; b < B; ++b) {
      float d = dOut[(size_t)b * K + k];
      if (act == 2) { const float o = out[(size_t)b * K + k]; d *= (1.f - o * o); }
      a += d;
      c = fmaf(d, xhat[(size_t)b * K + k], c);
    }
    s1[k] = a;
    s2[k] = c;
    if (ggamma) ggamma[k] += c;
    if (gbeta) gbeta[k] += a;
  }
}
// part 2: dX[b][k] (+)= gamma inv / B * (B dy - s1 - xhat s2)
__global__ void bn_bwd_input_kernel(const float* __restrict__ dOut, const float* __restrict__ out,
                                    const float* __restrict__ xhat, const float* __restrict__ inv,
                                    const float* __restrict__ gamma, const float* __restrict__ s1,
                                    const float* __restrict__ s2, int B, int K, int act, float* __restrict__ dX, int ldd,
                                    int accumulate) {
  const long long total = (long long)B * K;
  const float invB = 1.f / (float)B;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / K), k = (int)(i - (long long)b * K);
    float d = dOut[i];
    if (act == 2) { const float o = out[i]; d *= (1.f - o * o); }
    const float v = gamma[k] * inv[k] * invB * ((float)B * d - s1[k] - xhat[i] * s2[k]);
    float* dst = dX + (size_t)b * ldd + k;
    *dst = accumulate ? *dst + v : v;
  }
}

// dst[r][0..ncol) = src[idx ? idx[r] : r][col0 .. col0+ncol)
__global__ void gather_cols_kernel(const float* __restrict__ src, int lds, const int* __restrict__ idx, int col0, int ncol,
                                   int B, float* __restrict__ dst, int ldd, int dcol0) {
  const long long total = (long long)B * ncol;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / ncol), c = (int)(i - (long long)b * ncol);
    const long long r = idx ? idx[b] : b;
    dst[(size_t)b * ldd + dcol0 + c] = src[(size_t)r * lds + col0 + c];
  }
}
// dst[b][dcol0 + c] += src[b][col0 + c]
__global__ void add_cols_kernel(const float* __restrict__ src, int lds, int col0, int ncol, int B, float* __restrict__ dst,
                                int ldd, int dcol0) {
  const long long total = (long long)B * ncol;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / ncol), c = (int)(i - (long long)b * ncol);
    dst[(size_t)b * ldd + dcol0 + c] += src[(size_t)b * lds + col0 + c];
  }
}
__global__ void fill_kernel(float* __restrict__ p, long long n, float v) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}

// ---- losses: single block (training batches are small; evaluate has its own reduction) ----
__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
  return t;
}

struct GenLossArgs {
  int B, p, zd, binary;
  float use_z_rec;
  const float *z, *v, *x, *y;          // batch (B,zd), (B,p), (B), (B)
  const float *gB, *gC, *eB, *d, *fA, *fB, *hA, *hB;   // net outputs: (B,p+1), (B,p+1), (B,zd), (B), (B,2) x4
  float *dgB, *dgC, *deB, *dd, *dfA, *dfB, *dhA, *dhB;  // output gradients (same shapes), fully written
  float* losses;                        // [6]
};
// train_gen_step losses and their gradients w.r.t. the nets' outputs (causalbgm/base.py:336-368)
__global__ void gen_loss_kernel(const GenLossArgs A) {
  __shared__ float red[8];
  const int B = A.B, p = A.p, zd = A.zd;
  const float invB = 1.f / (float)B;
  float adv = 0.f, l2v = 0.f, l2z = 0.f, l2x = 0.f, l2y = 0.f, ssl = 0.f;
  for (int i = threadIdx.x; i < B * (p + 1); i += blockDim.x) {
    const int b = i / (p + 1), j = i - b * (p + 1);
    if (j < p) {
      const float df = A.gC[i] - A.v[(size_t)b * p + j];
      l2v += df * df;
      A.dgC[i] = 2.f * df / (float)(B * p);
      A.dgB[i] = 0.f;
    } else {
      const float r = A.gB[i];
      ssl += r * r * invB;
      A.dgB[i] = 0.001f * 2.f * r * invB;
      A.dgC[i] = 0.f;
    }
  }
  for (int i = threadIdx.x; i < B * zd; i += blockDim.x) {
    const float df = A.eB[i] - A.z[i];
    l2z += df * df;
    A.deB[i] = A.use_z_rec * 2.f * df / (float)(B * zd);
  }
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    adv -= A.d[b] * invB;
    A.dd[b] = -invB;
    const float xh = A.hA[b * 2], xv = A.x[b];
    if (A.binary) {
      l2x += (fmaxf(xh, 0.f) - xh * xv + log1pf(expf(-fabsf(xh)))) * invB;
      A.dhA[b * 2] = (sigmoid_l(xh) - xv) * invB;
    } else {
      l2x += (xh - xv) * (xh - xv) * invB;
      A.dhA[b * 2] = 2.f * (xh - xv) * invB;
    }
    A.dhA[b * 2 + 1] = 0.f;
    const float yh = A.fA[b * 2], yv = A.y[b];
    l2y += (yh - yv) * (yh - yv) * invB;
    A.dfA[b * 2] = 2.f * (yh - yv) * invB;
    A.dfA[b * 2 + 1] = 0.f;
    const float rf = A.fB[b * 2 + 1], rh = A.hB[b * 2 + 1];
    ssl += (rf * rf + rh * rh) * invB;
    A.dfB[b * 2] = 0.f;
    A.dfB[b * 2 + 1] = 0.001f * 2.f * rf * invB;
    A.dhB[b * 2] = 0.f;
    A.dhB[b * 2 + 1] = 0.001f * 2.f * rh * invB;
  }
  adv = block_sum(adv, red);
  l2v = block_sum(l2v, red) / (float)(B * p);
  l2z = block_sum(l2z, red) / (float)(B * zd);
  l2x = block_sum(l2x, red);
  l2y = block_sum(l2y, red);
  ssl = block_sum(ssl, red);
  if (threadIdx.x == 0) {
    A.losses[0] = adv; A.losses[1] = l2v; A.losses[2] = l2z; A.losses[3] = l2x; A.losses[4] = l2y;
    A.losses[5] = adv + (l2v + A.use_z_rec * l2z) + (l2x + l2y) + 0.001f * ssl;
  }
}

// Gaussian NLL of the iterative phase (causalbgm/base.py:166-170, :191-207, :229-233, :262-295):
//   loss = mean_b [ sum_j (t - mu)^2 / (2 s2) + D log(s2) / 2 ],  s2 = softplus(raw) + 1e-6 or fixed.
// mu from MU (B, ldo) columns [0, D), raw from RAW (B, ldo) column rcol (MU and RAW are the same call in
// update_*_net and two different calls in update_latent_variable_sgd).  Writes dMU / dRAW (full (B, ldo)
// matrices, other entries zero) and losses[0] += loss, losses[1] = mse.  binary != 0: sigmoid cross-entropy on
// column 0 instead (D = 1).
struct NllArgs {
  int B, D, ldo, rcol, binary;
  float s2_fixed;                  // >= 0: fixed variance
  const float *target;             // (B, D)
  const float *MU, *RAW;
  float *dMU, *dRAW;               // may alias (same call): then written once, both parts
  float* losses;                   // [2]: loss (+=), mse (=)
};
__global__ void nll_loss_kernel(const NllArgs A) {
  __shared__ float red[8];
  const int B = A.B, D = A.D, ldo = A.ldo;
  const float invB = 1.f / (float)B;
  const bool same = A.dMU == A.dRAW;
  float loss = 0.f, mse = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float* mu = A.MU + (size_t)b * ldo;
    const float* t = A.target + (size_t)b * D;
    float* dmu = A.dMU + (size_t)b * ldo;
    float* draw = A.dRAW + (size_t)b * ldo;
    for (int j = 0; j < ldo; ++j) {
      dmu[j] = 0.f;
      if (!same) draw[j] = 0.f;
    }
    if (A.binary) {
      const float l = mu[0], xv = t[0];
      loss += (fmaxf(l, 0.f) - l * xv + log1pf(expf(-fabsf(l)))) * invB;
      const float pr = sigmoid_l(l);
      dmu[0] = (pr - xv) * invB;
      continue;
    }
    const float raw = A.RAW[(size_t)b * ldo + A.rcol];
    const float s2 = A.s2_fixed >= 0.f ? A.s2_fixed : softplus_l(raw) + 1e-6f;
    float sse = 0.f;
    for (int j = 0; j < D; ++j) {
      const float df = t[j] - mu[j];
      sse = fmaf(df, df, sse);
      dmu[j] = -df / s2 * invB;
    }
    loss += (sse / (2.f * s2) + (float)D * logf(s2) / 2.f) * invB;
    mse += sse;
    if (A.s2_fixed < 0.f) draw[A.rcol] += (-sse / (2.f * s2 * s2) + (float)D / (2.f * s2)) * sigmoid_l(raw) * invB;
  }
  loss = block_sum(loss, red);
  mse = block_sum(mse, red);
  if (threadIdx.x == 0) {
    A.losses[0] += loss;
    A.losses[1] = A.binary ? loss : mse / (float)(B * D);
  }
}

// Keras Adam (TF 2.10 optimizer_v2): m, v, theta updated in place on scale * grad.
__global__ void adam_kernel(float* __restrict__ theta, const float* __restrict__ grad, float* __restrict__ m,
                            float* __restrict__ v, int n, float lr_t, float b1, float b2, float eps, float scale,
                            const float* __restrict__ lr_dev) {
  if (lr_dev) lr_t = *lr_dev;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float g = scale * grad[i];
    const float mi = b1 * m[i] + (1.f - b1) * g;
    const float vi = b2 * v[i] + (1.f - b2) * g * g;
    m[i] = mi;
    v[i] = vi;
    theta[i] -= lr_t * mi / (sqrtf(vi) + eps);
  }
}

// (the dense Keras-Adam sweep over the latent table is tr::latent_adam_sweep_kernel, train.cuh)
__global__ void set_slots_kernel(int* __restrict__ slot, const int* __restrict__ idx, int B, int value_is_pos) {
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) slot[idx[b]] = value_is_pos ? b : -1;
}

// prior of update_latent_variable_sgd (:291-292): loss += sum(z^2) / (2B); z (n values, in place) -> z / B
__global__ void prior_scale_kernel(float* __restrict__ z, int n, float invB, float* __restrict__ loss) {
  __shared__ float red[8];
  float part = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = z[i];
    part = fmaf(v, v, part);
    z[i] = v * invB;
  }
  part = block_sum(part, red);
  if (threadIdx.x == 0) loss[0] += 0.5f * part * invB;
}
// loss_postrior_z = loss_pv_z + loss_px_z + loss_py_zx + loss_prior_z (:294)
__global__ void sum_losses_kernel(const float* __restrict__ l, float* __restrict__ out) {
  if (threadIdx.x == 0) out[0] = ((l[0] + l[2]) + l[4]) + l[6];
}

// ---------------------------------------------------------------- BGM flavour (bgm/base.py:145-291) ----
// Keras BatchNormalization moving statistics, momentum .99 (biased batch variance = 1/inv^2 - 1e-3)
__global__ void moving_update_kernel(float* __restrict__ moving, const float* __restrict__ mean,
                                     const float* __restrict__ inv, int K) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < K; k += gridDim.x * blockDim.x) {
    const float var = 1.f / (inv[k] * inv[k]) - 1e-3f;
    moving[k] = moving[k] * 0.99f + mean[k] * 0.01f;
    moving[K + k] = moving[K + k] * 0.99f + var * 0.01f;
  }
}
// inference-mode BatchNorm: statistics = the moving ones
__global__ void moving_to_stats_kernel(const float* __restrict__ moving, int K, float* __restrict__ mean,
                                       float* __restrict__ inv) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < K; k += gridDim.x * blockDim.x) {
    mean[k] = moving[k];
    inv[k] = 1.f / sqrtf(moving[K + k] + 1e-3f);
  }
}
// BaseVariationalNet heads + reparameterize (networks/base.py:106-117): out (B, 2 xd) = [mu | raw],
// x = mu + sqrt(softplus(raw) + 1e-6) * noise
__global__ void reparam_kernel(const float* __restrict__ out, const float* __restrict__ noise, int B, int xd,
                               float* __restrict__ x) {
  const long long total = (long long)B * xd;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / xd), j = (int)(i - (long long)b * xd);
    const float mu = out[(size_t)b * 2 * xd + j], raw = out[(size_t)b * 2 * xd + xd + j];
    x[i] = fmaf(noise[i], sqrtf(softplus_l(raw) + 1e-6f), mu);
  }
}
// its backward: dOut[:, :xd] (+)= dX ; dOut[:, xd:] (+)= dX * noise * 0.5 / sqrt(s) * sigmoid(raw)
__global__ void reparam_bwd_kernel(const float* __restrict__ dX, const float* __restrict__ out,
                                   const float* __restrict__ noise, int B, int xd, float* __restrict__ dOut, int accumulate) {
  const long long total = (long long)B * xd;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / xd), j = (int)(i - (long long)b * xd);
    const float raw = out[(size_t)b * 2 * xd + xd + j];
    const float s = softplus_l(raw) + 1e-6f;
    const float d = dX[i];
    const float dr = d * noise[i] * 0.5f / sqrtf(s) * sigmoid_l(raw);
    float* dm = dOut + (size_t)b * 2 * xd + j;
    float* dv = dm + xd;
    if (accumulate) { *dm += d; *dv += dr; } else { *dm = d; *dv = dr; }
  }
}

// LSGAN discriminator losses (bgm/base.py:221-224): dz_loss, dx_loss, d_loss and d loss / d D of the four calls
__global__ void bgm_disc_loss_kernel(const float* __restrict__ dz, const float* __restrict__ dz_,
                                     const float* __restrict__ dx, const float* __restrict__ dx_, int B,
                                     float* __restrict__ ddz, float* __restrict__ ddz_, float* __restrict__ ddx,
                                     float* __restrict__ ddx_, float* __restrict__ losses) {
  __shared__ float red[8];
  const float invB = 1.f / (float)B;
  float lz = 0.f, lx = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float a = 0.9f - dz[b], c = 0.1f - dz_[b], e = 0.9f - dx[b], f = 0.1f - dx_[b];
    lz += (a * a + c * c) * invB * 0.5f;
    lx += (e * e + f * f) * invB * 0.5f;
    ddz[b] = -a * invB; ddz_[b] = -c * invB; ddx[b] = -e * invB; ddx_[b] = -f * invB;
  }
  lz = block_sum(lz, red);
  lx = block_sum(lx, red);
  if (threadIdx.x == 0) { losses[0] = lz; losses[1] = lx; losses[2] = lx + lz; }
}

struct BgmGenLossArgs {
  int B, xd, zd;
  float alpha;
  const float *x, *z;                 // batch (B, xd), (B, zd)
  const float *out1;                  // g(z) (B, 2 xd)
  const float *x2;                    // x__ = reparameterize(g(z_)) (B, xd)
  const float *z2;                    // z__ = e(x_) (B, zd)
  const float *dx_, *dz_;             // dx_net(x_), dz_net(z_) (B)
  float *dX2, *dZ2, *ddx, *ddz;       // gradients w.r.t. x__, z__, the two discriminator outputs
  float *dOut1;                       // (B, 2 xd): the alpha * reg_loss part (raw columns), mean columns zero
  float* losses;                      // [6]: g_loss_adv, e_loss_adv, l2_loss_z, l2_loss_x, reg_loss, g_e_loss (:291)
};
__global__ void bgm_gen_loss_kernel(const BgmGenLossArgs A) {
  __shared__ float red[8];
  const int B = A.B, xd = A.xd, zd = A.zd;
  const float invB = 1.f / (float)B;
  float gadv = 0.f, eadv = 0.f, l2z = 0.f, l2x = 0.f, reg = 0.f;
  for (int i = threadIdx.x; i < B * xd; i += blockDim.x) {
    const int b = i / xd, j = i - b * xd;
    const float df = A.x2[i] - A.x[i];
    l2x += df * df;
    A.dX2[i] = 10.f * 2.f * df / (float)(B * xd);
    const float raw = A.out1[(size_t)b * 2 * xd + xd + j];
    const float s = softplus_l(raw) + 1e-6f;
    reg += s * s;
    A.dOut1[(size_t)b * 2 * xd + j] = 0.f;
    A.dOut1[(size_t)b * 2 * xd + xd + j] = A.alpha * 2.f * s / (float)(B * xd) * sigmoid_l(raw);
  }
  for (int i = threadIdx.x; i < B * zd; i += blockDim.x) {
    const float df = A.z2[i] - A.z[i];
    l2z += df * df;
    A.dZ2[i] = 10.f * 2.f * df / (float)(B * zd);
  }
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float a = 0.9f - A.dx_[b], c = 0.9f - A.dz_[b];
    gadv += a * a * invB;
    eadv += c * c * invB;
    A.ddx[b] = -2.f * a * invB;
    A.ddz[b] = -2.f * c * invB;
  }
  gadv = block_sum(gadv, red);
  eadv = block_sum(eadv, red);
  l2z = block_sum(l2z, red) / (float)(B * zd);
  l2x = block_sum(l2x, red) / (float)(B * xd);
  reg = block_sum(reg, red) / (float)(B * xd);
  if (threadIdx.x == 0) {
    A.losses[0] = gadv; A.losses[1] = eadv; A.losses[2] = l2z; A.losses[3] = l2x; A.losses[4] = reg;
    A.losses[5] = gadv + eadv + 10.f * (l2x + l2z) + A.alpha * reg;
  }
}

// Gaussian NLL with a per-element variance (bgm/base.py:150-153, :173-177): loss = mean_b sum_j [(x-mu)^2/(2s) + log(s)/2],
// out (B, 2 xd) = [mu | raw], s = softplus(raw) + 1e-6; dOut fully written; losses[0] = loss, losses[1] = mean (x-mu)^2
__global__ void bgm_nll_kernel(const float* __restrict__ x, const float* __restrict__ out, int B, int xd,
                               float* __restrict__ dOut, float* __restrict__ losses) {
  __shared__ float red[8];
  const float invB = 1.f / (float)B;
  float loss = 0.f, mse = 0.f;
  for (int i = threadIdx.x; i < B * xd; i += blockDim.x) {
    const int b = i / xd, j = i - b * xd;
    const float mu = out[(size_t)b * 2 * xd + j], raw = out[(size_t)b * 2 * xd + xd + j];
    const float s = softplus_l(raw) + 1e-6f;
    const float df = x[i] - mu;
    loss += (df * df / (2.f * s) + 0.5f * logf(s)) * invB;
    mse += df * df;
    dOut[(size_t)b * 2 * xd + j] = -df / s * invB;
    dOut[(size_t)b * 2 * xd + xd + j] = (-df * df / (2.f * s * s) + 0.5f / s) * sigmoid_l(raw) * invB;
  }
  loss = block_sum(loss, red);
  mse = block_sum(mse, red);
  if (threadIdx.x == 0) { losses[0] = loss; losses[1] = mse / (float)(B * xd); }
}

// Adam on a FRESH variable per batch (bgm/base.py:402-413: the batch rows are wrapped in a new tf.Variable, so the
// slots start at zero; the optimizer's step count is shared): z[idx[b]] -= lr_t * m / (sqrt(v) + eps)
__global__ void fresh_adam_rows_kernel(float* __restrict__ zt, const int* __restrict__ idx, const float* __restrict__ gz,
                                       int B, int zd, float lr_t, float b1, float b2, float eps) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B * zd; i += gridDim.x * blockDim.x) {
    const int b = i / zd, d = i - b * zd;
    const float g = gz[i];
    const float m = (1.f - b1) * g, v = (1.f - b2) * g * g;
    zt[(size_t)idx[b] * zd + d] -= lr_t * m / (sqrtf(v) + eps);
  }
}

// sum over rows and columns of (T - P[:, :D])^2 -> float64 accumulator
__global__ void sq_err_kernel(const float* __restrict__ T, int ldt, const float* __restrict__ P, int ldp, long long B, int D,
                              int sigmoid_p, double* __restrict__ out) {
  double part = 0.0;
  const long long total = B * D;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / D;
    const int j = (int)(i - b * D);
    float pv = P[(size_t)b * ldp + j];
    if (sigmoid_p) pv = sigmoid_l(pv);
    const double df = (double)T[(size_t)b * ldt + j] - (double)pv;
    part += df * df;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, part);
}

// IdentifiableCausalBGM (causalbgm/identifiable.py:540-543): row r of the conditional prior =
// prior_net(one_hot(u_r)) = table[seg[r]]: mu_z[zd], then sigma^2 = softplus(last output) + 1e-6.
__global__ void prior_rows_kernel(const float* __restrict__ table, int zd, const int* __restrict__ seg, int n_seg,
                                  int n, float* __restrict__ prior, int ldprior) {
  const long long total = (long long)n * (zd + 1);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / (zd + 1)), d = (int)(i - (long long)r * (zd + 1));
    const int u = min(max(seg[r], 0), n_seg - 1);
    const float t = table[(size_t)u * (zd + 1) + d];
    prior[(size_t)r * ldprior + d] = d < zd ? t : softplus_l(t) + 1e-6f;
  }
}

__global__ void eye_kernel(float* __restrict__ a, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n * n; i += gridDim.x * blockDim.x) a[i] = (i / n == i % n) ? 1.f : 0.f;
}

}  // namespace lt
}  // namespace bgm
