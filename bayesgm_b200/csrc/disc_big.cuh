// train_disc_step (causalbgm/base.py:305-323) for mini-batches of MORE than 32 rows.
//
// The fused kernel of train.cuh keeps a 32-row batch in registers and shared memory; this one runs the same
// algorithm -- three discriminator passes on batch statistics, the WGAN-GP term and its hand-derived double
// backward through Dense -> BatchNormalization(batch statistics) -> tanh (disc_double_backward, same sweeps, same
// symbols) -- for any batch size on row-major (B, features) matrices in an L2-resident global workspace: ONE CTA,
// phases separated by __syncthreads, element-wise phases spread over all threads, per-feature reductions over the
// rows done by one thread per feature.  It is the functional path behind `fit(batch_size > 32)`; the discriminator
// is tiny (10 -> 64 -> 32 -> 8 -> 1), a step costs a few hundred microseconds at batch 256.
#pragma once
#include "train.cuh"

namespace bgm {
namespace tr {

struct BigDiscArgs {
  Disc dz;
  int B, zd;
  const float* theta_d;
  float* grad_d;            // written
  const float *z, *zenc;    // (B, zd) prior draws and z_ = e_net(v)
  float epsilon;
  const float* eps_dev;     // optional device-resident epsilon (graph replay)
  float gp_weight;
  float* losses;            // [2]: dz_loss, d_loss
  float* ws;                // workspace, disc_big_floats(dz, B) floats
};

__host__ __device__ inline size_t disc_big_floats(const Disc& dz, int B) {
  size_t ft = 0, hid = 0, md = 0;
  for (int l = 0; l <= dz.L; ++l) { ft += dz.dims[l]; md = dz.dims[l] > (int)md ? dz.dims[l] : md; }
  for (int l = 1; l <= dz.L; ++l) hid += dz.dims[l];
  // X (ft) | U (ft) | N, H, Q, XB, NB (5 hid) | UBa, UBb, ABa, ABb (4 md) | zhat (d0) | out, norm (2) per row; s, m2, sbar per feature
  return (size_t)B * (2 * ft + 5 * hid + 4 * md + dz.dims[0] + 2) + 3 * ft + 64;
}

struct BigBufs {
  float *X[MAXL + 1], *U[MAXL + 1];
  float *N[MAXL], *H[MAXL], *Q[MAXL], *XB[MAXL], *NB[MAXL];
  float *UBa, *UBb, *ABa, *ABb, *zhat, *outv, *normv;
  float *s, *m2, *sbar;
  int foff[MAXL + 1];
};

__device__ inline void big_carve(const Disc& dz, int B, float* p, BigBufs& K) {
  int f = 0, md = 0;
  for (int l = 0; l <= dz.L; ++l) { K.X[l] = p; p += (size_t)B * dz.dims[l]; K.foff[l] = f; f += dz.dims[l]; md = max(md, dz.dims[l]); }
  for (int l = 0; l <= dz.L; ++l) { K.U[l] = p; p += (size_t)B * dz.dims[l]; }
  for (int l = 1; l <= dz.L; ++l) {
    const size_t sz = (size_t)B * dz.dims[l];
    K.N[l - 1] = p; p += sz; K.H[l - 1] = p; p += sz; K.Q[l - 1] = p; p += sz; K.XB[l - 1] = p; p += sz; K.NB[l - 1] = p; p += sz;
  }
  K.UBa = p; p += (size_t)B * md; K.UBb = p; p += (size_t)B * md; K.ABa = p; p += (size_t)B * md; K.ABb = p; p += (size_t)B * md;
  K.zhat = p; p += (size_t)B * dz.dims[0];
  K.outv = p; p += B; K.normv = p; p += B;
  K.s = p; p += f; K.m2 = p; p += f; K.sbar = p; p += f;
}

// forward of all blocks on the rows in K.X[0]; out[r] -> K.outv
__device__ inline void big_forward(const Disc& dz, const float* th, const BigBufs& K, int B) {
  const float invB = 1.f / (float)B;
  for (int l = 1; l <= dz.L; ++l) {
    const int Kin = dz.dims[l - 1], Nf = dz.dims[l];
    const float* W = th + dz.w_off[l - 1];
    float* A = K.N[l - 1];                       // pre-activations first, normalised in place below
    for (int i = threadIdx.x; i < B * Nf; i += blockDim.x) {
      const int r = i / Nf, j = i - r * Nf;
      float a = th[dz.b_off[l - 1] + j];
      const float* x = K.X[l - 1] + (size_t)r * Kin;
      for (int k = 0; k < Kin; ++k) a = fmaf(x[k], W[k * Nf + j], a);
      A[i] = a;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < Nf; j += blockDim.x) {   // batch statistics (biased variance, two passes)
      float mu = 0.f;
      for (int r = 0; r < B; ++r) mu += A[(size_t)r * Nf + j];
      mu *= invB;
      float var = 0.f;
      for (int r = 0; r < B; ++r) { const float d = A[(size_t)r * Nf + j] - mu; var = fmaf(d, d, var); }
      var *= invB;
      K.s[K.foff[l] + j] = 1.f / sqrtf(var + BN_EPS);
      K.m2[K.foff[l] + j] = mu;                  // (mean parked here until the normalisation below)
    }
    __syncthreads();
    for (int i = threadIdx.x; i < B * Nf; i += blockDim.x) {
      const int j = i % Nf;
      const float n = (A[i] - K.m2[K.foff[l] + j]) * K.s[K.foff[l] + j];
      A[i] = n;
      K.X[l][i] = tanhf(fmaf(th[dz.g_off[l - 1] + j], n, th[dz.be_off[l - 1] + j]));
    }
    __syncthreads();
  }
  const int Kl = dz.dims[dz.L];
  for (int r = threadIdx.x; r < B; r += blockDim.x) {
    float o = th[dz.b_off[dz.L]];
    for (int k = 0; k < Kl; ++k) o = fmaf(K.X[dz.L][(size_t)r * Kl + k], th[dz.w_off[dz.L] + k], o);
    K.outv[r] = o;
  }
  __syncthreads();
}

// first-order backward with d loss / d out[r] = seed; gacc != NULL: += scale * parameter gradients.  Keeps H, U, Q, m2.
__device__ inline void big_backward(const Disc& dz, const float* th, const BigBufs& K, int B, float seed, float* gacc, float scale) {
  const float invB = 1.f / (float)B;
  const int L = dz.L;
  {
    const int Kl = dz.dims[L];
    for (int i = threadIdx.x; i < B * Kl; i += blockDim.x) K.U[L][i] = seed * th[dz.w_off[L] + (i % Kl)];
    if (gacc) {
      for (int j = threadIdx.x; j < Kl; j += blockDim.x) {
        float sx = 0.f;
        for (int r = 0; r < B; ++r) sx = fmaf(seed, K.X[L][(size_t)r * Kl + j], sx);
        gacc[dz.w_off[L] + j] += scale * sx;
      }
      if (threadIdx.x == 0) gacc[dz.b_off[L]] += scale * seed * (float)B;
    }
    __syncthreads();
  }
  for (int l = L; l >= 1; --l) {
    const int Kin = dz.dims[l - 1], Nf = dz.dims[l];
    float *N = K.N[l - 1], *H = K.H[l - 1], *Q = K.Q[l - 1];
    for (int j = threadIdx.x; j < Nf; j += blockDim.x) {
      const float gam = th[dz.g_off[l - 1] + j], s = K.s[K.foff[l] + j];
      float dgam = 0.f, dbet = 0.f, m1 = 0.f, m2 = 0.f;
      for (int r = 0; r < B; ++r) {
        const size_t i = (size_t)r * Nf + j;
        const float x = K.X[l][i];
        const float v = K.U[l][i] * (1.f - x * x);
        dgam = fmaf(v, N[i], dgam);
        dbet += v;
        const float q = v * gam;
        Q[i] = q;
        m1 += q;
        m2 = fmaf(q, N[i], m2);
      }
      m1 *= invB;
      m2 *= invB;
      K.m2[K.foff[l] + j] = m2;
      float dbias = 0.f;
      for (int r = 0; r < B; ++r) {
        const size_t i = (size_t)r * Nf + j;
        const float h = s * (Q[i] - m1 - N[i] * m2);
        H[i] = h;
        dbias += h;
      }
      if (gacc) {
        gacc[dz.g_off[l - 1] + j] += scale * dgam;
        gacc[dz.be_off[l - 1] + j] += scale * dbet;
        gacc[dz.b_off[l - 1] + j] += scale * dbias;
      }
    }
    __syncthreads();
    if (gacc)
      for (int i = threadIdx.x; i < Kin * Nf; i += blockDim.x) {   // dW[k][j] = sum_r X_{l-1}[r][k] H[r][j]
        const int k = i / Nf, j = i - k * Nf;
        float d = 0.f;
        for (int r = 0; r < B; ++r) d = fmaf(K.X[l - 1][(size_t)r * Kin + k], H[(size_t)r * Nf + j], d);
        gacc[dz.w_off[l - 1] + i] += scale * d;
      }
    const float* W = th + dz.w_off[l - 1];
    for (int i = threadIdx.x; i < B * Kin; i += blockDim.x) {       // U_{l-1}[r][k] = sum_j H[r][j] W[k][j]
      const int r = i / Kin, k = i - r * Kin;
      float u = 0.f;
      for (int j = 0; j < Nf; ++j) u = fmaf(H[(size_t)r * Nf + j], W[k * Nf + j], u);
      K.U[l - 1][i] = u;
    }
    __syncthreads();
  }
}

// double backward: given UB0 = d P / d U_0 (B, d0), gacc += scale * d P / d theta (train.cuh disc_double_backward)
__device__ inline void big_double_backward(const Disc& dz, const float* th, const BigBufs& K, int B, float* UB0, float* gacc,
                                           float scale) {
  const float invB = 1.f / (float)B;
  const int L = dz.L;
  const float* UBin = UB0;
  // ---- sweep 1: reverse of the first-order backward, blocks 1..L ----
  for (int l = 1; l <= L; ++l) {
    const int Kin = dz.dims[l - 1], Nf = dz.dims[l];
    const float* W = th + dz.w_off[l - 1];
    float* UBout = (UBin == K.UBa) ? K.UBb : K.UBa;
    float* HB = K.ABa;                                  // H-bar of this block (AB buffers are free during sweep 1)
    float *N = K.N[l - 1], *H = K.H[l - 1], *Q = K.Q[l - 1];
    for (int i = threadIdx.x; i < B * Nf; i += blockDim.x) {        // Hbar[r][j] = sum_k Ubar_{l-1}[r][k] W[k][j]
      const int r = i / Nf, j = i - r * Nf;
      float hb = 0.f;
      for (int k = 0; k < Kin; ++k) hb = fmaf(UBin[(size_t)r * Kin + k], W[k * Nf + j], hb);
      HB[i] = hb;
    }
    for (int i = threadIdx.x; i < Kin * Nf; i += blockDim.x) {      // Wbar[k][j] += sum_r Ubar_{l-1}[r][k] H[r][j]
      const int k = i / Nf, j = i - k * Nf;
      float d = 0.f;
      for (int r = 0; r < B; ++r) d = fmaf(UBin[(size_t)r * Kin + k], H[(size_t)r * Nf + j], d);
      gacc[dz.w_off[l - 1] + i] += scale * d;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < Nf; j += blockDim.x) {
      const float gam = th[dz.g_off[l - 1] + j], s = K.s[K.foff[l] + j], m2 = K.m2[K.foff[l] + j];
      float sb = 0.f, m1b = 0.f, m2b = 0.f;
      for (int r = 0; r < B; ++r) {
        const size_t i = (size_t)r * Nf + j;
        const float hb = HB[i];
        sb = fmaf(hb, H[i] / s, sb);                    // C = H / s
        const float cb = hb * s;                        // C-bar
        m1b -= cb;
        m2b = fmaf(-cb, N[i], m2b);
      }
      K.sbar[K.foff[l] + j] = sb;
      float gamb = 0.f;
      for (int r = 0; r < B; ++r) {
        const size_t i = (size_t)r * Nf + j;
        const float cb = HB[i] * s;
        const float qb = cb + m1b * invB + m2b * invB * N[i];                   // Q-bar
        K.NB[l - 1][i] = -cb * m2 + m2b * invB * Q[i];                          // N-bar (part 1)
        const float x = K.X[l][i], u = K.U[l][i];
        const float t = 1.f - x * x;
        gamb = fmaf(qb, u * t, gamb);                                           // V = U * T
        const float vb = qb * gam;                                              // V-bar
        UBout[i] = vb * t;                                                      // U_l-bar
        K.XB[l - 1][i] = -2.f * x * (vb * u);                                   // X_l-bar via T
      }
      gacc[dz.g_off[l - 1] + j] += scale * gamb;
    }
    __syncthreads();
    UBin = UBout;
  }
  {  // U_L = seed * w_out^T (seed 1): w_out-bar[j] += sum_r Ubar_L[r][j]
    const int Kl = dz.dims[L];
    for (int j = threadIdx.x; j < Kl; j += blockDim.x) {
      float sacc = 0.f;
      for (int r = 0; r < B; ++r) sacc += UBin[(size_t)r * Kl + j];
      gacc[dz.w_off[L] + j] += scale * sacc;
    }
    __syncthreads();
  }
  // ---- sweep 2: reverse of the forward pass, blocks L..1 ----
  float* ABcur = K.ABa;    // A-bar of block l+1
  float* ABnew = K.ABb;
  for (int l = L; l >= 1; --l) {
    const int Kin = dz.dims[l - 1], Nf = dz.dims[l];
    const int Nup = l < L ? dz.dims[l + 1] : 0;
    const float* Wup = l < L ? th + dz.w_off[l] : nullptr;          // W_{l+1}: [Nf][Nup]
    float *N = K.N[l - 1], *XB = K.XB[l - 1], *NB = K.NB[l - 1];
    if (l < L) {
      for (int i = threadIdx.x; i < B * Nf; i += blockDim.x) {      // X_l feeds A_{l+1} = X_l W_{l+1} + b
        const int r = i / Nf, j = i - r * Nf;
        float xb = XB[i];
        for (int k = 0; k < Nup; ++k) xb = fmaf(ABcur[(size_t)r * Nup + k], Wup[j * Nup + k], xb);
        XB[i] = xb;
      }
      __syncthreads();
    }
    for (int j = threadIdx.x; j < Nf; j += blockDim.x) {
      const float gam = th[dz.g_off[l - 1] + j], s = K.s[K.foff[l] + j];
      float gamb = 0.f, betb = 0.f, stot = K.sbar[K.foff[l] + j];
      for (int r = 0; r < B; ++r) {
        const size_t i = (size_t)r * Nf + j;
        const float x = K.X[l][i];
        const float yb = XB[i] * (1.f - x * x);                                 // Y-bar
        gamb = fmaf(yb, N[i], gamb);
        betb += yb;
        const float nb = NB[i] + yb * gam;                                      // N-bar total
        NB[i] = nb;
        stot = fmaf(nb, N[i] / s, stot);                                        // + sum N-bar * D, D = N / s
      }
      const float varb = -0.5f * stot * s * s * s;
      float mean_db = 0.f;
      for (int r = 0; r < B; ++r) {
        const size_t i = (size_t)r * Nf + j;
        const float db = NB[i] * s + varb * 2.f * invB * (N[i] / s);
        ABnew[i] = db;
        mean_db += db;
      }
      mean_db *= invB;
      float bb = 0.f;
      for (int r = 0; r < B; ++r) {
        const size_t i = (size_t)r * Nf + j;
        const float ab = ABnew[i] - mean_db;                                    // A-bar
        ABnew[i] = ab;
        bb += ab;
      }
      gacc[dz.g_off[l - 1] + j] += scale * gamb;
      gacc[dz.be_off[l - 1] + j] += scale * betb;
      gacc[dz.b_off[l - 1] + j] += scale * bb;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < Kin * Nf; i += blockDim.x) {      // W_l-bar[k][j] += sum_r X_{l-1}[r][k] A-bar[r][j]
      const int k = i / Nf, j = i - k * Nf;
      float d = 0.f;
      for (int r = 0; r < B; ++r) d = fmaf(K.X[l - 1][(size_t)r * Kin + k], ABnew[(size_t)r * Nf + j], d);
      gacc[dz.w_off[l - 1] + i] += scale * d;
    }
    __syncthreads();
    float* tmp = ABcur; ABcur = ABnew; ABnew = tmp;
  }
}

__global__ void __launch_bounds__(512, 1) disc_grad_big_kernel(const __grid_constant__ BigDiscArgs A) {
  const Disc& dz = A.dz;
  const int B = A.B, zd = A.zd;
  const float eps = A.eps_dev ? *A.eps_dev : A.epsilon;
  BigBufs K;
  big_carve(dz, B, A.ws, K);
  const float* th = A.theta_d;
  float* gacc = A.grad_d;
  const float invB = 1.f / (float)B;
  for (int i = threadIdx.x; i < dz.n_params; i += blockDim.x) gacc[i] = 0.f;
  for (int i = threadIdx.x; i < B * zd; i += blockDim.x) K.zhat[i] = A.z[i] * eps + A.zenc[i] * (1.f - eps);     // :311
  __syncthreads();
  __shared__ float red[3];
  auto mean_out = [&](int slot) {
    if (threadIdx.x == 0) {
      float m = 0.f;
      for (int r = 0; r < B; ++r) m += K.outv[r];
      red[slot] = m * invB;
    }
    __syncthreads();
  };
  auto set_input = [&](const float* src) {
    for (int i = threadIdx.x; i < B * zd; i += blockDim.x) K.X[0][i] = src[i];
    __syncthreads();
  };
  // ---- D(z): -mean ----
  set_input(A.z);
  big_forward(dz, th, K, B);
  mean_out(0);
  big_backward(dz, th, K, B, -invB, gacc, 1.f);
  // ---- D(z_): +mean ----
  set_input(A.zenc);
  big_forward(dz, th, K, B);
  mean_out(1);
  big_backward(dz, th, K, B, invB, gacc, 1.f);
  // ---- gradient penalty on z_hat (:319-321) ----
  set_input(K.zhat);
  big_forward(dz, th, K, B);
  big_backward(dz, th, K, B, 1.f, nullptr, 0.f);        // U[0] = d sum(D(z_hat)) / d z_hat
  for (int r = threadIdx.x; r < B; r += blockDim.x) {
    float s2 = 0.f;
    for (int d = 0; d < zd; ++d) { const float g = K.U[0][(size_t)r * zd + d]; s2 = fmaf(g, g, s2); }
    K.normv[r] = sqrtf(s2);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float gp = 0.f;
    for (int r = 0; r < B; ++r) { const float d = K.normv[r] - 1.f; gp = fmaf(d, d, gp); }
    red[2] = gp * invB;
  }
  float* UB0 = K.UBb;                                    // sweep 1 starts from here and writes UBa first
  for (int i = threadIdx.x; i < B * zd; i += blockDim.x) {
    const float nr = K.normv[i / zd];
    // d gp / d G[r][d] = (2 / B) (|G_r| - 1) G[r][d] / |G_r|
    UB0[i] = nr > 0.f ? 2.f * invB * (nr - 1.f) * K.U[0][i] / nr : 0.f;
  }
  __syncthreads();
  big_double_backward(dz, th, K, B, UB0, gacc, A.gp_weight);
  __syncthreads();
  if (threadIdx.x == 0) {
    const float dz_loss = -red[0] + red[1];               // :316
    A.losses[0] = dz_loss;
    A.losses[1] = dz_loss + A.gp_weight * red[2];         // :323
  }
}

}  // namespace tr
}  // namespace bgm
