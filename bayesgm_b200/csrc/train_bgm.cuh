// BGM EGM training steps (bgm/base.py:190-291) on the building blocks of train.cuh.
//
// Differences from the CausalBGM steps: the generator is a BaseVariationalNet in TRAINING
// mode (input BatchNormalization with batch statistics + moving-statistics update, mean and
// softplus variance heads, reparameterised draw), there are two discriminators (dz on the
// latent, dx on the data) with LSGAN targets 0.9 / 0.1, and the gradient penalties carry the
// weight `gamma` (0 in every shipped config -> that path is skipped).
// Device parameter layout of group 0: g = [gamma(zd) | beta(zd) | hidden kernel,bias ... |
// Wcat[last][2*xd] | bcat[2*xd]] with Wcat = [W_mean | W_var] side by side (the host
// wrapper converts from / to the Keras arrays), then e.  Group 1: [dz | dx].
#pragma once
#include "train.cuh"

namespace bgm {
namespace tr {

struct VarNet {
  Net mlp;                  // dims [zd, units..., 2*xd]; last layer = the two heads side by side
  int gamma_off, beta_off;  // input BatchNormalization
  int zd, xd;
};

struct BgmGenArgs {
  VarNet g;
  Net e;
  Disc dz, dx;
  int dx_base;               // float offset of dx inside theta_d
  int zd, xd, bs;
  float alpha;               // weight of reg_loss (:279)
  const float* theta;
  const float* theta_d;
  float* grad;
  float* tape;
  float* moving;             // [2*zd] BN moving mean | moving variance (updated)
  const float *z, *x, *noise1, *noise2;   // (bs,zd) (bs,xd) (bs,xd) (bs,xd)
  float* losses;             // [6] g_loss_adv, e_loss_adv, l2_loss_z, l2_loss_x, reg_loss, g_e_loss
  int wm;
  int disc_floats;           // shared-memory floats reserved for the discriminator machinery
};

struct BgmDiscArgs {
  VarNet g;
  Net e;
  Disc dz, dx;
  int dx_base;
  int zd, xd, bs;
  float gamma, eps_z, eps_x;
  const float* theta;
  const float* theta_d;
  float* grad_d;
  float* moving;
  const float *z, *x, *noise;
  float* losses;             // [3] dz_loss, dx_loss, d_loss
  int wm;
};

// ---- input BatchNormalization, training mode (Keras: biased batch variance, eps 1e-3,
// momentum .99).  One thread per feature. ----
__device__ void bn_train_fwd(const float* th, const VarNet& g, int bs, const float* zin, float* y, float* nsave,
                             float* ssave, float* moving) {
  const float inv_bs = 1.f / (float)bs;
  for (int d = threadIdx.x; d < g.zd; d += NTH) {
    float mu = 0.f;
    for (int r = 0; r < bs; ++r) mu += zin[d * LD + r];
    mu *= inv_bs;
    float var = 0.f;
    for (int r = 0; r < bs; ++r) { const float df = zin[d * LD + r] - mu; var = fmaf(df, df, var); }
    var *= inv_bs;
    const float s = 1.f / sqrtf(var + BN_EPS);
    const float gam = th[g.gamma_off + d], bet = th[g.beta_off + d];
    ssave[d] = s;
    for (int r = 0; r < 32; ++r) {
      const float nn = r < bs ? (zin[d * LD + r] - mu) * s : 0.f;
      nsave[d * LD + r] = nn;
      y[d * LD + r] = r < bs ? fmaf(gam, nn, bet) : 0.f;
    }
    if (moving) {
      moving[d] = moving[d] * 0.99f + mu * 0.01f;
      moving[g.zd + d] = moving[g.zd + d] * 0.99f + var * 0.01f;
    }
  }
  __syncthreads();
}
__device__ void bn_train_bwd(const float* th, const VarNet& g, int bs, const float* gy, const float* nsave,
                             const float* ssave, float* grad, bool accumulate, float* gz) {
  const float inv_bs = 1.f / (float)bs;
  for (int d = threadIdx.x; d < g.zd; d += NTH) {
    const float gam = th[g.gamma_off + d], s = ssave[d];
    float dgam = 0.f, dbet = 0.f, m1 = 0.f, m2 = 0.f;
    for (int r = 0; r < bs; ++r) {
      const float v = gy[d * LD + r], nn = nsave[d * LD + r];
      dgam = fmaf(v, nn, dgam);
      dbet += v;
      m1 += v * gam;
      m2 = fmaf(v * gam, nn, m2);
    }
    m1 *= inv_bs;
    m2 *= inv_bs;
    grad[g.gamma_off + d] = accumulate ? grad[g.gamma_off + d] + dgam : dgam;
    grad[g.beta_off + d] = accumulate ? grad[g.beta_off + d] + dbet : dbet;
    if (gz)
      for (int r = 0; r < 32; ++r)
        gz[d * LD + r] = r < bs ? s * (gy[d * LD + r] * gam - m1 - nsave[d * LD + r] * m2) : 0.f;
  }
  __syncthreads();
}

// ---- heads: o = [mu | raw] (2*xd features) -> x = mu + sqrt(softplus(raw)+1e-6) * eps
// (networks/base.py:108-117).  Returns this thread's partial sum of sigma^4 (reg_loss). ----
__device__ float heads_fwd(const float* o, int xd, int bs, const float* __restrict__ noise, float* xout) {
  float part = 0.f;
  for (int i = threadIdx.x; i < xd * 32; i += NTH) {
    const int c = i >> 5, r = i & 31;
    float xv = 0.f;
    if (r < bs) {
      const float s2 = softplus_f(o[(xd + c) * LD + r]) + 1e-6f;
      xv = fmaf(noise[(size_t)r * xd + c], sqrtf(s2), o[c * LD + r]);
      part = fmaf(s2, s2, part);
    }
    xout[c * LD + r] = xv;
  }
  return part;
}
// seeds of the head layer from gx = d loss / d x:  d/d mu = gx ;
// d/d raw = (gx * eps / (2 sqrt(s2)) + creg * s2) * sigmoid(raw)   (creg = alpha * 2 / (bs*xd))
__device__ void heads_bwd_seed(const float* o, int xd, int bs, const float* __restrict__ noise, const float* gx,
                               float creg, float* so) {
  for (int i = threadIdx.x; i < xd * 32; i += NTH) {
    const int c = i >> 5, r = i & 31;
    float dmu = 0.f, draw = 0.f;
    if (r < bs) {
      const float raw = o[(xd + c) * LD + r];
      const float s2 = softplus_f(raw) + 1e-6f;
      const float gxv = gx[c * LD + r];
      dmu = gxv;
      draw = (gxv * noise[(size_t)r * xd + c] * (0.5f / sqrtf(s2)) + creg * s2) * sigmoid_f(raw);
    }
    so[c * LD + r] = dmu;
    so[(xd + c) * LD + r] = draw;
  }
  __syncthreads();
}

// LSGAN term on one discriminator pass: adds sum_r (target - D_r)^2 to *acc (thread 0 reads
// it later) and writes the per-row seeds w * d/dD_r [(target - D)^2] / bs = -2 w (target-D_r)/bs.
__device__ void lsgan_seed(const float* outv, int bs, float target, float w, float* seedv, float* acc) {
  if (threadIdx.x < 32) {
    const int r = threadIdx.x;
    const float d = r < bs ? target - outv[r] : 0.f;
    seedv[r] = -2.f * w * d / (float)bs;
    float s = d * d;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (r == 0) *acc += s;
  }
  __syncthreads();
}

// bgm/base.py:247-291
__global__ void __launch_bounds__(NTH, 1) bgm_gen_grad_kernel(const __grid_constant__ BgmGenArgs A) {
  extern __shared__ __align__(16) float sm[];
  const int bs = A.bs, xd = A.xd, zd = A.zd;
  const float inv_bs = 1.f / (float)bs;
  float* bufA = sm;
  float* bufB = bufA + A.wm * LD;
  float* bufC = bufB + A.wm * LD;
  float* bufD = bufC + A.wm * LD;
  float* zmat = bufD + A.wm * LD;      // z            [zd][LD]
  float* zenc = zmat + zd * LD;        // z_ = e(x)
  float* ymat = zenc + zd * LD;        // BN output (input of the generator stack)
  float* n1 = ymat + zd * LD;          // normalised inputs of pass 1 / pass 2
  float* n2 = n1 + zd * LD;
  float* gy = n2 + zd * LD;            // d/d BN output
  float* gz = gy + zd * LD;            // d/d z_
  const int zr = (zd + 3) & ~3;        // keep everything behind 16-byte aligned
  float* s1 = gz + zd * LD;            // [zd] 1/sqrt(var+eps) of pass 1 / 2
  float* s2 = s1 + zr;
  float* red = s2 + zr;                // [16]
  float* seedv = red + 16;             // [32]
  float* dbase = seedv + 32;
  if (threadIdx.x < 16) red[threadIdx.x] = 0.f;
  const float* th = A.theta;
  const Net& G = A.g.mlp;
  auto last_off = [](const Net& n) { int t = 0; for (int l = 0; l < n.L - 1; ++l) t += n.dims[l + 1] * LD; return t; };
  auto block_sum = [&](float v, int slot) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(red + slot, v);
  };
  float* T_g1 = A.tape;
  float* T_e1 = T_g1 + net_tape_floats(G);
  float* T_e2 = T_e1 + net_tape_floats(A.e);
  float* T_g2 = T_e2 + net_tape_floats(A.e);
  float* T_x = T_g2 + net_tape_floats(G);            // x    [xd][LD]
  float* T_xg = T_x + xd * LD;                       // x_   [xd][LD]
  float* T_y1 = T_xg + xd * LD;                      // BN outputs of the two passes [zd][LD]
  float* T_y2 = T_y1 + zd * LD;

  load_cols(A.z, zd, 0, zd, bs, zmat);
  load_cols(A.x, xd, 0, xd, bs, bufC);
  __syncthreads();
  copy_mat(bufC, T_x, xd);
  // F1: (mu, sigma^2) = g(z), x_ = reparameterize, reg = mean(sigma^4)   (:258-260)
  bn_train_fwd(th, A.g, bs, zmat, ymat, n1, s1, A.moving);
  copy_mat(ymat, T_y1, zd);
  float* o1 = mlp_forward(G, th, ymat, bufA, bufB, T_g1);
  block_sum(heads_fwd(o1, xd, bs, A.noise1, bufD), 4);
  __syncthreads();
  copy_mat(bufD, T_xg, xd);
  // F2: z_ = e(x)   (:262)
  float* ze = mlp_forward(A.e, th, bufC, bufA, bufB, T_e1);
  copy_mat(ze, zenc, zd);
  __syncthreads();
  // F3: z__ = e(x_)  (:264) ; l2_loss_z and its seed 10 * 2 (z__ - z) / (bs*zd)
  float* z2 = mlp_forward(A.e, th, bufD, bufA, bufB, T_e2);
  {
    float s = 0.f;
    for (int i = threadIdx.x; i < zd * 32; i += NTH) {
      const int d = i >> 5, r = i & 31;
      const float df = r < bs ? z2[d * LD + r] - zmat[d * LD + r] : 0.f;
      s = fmaf(df, df, s);
      bufC[d * LD + r] = 10.f * 2.f * df / (float)(bs * zd);
    }
    block_sum(s, 2);
    __syncthreads();
  }
  // B2: back through e (input x_) -> e grads (first) and d/d x_ (part D) in bufD
  mlp_backward(A.e, th, A.grad, T_xg, T_e2, bufC, bufA, bufC, bufB, bufD, false);
  // F5: dx(x_): g_loss_adv = mean((0.9 - D)^2) (:277), its gradient w.r.t. x_ added to bufD
  {
    copy_mat(T_xg, bufA, xd);
    DiscBufs B;
    float *scr, *sbar, *outv;
    disc_carve(A.dx, dbase, false, B, scr, sbar, outv, bufA);
    const float* th_s = A.theta_d + A.dx_base;
    __syncthreads();
    disc_forward(A.dx, th_s, B, bs, outv);
    lsgan_seed(outv, bs, 0.9f, 1.f, seedv, red + 0);
    disc_backward(A.dx, th_s, B, bs, 0.f, nullptr, 0.f, seedv);
    for (int i = threadIdx.x; i < xd * (LD / 4); i += NTH) {
      float4 a = ld4(bufD + i * 4);
      const float4 b = ld4(B.U[0] + i * 4);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      st4(bufD + i * 4, a);
    }
    __syncthreads();
  }
  // F4: (mu, sigma^2) = g(z_), x__ = reparameterize (:266-267); l2_loss_x and its seed
  bn_train_fwd(th, A.g, bs, zenc, ymat, n2, s2, A.moving);
  copy_mat(ymat, T_y2, zd);
  float* o2 = mlp_forward(G, th, ymat, bufA, bufB, T_g2);
  {
    float part = heads_fwd(o2, xd, bs, A.noise2, bufC);
    (void)part;
    __syncthreads();
    float s = 0.f;
    for (int i = threadIdx.x; i < xd * 32; i += NTH) {
      const int c = i >> 5, r = i & 31;
      const float df = r < bs ? bufC[c * LD + r] - T_x[c * LD + r] : 0.f;
      s = fmaf(df, df, s);
      bufC[c * LD + r] = 10.f * 2.f * df / (float)(bs * xd);
    }
    block_sum(s, 3);
    __syncthreads();
  }
  // B1: heads + generator stack of pass 2 -> g grads (first), BN backward -> d/d z_ (part C)
  {
    float* other = o2 == bufA ? bufB : bufA;
    copy_mat(o2, other, 2 * xd);       // keep o2 readable while its buffer receives the seeds
    __syncthreads();
    heads_bwd_seed(other, xd, bs, A.noise2, bufC, 0.f, o2);
    mlp_backward(G, th, A.grad, T_y2, T_g2, o2, other, o2, bufC, gy, false);
    bn_train_bwd(th, A.g, bs, gy, n2, s2, A.grad, false, gz);
  }
  // F6: dz(z_): e_loss_adv = mean((0.9 - D)^2) (:278) -> d/d z_ (part B)
  {
    DiscBufs B;
    float *scr, *sbar, *outv;
    disc_carve(A.dz, dbase, false, B, scr, sbar, outv, zenc);
    const float* th_s = A.theta_d;
    __syncthreads();
    disc_forward(A.dz, th_s, B, bs, outv);
    lsgan_seed(outv, bs, 0.9f, 1.f, seedv, red + 1);
    disc_backward(A.dz, th_s, B, bs, 0.f, nullptr, 0.f, seedv);
    for (int i = threadIdx.x; i < zd * (LD / 4); i += NTH) {
      float4 a = ld4(gz + i * 4);
      const float4 b = ld4(B.U[0] + i * 4);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      st4(gz + i * 4, a);
    }
    __syncthreads();
  }
  // B3: heads + generator stack of pass 1 with d/d x_ = A + D and the reg term -> g grads (accumulate)
  {
    const float* o1t = T_g1 + last_off(G);
    copy_mat(o1t, bufB, 2 * xd);
    __syncthreads();
    heads_bwd_seed(bufB, xd, bs, A.noise1, bufD, A.alpha * 2.f / (float)(bs * xd), bufA);
    mlp_backward(G, th, A.grad, T_y1, T_g1, bufA, bufB, bufA, bufC, gy, true);
    bn_train_bwd(th, A.g, bs, gy, n1, s1, A.grad, true, nullptr);
  }
  // B4: back through e (input x) with d/d z_ = B + C -> e grads (accumulate)
  mlp_backward(A.e, th, A.grad, T_x, T_e1, gz, bufA, bufC, bufB, nullptr, true);
  if (threadIdx.x == 0) {
    const float g_adv = red[0] * inv_bs, e_adv = red[1] * inv_bs;
    const float l2_z = red[2] / (float)(bs * zd), l2_x = red[3] / (float)(bs * xd);
    const float reg = red[4] / (float)(bs * xd);
    A.losses[0] = g_adv; A.losses[1] = e_adv; A.losses[2] = l2_z; A.losses[3] = l2_x; A.losses[4] = reg;
    A.losses[5] = g_adv + e_adv + 10.f * (l2_x + l2_z) + A.alpha * reg;       // :279
  }
}

// gradient penalty helper: given U[0] of a seed-1 backward, returns gp (all threads) and
// writes d gp / d U0 into ub0 ([d0][LD]); normv: 32 floats of shared scratch.
__device__ float gp_value_and_seed(const float* U0, int d0, int bs, float* ub0, float* normv) {
  const float inv_bs = 1.f / (float)bs;
  if (threadIdx.x < 32) {
    const int r = threadIdx.x;
    float s2 = 0.f;
    for (int d = 0; d < d0; ++d) { const float g = U0[d * LD + r]; s2 = fmaf(g, g, s2); }
    normv[r] = sqrtf(s2);
  }
  __syncthreads();
  float gp = 0.f;
  for (int r = 0; r < bs; ++r) { const float d = normv[r] - 1.f; gp = fmaf(d, d, gp); }
  gp *= inv_bs;
  for (int i = threadIdx.x; i < d0 * 32; i += NTH) {
    const int d = i >> 5, r = i & 31;
    const float nr = normv[r];
    ub0[d * LD + r] = (r < bs && nr > 0.f) ? 2.f * inv_bs * (nr - 1.f) * U0[d * LD + r] / nr : 0.f;
  }
  __syncthreads();
  return gp;
}

// One discriminator of train_disc_step: LSGAN terms on `real` (target .9) and `fake`
// (target .1), optional gradient penalty on `hat`; gradients into grad_out (global).
// Returns (all threads) the loss (mean((.9-D(real))^2) + mean((.1-D(fake))^2)) / 2 and adds gp to *gp_out.
__device__ float disc_one(const Disc& D, const float* theta_d, float* grad_out, float* dbase, int bs,
                          float* real, float* fake, float* hat, float gamma, float* ub0, float* seedv,
                          float* accs, float* gp_out) {
  DiscBufs B;
  float *scr, *sbar, *outv;
  const bool gp_on = gamma != 0.f;
  disc_carve(D, dbase, gp_on, B, scr, sbar, outv, real);
  const float* th_s = theta_d;
  float* gacc = grad_out;
  for (int i = threadIdx.x; i < D.n_params; i += NTH) gacc[i] = 0.f;
  if (threadIdx.x == 0) { accs[0] = 0.f; accs[1] = 0.f; }
  __syncthreads();
  disc_forward(D, th_s, B, bs, outv);
  lsgan_seed(outv, bs, 0.9f, 0.5f, seedv, accs + 0);
  disc_backward(D, th_s, B, bs, 0.f, gacc, 1.f, seedv);
  B.X[0] = fake;
  disc_forward(D, th_s, B, bs, outv);
  lsgan_seed(outv, bs, 0.1f, 0.5f, seedv, accs + 1);
  disc_backward(D, th_s, B, bs, 0.f, gacc, 1.f, seedv);
  if (gp_on) {
    B.X[0] = hat;
    disc_forward(D, th_s, B, bs, outv);
    disc_backward(D, th_s, B, bs, 1.f, nullptr, 0.f);
    const float gp = gp_value_and_seed(B.U[0], D.dims[0], bs, ub0, outv + 32);
    disc_double_backward(D, th_s, B, bs, ub0, scr, sbar, gacc, gamma, disc_maxd(D));
    if (threadIdx.x == 0) *gp_out += gp;
  }
  __syncthreads();
  const float loss = 0.5f * (accs[0] + accs[1]) / (float)bs;
  __syncthreads();
  return loss;
}

// bgm/base.py:190-245
__global__ void __launch_bounds__(NTH, 1) bgm_disc_grad_kernel(const __grid_constant__ BgmDiscArgs A) {
  extern __shared__ __align__(16) float sm[];
  const int bs = A.bs, xd = A.xd, zd = A.zd;
  float* bufA = sm;
  float* bufB = bufA + A.wm * LD;
  float* xmat = bufB + A.wm * LD;      // x, x_, x_hat   [xd][LD] each
  float* xgen = xmat + xd * LD;
  float* xhat = xgen + xd * LD;
  float* zmat = xhat + xd * LD;        // z, z_, z_hat   [zd][LD] each
  float* zenc = zmat + zd * LD;
  float* zhat = zenc + zd * LD;
  float* ymat = zhat + zd * LD;
  float* nsv = ymat + zd * LD;
  float* ssv = nsv + zd * LD;          // [zd]
  float* seedv = ssv + ((zd + 3) & ~3);   // [32]   (16-byte aligned)
  float* accs = seedv + 32;            // [4]
  float* dbase = accs + 4;
  const float* th = A.theta;
  load_cols(A.z, zd, 0, zd, bs, zmat);
  load_cols(A.x, xd, 0, xd, bs, xmat);
  if (threadIdx.x == 0) accs[2] = 0.f;
  __syncthreads();
  // z_ = e(x) (:203)
  copy_mat(xmat, bufA, xd);
  __syncthreads();
  float* ze = mlp_forward(A.e, th, bufA, bufA, bufB, nullptr);
  copy_mat(ze, zenc, zd);
  __syncthreads();
  // x_ = reparameterize(g(z))  (:207-208), generator in training mode
  bn_train_fwd(th, A.g, bs, zmat, ymat, nsv, ssv, A.moving);
  float* o = mlp_forward(A.g.mlp, th, ymat, bufA, bufB, nullptr);
  heads_fwd(o, xd, bs, A.noise, xgen);
  __syncthreads();
  for (int i = threadIdx.x; i < zd * 32; i += NTH) {
    const int d = i >> 5, r = i & 31;
    zhat[d * LD + r] = r < bs ? zmat[d * LD + r] * A.eps_z + zenc[d * LD + r] * (1.f - A.eps_z) : 0.f;   // :204
  }
  for (int i = threadIdx.x; i < xd * 32; i += NTH) {
    const int c = i >> 5, r = i & 31;
    xhat[c * LD + r] = r < bs ? xmat[c * LD + r] * A.eps_x + xgen[c * LD + r] * (1.f - A.eps_x) : 0.f;   // :209
  }
  __syncthreads();
  float* ub0 = bufA;                   // free now; wm >= xd
  const float dz_loss = disc_one(A.dz, A.theta_d, A.grad_d, dbase, bs, zmat, zenc, zhat, A.gamma, ub0, seedv,
                                 accs, accs + 2);
  const float dx_loss = disc_one(A.dx, A.theta_d + A.dx_base, A.grad_d + A.dx_base, dbase, bs, xmat, xgen, xhat,
                                 A.gamma, ub0, seedv, accs, accs + 2);
  if (threadIdx.x == 0) {
    A.losses[0] = dz_loss;
    A.losses[1] = dx_loss;
    A.losses[2] = dx_loss + dz_loss + A.gamma * accs[2];     // :236
  }
}

// ---------------------------------------------------------------------------------------
// Iterative phase of BGM.fit (bgm/base.py:145-187, loop :397-415).
//  MODE 0  update_g_net: loss_x = mean_r sum_c [(x-mu)^2 / (2 s2) + log(s2)/2], gradients of all
//          generator parameters (input BN gamma/beta, hidden stack, mean / variance heads).
//  MODE 1  update_latent_variable_sgd: the same likelihood + mean_r |z_r|^2 / 2, gradient w.r.t. the
//          batch rows of z THROUGH the training-mode BatchNormalization (batch statistics couple
//          the rows), then Adam on a FRESH variable (the reference wraps every batch in a new
//          tf.Variable, so the slots start at zero: step = lr_t (1-b1) g / (sqrt((1-b2) g^2) + eps),
//          SURVEY A.4) written back into the latent table (scatter_nd_update, :410-413).
// Both calls run the generator in training mode, so both update the BN moving statistics.
struct BgmIterArgs {
  VarNet g;
  int zd, xd, bs;
  const float* theta;
  float* grad;
  float* tape;
  float* moving;
  float* zt;           // latent table (n, zd)
  const float* x;      // data (n, xd)
  const int* idx;      // (bs)
  float* losses;       // MODE 0: [2] loss_x, loss_mse_x ; MODE 1: [1] loss_postrior_z
  float* gz_out;       // MODE 1, optional: (bs, zd) gradient rows (parity tests)
  float lr_t, b1, b2, eps;
  int wm;
};

template <int MODE>
__global__ void __launch_bounds__(NTH, 1) bgm_iter_kernel(const __grid_constant__ BgmIterArgs A) {
  extern __shared__ __align__(16) float sm[];
  const int bs = A.bs, xd = A.xd, zd = A.zd;
  const float inv_bs = 1.f / (float)bs;
  float* bufA = sm;
  float* bufB = bufA + A.wm * LD;
  float* bufC = bufB + A.wm * LD;      // x batch, later layer-input scratch of the backward
  float* bufD = bufC + A.wm * LD;      // seeds [2*xd][LD]
  float* zmat = bufD + A.wm * LD;      // [zd][LD]
  float* ymat = zmat + zd * LD;        // BN output
  float* nsv = ymat + zd * LD;         // normalised input
  float* gy = nsv + zd * LD;           // d / d BN output
  float* gz = gy + zd * LD;            // d / d z
  float* ssv = gz + zd * LD;           // [zd]
  float* red = ssv + ((zd + 3) & ~3);  // [8]
  const float* th = A.theta;
  const Net& G = A.g.mlp;
  float* T_g = A.tape;
  float* T_y = T_g + net_tape_floats(G);
  gather_cols(A.zt, zd, A.idx, zd, bs, zmat);
  gather_cols(A.x, xd, A.idx, xd, bs, bufC);
  if (threadIdx.x < 8) red[threadIdx.x] = 0.f;
  __syncthreads();
  bn_train_fwd(th, A.g, bs, zmat, ymat, nsv, ssv, A.moving);
  copy_mat(ymat, T_y, zd);
  float* o = mlp_forward(G, th, ymat, bufA, bufB, T_g);
  // Gaussian NLL over every column; seeds d mean_r(loss_r) / d [mu | raw]
  {
    float l = 0.f, se = 0.f;
    for (int i = threadIdx.x; i < xd * 32; i += NTH) {
      const int c = i >> 5, r = i & 31;
      float dmu = 0.f, draw = 0.f;
      if (r < bs) {
        const float raw = o[(xd + c) * LD + r];
        const float s2 = softplus_f(raw) + 1e-6f;
        const float d = bufC[c * LD + r] - o[c * LD + r];
        l += (d * d) / (2.f * s2) + 0.5f * logf(s2);
        se = fmaf(d, d, se);
        dmu = -d / s2 * inv_bs;
        draw = (-(d * d) / (2.f * s2 * s2) + 0.5f / s2) * sigmoid_f(raw) * inv_bs;
      }
      bufD[c * LD + r] = dmu;
      bufD[(xd + c) * LD + r] = draw;
    }
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) {
      l += __shfl_xor_sync(0xffffffffu, l, k);
      se += __shfl_xor_sync(0xffffffffu, se, k);
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(red + 0, l); atomicAdd(red + 1, se); }
    __syncthreads();
  }
  mlp_backward(G, th, A.grad, T_y, T_g, bufD, bufA, bufB, bufC, gy, false);
  bn_train_bwd(th, A.g, bs, gy, nsv, ssv, A.grad, false, MODE == 1 ? gz : nullptr);
  if (MODE == 0) {
    if (threadIdx.x == 0) {
      A.losses[0] = red[0] * inv_bs;
      A.losses[1] = red[1] / (float)(bs * xd);
    }
  } else {
    float pr = 0.f;
    for (int i = threadIdx.x; i < zd * 32; i += NTH) {
      const int d = i >> 5, r = i & 31;
      if (r < bs) {
        const float z = zmat[d * LD + r];
        pr = fmaf(z, z, pr);
        const float g = gz[d * LD + r] + z * inv_bs;           // prior |z|^2/2, mean over the batch
        if (A.gz_out) A.gz_out[(size_t)r * zd + d] = g;
        const float m = (1.f - A.b1) * g, v = (1.f - A.b2) * g * g;
        A.zt[(size_t)A.idx[r] * zd + d] = z - A.lr_t * m / (sqrtf(v) + A.eps);
      }
    }
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) pr += __shfl_xor_sync(0xffffffffu, pr, k);
    if ((threadIdx.x & 31) == 0) atomicAdd(red + 2, pr);
    __syncthreads();
    if (threadIdx.x == 0) A.losses[0] = (red[0] + 0.5f * red[2]) * inv_bs;
  }
}

// evaluate (:446-471) with use_x_sd = False: sum over rows and columns of (x - mu(z))^2, generator in
// inference mode (BN moving statistics).  One CTA per 32 rows.
struct BgmEvalArgs {
  VarNet g;
  int zd, xd, n;
  const float* theta;
  const float* moving;
  const float* zt;
  const float* x;
  double* sum_out;
  int wm;
};
__global__ void __launch_bounds__(NTH, 1) bgm_eval_kernel(const __grid_constant__ BgmEvalArgs A) {
  extern __shared__ __align__(16) float sm[];
  const int xd = A.xd, zd = A.zd;
  float* bufA = sm;
  float* bufB = bufA + A.wm * LD;
  float* ymat = bufB + A.wm * LD;
  const float* th = A.theta;
  for (int row0 = blockIdx.x * 32; row0 < A.n; row0 += gridDim.x * 32) {
    const int bs = min(32, A.n - row0);
    for (int i = threadIdx.x; i < zd * 32; i += NTH) {
      const int d = i >> 5, r = i & 31;
      float y = 0.f;
      if (r < bs) {
        const float z = A.zt[(size_t)(row0 + r) * zd + d];
        y = (z - A.moving[d]) / sqrtf(A.moving[zd + d] + BN_EPS) * th[A.g.gamma_off + d] + th[A.g.beta_off + d];
      }
      ymat[d * LD + r] = y;
    }
    __syncthreads();
    float* o = mlp_forward(A.g.mlp, th, ymat, bufA, bufB, nullptr);
    float se = 0.f;
    for (int i = threadIdx.x; i < xd * 32; i += NTH) {
      const int c = i >> 5, r = i & 31;
      if (r < bs) {
        const float d = A.x[(size_t)(row0 + r) * xd + c] - o[c * LD + r];
        se = fmaf(d, d, se);
      }
    }
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) se += __shfl_xor_sync(0xffffffffu, se, k);
    if ((threadIdx.x & 31) == 0) atomicAdd(A.sum_out, (double)se);
    __syncthreads();
  }
}

}  // namespace tr
}  // namespace bgm
