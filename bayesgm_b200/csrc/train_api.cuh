// Host side of the EGM training entry points (include/bgm_b200.h).  Included by bgm_b200.cu.
#pragma once
#include <cmath>
#include <cstring>
#include <vector>

#include "train_bgm.cuh"

struct bgm_trainer {
  bgm::tr::Net g, e, f, h;
  bgm::tr::Disc dz;
  int z_dims[4];
  int zd = 0, p = 0, binary = 0;
  float use_z_rec = 1.f;
  int n_gen = 0, n_disc = 0;
  float *theta[2] = {nullptr, nullptr}, *grad[2] = {nullptr, nullptr};   // group 0: g|e|f|h ; group 1: dz
  float *m[2] = {nullptr, nullptr}, *v[2] = {nullptr, nullptr};
  float* tape = nullptr;
  int tape_floats = 0;
  long long step[2] = {0, 0};
  double lr = 0, b1 = 0.9, b2 = 0.99;
  int wm = 0, smem_gen = 0, smem_disc = 0, sm_count = 0;
  int stage_disc = 0;
  // BGM flavour (kind == 1): variational generator, two discriminators
  int kind = 0;
  bgm::tr::VarNet vg;
  bgm::tr::Disc dx;
  int dx_base = 0, xd = 0;
  float alpha = 0.f, gamma = 0.f;
  float* moving = nullptr;     // [2*zd] BN moving mean | variance of the generator input
  int disc_floats = 0;
  // iterative phase of CausalBGM.fit (bgm_trainer_set_iter)
  float *m_it = nullptr, *v_it = nullptr, *gz = nullptr;
  double lr_theta = 0, lr_z = 0;
  float s2v = -1.f, s2x = -1.f, s2y = -1.f;
  long long step_it = 0, step_z = 0;
  int smem_iter = 0, smem_eval = 0;
};

namespace bgm {

static int tr_fill_net(const bgm_net_desc* d, tr::Net& n, int& off, const char* name) {
  if (!d || d->n_layers < 2 || d->n_layers > tr::MAXL || !d->dims || !d->params)
    return fail(BGM_ERR_ARG, std::string(name) + ": need 2.." + std::to_string(tr::MAXL) + " Dense layers");
  n.L = d->n_layers;
  for (int l = 0; l <= n.L; ++l) n.dims[l] = d->dims[l];
  for (int l = 0; l < n.L; ++l) {
    if (n.dims[l] < 1 || n.dims[l + 1] < 1) return fail(BGM_ERR_ARG, std::string(name) + ": non-positive layer size");
    n.w_off[l] = off;
    off += n.dims[l] * n.dims[l + 1];
    n.b_off[l] = off;
    off += n.dims[l + 1];
  }
  return 0;
}
static int tr_net_params(const bgm_net_desc* d) {
  int t = 0;
  for (int l = 0; l < d->n_layers; ++l) t += d->dims[l] * d->dims[l + 1] + d->dims[l + 1];
  return t;
}

static int tr_fill_disc(const bgm_disc_desc* d, tr::Disc& D, int in_dim, const char* name) {
  if (!d || d->n_hidden < 1 || d->n_hidden >= tr::MAXL || !d->dims || !d->params || d->dims[0] != in_dim ||
      d->dims[d->n_hidden + 1] != 1)
    return fail(BGM_ERR_ARG, std::string(name) + ": discriminator must map its input -> 1 with 1..7 hidden blocks");
  D.L = d->n_hidden;
  int doff = 0;
  for (int l = 0; l <= D.L + 1; ++l) D.dims[l] = d->dims[l];
  for (int l = 0; l <= D.L; ++l) {
    D.w_off[l] = doff; doff += D.dims[l] * D.dims[l + 1];
    D.b_off[l] = doff; doff += D.dims[l + 1];
    if (l < D.L) {
      D.g_off[l] = doff; doff += D.dims[l + 1];
      D.be_off[l] = doff; doff += D.dims[l + 1];
    }
  }
  D.n_params = doff;
  return 0;
}

}  // namespace bgm

extern "C" {

void bgm_trainer_destroy(bgm_trainer* t);

int bgm_bgmtrainer_create(bgm_trainer** out, const bgm_varnet_desc* g, const bgm_net_desc* e_net,
                          const bgm_disc_desc* dz_net, const bgm_disc_desc* dx_net, float lr, float beta_1,
                          float beta_2, float alpha, float gamma) {
  using namespace bgm;
  if (!out || !g || !e_net || !dz_net || !dx_net) return fail(BGM_ERR_ARG, "bgm_bgmtrainer_create: null argument");
  *out = nullptr;
  if (!g->units || !g->bn || !g->hidden_params || !g->mean_params || !g->var_params || g->n_hidden < 1 ||
      g->n_hidden + 1 > tr::MAXL)
    return fail(BGM_ERR_ARG, "bgm_bgmtrainer_create: bad generator description");
  const int zd = g->z_dim, xd = g->x_dim, nh = g->n_hidden;
  bgm_trainer* t = new bgm_trainer();
  t->kind = 1;
  t->zd = zd; t->xd = xd; t->p = xd;
  t->alpha = alpha; t->gamma = gamma;
  int off = 0;
  t->vg.zd = zd; t->vg.xd = xd;
  t->vg.gamma_off = off; off += zd;
  t->vg.beta_off = off; off += zd;
  tr::Net& G = t->vg.mlp;
  G.L = nh + 1;
  G.dims[0] = zd;
  for (int l = 0; l < nh; ++l) G.dims[l + 1] = g->units[l];
  G.dims[nh + 1] = 2 * xd;
  for (int l = 0; l < G.L; ++l) {
    G.w_off[l] = off; off += G.dims[l] * G.dims[l + 1];
    G.b_off[l] = off; off += G.dims[l + 1];
  }
  const int g_params = off;
  int rc;
  if ((rc = tr_fill_net(e_net, t->e, off, "e_net"))) { delete t; return rc; }
  if (t->e.dims[0] != xd || t->e.dims[t->e.L] != zd) { delete t; return fail(BGM_ERR_ARG, "bgm_bgmtrainer_create: e_net must map x_dim -> z_dim"); }
  t->n_gen = off;
  if ((rc = tr_fill_disc(dz_net, t->dz, zd, "dz_net")) || (rc = tr_fill_disc(dx_net, t->dx, xd, "dx_net"))) { delete t; return rc; }
  t->dx_base = t->dz.n_params;
  t->n_disc = t->dz.n_params + t->dx.n_params;
  t->lr = lr; t->b1 = beta_1; t->b2 = beta_2;
  int maxw = 2 * xd;
  for (int l = 0; l <= G.L; ++l) maxw = std::max(maxw, G.dims[l]);
  for (int l = 0; l <= t->e.L; ++l) maxw = std::max(maxw, t->e.dims[l]);
  t->wm = (maxw + 3) / 4 * 4;
  const bool gp_on = gamma != 0.f;
  const int dgen = std::max(tr::disc_smem_floats(t->dz, false, true), tr::disc_smem_floats(t->dx, false, true));
  const int ddis = std::max(tr::disc_smem_floats(t->dz, gp_on, true), tr::disc_smem_floats(t->dx, gp_on, true));
  t->smem_gen = (4 * t->wm * tr::LD + 7 * zd * tr::LD + 2 * zd + 8 + 16 + 32 + dgen) * 4 + 64;
  t->smem_disc = (2 * t->wm * tr::LD + 3 * xd * tr::LD + 5 * zd * tr::LD + zd + 4 + 32 + 4 + ddis) * 4 + 64;
  int dev = 0, smem_max = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&t->sm_count, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) { delete t; return fail(BGM_ERR_CUDA, std::string("bgm_bgmtrainer_create: ") + cudaGetErrorString(e)); }
  if (t->smem_gen > smem_max - 1024 || t->smem_disc > smem_max - 1024) {
    delete t;
    return fail(BGM_ERR_NOMEM, "bgm_bgmtrainer_create: x_dim (with this gamma) too large for the single-CTA training kernels");
  }
  t->tape_floats = 2 * tr::net_tape_floats(G) + 2 * tr::net_tape_floats(t->e) + (2 * xd + 2 * zd) * tr::LD + 64;
  const int n[2] = {t->n_gen, t->n_disc};
  for (int gi = 0; gi < 2 && e == cudaSuccess; ++gi) {
    const size_t bytes = sizeof(float) * (size_t)n[gi];
    float** arrs[4] = {&t->theta[gi], &t->grad[gi], &t->m[gi], &t->v[gi]};
    for (float** a : arrs) {
      if (e == cudaSuccess) e = cudaMalloc(a, bytes);
      if (e == cudaSuccess) e = cudaMemset(*a, 0, bytes);
    }
  }
  if (e == cudaSuccess) e = cudaMalloc(&t->tape, sizeof(float) * (size_t)t->tape_floats);
  if (e == cudaSuccess) e = cudaMalloc(&t->moving, sizeof(float) * 2 * zd);
  if (e == cudaSuccess) {
    // pack group 0: gamma, beta, hidden..., [W_mean | W_var], [b_mean | b_var], then e
    std::vector<float> host(t->n_gen, 0.f);
    memcpy(host.data() + t->vg.gamma_off, g->bn, sizeof(float) * zd);
    memcpy(host.data() + t->vg.beta_off, g->bn + zd, sizeof(float) * zd);
    const float* p = g->hidden_params;
    for (int l = 0; l < nh; ++l) {
      const int cnt = G.dims[l] * G.dims[l + 1] + G.dims[l + 1];
      memcpy(host.data() + G.w_off[l], p, sizeof(float) * cnt);
      p += cnt;
    }
    const int last = G.dims[nh];
    for (int k = 0; k < last; ++k)
      for (int c = 0; c < xd; ++c) {
        host[G.w_off[nh] + (size_t)k * 2 * xd + c] = g->mean_params[(size_t)k * xd + c];
        host[G.w_off[nh] + (size_t)k * 2 * xd + xd + c] = g->var_params[(size_t)k * xd + c];
      }
    for (int c = 0; c < xd; ++c) {
      host[G.b_off[nh] + c] = g->mean_params[(size_t)last * xd + c];
      host[G.b_off[nh] + xd + c] = g->var_params[(size_t)last * xd + c];
    }
    memcpy(host.data() + g_params, e_net->params, sizeof(float) * tr_net_params(e_net));
    e = cudaMemcpy(t->theta[0], host.data(), sizeof(float) * t->n_gen, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(t->moving, g->bn + 2 * zd, sizeof(float) * 2 * zd, cudaMemcpyHostToDevice);
  }
  if (e == cudaSuccess) e = cudaMemcpy(t->theta[1], dz_net->params, sizeof(float) * t->dz.n_params, cudaMemcpyHostToDevice);
  if (e == cudaSuccess)
    e = cudaMemcpy(t->theta[1] + t->dx_base, dx_net->params, sizeof(float) * t->dx.n_params, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    bgm_trainer_destroy(t);
    return fail(BGM_ERR_CUDA, std::string("bgm_bgmtrainer_create: ") + cudaGetErrorString(e));
  }
  *out = t;
  return 0;
}

int bgm_trainer_bn_moving(bgm_trainer* t, float* host_inout, int set) {
  using namespace bgm;
  if (!t || t->kind != 1 || !host_inout) return fail(BGM_ERR_ARG, "bgm_trainer_bn_moving: needs a BGM trainer and a buffer");
  if (set) BGM_CUDA_OK(cudaMemcpy(t->moving, host_inout, sizeof(float) * 2 * t->zd, cudaMemcpyHostToDevice));
  else BGM_CUDA_OK(cudaMemcpy(host_inout, t->moving, sizeof(float) * 2 * t->zd, cudaMemcpyDeviceToHost));
  return 0;
}

int bgm_bgm_train_disc_grad(bgm_trainer* t, const float* z_dev, const float* x_dev, int bs, float eps_z, float eps_x,
                            const float* noise_dev, float* losses_dev, void* stream) {
  using namespace bgm;
  if (!t || t->kind != 1 || !z_dev || !x_dev || !noise_dev || !losses_dev)
    return fail(BGM_ERR_ARG, "bgm_bgm_train_disc_grad: null argument / not a BGM trainer");
  if (bs < 2 || bs > 32) return fail(BGM_ERR_UNSUPPORTED, "bgm_bgm_train_disc_grad: batch size must be in [2, 32]");
  tr::BgmDiscArgs A;
  memset(&A, 0, sizeof(A));
  A.g = t->vg; A.e = t->e; A.dz = t->dz; A.dx = t->dx; A.dx_base = t->dx_base;
  A.zd = t->zd; A.xd = t->xd; A.bs = bs; A.gamma = t->gamma; A.eps_z = eps_z; A.eps_x = eps_x;
  A.theta = t->theta[0]; A.theta_d = t->theta[1]; A.grad_d = t->grad[1]; A.moving = t->moving;
  A.z = z_dev; A.x = x_dev; A.noise = noise_dev; A.losses = losses_dev; A.wm = t->wm;
  BGM_CUDA_OK(cudaFuncSetAttribute(tr::bgm_disc_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, t->smem_disc));
  tr::bgm_disc_grad_kernel<<<1, tr::NTH, t->smem_disc, (cudaStream_t)stream>>>(A);
  BGM_CUDA_OK(cudaGetLastError());
  return 0;
}

int bgm_bgm_train_gen_grad(bgm_trainer* t, const float* z_dev, const float* x_dev, int bs, const float* noise1_dev,
                           const float* noise2_dev, float* losses_dev, void* stream) {
  using namespace bgm;
  if (!t || t->kind != 1 || !z_dev || !x_dev || !noise1_dev || !noise2_dev || !losses_dev)
    return fail(BGM_ERR_ARG, "bgm_bgm_train_gen_grad: null argument / not a BGM trainer");
  if (bs < 2 || bs > 32) return fail(BGM_ERR_UNSUPPORTED, "bgm_bgm_train_gen_grad: batch size must be in [2, 32]");
  tr::BgmGenArgs A;
  memset(&A, 0, sizeof(A));
  A.g = t->vg; A.e = t->e; A.dz = t->dz; A.dx = t->dx; A.dx_base = t->dx_base;
  A.zd = t->zd; A.xd = t->xd; A.bs = bs; A.alpha = t->alpha;
  A.theta = t->theta[0]; A.theta_d = t->theta[1]; A.grad = t->grad[0]; A.tape = t->tape; A.moving = t->moving;
  A.z = z_dev; A.x = x_dev; A.noise1 = noise1_dev; A.noise2 = noise2_dev; A.losses = losses_dev; A.wm = t->wm;
  BGM_CUDA_OK(cudaFuncSetAttribute(tr::bgm_gen_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, t->smem_gen));
  tr::bgm_gen_grad_kernel<<<1, tr::NTH, t->smem_gen, (cudaStream_t)stream>>>(A);
  BGM_CUDA_OK(cudaGetLastError());
  return 0;
}

int bgm_trainer_create(bgm_trainer** out, const int z_dims[4], int v_dim, int binary_treatment, int use_z_rec,
                       const bgm_net_desc* g_net, const bgm_net_desc* e_net, const bgm_net_desc* f_net,
                       const bgm_net_desc* h_net, const bgm_disc_desc* dz_net, float lr, float beta_1,
                       float beta_2) {
  using namespace bgm;
  if (!out || !z_dims || !dz_net) return fail(BGM_ERR_ARG, "bgm_trainer_create: null argument");
  *out = nullptr;
  bgm_trainer* t = new bgm_trainer();
  int off = 0, rc;
  if ((rc = tr_fill_net(g_net, t->g, off, "g_net")) || (rc = tr_fill_net(e_net, t->e, off, "e_net")) ||
      (rc = tr_fill_net(f_net, t->f, off, "f_net")) || (rc = tr_fill_net(h_net, t->h, off, "h_net"))) {
    delete t;
    return rc;
  }
  t->n_gen = off;
  const int zd = z_dims[0] + z_dims[1] + z_dims[2] + z_dims[3];
  for (int i = 0; i < 4; ++i) t->z_dims[i] = z_dims[i];
  t->zd = zd;
  t->p = v_dim;
  t->binary = binary_treatment ? 1 : 0;
  t->use_z_rec = use_z_rec ? 1.f : 0.f;
  bool ok = t->g.dims[0] == zd && t->g.dims[t->g.L] == v_dim + 1 && t->e.dims[0] == v_dim &&
            t->e.dims[t->e.L] == zd && t->f.dims[0] == z_dims[0] + z_dims[1] + 1 && t->f.dims[t->f.L] == 2 &&
            t->h.dims[0] == z_dims[0] + z_dims[2] && t->h.dims[t->h.L] == 2;
  if (!ok) { delete t; return fail(BGM_ERR_ARG, "bgm_trainer_create: net shapes do not match z_dims / v_dim (causalbgm/base.py:74-81)"); }
  // discriminator
  if ((rc = tr_fill_disc(dz_net, t->dz, zd, "dz_net"))) { delete t; return rc; }
  tr::Disc& D = t->dz;
  t->n_disc = D.n_params;
  t->lr = lr; t->b1 = beta_1; t->b2 = beta_2;
  int maxw = v_dim + 1;
  for (const tr::Net* n : {&t->g, &t->e, &t->f, &t->h})
    for (int l = 0; l <= n->L; ++l) maxw = std::max(maxw, n->dims[l]);
  t->wm = (maxw + 3) / 4 * 4;
  const int gen_floats = 4 * t->wm * tr::LD + 2 * zd * tr::LD + 2 * tr::LD + (zd + 1) * tr::LD + 16 +
                         tr::disc_smem_floats(D, false);
  const int disc_floats = 2 * t->wm * tr::LD + 3 * zd * tr::LD + tr::disc_smem_floats(D, true);
  t->smem_gen = gen_floats * 4 + 64;
  t->smem_disc = disc_floats * 4 + 64;
  int dev = 0, smem_max = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&t->sm_count, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) { delete t; return fail(BGM_ERR_CUDA, std::string("bgm_trainer_create: ") + cudaGetErrorString(e)); }
  if (t->smem_gen > smem_max - 1024 || t->smem_disc > smem_max - 1024) {
    delete t;
    return fail(BGM_ERR_NOMEM, "bgm_trainer_create: v_dim / layer widths too large for the single-CTA training kernels");
  }
  // discriminator parameters + gradient accumulators in shared memory when they fit (DiscArgs.stage)
  t->stage_disc = t->smem_disc + 8 * D.n_params <= smem_max - 1024;
  if (t->stage_disc) t->smem_disc += 8 * D.n_params;
  t->tape_floats = 2 * tr::net_tape_floats(t->g) + 2 * tr::net_tape_floats(t->e) + tr::net_tape_floats(t->f) +
                   tr::net_tape_floats(t->h) + (v_dim + zd) * tr::LD + 64;
  const int n[2] = {t->n_gen, t->n_disc};
  for (int gidx = 0; gidx < 2 && e == cudaSuccess; ++gidx) {
    const size_t bytes = sizeof(float) * (size_t)n[gidx];
    float** arrs[4] = {&t->theta[gidx], &t->grad[gidx], &t->m[gidx], &t->v[gidx]};
    for (float** a : arrs) {
      if (e == cudaSuccess) e = cudaMalloc(a, bytes);
      if (e == cudaSuccess) e = cudaMemset(*a, 0, bytes);
    }
  }
  if (e == cudaSuccess) e = cudaMalloc(&t->tape, sizeof(float) * (size_t)t->tape_floats);
  // initial parameters
  if (e == cudaSuccess) {
    std::vector<float> host(t->n_gen);
    size_t o = 0;
    for (const bgm_net_desc* d : {g_net, e_net, f_net, h_net}) {
      const int cnt = tr_net_params(d);
      memcpy(host.data() + o, d->params, sizeof(float) * cnt);
      o += cnt;
    }
    e = cudaMemcpy(t->theta[0], host.data(), sizeof(float) * t->n_gen, cudaMemcpyHostToDevice);
  }
  if (e == cudaSuccess) e = cudaMemcpy(t->theta[1], dz_net->params, sizeof(float) * t->n_disc, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    bgm_trainer_destroy(t);
    return fail(BGM_ERR_CUDA, std::string("bgm_trainer_create: ") + cudaGetErrorString(e));
  }
  *out = t;
  return 0;
}

void bgm_trainer_destroy(bgm_trainer* t) {
  if (!t) return;
  for (int g = 0; g < 2; ++g) {
    if (t->theta[g]) cudaFree(t->theta[g]);
    if (t->grad[g]) cudaFree(t->grad[g]);
    if (t->m[g]) cudaFree(t->m[g]);
    if (t->v[g]) cudaFree(t->v[g]);
  }
  if (t->tape) cudaFree(t->tape);
  if (t->moving) cudaFree(t->moving);
  if (t->m_it) cudaFree(t->m_it);
  if (t->v_it) cudaFree(t->v_it);
  if (t->gz) cudaFree(t->gz);
  delete t;
}

int bgm_trainer_buffers(bgm_trainer* t, int group, int* n_params, float** theta_dev, float** grad_dev) {
  using namespace bgm;
  if (!t || group < 0 || group > 1) return fail(BGM_ERR_ARG, "bgm_trainer_buffers: bad trainer / group");
  if (n_params) *n_params = group == 0 ? t->n_gen : t->n_disc;
  if (theta_dev) *theta_dev = t->theta[group];
  if (grad_dev) *grad_dev = t->grad[group];
  return 0;
}

int bgm_trainer_get_params(bgm_trainer* t, int group, float* host_out) {
  using namespace bgm;
  if (!t || group < 0 || group > 1 || !host_out) return fail(BGM_ERR_ARG, "bgm_trainer_get_params: bad argument");
  BGM_CUDA_OK(cudaMemcpy(host_out, t->theta[group], sizeof(float) * (group == 0 ? t->n_gen : t->n_disc),
                         cudaMemcpyDeviceToHost));
  return 0;
}

int bgm_trainer_set_params(bgm_trainer* t, int group, const float* host_in) {
  using namespace bgm;
  if (!t || group < 0 || group > 1 || !host_in) return fail(BGM_ERR_ARG, "bgm_trainer_set_params: bad argument");
  BGM_CUDA_OK(cudaMemcpy(t->theta[group], host_in, sizeof(float) * (group == 0 ? t->n_gen : t->n_disc),
                         cudaMemcpyHostToDevice));
  return 0;
}

int bgm_train_disc_grad(bgm_trainer* t, const float* z_dev, const float* v_dev, int bs, float epsilon,
                        float gp_weight, float* losses_dev, void* stream) {
  using namespace bgm;
  if (!t || t->kind != 0 || !z_dev || !v_dev || !losses_dev) return fail(BGM_ERR_ARG, "bgm_train_disc_grad: null argument / not a CausalBGM trainer");
  if (bs < 2 || bs > 32) return fail(BGM_ERR_UNSUPPORTED, "bgm_train_disc_grad: batch size must be in [2, 32]");
  tr::DiscArgs A;
  memset(&A, 0, sizeof(A));
  A.e = t->e; A.dz = t->dz; A.zd = t->zd; A.p = t->p; A.bs = bs;
  A.theta = t->theta[0]; A.theta_d = t->theta[1]; A.grad_d = t->grad[1];
  A.z = z_dev; A.v = v_dev; A.epsilon = epsilon; A.gp_weight = gp_weight; A.losses = losses_dev; A.wm = t->wm;
  A.stage = t->stage_disc;
  BGM_CUDA_OK(cudaFuncSetAttribute(tr::disc_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, t->smem_disc));
  tr::disc_grad_kernel<<<1, tr::NTH, t->smem_disc, (cudaStream_t)stream>>>(A);
  BGM_CUDA_OK(cudaGetLastError());
  return 0;
}

int bgm_train_gen_grad(bgm_trainer* t, const float* z_dev, const float* v_dev, const float* x_dev,
                       const float* y_dev, int bs, float* losses_dev, void* stream) {
  using namespace bgm;
  if (!t || t->kind != 0 || !z_dev || !v_dev || !x_dev || !y_dev || !losses_dev)
    return fail(BGM_ERR_ARG, "bgm_train_gen_grad: null argument / not a CausalBGM trainer");
  if (bs < 2 || bs > 32) return fail(BGM_ERR_UNSUPPORTED, "bgm_train_gen_grad: batch size must be in [2, 32]");
  tr::GenArgs A;
  memset(&A, 0, sizeof(A));
  A.g = t->g; A.e = t->e; A.f = t->f; A.h = t->h; A.dz = t->dz;
  for (int i = 0; i < 4; ++i) A.z_dims[i] = t->z_dims[i];
  A.zd = t->zd; A.p = t->p; A.binary = t->binary; A.use_z_rec = t->use_z_rec; A.bs = bs;
  A.theta = t->theta[0]; A.theta_d = t->theta[1]; A.grad = t->grad[0]; A.tape = t->tape;
  A.tape_floats = t->tape_floats;
  A.z = z_dev; A.v = v_dev; A.x = x_dev; A.y = y_dev; A.losses = losses_dev; A.wm = t->wm;
  BGM_CUDA_OK(cudaFuncSetAttribute(tr::gen_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, t->smem_gen));
  tr::gen_grad_kernel<<<1, tr::NTH, t->smem_gen, (cudaStream_t)stream>>>(A);
  BGM_CUDA_OK(cudaGetLastError());
  return 0;
}

int bgm_train_adam(bgm_trainer* t, int group, float grad_scale, void* stream) {
  using namespace bgm;
  if (!t || group < 0 || group > 1) return fail(BGM_ERR_ARG, "bgm_train_adam: bad trainer / group");
  const int n = group == 0 ? t->n_gen : t->n_disc;
  t->step[group] += 1;
  const double k = (double)t->step[group];
  const float lr_t = (float)(t->lr * std::sqrt(1.0 - std::pow(t->b2, k)) / (1.0 - std::pow(t->b1, k)));
  const int grid = std::max(1, std::min((n + 255) / 256, t->sm_count * 4));
  tr::adam_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(t->theta[group], t->grad[group], t->m[group], t->v[group],
                                                          n, lr_t, (float)t->b1, (float)t->b2, 1e-7f, grad_scale);
  BGM_CUDA_OK(cudaGetLastError());
  return 0;
}

int bgm_gather_rows(const float* src_dev, int ld, const int* idx_dev, int bs, int dim, float* dst_dev, void* stream) {
  using namespace bgm;
  if (!src_dev || !idx_dev || !dst_dev || bs < 1 || dim < 1 || ld < dim)
    return fail(BGM_ERR_ARG, "bgm_gather_rows: bad argument");
  const int grid = std::max(1, std::min((bs * dim + 255) / 256, 148));
  tr::gather_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src_dev, ld, idx_dev, bs, dim, dst_dev);
  BGM_CUDA_OK(cudaGetLastError());
  return 0;
}


// ---------------------------------------------- iterative phase of CausalBGM.fit ----
int bgm_trainer_set_iter(bgm_trainer* t, float lr_theta, float lr_z, float sigma_v, float sigma_x, float sigma_y) {
  using namespace bgm;
  if (!t || t->kind != 0) return fail(BGM_ERR_ARG, "bgm_trainer_set_iter: needs a CausalBGM trainer");
  t->lr_theta = lr_theta; t->lr_z = lr_z;
  t->s2v = sigma_v >= 0.f ? sigma_v * sigma_v : -1.f;
  t->s2x = sigma_x >= 0.f ? sigma_x * sigma_x : -1.f;
  t->s2y = sigma_y >= 0.f ? sigma_y * sigma_y : -1.f;
  if (!t->m_it) {
    const size_t bytes = sizeof(float) * (size_t)t->n_gen;
    BGM_CUDA_OK(cudaMalloc(&t->m_it, bytes));
    BGM_CUDA_OK(cudaMalloc(&t->v_it, bytes));
    BGM_CUDA_OK(cudaMalloc(&t->gz, sizeof(float) * 32 * t->zd));
  }
  BGM_CUDA_OK(cudaMemset(t->m_it, 0, sizeof(float) * (size_t)t->n_gen));
  BGM_CUDA_OK(cudaMemset(t->v_it, 0, sizeof(float) * (size_t)t->n_gen));
  t->step_it = t->step_z = 0;
  t->smem_iter = (4 * t->wm * tr::LD + 2 * t->zd * tr::LD + 2 * tr::LD + (t->zd + 1) * tr::LD + 16 + 32) * 4 + 64;
  t->smem_eval = (3 * t->wm * tr::LD + t->zd * tr::LD + 2 * tr::LD + (t->zd + 1) * tr::LD) * 4 + 64;
  return 0;
}

static void iter_fill(const bgm_trainer* t, bgm::tr::IterArgs& A, const float* zt, const float* x, const float* y,
                      const float* v, const int* idx, int bs) {
  memset(&A, 0, sizeof(A));
  A.g = t->g; A.f = t->f; A.h = t->h;
  for (int i = 0; i < 4; ++i) A.z_dims[i] = t->z_dims[i];
  A.zd = t->zd; A.p = t->p; A.binary = t->binary; A.bs = bs;
  A.s2v = t->s2v; A.s2x = t->s2x; A.s2y = t->s2y;
  A.theta = t->theta[0]; A.grad = t->grad[0]; A.tape = t->tape;
  A.zt = zt; A.x = x; A.y = y; A.v = v; A.idx = idx; A.wm = t->wm;
}

int bgm_train_iter_nets(bgm_trainer* t, const float* zt_dev, const float* x_dev, const float* y_dev,
                        const float* v_dev, const int* idx_dev, int bs, int apply, float grad_scale,
                        float* losses_dev, void* stream) {
  using namespace bgm;
  if (!t || t->kind != 0 || !t->m_it) return fail(BGM_ERR_ARG, "bgm_train_iter_nets: call bgm_trainer_set_iter first");
  if (!zt_dev || !x_dev || !y_dev || !v_dev || !idx_dev || !losses_dev) return fail(BGM_ERR_ARG, "bgm_train_iter_nets: null argument");
  if (bs < 1 || bs > 32) return fail(BGM_ERR_UNSUPPORTED, "bgm_train_iter_nets: batch size must be in [1, 32]");
  cudaStream_t st = (cudaStream_t)stream;
  if (apply != 2) {
    tr::IterArgs A;
    iter_fill(t, A, zt_dev, x_dev, y_dev, v_dev, idx_dev, bs);
    A.losses = losses_dev;
    BGM_CUDA_OK(cudaFuncSetAttribute(tr::iter_grad_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, t->smem_iter));
    tr::iter_grad_kernel<0><<<1, tr::NTH, t->smem_iter, st>>>(A);
    BGM_CUDA_OK(cudaGetLastError());
  }
  if (apply) {   // g_optimizer, h_optimizer, f_optimizer (:89-91): same hyper-parameters, same step count
    t->step_it += 1;
    const double k = (double)t->step_it;
    const float lr_t = (float)(t->lr_theta * std::sqrt(1.0 - std::pow(0.99, k)) / (1.0 - std::pow(0.9, k)));
    const int ranges[2][2] = {{t->g.w_off[0], t->e.w_off[0]}, {t->f.w_off[0], t->n_gen}};
    for (auto& r : ranges) {
      const int n = r[1] - r[0];
      const int grid = std::max(1, std::min((n + 255) / 256, t->sm_count * 4));
      tr::adam_kernel<<<grid, 256, 0, st>>>(t->theta[0] + r[0], t->grad[0] + r[0], t->m_it + r[0], t->v_it + r[0], n,
                                            lr_t, 0.9f, 0.99f, 1e-7f, grad_scale);
      BGM_CUDA_OK(cudaGetLastError());
    }
  }
  return 0;
}

int bgm_train_iter_latent(bgm_trainer* t, float* zt_dev, float* m_dev, float* v_dev_adam, int* slot_dev,
                          long long n, const float* x_dev, const float* y_dev, const float* v_dev,
                          const int* idx_dev, int bs, float* loss_dev, void* stream) {
  using namespace bgm;
  if (!t || t->kind != 0 || !t->m_it) return fail(BGM_ERR_ARG, "bgm_train_iter_latent: call bgm_trainer_set_iter first");
  if (!zt_dev || !m_dev || !v_dev_adam || !slot_dev || !x_dev || !y_dev || !v_dev || !idx_dev || !loss_dev)
    return fail(BGM_ERR_ARG, "bgm_train_iter_latent: null argument");
  if (bs < 1 || bs > 32 || n < bs) return fail(BGM_ERR_UNSUPPORTED, "bgm_train_iter_latent: batch size must be in [1, 32]");
  cudaStream_t st = (cudaStream_t)stream;
  tr::IterArgs A;
  iter_fill(t, A, zt_dev, x_dev, y_dev, v_dev, idx_dev, bs);
  A.losses = loss_dev;
  A.gz_out = t->gz;
  BGM_CUDA_OK(cudaFuncSetAttribute(tr::iter_grad_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, t->smem_iter));
  tr::iter_grad_kernel<1><<<1, tr::NTH, t->smem_iter, st>>>(A);
  BGM_CUDA_OK(cudaGetLastError());
  t->step_z += 1;
  const double k = (double)t->step_z;
  const float lr_t = (float)(t->lr_z * std::sqrt(1.0 - std::pow(0.99, k)) / (1.0 - std::pow(0.9, k)));
  tr::latent_mark_kernel<<<1, 32, 0, st>>>(idx_dev, bs, slot_dev);
  tr::launch_latent_adam_sweep(zt_dev, m_dev, v_dev_adam, slot_dev, t->gz, n, t->zd, lr_t, 0.9f, 0.99f, 1e-7f, nullptr, t->sm_count, st);
  tr::latent_unmark_kernel<<<1, 32, 0, st>>>(idx_dev, bs, slot_dev);
  BGM_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------- iterative phase of BGM.fit ----
int bgm_bgmtrainer_set_iter(bgm_trainer* t, float lr_theta, float lr_z) {
  using namespace bgm;
  if (!t || t->kind != 1) return fail(BGM_ERR_ARG, "bgm_bgmtrainer_set_iter: needs a BGM trainer");
  t->lr_theta = lr_theta; t->lr_z = lr_z;
  const int g_params = t->e.w_off[0];
  if (!t->m_it) {
    BGM_CUDA_OK(cudaMalloc(&t->m_it, sizeof(float) * (size_t)g_params));
    BGM_CUDA_OK(cudaMalloc(&t->v_it, sizeof(float) * (size_t)g_params));
  }
  BGM_CUDA_OK(cudaMemset(t->m_it, 0, sizeof(float) * (size_t)g_params));
  BGM_CUDA_OK(cudaMemset(t->v_it, 0, sizeof(float) * (size_t)g_params));
  t->step_it = t->step_z = 0;
  t->smem_iter = (4 * t->wm * tr::LD + 5 * t->zd * tr::LD + ((t->zd + 3) & ~3) + 8) * 4 + 64;
  t->smem_eval = (2 * t->wm * tr::LD + t->zd * tr::LD) * 4 + 64;
  return 0;
}

static void bgm_iter_fill(const bgm_trainer* t, bgm::tr::BgmIterArgs& A, float* zt, const float* x, const int* idx,
                          int bs) {
  memset(&A, 0, sizeof(A));
  A.g = t->vg; A.zd = t->zd; A.xd = t->xd; A.bs = bs;
  A.theta = t->theta[0]; A.grad = t->grad[0]; A.tape = t->tape; A.moving = t->moving;
  A.zt = zt; A.x = x; A.idx = idx; A.wm = t->wm;
}

int bgm_bgm_iter_g(bgm_trainer* t, const float* zt_dev, const float* x_dev, const int* idx_dev, int bs, int apply,
                   float grad_scale, float* losses_dev, void* stream) {
  using namespace bgm;
  if (!t || t->kind != 1 || !t->m_it) return fail(BGM_ERR_ARG, "bgm_bgm_iter_g: call bgm_bgmtrainer_set_iter first");
  if (!zt_dev || !x_dev || !idx_dev || !losses_dev) return fail(BGM_ERR_ARG, "bgm_bgm_iter_g: null argument");
  if (bs < 1 || bs > 32) return fail(BGM_ERR_UNSUPPORTED, "bgm_bgm_iter_g: batch size must be in [1, 32]");
  cudaStream_t st = (cudaStream_t)stream;
  if (apply != 2) {
    tr::BgmIterArgs A;
    bgm_iter_fill(t, A, const_cast<float*>(zt_dev), x_dev, idx_dev, bs);
    A.losses = losses_dev;
    BGM_CUDA_OK(cudaFuncSetAttribute(tr::bgm_iter_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, t->smem_iter));
    tr::bgm_iter_kernel<0><<<1, tr::NTH, t->smem_iter, st>>>(A);
    BGM_CUDA_OK(cudaGetLastError());
  }
  if (apply) {   // g_optimizer (:87): Adam(lr_theta, 0.9, 0.99) on the generator parameters only
    t->step_it += 1;
    const double k = (double)t->step_it;
    const float lr_t = (float)(t->lr_theta * std::sqrt(1.0 - std::pow(0.99, k)) / (1.0 - std::pow(0.9, k)));
    const int n = t->e.w_off[0];
    const int grid = std::max(1, std::min((n + 255) / 256, t->sm_count * 4));
    tr::adam_kernel<<<grid, 256, 0, st>>>(t->theta[0], t->grad[0], t->m_it, t->v_it, n, lr_t, 0.9f, 0.99f, 1e-7f,
                                          grad_scale);
    BGM_CUDA_OK(cudaGetLastError());
  }
  return 0;
}

int bgm_bgm_iter_latent(bgm_trainer* t, float* zt_dev, const float* x_dev, const int* idx_dev, int bs,
                        float* loss_dev, float* gz_out_dev, void* stream) {
  using namespace bgm;
  if (!t || t->kind != 1 || !t->m_it) return fail(BGM_ERR_ARG, "bgm_bgm_iter_latent: call bgm_bgmtrainer_set_iter first");
  if (!zt_dev || !x_dev || !idx_dev || !loss_dev) return fail(BGM_ERR_ARG, "bgm_bgm_iter_latent: null argument");
  if (bs < 1 || bs > 32) return fail(BGM_ERR_UNSUPPORTED, "bgm_bgm_iter_latent: batch size must be in [1, 32]");
  tr::BgmIterArgs A;
  bgm_iter_fill(t, A, zt_dev, x_dev, idx_dev, bs);
  A.losses = loss_dev;
  A.gz_out = gz_out_dev;
  // posterior_optimizer (:88): one optimizer (shared step count), a fresh variable per batch
  t->step_z += 1;
  const double k = (double)t->step_z;
  A.lr_t = (float)(t->lr_z * std::sqrt(1.0 - std::pow(0.99, k)) / (1.0 - std::pow(0.9, k)));
  A.b1 = 0.9f; A.b2 = 0.99f; A.eps = 1e-7f;
  BGM_CUDA_OK(cudaFuncSetAttribute(tr::bgm_iter_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, t->smem_iter));
  tr::bgm_iter_kernel<1><<<1, tr::NTH, t->smem_iter, (cudaStream_t)stream>>>(A);
  BGM_CUDA_OK(cudaGetLastError());
  return 0;
}

int bgm_bgm_evaluate(bgm_trainer* t, const float* zt_dev, const float* x_dev, int n, double* sum_dev, void* stream) {
  using namespace bgm;
  if (!t || t->kind != 1) return fail(BGM_ERR_ARG, "bgm_bgm_evaluate: needs a BGM trainer");
  if (!zt_dev || !x_dev || !sum_dev || n < 1) return fail(BGM_ERR_ARG, "bgm_bgm_evaluate: bad argument");
  if (!t->smem_eval) t->smem_eval = (2 * t->wm * tr::LD + t->zd * tr::LD) * 4 + 64;
  tr::BgmEvalArgs A;
  memset(&A, 0, sizeof(A));
  A.g = t->vg; A.zd = t->zd; A.xd = t->xd; A.n = n; A.theta = t->theta[0]; A.moving = t->moving;
  A.zt = zt_dev; A.x = x_dev; A.sum_out = sum_dev; A.wm = t->wm;
  BGM_CUDA_OK(cudaMemsetAsync(sum_dev, 0, sizeof(double), (cudaStream_t)stream));
  BGM_CUDA_OK(cudaFuncSetAttribute(tr::bgm_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, t->smem_eval));
  const int grid = std::max(1, std::min((n + 31) / 32, t->sm_count * 2));
  tr::bgm_eval_kernel<<<grid, tr::NTH, t->smem_eval, (cudaStream_t)stream>>>(A);
  BGM_CUDA_OK(cudaGetLastError());
  return 0;
}

int bgm_causal_evaluate(bgm_trainer* t, const float* zt_dev, const float* x_dev, const float* y_dev,
                        const float* v_dev, int n, double* sums_dev, float* z_out_dev, void* stream) {
  using namespace bgm;
  if (!t || t->kind != 0) return fail(BGM_ERR_ARG, "bgm_causal_evaluate: needs a CausalBGM trainer");
  if (!x_dev || !y_dev || !v_dev || !sums_dev || n < 1) return fail(BGM_ERR_ARG, "bgm_causal_evaluate: bad argument");
  if (!t->smem_eval) t->smem_eval = (3 * t->wm * tr::LD + t->zd * tr::LD + 2 * tr::LD + (t->zd + 1) * tr::LD) * 4 + 64;
  tr::EvalArgs A;
  memset(&A, 0, sizeof(A));
  A.g = t->g; A.f = t->f; A.h = t->h; A.e = t->e;
  for (int i = 0; i < 4; ++i) A.z_dims[i] = t->z_dims[i];
  A.zd = t->zd; A.p = t->p; A.binary = t->binary; A.n = n;
  A.theta = t->theta[0]; A.zt = zt_dev; A.x = x_dev; A.y = y_dev; A.v = v_dev; A.sums = sums_dev; A.z_out = z_out_dev;
  A.wm = t->wm;
  BGM_CUDA_OK(cudaMemsetAsync(sums_dev, 0, 3 * sizeof(double), (cudaStream_t)stream));
  BGM_CUDA_OK(cudaFuncSetAttribute(tr::eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, t->smem_eval));
  const int grid = std::max(1, std::min((n + 31) / 32, t->sm_count * 2));
  tr::eval_kernel<<<grid, tr::NTH, t->smem_eval, (cudaStream_t)stream>>>(A);
  BGM_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
