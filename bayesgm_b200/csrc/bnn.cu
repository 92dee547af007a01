// Host side of the Bayesian-network entry points (include/bgm_b200.h, bgm_bnn_*): packs the
// DenseFlipout parameters, validates, launches the kernels of bnn.cuh.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "bnn2.cuh"

struct bgm_bnn {
  bgm::bnn::BnnProgram prog;
  float* image_dev = nullptr;
  float* dw_dev = nullptr;      // plan 2: kernel perturbations of the running iteration (one sampler run at a time per model)
  int smem_bytes = 0;
  int zmax = 0;
  int plan = 2;                 // 1: thread = row (bnn.cuh), 2: two threads per row, chunk program (bnn2.cuh)
  bgm::bnn::BnnProgram2 prog2;
  int sm_count = 148;
  long long macs = 0;
};

namespace bgm {
namespace bnn {

static int pack_net(const bgm_bnn_net_desc* d, const char* name, int want_in, int want_out, std::vector<float>& image,
                    BnnNet& net, long long& macs, std::string& err) {
  if (!d || !d->dims || !d->params || !d->bn) { err = std::string(name) + ": null descriptor"; return -1; }
  if (d->n_layers < 1 || d->n_layers > BNN_MAXL) { err = std::string(name) + ": 1..8 layers supported"; return -2; }
  if (d->dims[0] != want_in || d->dims[d->n_layers] != want_out) { err = std::string(name) + ": input / output width mismatch"; return -1; }
  memset(&net, 0, sizeof(net));
  net.L = d->n_layers;
  net.kin = d->dims[0];
  net.bn_off = (int)image.size();
  image.insert(image.end(), d->bn, d->bn + 2 * net.kin);
  while (image.size() % 4) image.push_back(0.f);
  const float* src = d->params;
  const float feps = 1.1920928955078125e-07f;   // np.finfo(float32).eps (tfp default_mean_field_normal_fn)
  for (int l = 0; l < net.L; ++l) {
    const int K = d->dims[l], N = d->dims[l + 1];
    if (K < 1 || K > BNN_MAXK || N < 1) { err = std::string(name) + ": layer input widths up to 64 are supported"; return -2; }
    if (l < net.L - 1 && N > BNN_MAXK) { err = std::string(name) + ": hidden widths up to 64 are supported"; return -2; }
    BnnLayer& Ly = net.layer[l];
    Ly.K = K; Ly.N = N; Ly.N32 = (N + 31) / 32 * 32;
    Ly.loc_off = (int)image.size();
    image.resize(image.size() + (size_t)K * Ly.N32, 0.f);
    Ly.scale_off = (int)image.size();
    image.resize(image.size() + (size_t)K * Ly.N32, 0.f);
    Ly.bias_off = (int)image.size();
    image.resize(image.size() + Ly.N32, 0.f);
    const float* loc = src;
    const float* rho = src + (size_t)K * N;
    const float* bias = src + (size_t)2 * K * N;
    for (int k = 0; k < K; ++k)
      for (int c = 0; c < N; ++c) {
        image[Ly.loc_off + (size_t)k * Ly.N32 + c] = loc[(size_t)k * N + c];
        const float r = rho[(size_t)k * N + c];
        const float sp = std::max(r, 0.f) + log1pf(expf(-fabsf(r)));
        image[Ly.scale_off + (size_t)k * Ly.N32 + c] = feps + sp;
      }
    for (int c = 0; c < N; ++c) image[Ly.bias_off + c] = bias[c];
    src += (size_t)2 * K * N + N;
    macs += (long long)2 * K * N;
  }
  return 0;
}

static int zmax_of(int zd) { return zd <= 8 ? 8 : (zd <= 16 ? 16 : 32); }

// Rows (= threads) per CTA.  One CTA per SM (shared memory): a slice smaller than SMs x 256 rows is cut into one
// wave of equal CTAs (n = 20000 on 148 SMs: 125 CTAs of 160 rows instead of 79 of 256), larger slices use 256.
static int rows_per_cta2(const bgm_bnn* m, int n);
static int rows_per_cta(const bgm_bnn* m, int n) {
  if (m->plan == 2) return rows_per_cta2(m, n);
  const int per_sm = (n + m->sm_count - 1) / m->sm_count;
  const int warps = std::min(8, std::max(2, (per_sm + 31) / 32));
  return warps * 32;
}
template <int ZMAX>
static int launch_mh(const bgm_bnn* m, const BnnMhDev& D, int nth, int ncta, cudaStream_t st) {
  BGM_CUDA_OK(cudaFuncSetAttribute(bnn_mh_kernel<ZMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, m->smem_bytes));
  bnn_mh_kernel<ZMAX><<<ncta, nth, m->smem_bytes, st>>>(m->prog, m->image_dev, D);
  BGM_CUDA_OK(cudaGetLastError());
  return 0;
}
// ---- plan 2 ----
static int add_chunks(BnnProgram2& Q, const BnnNet& net, int net_id) {
  for (int l = 0; l < net.L; ++l) {
    const BnnLayer& Ly = net.layer[l];
    const int nchunk = Ly.N32 / 32;
    for (int c = 0; c < nchunk; ++c) {
      if (Q.nchunks >= B2_MAXCH) return -1;
      BnnChunk& C = Q.ch[Q.nchunks++];
      C.loc_off = Ly.loc_off; C.scale_off = Ly.scale_off; C.bias_off = Ly.bias_off;
      C.K = (short)Ly.K; C.N = (short)Ly.N; C.N32 = (short)Ly.N32; C.N4 = (short)((Ly.N + 3) & ~3);
      C.c = (unsigned char)c; C.net = (unsigned char)net_id; C.layer = (unsigned char)l;
      C.flags = 0;
      if (c == 0) C.flags |= CH_LAYER_FIRST;
      if (c == nchunk - 1) C.flags |= CH_LAYER_LAST;
      if (l == net.L - 1) C.flags |= CH_FINAL;
      if (Ly.N <= 8) C.flags |= CH_NARROW;
      if (l == 0 && c == 0) C.flags |= CH_NET_FIRST;
    }
  }
  return 0;
}
static int add_segments(BnnProgram2& Q, const BnnNet& net, int net_id) {
  for (int l = 0; l < net.L; ++l) {
    if (Q.nseg >= 24) return -1;
    const BnnLayer& Ly = net.layer[l];
    BnnSeg& S = Q.seg[Q.nseg++];
    S.scale_off = Ly.scale_off; S.K = Ly.K; S.N32 = Ly.N32; S.N4 = (Ly.N + 3) & ~3; S.net = net_id; S.layer = l;
    S.g0 = Q.ngroups;
    Q.ngroups += Ly.K * Ly.N32 / 4;
  }
  return 0;
}
static int rows_per_cta2(const bgm_bnn* m, int n) {      // two CTAs per SM, 16 rows per warp
  const int per_slot = (n + 2 * m->sm_count - 1) / (2 * m->sm_count);
  const int warps = std::min(8, std::max(2, (per_slot + 15) / 16));
  return warps * 16;
}
static int smem2(const bgm_bnn* m, int nrows) {
  const int NP = 4 * m->zmax + 2;
  return (2 * BNN_MAXK * nrows + 4 * W_FLOATS + 8 * NP + NP + 8) * 4;
}
template <int ZH>
static int launch_mh2(const bgm_bnn* m, const BnnMhDev& D, int nrows, int ncta, cudaStream_t st) {
  const int smem = smem2(m, nrows);
  BGM_CUDA_OK(cudaFuncSetAttribute(bnn_mh2_kernel<ZH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2(m, 128)));
  bnn_mh2_kernel<ZH><<<ncta, 2 * nrows, smem, st>>>(m->prog2, m->image_dev, D);
  BGM_CUDA_OK(cudaGetLastError());
  return 0;
}
static int launch(const bgm_bnn* m, const BnnMhDev& D, int nth, int ncta, cudaStream_t st) {
  if (m->plan == 2) {
    const int nrows = nth;      // the callers pass rows per CTA
    switch (m->zmax) {
      case 8: return launch_mh2<4>(m, D, nrows, ncta, st);
      case 16: return launch_mh2<8>(m, D, nrows, ncta, st);
      default: return launch_mh2<16>(m, D, nrows, ncta, st);
    }
  }
  switch (m->zmax) {
    case 8: return launch_mh<8>(m, D, nth, ncta, st);
    case 16: return launch_mh<16>(m, D, nth, ncta, st);
    default: return launch_mh<32>(m, D, nth, ncta, st);
  }
}

}  // namespace bnn
}  // namespace bgm

extern "C" {

int bgm_bnn_create(bgm_bnn** out, const int z_dims[4], int v_dim, int binary_treatment, float sigma_v, float sigma_x,
                   float sigma_y, const bgm_bnn_net_desc* g_net, const bgm_bnn_net_desc* f_net,
                   const bgm_bnn_net_desc* h_net) {
  using namespace bgm;
  using namespace bgm::bnn;
  if (!out || !z_dims || !g_net || !f_net || !h_net) return fail(BGM_ERR_ARG, "bgm_bnn_create: null pointer");
  *out = nullptr;
  const int d0 = z_dims[0], d1 = z_dims[1], d2 = z_dims[2], d3 = z_dims[3];
  const int zd = d0 + d1 + d2 + d3;
  if (d0 < 0 || d1 < 0 || d2 < 0 || d3 < 0 || zd < 1 || zd > 32)
    return fail(BGM_ERR_UNSUPPORTED, "bgm_bnn_create: 1 <= sum(z_dims) <= 32");
  if (d0 + d1 > 31) return fail(BGM_ERR_UNSUPPORTED, "bgm_bnn_create: z0 + z1 <= 31");
  if (v_dim < 1) return fail(BGM_ERR_ARG, "bgm_bnn_create: v_dim < 1");
  bgm_bnn* m = new bgm_bnn();
  BnnProgram& P = m->prog;
  memset(&P, 0, sizeof(P));
  P.zd = zd; P.d0 = d0; P.d1 = d1; P.d2 = d2; P.p = v_dim; P.binary = binary_treatment ? 1 : 0;
  P.s2v = sigma_v >= 0.f ? sigma_v * sigma_v : -1.f;
  P.s2x = sigma_x >= 0.f ? sigma_x * sigma_x : -1.f;
  P.s2y = sigma_y >= 0.f ? sigma_y * sigma_y : -1.f;
  std::vector<float> image;
  std::string err;
  int rc = pack_net(g_net, "g_net", zd, v_dim + 1, image, P.g, m->macs, err);
  if (!rc) rc = pack_net(f_net, "f_net", d0 + d1 + 1, 2, image, P.f, m->macs, err);
  if (!rc) rc = pack_net(h_net, "h_net", d0 + d2, 2, image, P.h, m->macs, err);
  if (rc) {
    delete m;
    return fail(rc == -2 ? BGM_ERR_UNSUPPORTED : BGM_ERR_ARG, "bgm_bnn_create: " + err);
  }
  m->zmax = zmax_of(zd);
  m->prog2.P = P;
  m->prog2.nchunks = 0;
  m->prog2.nseg = 0; m->prog2.ngroups = 0; m->prog2.image_floats = (int)image.size();
  if (add_chunks(m->prog2, P.g, NET_G) || add_chunks(m->prog2, P.h, NET_H) || add_chunks(m->prog2, P.f, NET_F) ||
      add_segments(m->prog2, P.g, NET_G) || add_segments(m->prog2, P.h, NET_H) || add_segments(m->prog2, P.f, NET_F)) {
    m->plan = 1;
    m->prog2.nchunks = 0;
  }
  {
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&m->sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (m->sm_count < 1) m->sm_count = 148;
  }
  const int NP = 4 * m->zmax + 2;
  m->smem_bytes = (2 * ACT_FLOATS + 2 * W_FLOATS + 8 * NP + NP + 8) * 4;
  cudaError_t e = cudaMalloc(&m->image_dev, image.size() * sizeof(float));
  if (e == cudaSuccess) e = cudaMemcpy(m->image_dev, image.data(), image.size() * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMalloc(&m->dw_dev, 4 * image.size() * sizeof(float));
  if (e != cudaSuccess) {
    if (m->image_dev) cudaFree(m->image_dev);
    if (m->dw_dev) cudaFree(m->dw_dev);
    delete m;
    return fail(BGM_ERR_CUDA, std::string("bgm_bnn_create: ") + cudaGetErrorString(e));
  }
  *out = m;
  return 0;
}

void bgm_bnn_destroy(bgm_bnn* m) {
  if (!m) return;
  if (m->image_dev) cudaFree(m->image_dev);
  if (m->dw_dev) cudaFree(m->dw_dev);
  delete m;
}

int bgm_bnn_info(const bgm_bnn* m, int* smem_bytes, int* rows_per_cta, long long* macs_per_eval) {
  if (!m) return bgm::fail(BGM_ERR_ARG, "bgm_bnn_info: null model");
  if (smem_bytes) *smem_bytes = m->plan == 2 ? bgm::bnn::smem2(m, 128) : m->smem_bytes;
  if (rows_per_cta) *rows_per_cta = m->plan == 2 ? 128 : bgm::bnn::BNN_THREADS;
  if (macs_per_eval) *macs_per_eval = m->macs;
  return 0;
}

int bgm_bnn_set_plan(bgm_bnn* m, int plan) {
  if (!m || (plan != 1 && plan != 2)) return bgm::fail(BGM_ERR_ARG, "bgm_bnn_set_plan: plan must be 1 or 2");
  if (plan == 2 && m->prog2.nchunks == 0) return bgm::fail(BGM_ERR_UNSUPPORTED, "bgm_bnn_set_plan: no chunk program for this model");
  m->plan = plan;
  return 0;
}

long long bgm_bnn_scratch_doubles(const bgm_bnn* m, int n) {
  if (!m || n < 1) return -1;
  const long long ncta = (n + 31) / 32;     // upper bound: the smallest CTA holds 32 rows
  return 2 * ncta * (4 * m->zmax + 2);
}

static int bnn_check(const char* fn, const bgm_bnn* m, const float* x, const float* y, const float* v, int ldv, int n) {
  using namespace bgm;
  if (!m) return fail(BGM_ERR_ARG, std::string(fn) + ": null model");
  if (!x || !y || !v) return fail(BGM_ERR_ARG, std::string(fn) + ": null data pointer");
  if (n < 1) return fail(BGM_ERR_ARG, std::string(fn) + ": n < 1");
  if (ldv < m->prog.p || ldv % 4 != 0 || reinterpret_cast<uintptr_t>(v) % 16 != 0)
    return fail(BGM_ERR_ARG, std::string(fn) + ": v_dev must be 16-byte aligned with ldv >= v_dim, ldv % 4 == 0");
  return 0;
}

int bgm_bnn_logpost(const bgm_bnn* m, const float* x_dev, const float* y_dev, const float* v_dev, int ldv,
                    const float* z_dev, int n, uint64_t seed, int slice, int64_t row_offset, uint32_t call,
                    double* scratch_dev, float* out_logp_dev, void* stream) {
  using namespace bgm;
  using namespace bgm::bnn;
  int rc = bnn_check("bgm_bnn_logpost", m, x_dev, y_dev, v_dev, ldv, n);
  if (rc) return rc;
  if (!z_dev || !out_logp_dev || !scratch_dev) return fail(BGM_ERR_ARG, "bgm_bnn_logpost: null pointer");
  BnnMhDev D;
  memset(&D, 0, sizeof(D));
  D.a.x_dev = x_dev; D.a.y_dev = y_dev; D.a.v_dev = v_dev; D.a.ldv = ldv; D.a.n = n;
  D.a.seed = seed; D.a.row_offset = row_offset; D.a.init_mode = 1;
  D.slice = slice; D.call0 = call; D.z_in = z_dev; D.out_lp = out_logp_dev; D.part = scratch_dev; D.dw = m->dw_dev;
  D.t = 0;
  const int nth = rows_per_cta(m, n);
  const int ncta = (n + nth - 1) / nth;
  D.mode = 2;
  rc = launch(m, D, nth, ncta, (cudaStream_t)stream);
  if (rc) return rc;
  D.mode = 1;
  return launch(m, D, nth, ncta, (cudaStream_t)stream);
}

int bgm_bnn_mh(const bgm_bnn* m, const bgm_mh_args* a, int slice, double* scratch_dev, float* lp_cur_trace_dev,
               void* stream) {
  using namespace bgm;
  using namespace bgm::bnn;
  if (!a) return fail(BGM_ERR_ARG, "bgm_bnn_mh: null args");
  int rc = bnn_check("bgm_bnn_mh", m, a->x_dev, a->y_dev, a->v_dev, a->ldv, a->n);
  if (rc) return rc;
  if (!a->z_state_dev || !scratch_dev) return fail(BGM_ERR_ARG, "bgm_bnn_mh: z_state_dev / scratch_dev required");
  if (a->t_begin < 0 || a->t_end < a->t_begin) return fail(BGM_ERR_ARG, "bgm_bnn_mh: bad iteration range");
  if ((a->eps_dev == nullptr) != (a->u_dev == nullptr))
    return fail(BGM_ERR_ARG, "bgm_bnn_mh: eps_dev and u_dev must be given together");
  if (a->t_end == a->t_begin && a->init_mode != 2) return 0;
  BnnMhDev D;
  memset(&D, 0, sizeof(D));
  D.a = *a;
  D.slice = slice;
  D.part = scratch_dev;
  D.dw = m->dw_dev;
  D.lp_cur_trace = lp_cur_trace_dev;
  const int nth = rows_per_cta(m, a->n);
  const int ncta = (a->n + nth - 1) / nth;
  cudaStream_t st = (cudaStream_t)stream;
  // statistics of the first iteration (and the initial draw for init_mode 2), then one launch per iteration
  D.mode = 2;
  D.t = a->t_begin;
  rc = launch(m, D, nth, ncta, st);
  if (rc) return rc;
  D.mode = 0;
  D.a.init_mode = 0;
  for (int t = a->t_begin; t < a->t_end; ++t) {
    D.t = t;
    rc = launch(m, D, nth, ncta, st);
    if (rc) return rc;
  }
  return 0;
}

int bgm_bnn_effect(const bgm_bnn* m, const float* z_samples_dev, int n_keep, int n, const float* x_values_dev, int n_x,
                   int sample_y, uint64_t seed, int64_t row_offset, const float* noise_dev, float* stats_scratch_dev,
                   double* adrf_sum_dev, float* ite_dev, void* stream) {
  using namespace bgm;
  using namespace bgm::bnn;
  if (!m || !z_samples_dev || !stats_scratch_dev) return fail(BGM_ERR_ARG, "bgm_bnn_effect: null model / pointer");
  if (n_keep < 1 || n < 1 || n_keep > 65535) return fail(BGM_ERR_ARG, "bgm_bnn_effect: 1 <= n_keep <= 65535, n >= 1");
  const BnnProgram& P = m->prog;
  if (P.binary) {
    if (!ite_dev) return fail(BGM_ERR_ARG, "bgm_bnn_effect: ite_dev required for a binary treatment");
    n_x = 2;
  } else {
    if (!adrf_sum_dev || n_x < 1) return fail(BGM_ERR_ARG, "bgm_bnn_effect: adrf_sum_dev and n_x >= 1 required");
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int nc = P.d0 + P.d1;
  bnn_sample_stats_kernel<<<n_keep, 256, 0, st>>>(z_samples_dev, n_keep, n, P.zd, nc, stats_scratch_dev);
  BGM_CUDA_OK(cudaGetLastError());
  BnnEffectDev E;
  memset(&E, 0, sizeof(E));
  E.zs = z_samples_dev; E.stats = stats_scratch_dev; E.x_values = x_values_dev; E.noise = noise_dev;
  E.adrf_sum = adrf_sum_dev; E.ite = ite_dev; E.n_keep = n_keep; E.n = n; E.n_x = n_x; E.sample_y = sample_y;
  E.seed = seed; E.row_offset = row_offset;
  const int smem = (2 * ACT_FLOATS + 2 * W_FLOATS + 16) * 4;
  BGM_CUDA_OK(cudaFuncSetAttribute(bnn_effect_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  dim3 grid((n + BNN_THREADS - 1) / BNN_THREADS, n_keep);
  bnn_effect_kernel<<<grid, BNN_THREADS, smem, st>>>(P, m->image_dev, E);
  BGM_CUDA_OK(cudaGetLastError());
  return 0;
}

int bgm_bnn_noise(const bgm_bnn* m, uint64_t seed, int slice, int net, int layer, uint32_t call, int64_t row_offset,
                  int rows, float* eps_dev, signed char* sign_in_dev, signed char* sign_out_dev, void* stream) {
  using namespace bgm;
  using namespace bgm::bnn;
  if (!m) return fail(BGM_ERR_ARG, "bgm_bnn_noise: null model");
  const BnnNet* nets[3] = {&m->prog.g, &m->prog.f, &m->prog.h};
  if (net < 0 || net > 2 || layer < 0 || layer >= nets[net]->L) return fail(BGM_ERR_ARG, "bgm_bnn_noise: bad net / layer");
  const BnnLayer& Ly = nets[net]->layer[layer];
  bnn_noise_kernel<<<64, 256, 0, (cudaStream_t)stream>>>(seed, slice, net, layer, call, Ly.K, Ly.N, row_offset, rows,
                                                        eps_dev, sign_in_dev, sign_out_dev);
  BGM_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
