// Host side of the BGM / HMC entry points (include/bgm_b200.h): cuts the generator's
// Keras arrays into the streamed tile program of hmc.cuh, validates, launches.
// Included by bgm_b200.cu.
#pragma once
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "hmc.cuh"
#include "hmc_tc.cuh"

struct bgm_hmc {
  bgm::HmcProgram prog;      // forward + backward (one gradient evaluation)
  bgm::HmcProgram prog_fwd;  // forward only (posterior-predictive draws)
  float* image_dev = nullptr;
  int sm_count = 0;
  int smem_max = 0;
  long long macs = 0, issued = 0;
  // tensor-core engine (hmc_tc.cuh): pre-split operand images of one evaluation + the resident small arrays
  bgm::HmcTcProgram tc;
  float* tc_stream_dev = nullptr;
  float* tc_stream_fwd_dev = nullptr;   // forward-only stream (posterior-predictive draws)
  float* tc_small_dev = nullptr;
  long long tc_issued = 0;
  int engine = 0;            // 0 auto (tensor when available), 1 SIMT, 2 tensor
};

namespace bgm {

constexpr int HMC_SMEM_RESERVE = 1024;  // static shared memory (barriers) + alignment slack

struct HmcPacker {
  std::vector<float> image;
  long long issued = 0;
  // appends a tile [kp][NT] | bias[NT]; `fill(k, c)` gives W, `bias(c)` the bias
  template <class FW, class FB>
  HmcOp add(int kp, int NT, unsigned char kind, int layer, int c0, int flags, FW fill, FB bias) {
    HmcOp op;
    memset(&op, 0, sizeof(op));
    op.g_off = (int)image.size();
    image.resize(image.size() + (size_t)kp * NT + NT, 0.f);
    for (int k = 0; k < kp; ++k)
      for (int c = 0; c < NT; ++c) image[op.g_off + (size_t)k * NT + c] = fill(k, c);
    for (int c = 0; c < NT; ++c) image[op.g_off + (size_t)kp * NT + c] = bias(c);
    op.bytes = (kp * NT + NT) * 4;
    op.kp = (short)kp;
    op.kind = kind;
    op.layer = (unsigned char)layer;
    op.c0 = (short)c0;
    op.flags = (short)flags;
    issued += (long long)kp * NT;
    return op;
  }
};

static int hmc_pick_ncons(const bgm_hmc* m, int n_rows) {
  const HmcProgram& P = m->prog;
  const int per_warp = (P.kin * TILE_ROWS + 2 * HMC_BUF + 3 * P.zd * TILE_ROWS) * 4;
  const int ring = HMC_STAGES * HMC_STAGE_FLOATS * 4;
  int fit = (m->smem_max - HMC_SMEM_RESERVE - ring) / per_warp;
  int ncons = fit >= 8 ? 8 : (fit >= 4 ? 4 : (fit >= 2 ? 2 : (fit >= 1 ? 1 : 0)));
  // small problems: spread the 32-row tiles over more CTAs instead of more warps per CTA
  const int ntiles = (n_rows + TILE_ROWS - 1) / TILE_ROWS;
  while (ncons > 1 && (ntiles + ncons - 1) / ncons < m->sm_count) ncons >>= 1;
  return ncons;
}

static int hmc_launch(const bgm_hmc* m, const HmcProgram& P, HmcDev& D, int n_rows, cudaStream_t st) {
  const int ncons = hmc_pick_ncons(m, n_rows);
  if (ncons < 1) return fail(BGM_ERR_NOMEM, "bgm_hmc: per-warp buffers do not fit in shared memory");
  D.ncons = ncons;
  const int per_warp = P.kin * TILE_ROWS + 2 * HMC_BUF + 3 * P.zd * TILE_ROWS;
  const int smem = (HMC_STAGES * HMC_STAGE_FLOATS + ncons * per_warp) * 4;
  const int ntiles = (n_rows + TILE_ROWS - 1) / TILE_ROWS;
  const int nblocks = (ntiles + ncons - 1) / ncons;
  const int grid = std::max(1, std::min(nblocks, m->sm_count));
  BGM_CUDA_OK(cudaFuncSetAttribute(hmc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, m->smem_max - HMC_SMEM_RESERVE));
  hmc_kernel<<<grid, ncons * 32, smem, st>>>(P, m->image_dev, D);
  BGM_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace bgm

namespace bgm {

// host twin of umma::split_tf32
static inline void ht_split_host(float v, float& hi, float& lo) {
  uint32_t b;
  memcpy(&b, &v, 4);
  b = (b + 0x1000u) & 0xffffe000u;
  memcpy(&hi, &b, 4);
  lo = v - hi;
}
// appends the [hi | lo] UMMA image of B (K = 64 x N = 64, element (k, n) = fill(k, n)) in layout [K/4][N][4]
template <class F>
static void ht_add_image(std::vector<float>& stream, F fill) {
  const size_t base = stream.size();
  stream.resize(base + HT_IMG_FLOATS, 0.f);
  for (int k = 0; k < 64; ++k)
    for (int n = 0; n < 64; ++n) {
      float hi, lo;
      ht_split_host(fill(k, n), hi, lo);
      const size_t idx = (size_t)(k / 4) * (64 * 4) + (size_t)n * 4 + (k % 4);
      stream[base + idx] = hi;
      stream[base + 64 * 64 + idx] = lo;
    }
}
static bool hmc_use_tc(const bgm_hmc* m) { return m->tc.enabled && m->engine != 1; }

static int hmc_tc_zmax(int zd) { return zd <= 4 ? 4 : (zd <= 8 ? 8 : (zd <= 12 ? 12 : 16)); }
static int hmc_tc_smem(const bgm_hmc* m) {
  return (HT_SLOTS * HT_IMG_FLOATS + m->tc.small_floats + HT_TPR * (hmc_tc_zmax(m->tc.zd) + 1) * HT_ROWS) * 4;
}
static int hmc_tc_launch(const bgm_hmc* m, HmcDev& D, int n_rows, cudaStream_t st) {
  const int smem = hmc_tc_smem(m);
  const int nblocks = (n_rows + HT_ROWS - 1) / HT_ROWS;
  const int grid = std::max(1, std::min(nblocks, m->sm_count));
  const int zmax = hmc_tc_zmax(m->tc.zd);
#define BGM_HT_LAUNCH(Z)                                                                                              \
  do {                                                                                                                \
    BGM_CUDA_OK(cudaFuncSetAttribute(hmc_tc_kernel<Z>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));            \
    hmc_tc_kernel<Z><<<grid, HT_THREADS + 32, smem, st>>>(m->tc, m->tc_stream_dev, m->tc_stream_fwd_dev, m->tc_small_dev, D);                     \
  } while (0)
  if (zmax == 4) BGM_HT_LAUNCH(4);
  else if (zmax == 8) BGM_HT_LAUNCH(8);
  else if (zmax == 12) BGM_HT_LAUNCH(12);
  else BGM_HT_LAUNCH(16);
#undef BGM_HT_LAUNCH
  BGM_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace bgm

namespace bgm {
struct ColQ {
  const float* draws;
  int S;
  long long M;
  float *mean, *lo, *hi;
  int lo_idx, hi_idx;
  float lo_frac, hi_frac;
};
template <int K>
__global__ void __launch_bounds__(128) column_quantiles_kernel(const ColQ Q) {
  for (long long m = blockIdx.x * 128ll + threadIdx.x; m < Q.M; m += (long long)gridDim.x * 128ll) {
    float small[K], large[K];     // ascending: small[0] is the minimum; large[0] is the maximum (descending)
#pragma unroll
    for (int k = 0; k < K; ++k) { small[k] = __int_as_float(0x7f800000); large[k] = __int_as_float(0xff800000); }
    float sum = 0.f;
    for (int s = 0; s < Q.S; ++s) {
      float v = __ldg(Q.draws + (size_t)s * Q.M + m);
      sum += v;
      if (Q.lo) {
        if (v < small[K - 1]) {
          float c = v;
#pragma unroll
          for (int k = 0; k < K; ++k) { const float a = small[k]; small[k] = fminf(a, c); c = fmaxf(a, c); }
        }
        if (v > large[K - 1]) {
          float c = v;
#pragma unroll
          for (int k = 0; k < K; ++k) { const float a = large[k]; large[k] = fmaxf(a, c); c = fminf(a, c); }
        }
      }
    }
    Q.mean[m] = sum / (float)Q.S;
    if (Q.lo) {
      auto ord = [&](int idx) -> float {          // idx-th order statistic (0 = minimum), taken from the nearer end
        float r = 0.f;
        if (idx < K && idx < Q.S) {
#pragma unroll
          for (int k = 0; k < K; ++k) if (k == idx) r = small[k];
        } else {
          const int t = Q.S - 1 - idx;
#pragma unroll
          for (int k = 0; k < K; ++k) if (k == t) r = large[k];
        }
        return r;
      };
      const float a = ord(Q.lo_idx), b = ord(min(Q.lo_idx + 1, Q.S - 1));
      Q.lo[m] = a + (b - a) * Q.lo_frac;
      const float c = ord(Q.hi_idx), d = ord(min(Q.hi_idx + 1, Q.S - 1));
      Q.hi[m] = c + (d - c) * Q.hi_frac;
    }
  }
}
}  // namespace bgm

extern "C" {

int bgm_hmc_create(bgm_hmc** out, const bgm_varnet_desc* g) {
  using namespace bgm;
  if (!out || !g) return fail(BGM_ERR_ARG, "bgm_hmc_create: null argument");
  *out = nullptr;
  if (!g->units || !g->bn || !g->hidden_params || !g->mean_params || !g->var_params)
    return fail(BGM_ERR_ARG, "bgm_hmc_create: null array in the net description");
  const int zd = g->z_dim, xd = g->x_dim, nh = g->n_hidden;
  if (zd < 1 || zd > HMC_MAXZ) return fail(BGM_ERR_UNSUPPORTED, "bgm_hmc_create: z_dim must be in [1, 16]");
  if (xd < 1) return fail(BGM_ERR_ARG, "bgm_hmc_create: x_dim must be >= 1");
  if (nh < 1 || nh > HMC_MAXL) return fail(BGM_ERR_UNSUPPORTED, "bgm_hmc_create: 1..6 hidden layers are supported");
  for (int l = 0; l < nh; ++l)
    if (g->units[l] < 1 || g->units[l] > 64)
      return fail(BGM_ERR_UNSUPPORTED, "bgm_hmc_create: hidden widths must be in [1, 64]");
  const int ntile = (xd + 31) / 32, ndz = (zd + 7) / 8;
  if (nh + 2 * ntile + (nh - 1) + ndz > HMC_MAX_OPS)
    return fail(BGM_ERR_UNSUPPORTED, "bgm_hmc_create: x_dim too large for the tile program");

  // unpack the Keras arrays
  std::vector<int> dims(nh + 1);
  dims[0] = zd;
  for (int l = 0; l < nh; ++l) dims[l + 1] = g->units[l];
  std::vector<const float*> Wl(nh), bl(nh);
  const float* p = g->hidden_params;
  long long macs = 0;
  for (int l = 0; l < nh; ++l) {
    Wl[l] = p;
    p += (size_t)dims[l] * dims[l + 1];
    bl[l] = p;
    p += dims[l + 1];
    macs += (long long)dims[l] * dims[l + 1];
  }
  const int last = dims[nh];
  const float *Wm = g->mean_params, *bm = g->mean_params + (size_t)last * xd;
  const float *Wv = g->var_params, *bv = g->var_params + (size_t)last * xd;
  macs += 2LL * last * xd;
  macs *= 2;  // forward + gradient w.r.t. the inputs

  bgm_hmc* m = new bgm_hmc();
  HmcProgram& P = m->prog;
  memset(&P, 0, sizeof(P));
  P.zd = zd;
  P.kin = (zd + 3) / 4 * 4;
  P.x_dim = xd;
  P.nh = nh;
  for (int d = 0; d < zd; ++d) {   // Keras BN inference: gamma*(z-mean)/sqrt(var+1e-3)+beta
    const float gamma = g->bn[d], beta = g->bn[zd + d], mean = g->bn[2 * zd + d], var = g->bn[3 * zd + d];
    P.bn_mean[d] = mean;
    P.bn_inv[d] = gamma / sqrtf(var + 1e-3f);
    P.bn_beta[d] = beta;
  }
  HmcPacker pk;
  std::vector<HmcOp> ops, fops;
  auto zero = [](int) { return 0.f; };
  auto r4 = [](int a) { return (a + 3) / 4 * 4; };
  for (int l = 0; l < nh; ++l) {  // forward hidden layers
    const int K = dims[l], N = dims[l + 1];
    const int kp = l == 0 ? P.kin : r4(K);
    ops.push_back(pk.add(kp, 64, HK_FWD, l, 0, 0,
                         [&](int k, int c) { return (k < K && c < N) ? Wl[l][(size_t)k * N + c] : 0.f; },
                         [&](int c) { return c < N ? bl[l][c] : 0.f; }));
    fops.push_back(ops.back());
  }
  for (int t = 0; t < ntile; ++t) {  // fused head tiles: 32 data columns, [mean | var]
    const int c0 = t * 32;
    auto col = [&](int c) { return c0 + (c & 31); };
    ops.push_back(pk.add(r4(last), 64, HK_HEADF, 0, c0, 0,
                         [&](int k, int c) {
                           if (k >= last || col(c) >= xd) return 0.f;
                           return (c < 32 ? Wm : Wv)[(size_t)k * xd + col(c)];
                         },
                         [&](int c) { return col(c) < xd ? (c < 32 ? bm : bv)[col(c)] : 0.f; }));
    fops.push_back(ops.back());
    const int flags = (t == 0 ? 1 : 0) | (t == ntile - 1 ? 2 : 0);
    ops.push_back(pk.add(64, 64, HK_HEADB, nh - 1, c0, flags,
                         [&](int c, int k) {   // row = tile column c, output = hidden unit k
                           if (k >= last || col(c) >= xd) return 0.f;
                           return (c < 32 ? Wm : Wv)[(size_t)k * xd + col(c)];
                         },
                         zero));
  }
  for (int l = nh - 1; l >= 1; --l) {  // backward hidden layers: d/d(output of layer l-1)
    const int K = dims[l], N = dims[l + 1];
    ops.push_back(pk.add(r4(N), 64, HK_BWD, l - 1, 0, 0,
                         [&](int j, int i) { return (j < N && i < K) ? Wl[l][(size_t)i * N + j] : 0.f; }, zero));
  }
  for (int t = 0; t < ndz; ++t) {  // backward of layer 0: d/d BN(z)
    const int N = dims[1];
    ops.push_back(pk.add(r4(N), 8, HK_DZ, 0, t * 8, 0,
                         [&](int j, int c) { return (j < N && t * 8 + c < zd) ? Wl[0][(size_t)(t * 8 + c) * N + j] : 0.f; },
                         zero));
  }
  P.n_ops = (int)ops.size();
  for (int i = 0; i < P.n_ops; ++i) P.ops[i] = ops[i];
  m->prog_fwd = P;
  m->prog_fwd.n_ops = (int)fops.size();
  for (size_t i = 0; i < fops.size(); ++i) m->prog_fwd.ops[i] = fops[i];
  m->macs = macs;
  m->issued = pk.issued;

  // ---- tensor-core engine: every hidden layer 64 wide, at least two of them ----
  std::vector<float> tstream, tstream_fwd, tsmall;
  {
    HmcTcProgram& T = m->tc;
    memset(&T, 0, sizeof(T));
    bool ok = nh >= 2;
    for (int l = 0; l < nh; ++l) ok = ok && g->units[l] == 64;
    if (ok) {
      T.enabled = 1;
      T.zd = zd; T.kin = P.kin; T.x_dim = xd; T.nh = nh;
      T.n_chunks = (xd + 31) / 32;
      T.n_img = 2 * (nh - 1) + 2 * T.n_chunks;
      for (int d = 0; d < zd; ++d) { T.bn_mean[d] = P.bn_mean[d]; T.bn_inv[d] = P.bn_inv[d]; T.bn_beta[d] = P.bn_beta[d]; }
      for (int l = 1; l < nh; ++l)                              // forward hidden layers 2..nh: B[k][n] = W_l[k][n]
        ht_add_image(tstream, [&](int k, int n) { return Wl[l][(size_t)k * 64 + n]; });
      // head images in the order the software pipeline of hmc_tc.cuh consumes them:
      // HF_0; then per chunk k: HB_{k-1} (k > 0), HF_{k+1} (k + 1 < NC); finally HB_{NC-1}
      auto add_hf = [&](int c) {   // head forward: B[k][n] = (n < 32 ? Wm : Wv)[k][feature]
        auto feat = [&](int q) { return c * 32 + (q & 31); };
        ht_add_image(tstream, [&](int k, int n) { return feat(n) < xd ? (n < 32 ? Wm : Wv)[(size_t)k * xd + feat(n)] : 0.f; });
      };
      auto add_hb = [&](int c) {   // head backward: B[k][n] = (k < 32 ? Wm : Wv)[n][feature of k]
        auto feat = [&](int q) { return c * 32 + (q & 31); };
        ht_add_image(tstream, [&](int k, int n) { return feat(k) < xd ? (k < 32 ? Wm : Wv)[(size_t)n * xd + feat(k)] : 0.f; });
      };
      // forward-only stream: hidden layers, then the head-forward images in chunk order
      for (int l = 1; l < nh; ++l)
        ht_add_image(tstream_fwd, [&](int k, int n) { return Wl[l][(size_t)k * 64 + n]; });
      for (int c = 0; c < T.n_chunks; ++c) {
        auto feat = [&](int q) { return c * 32 + (q & 31); };
        ht_add_image(tstream_fwd, [&](int k, int n) { return feat(n) < xd ? (n < 32 ? Wm : Wv)[(size_t)k * xd + feat(n)] : 0.f; });
      }
      T.n_img_fwd = (nh - 1) + T.n_chunks;
      add_hf(0);
      for (int c = 0; c < T.n_chunks; ++c) {
        if (c > 0) add_hb(c - 1);
        if (c + 1 < T.n_chunks) add_hf(c + 1);
      }
      add_hb(T.n_chunks - 1);
      for (int l = nh - 1; l >= 1; --l)                         // backward hidden layers: B[k][n] = W_l[n][k]
        ht_add_image(tstream, [&](int k, int n) { return Wl[l][(size_t)n * 64 + k]; });
      T.off_W1 = 0;
      tsmall.resize((size_t)zd * 64, 0.f);
      for (int d = 0; d < zd; ++d)
        for (int j = 0; j < 64; ++j) tsmall[(size_t)d * 64 + j] = Wl[0][(size_t)d * 64 + j];
      T.off_b1 = (int)tsmall.size();
      for (int j = 0; j < 64; ++j) tsmall.push_back(bl[0][j]);
      T.off_bh = (int)tsmall.size();
      for (int l = 1; l < nh; ++l)
        for (int j = 0; j < 64; ++j) tsmall.push_back(bl[l][j]);
      T.off_bm = (int)tsmall.size();
      for (int q = 0; q < 32 * T.n_chunks; ++q) tsmall.push_back(q < xd ? bm[q] : 0.f);
      T.off_bv = (int)tsmall.size();
      for (int q = 0; q < 32 * T.n_chunks; ++q) tsmall.push_back(q < xd ? bv[q] : 0.f);
      while (tsmall.size() % 4) tsmall.push_back(0.f);
      T.small_floats = (int)tsmall.size();
      m->tc_issued = 3LL * 64 * 64 * T.n_img + 2LL * zd * 64;   // 3 TF32 passes per 64 x 64 product + the first layer twice
    }
  }

  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&m->smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&m->sm_count, cudaDevAttrMultiProcessorCount, dev);
  if (e == cudaSuccess) e = cudaMalloc(&m->image_dev, pk.image.size() * sizeof(float));
  if (e == cudaSuccess)
    e = cudaMemcpy(m->image_dev, pk.image.data(), pk.image.size() * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess && m->tc.enabled) {
    if (hmc_tc_smem(m) > m->smem_max - HMC_SMEM_RESERVE) m->tc.enabled = 0;
  }
  if (e == cudaSuccess && m->tc.enabled) {
    e = cudaMalloc(&m->tc_stream_dev, tstream.size() * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&m->tc_stream_fwd_dev, tstream_fwd.size() * sizeof(float));
    if (e == cudaSuccess)
      e = cudaMemcpy(m->tc_stream_fwd_dev, tstream_fwd.data(), tstream_fwd.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&m->tc_small_dev, tsmall.size() * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(m->tc_stream_dev, tstream.data(), tstream.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(m->tc_small_dev, tsmall.data(), tsmall.size() * sizeof(float), cudaMemcpyHostToDevice);
  }
  {
    const char* eng = getenv("BGM_HMC_ENGINE");                 // "simt" / "tensor": override for experiments
    if (eng && eng[0] == 's') m->engine = 1;
  }
  if (e != cudaSuccess) {
    if (m->image_dev) cudaFree(m->image_dev);
    if (m->tc_stream_dev) cudaFree(m->tc_stream_dev);
    if (m->tc_stream_fwd_dev) cudaFree(m->tc_stream_fwd_dev);
    if (m->tc_small_dev) cudaFree(m->tc_small_dev);
    delete m;
    return fail(BGM_ERR_CUDA, std::string("bgm_hmc_create: ") + cudaGetErrorString(e));
  }
  if (hmc_pick_ncons(m, 1 << 30) < 1) {
    cudaFree(m->image_dev);
    delete m;
    return fail(BGM_ERR_NOMEM, "bgm_hmc_create: per-warp buffers do not fit in shared memory");
  }
  *out = m;
  return 0;
}

void bgm_hmc_destroy(bgm_hmc* m) {
  if (!m) return;
  if (m->image_dev) cudaFree(m->image_dev);
  if (m->tc_stream_dev) cudaFree(m->tc_stream_dev);
  if (m->tc_stream_fwd_dev) cudaFree(m->tc_stream_fwd_dev);
  if (m->tc_small_dev) cudaFree(m->tc_small_dev);
  delete m;
}

int bgm_hmc_set_engine(bgm_hmc* m, int kind) {
  using namespace bgm;
  if (!m) return fail(BGM_ERR_ARG, "bgm_hmc_set_engine: null model");
  if (kind < 0 || kind > 2) return fail(BGM_ERR_ARG, "bgm_hmc_set_engine: kind must be 0 (auto), 1 (SIMT) or 2 (tensor)");
  if (kind == 2 && !m->tc.enabled)
    return fail(BGM_ERR_UNSUPPORTED, "bgm_hmc_set_engine: the tensor engine needs >= 2 hidden layers, all 64 wide");
  m->engine = kind;
  return 0;
}
int bgm_hmc_engine_info(const bgm_hmc* m, int* active_kind, int* tensor_available, int* tensor_smem_bytes,
                        long long* tensor_issued_macs_per_grad) {
  using namespace bgm;
  if (!m) return fail(BGM_ERR_ARG, "bgm_hmc_engine_info: null model");
  if (active_kind) *active_kind = hmc_use_tc(m) ? 2 : 1;
  if (tensor_available) *tensor_available = m->tc.enabled;
  if (tensor_smem_bytes) *tensor_smem_bytes = m->tc.enabled ? hmc_tc_smem(m) : 0;
  if (tensor_issued_macs_per_grad) *tensor_issued_macs_per_grad = m->tc_issued;
  return 0;
}

int bgm_hmc_info(const bgm_hmc* m, int* smem_bytes, int* n_ops, long long* macs_per_grad,
                 long long* issued_macs_per_grad) {
  using namespace bgm;
  if (!m) return fail(BGM_ERR_ARG, "bgm_hmc_info: null model");
  const HmcProgram& P = m->prog;
  const int ncons = hmc_pick_ncons(m, 1 << 30);
  if (smem_bytes)
    *smem_bytes = (HMC_STAGES * HMC_STAGE_FLOATS + ncons * (P.kin * TILE_ROWS + 2 * HMC_BUF + 3 * P.zd * TILE_ROWS)) * 4;
  if (n_ops) *n_ops = P.n_ops;
  if (macs_per_grad) *macs_per_grad = m->macs;
  if (issued_macs_per_grad) *issued_macs_per_grad = m->issued;
  return 0;
}

static int hmc_check_x(const char* fn, const float* x, int ldx, int n, int x_dim) {
  using namespace bgm;
  if (!x) return fail(BGM_ERR_ARG, std::string(fn) + ": null x_dev");
  if (n < 1) return fail(BGM_ERR_ARG, std::string(fn) + ": n must be >= 1");
  if (ldx < x_dim || ldx % 4 != 0) return fail(BGM_ERR_ARG, std::string(fn) + ": ldx must be >= x_dim and a multiple of 4");
  if (reinterpret_cast<uintptr_t>(x) % 16 != 0) return fail(BGM_ERR_ARG, std::string(fn) + ": x_dev must be 16-byte aligned");
  return 0;
}

int bgm_hmc_logpost_grad(const bgm_hmc* m, const float* x_dev, int ldx, const float* z_dev, int n,
                         float* out_logp_dev, float* out_grad_dev, void* stream) {
  using namespace bgm;
  if (!m) return fail(BGM_ERR_ARG, "bgm_hmc_logpost_grad: null model");
  int rc = hmc_check_x("bgm_hmc_logpost_grad", x_dev, ldx, n, m->prog.x_dim);
  if (rc) return rc;
  if (!z_dev || !out_logp_dev) return fail(BGM_ERR_ARG, "bgm_hmc_logpost_grad: null z / out pointer");
  HmcDev D;
  memset(&D, 0, sizeof(D));
  D.a.x_dev = x_dev; D.a.ldx = ldx; D.a.n = n; D.a.lp_state_dev = out_logp_dev;
  D.mode = HMC_EVAL;
  D.z_in = z_dev;
  D.out_grad = out_grad_dev;
  if (hmc_use_tc(m)) return hmc_tc_launch(m, D, n, (cudaStream_t)stream);
  return hmc_launch(m, m->prog, D, n, (cudaStream_t)stream);
}

int bgm_hmc_run(const bgm_hmc* m, const bgm_hmc_args* a, void* stream) {
  using namespace bgm;
  if (!m || !a) return fail(BGM_ERR_ARG, "bgm_hmc_run: null model / args");
  int rc = hmc_check_x("bgm_hmc_run", a->x_dev, a->ldx, a->n, m->prog.x_dim);
  if (rc) return rc;
  if (!a->z_state_dev || !a->g_state_dev || !a->lp_state_dev)
    return fail(BGM_ERR_ARG, "bgm_hmc_run: z_state_dev, g_state_dev and lp_state_dev are required");
  if (a->init_mode < 0 || a->init_mode > 2) return fail(BGM_ERR_ARG, "bgm_hmc_run: init_mode must be 0, 1 or 2");
  if (a->t_begin < 0 || a->t_end < a->t_begin) return fail(BGM_ERR_ARG, "bgm_hmc_run: bad step range");
  if (a->num_leapfrog < 1) return fail(BGM_ERR_ARG, "bgm_hmc_run: num_leapfrog must be >= 1");
  if (!a->step_dev) return fail(BGM_ERR_ARG, "bgm_hmc_run: step_dev is required");
  if ((a->mom_dev == nullptr) != (a->logu_dev == nullptr))
    return fail(BGM_ERR_ARG, "bgm_hmc_run: mom_dev and logu_dev must be given together");
  if (a->init_mode == 2 && a->mom_dev)
    return fail(BGM_ERR_ARG, "bgm_hmc_run: init_mode 2 draws from Philox; pass z_state with injected noise");
  HmcDev D;
  memset(&D, 0, sizeof(D));
  D.a = *a;
  D.mode = HMC_RUN;
  if (hmc_use_tc(m)) return hmc_tc_launch(m, D, a->n, (cudaStream_t)stream);
  return hmc_launch(m, m->prog, D, a->n, (cudaStream_t)stream);
}

int bgm_hmc_adapt(const double* accept_stat_dev, int t, long long n_total, float target, float rate,
                  float* step_dev, void* stream) {
  using namespace bgm;
  if (!accept_stat_dev || !step_dev) return fail(BGM_ERR_ARG, "bgm_hmc_adapt: null pointer");
  if (t < 0 || n_total < 1) return fail(BGM_ERR_ARG, "bgm_hmc_adapt: bad t / n_total");
  hmc_adapt_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(accept_stat_dev, t, n_total, target, rate, step_dev);
  BGM_CUDA_OK(cudaGetLastError());
  return 0;
}

int bgm_hmc_noise(uint64_t seed, int64_t row_offset, int n, int z_dim, int t_begin, int t_end,
                  float* z0_dev, float* mom_dev, float* logu_dev, void* stream) {
  using namespace bgm;
  if (n < 1 || z_dim < 1 || t_end < t_begin) return fail(BGM_ERR_ARG, "bgm_hmc_noise: bad sizes");
  const long long total = (long long)(t_end - t_begin + 1) * n;
  const int grid = (int)std::min<long long>((total + 255) / 256, 148 * 16);
  hmc_noise_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(seed, row_offset, n, z_dim, t_begin, t_end, z0_dev,
                                                          mom_dev, logu_dev);
  BGM_CUDA_OK(cudaGetLastError());
  return 0;
}

int bgm_hmc_predict(const bgm_hmc* m, const float* z_samples_dev, int n_keep, int n, int sample0,
                    uint64_t seed, int64_t row_offset, const float* noise_dev, float* out_x_dev,
                    void* stream) {
  using namespace bgm;
  if (!m || !z_samples_dev || !out_x_dev) return fail(BGM_ERR_ARG, "bgm_hmc_predict: null model / pointer");
  if (n_keep < 1 || n < 1) return fail(BGM_ERR_ARG, "bgm_hmc_predict: n_keep and n must be >= 1");
  if ((long long)n_keep * n > 0x7fffffffLL) return fail(BGM_ERR_ARG, "bgm_hmc_predict: n_keep*n exceeds 2^31-1; call per chunk of samples");
  HmcDev D;
  memset(&D, 0, sizeof(D));
  D.a.n = n_keep * n;
  D.a.ldx = 0;
  D.a.seed = seed;
  D.a.row_offset = row_offset;
  D.mode = HMC_PREDICT;
  D.z_in = z_samples_dev;
  D.out_x = out_x_dev;
  D.noise_x = noise_dev;
  D.n_per_sample = n;
  D.sample0 = sample0;
  if (hmc_use_tc(m)) return hmc_tc_launch(m, D, n_keep * n, (cudaStream_t)stream);
  return hmc_launch(m, m->prog_fwd, D, n_keep * n, (cudaStream_t)stream);
}

int bgm_hmc_heads(const bgm_hmc* m, const float* z_dev, int n, float* out_mu_dev, float* out_var_dev,
                  void* stream) {
  using namespace bgm;
  if (!m || !z_dev || !out_mu_dev || !out_var_dev) return fail(BGM_ERR_ARG, "bgm_hmc_heads: null model / pointer");
  if (n < 1) return fail(BGM_ERR_ARG, "bgm_hmc_heads: n must be >= 1");
  HmcDev D;
  memset(&D, 0, sizeof(D));
  D.a.n = n;
  D.a.ldx = 0;
  D.mode = HMC_PREDICT;
  D.z_in = z_dev;
  D.out_x = out_mu_dev;
  D.out_var = out_var_dev;
  D.n_per_sample = n;
  if (hmc_use_tc(m)) return hmc_tc_launch(m, D, n, (cudaStream_t)stream);
  return hmc_launch(m, m->prog_fwd, D, n, (cudaStream_t)stream);
}

// Posterior-predictive reduction of BGM.predict (bgm/base.py:640-660): for every column m of draws (S, M) the mean
// over the S kept states and np.quantile(., q) at two levels (linear interpolation between the order statistics
// lo = floor(q (S - 1)) and lo + 1, the same fp32 formula as a sort would feed).  Thread = column: one pass over
// the draws (coalesced across columns) keeping the K smallest and K largest values in sorted register arrays,
// instead of sorting S x M values.  K = order statistics needed from either end (<= 16).
int bgm_column_quantiles(const float* draws_dev, int S, long long M, double q_lo, double q_hi, float* mean_dev,
                         float* lo_dev, float* hi_dev, void* stream) {
  using namespace bgm;
  if (!draws_dev || !mean_dev || S < 1 || M < 1) return fail(BGM_ERR_ARG, "bgm_column_quantiles: bad argument");
  if (!(q_lo >= 0.0 && q_lo <= 1.0 && q_hi >= 0.0 && q_hi <= 1.0)) return fail(BGM_ERR_ARG, "bgm_column_quantiles: q outside [0, 1]");
  ColQ Q;
  Q.draws = draws_dev; Q.S = S; Q.M = M; Q.mean = mean_dev; Q.lo = lo_dev; Q.hi = hi_dev;
  const double pl = q_lo * (S - 1), ph = q_hi * (S - 1);
  Q.lo_idx = (int)std::floor(pl); Q.lo_frac = (float)(pl - Q.lo_idx);
  Q.hi_idx = (int)std::floor(ph); Q.hi_frac = (float)(ph - Q.hi_idx);
  const int need_lo = std::min(Q.lo_idx + 2, S);                 // smallest values needed
  const int need_hi = std::min(S - Q.hi_idx, S);                 // largest values needed (index hi_idx and hi_idx + 1)
  const int K = (lo_dev && hi_dev) ? std::max(need_lo, need_hi) : 1;
  const unsigned grid = (unsigned)std::min<long long>((M + 127) / 128, 1 << 20);
  cudaStream_t st = (cudaStream_t)stream;
  if (K <= 4) column_quantiles_kernel<4><<<grid, 128, 0, st>>>(Q);
  else if (K <= 8) column_quantiles_kernel<8><<<grid, 128, 0, st>>>(Q);
  else if (K <= 16) column_quantiles_kernel<16><<<grid, 128, 0, st>>>(Q);
  else return fail(BGM_ERR_UNSUPPORTED, "bgm_column_quantiles: more than 16 order statistics from one end; sort instead");
  BGM_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
