// libbgm_b200.so -- C ABI (include/bgm_b200.h) over the sm_100a kernels.
// Host code here only packs weights, validates arguments and launches.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "causal.cuh"
#include "causal_tc.cuh"

namespace bgm {

thread_local std::string g_last_error;
int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

static int round_up(int a, int b) { return (a + b - 1) / b * b; }

// ----------------------------------------------------------- weight packer --
// A Dense layer W[K][N] (+bias) whose K input rows have already been mapped onto
// rows of the source buffer (`krows`), cut into column tiles [kp][NT] | bias[NT].
struct Packer {
  std::vector<float> image;
  std::vector<TileOp> ops;
  long long macs = 0, issued = 0;

  // Wm: dense [krows][N] row-major (krows = rows of the source buffer actually used)
  void add_tile(const std::vector<float>& Wm, const std::vector<float>& b, int krows, int N,
                int col_begin, int ncols, int ctype, int src, int epi, int c0) {
    static const int NTS[3] = {64, 32, 8};
    const int NT = NTS[ctype];
    const int kp = round_up(krows, 4);
    TileOp op;
    memset(&op, 0, sizeof(op));
    op.w_off = (int)image.size();
    image.resize(image.size() + (size_t)kp * NT, 0.f);
    for (int k = 0; k < krows; ++k)
      for (int c = 0; c < ncols; ++c)
        image[op.w_off + (size_t)k * NT + c] = Wm[(size_t)k * N + col_begin + c];
    op.b_off = (int)image.size();
    image.resize(image.size() + NT, 0.f);
    for (int c = 0; c < ncols; ++c) image[op.b_off + c] = b[col_begin + c];
    op.kp = (short)kp;
    op.ctype = (unsigned char)ctype;
    op.src = (unsigned char)src;
    op.epi = (unsigned char)epi;
    op.c0 = (short)c0;
    op.nvalid = (short)ncols;
    ops.push_back(op);
    issued += (long long)kp * NT;
  }

  // hidden layer: one tile, LeakyReLU, output to the activation buffer
  int add_hidden(const std::vector<float>& Wm, const std::vector<float>& b, int krows, int N,
                 int src, int epi = EPI_ACT) {
    if (N > 64) return -1;
    const int ctype = N > 32 ? 0 : (N > 8 ? 1 : 2);
    add_tile(Wm, b, krows, N, 0, N, ctype, src, epi, 0);
    return 0;
  }

  // output columns [col_begin, col_begin+ncols) of a final layer, cut by issue cost
  // (per k: NT=64 costs 68 slots, NT=32 35, NT=8 11)
  void add_final(const std::vector<float>& Wm, const std::vector<float>& b, int krows, int N,
                 int col_begin, int ncols, int src, int epi, int c0) {
    int done = 0;
    while (done < ncols) {
      const int rem = ncols - done;
      int ctype, take;
      if (rem > 48) { ctype = 0; take = std::min(rem, 64); }
      else if (rem > 24) { ctype = 1; take = std::min(rem, 32); }
      else { ctype = 2; take = std::min(rem, 8); }
      add_tile(Wm, b, krows, N, col_begin + done, take, ctype, src, epi, c0 + done);
      done += take;
    }
  }
};

// Thin QR of A (rows x cols, rows >= cols, row-major, float64) by modified Gram-Schmidt
// with one re-orthogonalisation pass: A = Q R, Q rows x cols with orthonormal (or zero,
// if A is rank deficient) columns, R cols x cols upper triangular.
static void thin_qr(const std::vector<double>& A, int rows, int cols, std::vector<double>& Q,
                    std::vector<double>& R) {
  Q.assign((size_t)rows * cols, 0.0);
  R.assign((size_t)cols * cols, 0.0);
  std::vector<double> u(rows);
  double scale = 0.0;
  for (double a : A) scale = std::max(scale, std::fabs(a));
  for (int j = 0; j < cols; ++j) {
    for (int r = 0; r < rows; ++r) u[r] = A[(size_t)r * cols + j];
    for (int pass = 0; pass < 2; ++pass)
      for (int i = 0; i < j; ++i) {
        double d = 0.0;
        for (int r = 0; r < rows; ++r) d += u[r] * Q[(size_t)r * cols + i];
        R[(size_t)i * cols + j] += d;
        for (int r = 0; r < rows; ++r) u[r] -= d * Q[(size_t)r * cols + i];
      }
    double nrm = 0.0;
    for (int r = 0; r < rows; ++r) nrm += u[r] * u[r];
    nrm = std::sqrt(nrm);
    if (nrm > 1e-12 * std::max(scale, 1e-300) * std::sqrt((double)rows)) {
      R[(size_t)j * cols + j] = nrm;
      for (int r = 0; r < rows; ++r) Q[(size_t)r * cols + j] = u[r] / nrm;
    }
  }
}

struct HostNet {
  int L;
  std::vector<int> dims;
  std::vector<std::vector<float>> W, b;
};
static int read_net(const bgm_net_desc* d, HostNet& n, const char* name) {
  if (!d || d->n_layers < 1 || !d->dims || !d->params)
    return fail(BGM_ERR_ARG, std::string(name) + ": null / empty net description");
  n.L = d->n_layers;
  n.dims.assign(d->dims, d->dims + n.L + 1);
  const float* p = d->params;
  for (int l = 0; l < n.L; ++l) {
    const int K = n.dims[l], N = n.dims[l + 1];
    if (K < 1 || N < 1) return fail(BGM_ERR_ARG, std::string(name) + ": non-positive layer size");
    n.W.emplace_back(p, p + (size_t)K * N);
    p += (size_t)K * N;
    n.b.emplace_back(p, p + N);
    p += N;
  }
  return 0;
}

// ---- tensor-core engine image (causal_tc.cuh) ----
static void split_tf32_host(float w, float& hi, float& lo) {
  uint32_t b;
  memcpy(&b, &w, 4);
  uint32_t h = (b + 0x1000u) & 0xffffe000u;
  memcpy(&hi, &h, 4);
  const float l = w - hi;
  memcpy(&b, &l, 4);
  b = (b + 0x1000u) & 0xffffe000u;
  memcpy(&lo, &b, 4);
}
// Wm [K][N] (Keras [in][out]) -> two [K/4][N][4] blocks (umma.cuh layout), hi and lo
static void pack_mma_k64(std::vector<float>& img, const std::vector<float>& Wm, int N, int& off_hi, int& off_lo,
                         int K = 64) {
  off_hi = (int)img.size();
  img.resize(img.size() + (size_t)K * N, 0.f);
  off_lo = (int)img.size();
  img.resize(img.size() + (size_t)K * N, 0.f);
  for (int k = 0; k < K; ++k)
    for (int n = 0; n < N; ++n) {
      float hi, lo;
      split_tf32_host(Wm[(size_t)k * N + n], hi, lo);
      const int idx = (k >> 2) * (N * 4) + n * 4 + (k & 3);
      img[off_hi + idx] = hi;
      img[off_lo + idx] = lo;
    }
}
static int push_floats(std::vector<float>& img, const float* p, size_t count) {
  while (img.size() % 4) img.push_back(0.f);   // float4 loads
  const int off = (int)img.size();
  img.insert(img.end(), p, p + count);
  return off;
}
static bool small_net_shape(const HostNet& n) {
  return n.L == 4 && n.dims[1] == 64 && n.dims[2] == 32 && n.dims[3] == 8 && n.dims[4] == 2;
}
// first layer [in][64] as is (its rows in order) + the mask of the entries of the input vector
// [z.., x] those rows multiply (rows[] is ascending)
static int pack_first_layer(std::vector<float>& img, const HostNet& net, const std::vector<int>& rows, int zd,
                            unsigned long long& mask) {
  (void)zd;
  mask = 0;
  for (size_t r = 0; r < rows.size(); ++r) mask |= 1ull << rows[r];
  return push_floats(img, net.W[0].data(), rows.size() * 64);
}

}  // namespace bgm

using namespace bgm;

struct bgm_causal {
  CausalProgram prog;
  float* image_dev = nullptr;
  float* proj_dev = nullptr;   // U [p][HP] | b [p] of the covariate projection (proj_dim > 0)
  int proj_dim = 0, proj_hp = 0;
  int warps = 0;
  int smem_bytes = 0;
  int effect_warps = 0;
  int effect_smem_bytes = 0;
  int sm_count = 0;
  int smem_max = 0;            // opt-in shared memory per block of the device
  long long macs = 0, issued = 0;
  // tensor-core engine (causal_tc.cuh); tc.enabled == 0 when the net shape is outside it
  TcProgram tc;
  float* tc_image_dev = nullptr;
  int tc_smem_bytes = 0;
  long long tc_issued = 0;     // FMA-equivalents per row per evaluation (tensor + FMA pipe)
  int sampler = 0;             // 0: auto, 1: SIMT engine, 2: tensor-core engine
  int tc16 = 0;                // tensor engine with 8 warps per tile (fits in shared memory)
  // 8-warp-per-tile sampler with the first layers of f and h on the tensor cores too (z_dim <= 6): its own image
  TcProgram tc16p;
  float* tc16_image_dev = nullptr;
  int tc16_l1 = 0;
};

static int check_data(const char* fn, const bgm_causal* m, const float* x, const float* y, const float* v,
                      int ldv, const float* vproj, int ldvproj, const float* r0, const int* sched, int n) {
  if (!x || !y) return fail(BGM_ERR_ARG, std::string(fn) + ": null data pointer");
  if (n < 1) return fail(BGM_ERR_ARG, std::string(fn) + ": n must be >= 1");
  if (!sched) return fail(BGM_ERR_ARG, std::string(fn) + ": sched_dev is required");
  if (m->proj_dim) {
    if (!vproj || !r0)
      return fail(BGM_ERR_ARG, std::string(fn) + ": this model projects the covariates: vproj_dev and r0_dev "
                                                 "(bgm_causal_project) are required");
    if (ldvproj < m->proj_dim || ldvproj % 4 != 0 || reinterpret_cast<uintptr_t>(vproj) % 16 != 0)
      return fail(BGM_ERR_ARG, std::string(fn) + ": vproj_dev must be 16-byte aligned, ldvproj >= proj_dim, % 4 == 0");
    return 0;
  }
  if (!v) return fail(BGM_ERR_ARG, std::string(fn) + ": null data pointer");
  if (ldv < m->prog.p || ldv % 4 != 0)
    return fail(BGM_ERR_ARG, std::string(fn) + ": ldv must be >= v_dim and a multiple of 4");
  if (reinterpret_cast<uintptr_t>(v) % 16 != 0)
    return fail(BGM_ERR_ARG, std::string(fn) + ": v_dev must be 16-byte aligned");
  return 0;
}

template <int ZMAX>
static int launch_mh_t(const bgm_causal* m, const MhDev& D, int grid, cudaStream_t st) {
  auto k = causal_mh_kernel<ZMAX>;
  BGM_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, m->smem_bytes));
  k<<<grid, m->warps * 32, m->smem_bytes, st>>>(m->prog, m->image_dev, D);
  BGM_CUDA_OK(cudaGetLastError());
  return 0;
}
// dynamic shared memory of causal_mh_tc16_kernel<8, true>: image, partials [2][2][128][4], noise [2][128][4] + [2][128][zd - 4]
static int tc16_l1_smem(const TcProgram& T) {
  const int n1s = T.zd > 4 ? T.zd - 4 : 0;
  return T.image_floats * 4 + TC16_XCH_FLOATS * 4 + 2 * TC_ROWS * 16 + 2 * TC_ROWS * n1s * 4;
}
template <int ZMAX>
static int launch_mh_tc_t(const bgm_causal* m, const MhDev& D, int grid, cudaStream_t st) {
  if (m->tc16 && ZMAX <= 12) {   // 8 warps per tile (128 registers per thread: spills beyond zd = 12)
    if constexpr (ZMAX == 8) {
      if (m->tc16_l1) {          // first layers of f and h on the tensor cores as well
        auto k = causal_mh_tc16_kernel<ZMAX, true>;
        const int smem = tc16_l1_smem(m->tc16p);
        BGM_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        k<<<grid, 512, smem, st>>>(m->tc16p, m->tc16_image_dev, D);
        BGM_CUDA_OK(cudaGetLastError());
        return 0;
      }
    }
    auto k = causal_mh_tc16_kernel<ZMAX, false>;
    const int smem = m->tc_smem_bytes + (ZMAX == 8 ? 2 : 1) * TC16_XCH_FLOATS * 4;
    BGM_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k<<<grid, 512, smem, st>>>(m->tc, m->tc_image_dev, D);
    BGM_CUDA_OK(cudaGetLastError());
    return 0;
  }
  auto k = causal_mh_tc_kernel<ZMAX>;
  BGM_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, m->tc_smem_bytes));
  k<<<grid, 256, m->tc_smem_bytes, st>>>(m->tc, m->tc_image_dev, D);
  BGM_CUDA_OK(cudaGetLastError());
  return 0;
}
static bool use_tc(const bgm_causal* m) { return m->tc.enabled && m->sampler != 1; }
static int launch_mh(const bgm_causal* m, MhDev& D, cudaStream_t st) {
  // the conditional prior (bgm_mh_args.prior_dev) is a flag of the SIMT engine only
  if (use_tc(m) && !D.a.prior_dev) {
    // 128-row tiles, two warpgroups per CTA
    const int nt = (D.a.n + TC_ROWS - 1) / TC_ROWS;
    const int grid = std::max(1, std::min((nt + 1) / 2, m->sm_count));
    const bool need_init = !(D.a.init_mode == 0 && D.mode == 0);
    const int n_iter = (D.mode == 1 ? 0 : D.a.t_end - D.a.t_begin) + (need_init ? 1 : 0);
    D.nchunks = std::max(1, (n_iter + 15) / 16);
    BGM_CUDA_OK(cudaMemsetAsync(D.a.sched_dev, 0, sizeof(int) * (size_t)(nt + 1), st));
    const int zd = m->prog.zd;
    if (zd <= 8) return launch_mh_tc_t<8>(m, D, grid, st);
    if (zd <= 12) return launch_mh_tc_t<12>(m, D, grid, st);
    if (zd <= 16) return launch_mh_tc_t<16>(m, D, grid, st);
    if (zd <= 20) return launch_mh_tc_t<20>(m, D, grid, st);
    return launch_mh_tc_t<32>(m, D, grid, st);
  }
  const int ntiles = (D.a.n + TILE_ROWS - 1) / TILE_ROWS;
  const int grid = std::max(1, std::min((ntiles + m->warps - 1) / m->warps, m->sm_count));
  const int zd = m->prog.zd;
  // iteration chunks of ~16 per work unit, fewer when there are many tiles per warp anyway
  const bool need_init = !(D.a.init_mode == 0 && D.mode == 0);
  const int n_iter = (D.mode == 1 ? 0 : D.a.t_end - D.a.t_begin) + (need_init ? 1 : 0);
  D.nchunks = std::max(1, (n_iter + 15) / 16);
  BGM_CUDA_OK(cudaMemsetAsync(D.a.sched_dev, 0, sizeof(int) * (size_t)(ntiles + 1), st));
  if (zd <= 8) return launch_mh_t<8>(m, D, grid, st);
  if (zd <= 16) return launch_mh_t<16>(m, D, grid, st);
  return launch_mh_t<32>(m, D, grid, st);
}

extern "C" {

const char* bgm_last_error(void) { return g_last_error.c_str(); }
int bgm_version(void) { return 100; }

int bgm_device_info(int* sm_count, int* smem_optin_bytes, int* clock_khz) {
  int dev = 0;
  BGM_CUDA_OK(cudaGetDevice(&dev));
  if (sm_count) BGM_CUDA_OK(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev));
  if (smem_optin_bytes)
    BGM_CUDA_OK(cudaDeviceGetAttribute(smem_optin_bytes, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  if (clock_khz) BGM_CUDA_OK(cudaDeviceGetAttribute(clock_khz, cudaDevAttrClockRate, dev));
  return 0;
}

int bgm_causal_create(bgm_causal** out, const int z_dims[4], int v_dim, int binary_treatment,
                      float sigma_v, float sigma_x, float sigma_y, const bgm_net_desc* g_net,
                      const bgm_net_desc* f_net, const bgm_net_desc* h_net) {
  if (!out || !z_dims) return fail(BGM_ERR_ARG, "bgm_causal_create: null argument");
  *out = nullptr;
  const int d0 = z_dims[0], d1 = z_dims[1], d2 = z_dims[2], d3 = z_dims[3];
  if (d0 < 0 || d1 < 0 || d2 < 0 || d3 < 0) return fail(BGM_ERR_ARG, "z_dims must be >= 0");
  const int zd = d0 + d1 + d2 + d3;
  if (zd < 1 || zd > 32) return fail(BGM_ERR_UNSUPPORTED, "sum(z_dims) must be in [1, 32]");
  if (v_dim < 1) return fail(BGM_ERR_ARG, "v_dim must be >= 1");
  HostNet g, f, h;
  int rc;
  if ((rc = read_net(g_net, g, "g_net"))) return rc;
  if ((rc = read_net(f_net, f, "f_net"))) return rc;
  if ((rc = read_net(h_net, h, "h_net"))) return rc;
  if (g.dims[0] != zd || g.dims[g.L] != v_dim + 1)
    return fail(BGM_ERR_ARG, "g_net must map sum(z_dims) -> v_dim+1 (causalbgm/base.py:74)");
  if (f.dims[0] != d0 + d1 + 1 || f.dims[f.L] != 2)
    return fail(BGM_ERR_ARG, "f_net must map z0+z1+1 -> 2 (causalbgm/base.py:78)");
  if (h.dims[0] != d0 + d2 || h.dims[h.L] != 2)
    return fail(BGM_ERR_ARG, "h_net must map z0+z2 -> 2 (causalbgm/base.py:80)");
  if (g.L < 2 || f.L < 2 || h.L < 2)
    return fail(BGM_ERR_UNSUPPORTED, "each net needs at least one hidden layer");

  const int kin = round_up(zd + 1, 4);
  Packer pk;
  CausalProgram P;
  memset(&P, 0, sizeof(P));

  int proj_dim = 0, proj_hp = 0;
  std::vector<float> proj_host, proj_RT;
  // maps each net's first-layer input rows onto rows of the shared input buffer
  // zin = [z (zd rows), x (row zd), zero pad]
  auto remap_first = [&](const HostNet& net, const std::vector<int>& rows) {
    const int N = net.dims[1];
    std::vector<float> Wm((size_t)(zd + 1) * N, 0.f);
    for (size_t k = 0; k < rows.size(); ++k)
      for (int c = 0; c < N; ++c) Wm[(size_t)rows[k] * N + c] = net.W[0][k * N + c];
    return Wm;
  };
  auto add_net = [&](const HostNet& net, const std::vector<int>& rows, int kind) -> int {
    // kind 0: g (SSE over v_dim columns + sigma head), 1: f / h (mu, sigma -> scratch)
    for (int l = 0; l < net.L; ++l) {
      const int K = net.dims[l], N = net.dims[l + 1];
      pk.macs += (long long)K * N;
      const bool last = l == net.L - 1;
      const int src = l == 0 ? 0 : 1;
      const std::vector<float>& Wm = l == 0 ? remap_first(net, rows) : net.W[l];
      const int krows = l == 0 ? zd + 1 : K;
      if (l > 0 && K > ACT_ROWS) return -1;
      if (!last) {
        if (pk.add_hidden(Wm, net.b[l], krows, N, src)) return -1;
      } else if (kind == 0) {
        const int H = K;
        if (l > 0 && v_dim > H + 8) {
          // covariate likelihood through the row space of this layer (bgm_causal_project):
          // M^T = U R; the SSE tiles see R^T (H x H, zero bias) and the projected data.
          std::vector<double> A((size_t)v_dim * H), Q, R;
          for (int k = 0; k < H; ++k)
            for (int c = 0; c < v_dim; ++c) A[(size_t)c * H + k] = Wm[(size_t)k * N + c];
          thin_qr(A, v_dim, H, Q, R);
          std::vector<float> RT((size_t)H * H), zb(H, 0.f);
          for (int j = 0; j < H; ++j)
            for (int i = 0; i < H; ++i) RT[(size_t)j * H + i] = (float)R[(size_t)i * H + j];
          pk.add_final(RT, zb, H, H, 0, H, src, EPI_SSE, 0);
          proj_RT = RT;
          proj_dim = H;
          proj_hp = H > 32 ? 64 : 32;
          proj_host.assign((size_t)v_dim * proj_hp + v_dim, 0.f);
          for (int c = 0; c < v_dim; ++c) {
            for (int i = 0; i < H; ++i) proj_host[(size_t)c * proj_hp + i] = (float)Q[(size_t)c * H + i];
            proj_host[(size_t)v_dim * proj_hp + c] = net.b[l][c];
          }
        } else {
          pk.add_final(Wm, net.b[l], krows, N, 0, v_dim, src, EPI_SSE, 0);
        }
        if (sigma_v < 0.f) pk.add_final(Wm, net.b[l], krows, N, v_dim, 1, src, EPI_OUT, 1);
      } else {
        pk.add_final(Wm, net.b[l], krows, N, 0, 2, src, EPI_OUT, 0);
      }
    }
    return 0;
  };
  std::vector<int> g_rows, f_rows, h_rows;
  for (int k = 0; k < zd; ++k) g_rows.push_back(k);
  for (int k = 0; k < d0 + d1; ++k) f_rows.push_back(k);
  f_rows.push_back(zd);  // x
  for (int k = 0; k < d0; ++k) h_rows.push_back(k);
  for (int k = 0; k < d2; ++k) h_rows.push_back(d0 + d1 + k);
  const char* wide = "hidden layers wider than 64 units are not supported by the sm_100a sampler kernel";
  if (add_net(g, g_rows, 0)) return fail(BGM_ERR_UNSUPPORTED, wide);
  pk.ops.back().post = POST_G;
  P.g_end = (int)pk.ops.size();
  P.f_img_begin = (int)pk.image.size();
  if (add_net(f, f_rows, 1)) return fail(BGM_ERR_UNSUPPORTED, wide);
  pk.ops.back().post = POST_F;
  P.f_end = (int)pk.ops.size();
  P.f_img_end = (int)pk.image.size();
  if (add_net(h, h_rows, 1)) return fail(BGM_ERR_UNSUPPORTED, wide);
  pk.ops.back().post = POST_H;
  P.h_end = (int)pk.ops.size();
  P.n_ops = P.h_end;
  if (P.n_ops > MAX_OPS) return fail(BGM_ERR_UNSUPPORTED, "too many column tiles (v_dim too large)");
  for (int i = 0; i < P.n_ops; ++i) P.ops[i] = pk.ops[i];
  P.zd = zd;
  P.kin = kin;
  P.p = v_dim;
  P.p_data = proj_dim ? proj_dim : v_dim;
  P.proj = proj_dim ? 1 : 0;
  P.binary = binary_treatment ? 1 : 0;
  P.s2v = sigma_v >= 0.f ? sigma_v * sigma_v : -1.f;
  P.s2x = sigma_x >= 0.f ? sigma_x * sigma_x : -1.f;
  P.s2y = sigma_y >= 0.f ? sigma_y * sigma_y : -1.f;
  pk.image.resize(pk.image.size() + IMG_PAD, 0.f);
  P.image_floats = (int)pk.image.size();
  P.per_warp_floats = (ACT_ROWS + kin + SCR_SLOTS) * TILE_ROWS;

  // ---- tensor-core engine: g hidden layers 64 wide, projected likelihood, standard f / h ----
  TcProgram T, T16;
  memset(&T, 0, sizeof(T));
  memset(&T16, 0, sizeof(T16));
  std::vector<float> tc_image, tc16_image;
  long long tc_issued = 0;
  bool tc_ok = proj_dim == 64 && g.L >= 3 && g.L - 1 <= TC_MAX_MMA && small_net_shape(f) && small_net_shape(h);
  for (int l = 1; l < g.L; ++l) tc_ok = tc_ok && g.dims[l] == 64;
  // l1 = false: the image of causal_mh_tc_kernel / causal_effect_tc_kernel / causal_mh_tc16_kernel<ZMAX, false>;
  // l1 = true: the image of causal_mh_tc16_kernel<8, true> -- no fp32 first layers of f / h and no effect-only pieces,
  // but [z.., x, 0.., 1] (8) -> [f_h1 | h_h1] (128) as tf32 hi / lo images
  auto pack_tc = [&](TcProgram& T, std::vector<float>& tc_image, bool l1) {
    {
      T.zd = zd; T.p = v_dim; T.binary = P.binary;
      T.s2v = P.s2v; T.s2x = P.s2x; T.s2y = P.s2y;
      T.n_mma = g.L - 1;
      for (int m = 0; m < T.n_mma; ++m) {
        const bool last = m == T.n_mma - 1;
        pack_mma_k64(tc_image, last ? proj_RT : g.W[m + 1], 64, T.w_hi[m], T.w_lo[m]);
      }
      pack_mma_k64(tc_image, f.W[1], 32, T.f2_hi, T.f2_lo);
      pack_mma_k64(tc_image, h.W[1], 32, T.h2_hi, T.h2_lo);
      {
        // [f_h2 | h_h2] (64) -> [f_h3 (8) | h_h3 (8)], block diagonal
        std::vector<float> w3((size_t)64 * 16, 0.f), b3(16);
        for (int k = 0; k < 32; ++k)
          for (int c = 0; c < 8; ++c) {
            w3[(size_t)k * 16 + c] = f.W[2][(size_t)k * 8 + c];
            w3[(size_t)(32 + k) * 16 + 8 + c] = h.W[2][(size_t)k * 8 + c];
          }
        for (int c = 0; c < 8; ++c) { b3[c] = f.b[2][c]; b3[8 + c] = h.b[2][c]; }
        pack_mma_k64(tc_image, w3, 16, T.w3_hi, T.w3_lo);
        T.b3 = push_floats(tc_image, b3.data(), 16);
        if (!l1) {
          // the effect kernel runs f alone: f_h2 (32) -> [f_h3 (8) | 0 (8)]
          std::vector<float> f3((size_t)32 * 16, 0.f);
          for (int k = 0; k < 32; ++k)
            for (int c = 0; c < 8; ++c) f3[(size_t)k * 16 + c] = f.W[2][(size_t)k * 8 + c];
          pack_mma_k64(tc_image, f3, 16, T.f3_hi, T.f3_lo, 32);
          T.fb3 = push_floats(tc_image, f.b[2].data(), 8);
        } else {
          // rows: input j of [z_0 .. z_{zd-1}, x, 0.., 1 (row 7)]; columns: f_h1 (64) | h_h1 (64)
          std::vector<float> w1((size_t)8 * 128, 0.f);
          for (size_t r = 0; r < f_rows.size(); ++r)
            for (int c = 0; c < 64; ++c) w1[(size_t)f_rows[r] * 128 + c] = f.W[0][r * 64 + c];
          for (size_t r = 0; r < h_rows.size(); ++r)
            for (int c = 0; c < 64; ++c) w1[(size_t)h_rows[r] * 128 + 64 + c] = h.W[0][r * 64 + c];
          for (int c = 0; c < 64; ++c) {
            w1[(size_t)7 * 128 + c] = f.b[0][c];
            w1[(size_t)7 * 128 + 64 + c] = h.b[0][c];
          }
          T.l1_k = 8;
          pack_mma_k64(tc_image, w1, 128, T.l1_hi, T.l1_lo, 8);
        }
      }
      for (int m = 0; m + 1 < T.n_mma; ++m) T.gb[m] = push_floats(tc_image, g.b[m + 1].data(), 64);
      T.gW1 = push_floats(tc_image, g.W[0].data(), (size_t)zd * 64);
      T.gb1 = push_floats(tc_image, g.b[0].data(), 64);
      std::vector<float> wsig(64);
      const int NL = v_dim + 1;
      for (int k = 0; k < 64; ++k) wsig[k] = g.W[g.L - 1][(size_t)k * NL + v_dim];
      T.wsig = push_floats(tc_image, wsig.data(), 64);
      T.bsig = push_floats(tc_image, &g.b[g.L - 1][v_dim], 1);
      if (!l1) {
        T.fW1 = pack_first_layer(tc_image, f, f_rows, zd, T.fmask);
        T.fb1 = push_floats(tc_image, f.b[0].data(), 64);
        T.hW1 = pack_first_layer(tc_image, h, h_rows, zd, T.hmask);
        T.hb1 = push_floats(tc_image, h.b[0].data(), 64);
      }
      T.fb2 = push_floats(tc_image, f.b[1].data(), 32);
      T.hb2 = push_floats(tc_image, h.b[1].data(), 32);
      T.fW4 = push_floats(tc_image, f.W[3].data(), 16);
      T.fb4 = push_floats(tc_image, f.b[3].data(), 2);
      T.hW4 = push_floats(tc_image, h.W[3].data(), 16);
      T.hb4 = push_floats(tc_image, h.b[3].data(), 2);
      while (tc_image.size() % 4) tc_image.push_back(0.f);
      T.image_floats = (int)tc_image.size();
      T.enabled = 1;
    }
  };
  if (tc_ok) {
    pack_tc(T, tc_image, false);
    // tensor pipe: 3 TF32 products per FMA of the 64x64 layers; FMA pipe: the narrow layers
    tc_issued = 3LL * (4096LL * T.n_mma + 2 * 64 * 32 + 64 * 16) + (long long)zd * 64 + 64 +
                (long long)(f.dims[0] + h.dims[0]) * 64 + 2 * 16;
    if (zd + 2 <= 8) pack_tc(T16, tc16_image, true);
  }

  int dev = 0, smem_max = 0, sms = 0;
  BGM_CUDA_OK(cudaGetDevice(&dev));
  BGM_CUDA_OK(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  BGM_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int static_smem = 64;  // mbarrier + alignment slack
  auto fit = [&](int img_floats) {
    int w = (smem_max - static_smem - img_floats * 4) / (P.per_warp_floats * 4);
    return std::min(w, MAX_WARPS);
  };
  bgm_causal* m = new bgm_causal();
  m->warps = fit(P.image_floats);
  if (m->warps < 1) {
    delete m;
    return fail(BGM_ERR_NOMEM, "packed nets do not fit in shared memory");
  }
  if (m->warps > 8) m->warps = (m->warps / 4) * 4;  // keep the 4 SM sub-partitions balanced
  m->smem_bytes = (P.image_floats + m->warps * P.per_warp_floats) * 4;
  const int f_floats = P.f_img_end - P.f_img_begin + IMG_PAD;
  m->effect_warps = std::min(fit(f_floats), MAX_WARPS);
  m->effect_smem_bytes = (f_floats + m->effect_warps * P.per_warp_floats) * 4;
  m->sm_count = sms;
  m->smem_max = smem_max;
  m->prog = P;
  m->macs = pk.macs;
  m->issued = pk.issued;
  m->proj_dim = proj_dim;
  m->proj_hp = proj_hp;
  cudaError_t e = cudaMalloc(&m->image_dev, pk.image.size() * sizeof(float));
  if (e == cudaSuccess)
    e = cudaMemcpy(m->image_dev, pk.image.data(), pk.image.size() * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess && proj_dim) {
    e = cudaMalloc(&m->proj_dev, proj_host.size() * sizeof(float));
    if (e == cudaSuccess)
      e = cudaMemcpy(m->proj_dev, proj_host.data(), proj_host.size() * sizeof(float), cudaMemcpyHostToDevice);
  }
  m->tc = T;
  m->tc_issued = tc_issued;
  m->tc_smem_bytes = T.image_floats * 4;
  if (T.enabled && m->tc_smem_bytes + 256 > smem_max) m->tc.enabled = 0;
  m->tc16 = m->tc.enabled && m->tc_smem_bytes + (zd <= 8 ? 2 : 1) * TC16_XCH_FLOATS * 4 + 128 <= smem_max &&
            !getenv("BGM_TC8");
  m->tc16p = T16;
  m->tc16_l1 = m->tc16 && T16.enabled && tc16_l1_smem(T16) + 128 <= smem_max &&
               !getenv("BGM_TC16_NO_L1");
  if (e == cudaSuccess && m->tc.enabled) {
    e = cudaMalloc(&m->tc_image_dev, tc_image.size() * sizeof(float));
    if (e == cudaSuccess)
      e = cudaMemcpy(m->tc_image_dev, tc_image.data(), tc_image.size() * sizeof(float), cudaMemcpyHostToDevice);
  }
  if (e == cudaSuccess && m->tc16_l1) {
    e = cudaMalloc(&m->tc16_image_dev, tc16_image.size() * sizeof(float));
    if (e == cudaSuccess)
      e = cudaMemcpy(m->tc16_image_dev, tc16_image.data(), tc16_image.size() * sizeof(float), cudaMemcpyHostToDevice);
  }
  if (e != cudaSuccess) {
    if (m->tc16_image_dev) cudaFree(m->tc16_image_dev);
    if (m->tc_image_dev) cudaFree(m->tc_image_dev);
    if (m->proj_dev) cudaFree(m->proj_dev);
    if (m->image_dev) cudaFree(m->image_dev);
    delete m;
    return fail(BGM_ERR_CUDA, std::string("uploading weight image: ") + cudaGetErrorString(e));
  }
  *out = m;
  return 0;
}

void bgm_causal_destroy(bgm_causal* m) {
  if (!m) return;
  if (m->image_dev) cudaFree(m->image_dev);
  if (m->proj_dev) cudaFree(m->proj_dev);
  if (m->tc_image_dev) cudaFree(m->tc_image_dev);
  if (m->tc16_image_dev) cudaFree(m->tc16_image_dev);
  delete m;
}

int bgm_causal_info(const bgm_causal* m, int* smem_bytes, int* warps_per_cta, int* n_ops,
                    long long* macs_per_row, long long* issued_macs_per_row, int* proj_dim) {
  if (!m) return fail(BGM_ERR_ARG, "null model");
  if (smem_bytes) *smem_bytes = m->smem_bytes;
  if (warps_per_cta) *warps_per_cta = m->warps;
  if (n_ops) *n_ops = m->prog.n_ops;
  if (macs_per_row) *macs_per_row = m->macs;
  if (issued_macs_per_row) *issued_macs_per_row = m->issued;
  if (proj_dim) *proj_dim = m->proj_dim;
  return 0;
}

int bgm_causal_set_sampler(bgm_causal* m, int kind) {
  if (!m) return fail(BGM_ERR_ARG, "bgm_causal_set_sampler: null model");
  if (kind < 0 || kind > 2) return fail(BGM_ERR_ARG, "bgm_causal_set_sampler: kind must be 0 (auto), 1 (SIMT) or 2 (tensor)");
  if (kind == 2 && !m->tc.enabled)
    return fail(BGM_ERR_UNSUPPORTED, "bgm_causal_set_sampler: the tensor-core engine needs g hidden layers of 64 "
                                     "units, v_dim > 72 and f/h units [64,32,8]");
  m->sampler = kind;
  return 0;
}

int bgm_causal_sampler_info(const bgm_causal* m, int* active_kind, int* tensor_available, int* tensor_smem_bytes,
                            long long* tensor_issued_macs_per_row) {
  if (!m) return fail(BGM_ERR_ARG, "bgm_causal_sampler_info: null model");
  if (active_kind) *active_kind = use_tc(m) ? 2 : 1;
  if (tensor_available) *tensor_available = m->tc.enabled;
  if (tensor_smem_bytes) {
    // dynamic shared memory of the launch: weight image (+ the exchange buffers of the 16-warp kernel)
    const int zd = m->prog.zd;
    int smem = m->tc_smem_bytes;
    if (m->tc16 && zd <= 12) smem += (zd <= 8 ? 2 : 1) * TC16_XCH_FLOATS * 4;
    if (m->tc16_l1) smem = tc16_l1_smem(m->tc16p);
    *tensor_smem_bytes = smem;
  }
  if (tensor_issued_macs_per_row) *tensor_issued_macs_per_row = m->tc_issued;
  return 0;
}

int bgm_causal_kernel_name(const bgm_causal* m, char* buf, int len) {
  if (!m || !buf || len < 1) return fail(BGM_ERR_ARG, "bgm_causal_kernel_name: null argument");
  const int zd = m->prog.zd;
  int zmax = zd <= 8 ? 8 : (zd <= 16 ? 16 : 32);
  if (use_tc(m) && zd > 8 && zd <= 12) zmax = 12;
  if (use_tc(m) && zd > 16 && zd <= 20) zmax = 20;
  const char* base = !use_tc(m) ? "causal_mh_kernel" : ((m->tc16 && zmax <= 12) ? "causal_mh_tc16_kernel" : "causal_mh_tc_kernel");
  if (use_tc(m) && m->tc16 && zmax <= 12) snprintf(buf, (size_t)len, "%s<%d, %s>", base, zmax, m->tc16_l1 ? "true" : "false");
  else snprintf(buf, (size_t)len, "%s<%d>", base, zmax);
  return 0;
}

int bgm_causal_project(const bgm_causal* m, const float* v_dev, int ldv, int n, float* vproj_dev,
                       int ldvproj, float* r0_dev, void* stream) {
  if (!m) return fail(BGM_ERR_ARG, "bgm_causal_project: null model");
  if (!m->proj_dim) return fail(BGM_ERR_ARG, "bgm_causal_project: this model does not project (proj_dim == 0)");
  if (!v_dev || !vproj_dev || !r0_dev) return fail(BGM_ERR_ARG, "bgm_causal_project: null pointer");
  if (n < 1 || ldv < m->prog.p) return fail(BGM_ERR_ARG, "bgm_causal_project: bad n / ldv");
  if (ldvproj < m->proj_dim || ldvproj % 4 != 0)
    return fail(BGM_ERR_ARG, "bgm_causal_project: ldvproj must be >= proj_dim and a multiple of 4");
  const int p = m->prog.p, HP = m->proj_hp;
  int smem = (p * HP + p) * 4;
  const int stage = smem <= m->smem_max - 1024;   // else U is read through L1/L2 (v_dim >~ 890)
  if (!stage) smem = 0;
  BGM_CUDA_OK(cudaFuncSetAttribute(causal_project_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int grid = std::max(1, std::min((n + 7) / 8, m->sm_count * 2));
  causal_project_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(v_dev, ldv, n, p, HP, m->proj_dev, vproj_dev,
                                                                   ldvproj, r0_dev, stage);
  BGM_CUDA_OK(cudaGetLastError());
  return 0;
}

int bgm_causal_logpost(const bgm_causal* m, const float* x_dev, const float* y_dev,
                       const float* v_dev, int ldv, const float* vproj_dev, int ldvproj,
                       const float* r0_dev, const float* z_dev, int n, float* out_logp_dev,
                       int* sched_dev, void* stream) {
  return bgm_causal_logpost_cond(m, x_dev, y_dev, v_dev, ldv, vproj_dev, ldvproj, r0_dev, z_dev, n, nullptr, 0,
                                 out_logp_dev, sched_dev, stream);
}

int bgm_causal_logpost_cond(const bgm_causal* m, const float* x_dev, const float* y_dev,
                            const float* v_dev, int ldv, const float* vproj_dev, int ldvproj,
                            const float* r0_dev, const float* z_dev, int n, const float* prior_dev,
                            int ldprior, float* out_logp_dev, int* sched_dev, void* stream) {
  if (!m) return fail(BGM_ERR_ARG, "bgm_causal_logpost: null model");
  if (prior_dev && ldprior < m->prog.zd + 1)
    return fail(BGM_ERR_ARG, "bgm_causal_logpost_cond: ldprior must be >= zd + 1");
  int rc = check_data("bgm_causal_logpost", m, x_dev, y_dev, v_dev, ldv, vproj_dev, ldvproj, r0_dev, sched_dev, n);
  if (rc) return rc;
  if (!z_dev || !out_logp_dev) return fail(BGM_ERR_ARG, "bgm_causal_logpost: null z / out pointer");
  MhDev D;
  memset(&D, 0, sizeof(D));
  D.a.x_dev = x_dev; D.a.y_dev = y_dev; D.a.v_dev = v_dev; D.a.ldv = ldv; D.a.n = n;
  D.a.vproj_dev = vproj_dev; D.a.ldvproj = ldvproj; D.a.r0_dev = r0_dev; D.a.sched_dev = sched_dev;
  D.a.z_state_dev = const_cast<float*>(z_dev);
  D.a.lp_state_dev = out_logp_dev;
  D.a.init_mode = 1;
  D.a.prior_dev = prior_dev;
  D.a.ldprior = ldprior;
  D.mode = 1;
  return launch_mh(m, D, (cudaStream_t)stream);
}

int bgm_causal_mh(const bgm_causal* m, const bgm_mh_args* a, void* stream) {
  if (!m || !a) return fail(BGM_ERR_ARG, "bgm_causal_mh: null model / args");
  int rc = check_data("bgm_causal_mh", m, a->x_dev, a->y_dev, a->v_dev, a->ldv, a->vproj_dev, a->ldvproj,
                      a->r0_dev, a->sched_dev, a->n);
  if (rc) return rc;
  if (!a->z_state_dev || !a->lp_state_dev)
    return fail(BGM_ERR_ARG, "bgm_causal_mh: z_state_dev and lp_state_dev are required");
  if (a->init_mode < 0 || a->init_mode > 2) return fail(BGM_ERR_ARG, "bgm_causal_mh: init_mode must be 0, 1 or 2");
  if (a->t_begin < 0 || a->t_end < a->t_begin) return fail(BGM_ERR_ARG, "bgm_causal_mh: bad iteration range");
  if ((a->eps_dev == nullptr) != (a->u_dev == nullptr))
    return fail(BGM_ERR_ARG, "bgm_causal_mh: eps_dev and u_dev must be given together");
  if (a->init_mode == 2 && a->eps_dev)
    return fail(BGM_ERR_ARG, "bgm_causal_mh: init_mode 2 draws from Philox; pass z_state with injected noise");
  if (!a->q_sd_dev) return fail(BGM_ERR_ARG, "bgm_causal_mh: q_sd_dev is required");
  if (a->prior_dev && a->ldprior < m->prog.zd + 1)
    return fail(BGM_ERR_ARG, "bgm_causal_mh: ldprior must be >= zd + 1");
  MhDev D;
  D.a = *a;
  D.mode = 0;
  return launch_mh(m, D, (cudaStream_t)stream);
}

int bgm_mh_adapt_qsd(const int* accept_count_dev, int t, int window, long long n_total, double target,
                     double tolerance, double* q_sd_dev, void* stream) {
  if (!accept_count_dev || !q_sd_dev) return fail(BGM_ERR_ARG, "bgm_mh_adapt_qsd: null pointer");
  if (t < 0 || window < 1 || n_total < 1) return fail(BGM_ERR_ARG, "bgm_mh_adapt_qsd: bad t / window / n");
  mh_adapt_qsd_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(accept_count_dev, t, window, n_total, target,
                                                        tolerance, q_sd_dev);
  BGM_CUDA_OK(cudaGetLastError());
  return 0;
}

int bgm_mh_noise(uint64_t seed, int64_t row_offset, int n, int zd, int t_begin, int t_end,
                 float* z0_dev, float* eps_dev, double* u_dev, void* stream) {
  if (n < 1 || zd < 1 || t_end < t_begin) return fail(BGM_ERR_ARG, "bgm_mh_noise: bad sizes");
  const long long total = (long long)(t_end - t_begin + 1) * n;
  const int grid = (int)std::min<long long>((total + 255) / 256, 148 * 16);
  mh_noise_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(seed, row_offset, n, zd, t_begin, t_end,
                                                         z0_dev, eps_dev, u_dev);
  BGM_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---- memoised effect evaluation (causal.cuh) ----
int bgm_causal_effect_index(const float* z_samples_dev, int n_keep, int n, int zd, int* local_dev, int* rowtot_dev,
                            int* rowend_dev, int* scratch_dev, void* stream) {
  if (!z_samples_dev || !local_dev || !rowtot_dev || !rowend_dev || !scratch_dev || n_keep < 1 || n < 1 || zd < 1 || zd > 32)
    return fail(BGM_ERR_ARG, "bgm_causal_effect_index: bad argument");
  if ((long long)n_keep * n > 0x7fffffffLL) return fail(BGM_ERR_UNSUPPORTED, "bgm_causal_effect_index: n_keep * n must be < 2^31");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = std::max(1, std::min((n + 127) / 128, 148 * 8));
  if (zd <= 8) effect_rowscan_kernel<8><<<grid, 128, 0, st>>>(z_samples_dev, n_keep, n, zd, local_dev, rowtot_dev);
  else if (zd <= 16) effect_rowscan_kernel<16><<<grid, 128, 0, st>>>(z_samples_dev, n_keep, n, zd, local_dev, rowtot_dev);
  else effect_rowscan_kernel<32><<<grid, 128, 0, st>>>(z_samples_dev, n_keep, n, zd, local_dev, rowtot_dev);
  const int nblocks = (n + SCAN_TILE - 1) / SCAN_TILE;
  scan_block_kernel<<<nblocks, 256, 0, st>>>(rowtot_dev, rowend_dev, n, scratch_dev);
  scan_totals_kernel<<<1, 1024, 0, st>>>(scratch_dev, nblocks);
  scan_add_kernel<<<nblocks, 256, 0, st>>>(rowend_dev, n, scratch_dev);
  BGM_CUDA_OK(cudaGetLastError());
  return 0;
}

int bgm_causal_effect_compact(const float* z_samples_dev, int n_keep, int n, int zd, const int* local_dev,
                              const int* rowend_dev, float* zlist_dev, void* stream) {
  if (!z_samples_dev || !local_dev || !rowend_dev || !zlist_dev) return fail(BGM_ERR_ARG, "bgm_causal_effect_compact: null pointer");
  const int grid = std::max(1, std::min((n + 127) / 128, 148 * 8));
  effect_compact_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(z_samples_dev, n_keep, n, zd, local_dev, rowend_dev,
                                                               zlist_dev);
  BGM_CUDA_OK(cudaGetLastError());
  return 0;
}

int bgm_causal_effect_combine(const bgm_causal* m, const float* heads_dev, const int* local_dev, const int* rowend_dev,
                              int n_keep, int n, int n_x, int sample_y, uint64_t seed, int64_t row_offset,
                              const float* noise_dev, double* adrf_sum_dev, float* ite_dev, void* stream) {
  if (!m || !heads_dev || !local_dev || !rowend_dev || n_keep < 1 || n < 1)
    return fail(BGM_ERR_ARG, "bgm_causal_effect_combine: bad argument");
  CombineDev C;
  memset(&C, 0, sizeof(C));
  C.heads = heads_dev; C.local = local_dev; C.rowend = rowend_dev; C.n_keep = n_keep; C.n = n;
  C.binary = m->prog.binary; C.n_x = C.binary ? 2 : n_x;
  if (C.binary ? !ite_dev : (!adrf_sum_dev || n_x < 1))
    return fail(BGM_ERR_ARG, "bgm_causal_effect_combine: binary needs ite_dev, continuous adrf_sum_dev and n_x >= 1");
  C.sample_y = sample_y ? 1 : 0; C.s2y = m->prog.s2y; C.seed = seed; C.row_offset = row_offset; C.noise = noise_dev;
  C.adrf_sum = adrf_sum_dev; C.ite = ite_dev;
  if (n >= 4096) {
    // row-major walk: enough rows to fill the GPU with (row block, dose group) CTAs
    const long long blocks = (long long)((n + 127) / 128) * ((C.n_x + 3) / 4);
    effect_combine_rows_kernel<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(C);
    BGM_CUDA_OK(cudaGetLastError());
    return 0;
  }
  const long long ntiles = (long long)((n + 31) / 32) * n_keep;
  const int grid = (int)std::max<long long>(1, std::min<long long>((ntiles + 7) / 8, (long long)m->sm_count * 16));
  effect_combine_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(C);
  BGM_CUDA_OK(cudaGetLastError());
  return 0;
}

static int effect_launch(const bgm_causal* m, const float* z_samples_dev, int n_keep, int n,
                         const float* x_values_dev, int n_x, int sample_y, uint64_t seed, int64_t row_offset,
                         const float* noise_dev, double* adrf_sum_dev, float* ite_dev, float* heads_dev, void* stream);

int bgm_causal_effect(const bgm_causal* m, const float* z_samples_dev, int n_keep, int n,
                      const float* x_values_dev, int n_x, int sample_y, uint64_t seed,
                      int64_t row_offset, const float* noise_dev, double* adrf_sum_dev, float* ite_dev,
                      void* stream) {
  return effect_launch(m, z_samples_dev, n_keep, n, x_values_dev, n_x, sample_y, seed, row_offset, noise_dev,
                       adrf_sum_dev, ite_dev, nullptr, stream);
}

int bgm_causal_effect_heads(const bgm_causal* m, const float* states_dev, int n_states, const float* x_values_dev,
                            int n_x, float* heads_dev, void* stream) {
  if (!heads_dev) return fail(BGM_ERR_ARG, "bgm_causal_effect_heads: null output");
  return effect_launch(m, states_dev, 1, n_states, x_values_dev, n_x, 0, 0, 0, nullptr, nullptr, nullptr, heads_dev,
                       stream);
}

static int effect_launch(const bgm_causal* m, const float* z_samples_dev, int n_keep, int n,
                         const float* x_values_dev, int n_x, int sample_y, uint64_t seed, int64_t row_offset,
                         const float* noise_dev, double* adrf_sum_dev, float* ite_dev, float* heads_dev, void* stream) {
  if (!m || !z_samples_dev) return fail(BGM_ERR_ARG, "bgm_causal_effect: null model / samples");
  if (n_keep < 1 || n < 1) return fail(BGM_ERR_ARG, "bgm_causal_effect: n_keep and n must be >= 1");
  EffectDev E;
  memset(&E, 0, sizeof(E));
  E.heads = heads_dev;
  if (m->prog.binary) {
    if (!ite_dev && !heads_dev) return fail(BGM_ERR_ARG, "bgm_causal_effect: binary treatment needs ite_dev");
    E.x_values = nullptr;
    E.n_x = 2;
  } else {
    if (!x_values_dev || n_x < 1 || (!adrf_sum_dev && !heads_dev))
      return fail(BGM_ERR_ARG, "bgm_causal_effect: continuous treatment needs x_values_dev, n_x >= 1 and adrf_sum_dev");
    E.x_values = x_values_dev;
    E.n_x = n_x;
  }
  E.z_samples = z_samples_dev; E.n_keep = n_keep; E.n = n; E.sample_y = sample_y ? 1 : 0;
  E.seed = seed; E.row_offset = row_offset; E.noise = noise_dev; E.adrf_sum = adrf_sum_dev; E.ite = ite_dev;
  if (use_tc(m)) {
    const long long nt = (long long)((n + TC_ROWS - 1) / TC_ROWS) * n_keep;
    const int grid = (int)std::max<long long>(1, std::min<long long>((nt + 1) / 2, m->sm_count));
    const int zd = m->prog.zd;
    auto launch = [&](auto kern) -> int {
      BGM_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, m->tc_smem_bytes));
      kern<<<grid, 512, m->tc_smem_bytes, (cudaStream_t)stream>>>(m->tc, m->tc_image_dev, E);
      BGM_CUDA_OK(cudaGetLastError());
      return 0;
    };
    if (zd <= 8) return launch(causal_effect_tc_kernel<8>);
    if (zd <= 16) return launch(causal_effect_tc_kernel<16>);
    return launch(causal_effect_tc_kernel<32>);
  }
  const long long ntiles = (long long)((n + TILE_ROWS - 1) / TILE_ROWS) * n_keep;
  const int grid = (int)std::max<long long>(1, std::min<long long>((ntiles + m->effect_warps - 1) / m->effect_warps, m->sm_count));
  BGM_CUDA_OK(cudaFuncSetAttribute(causal_effect_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   m->effect_smem_bytes));
  causal_effect_kernel<<<grid, m->effect_warps * 32, m->effect_smem_bytes, (cudaStream_t)stream>>>(
      m->prog, m->image_dev, E);
  BGM_CUDA_OK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------- fp32 peak ----
}  // extern "C"

namespace bgm {
// 16 independent FFMA chains per thread, operands in registers: the issue-bound
// ceiling of the fp32 pipe that the sampler's inner loop runs on.
__global__ void __launch_bounds__(256) fp32_peak_kernel(float* out, int iters, float a, float b) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  if (s == 123.456f) out[0] = s;  // never true; keeps the chain alive
}
}  // namespace bgm

extern "C" int bgm_fp32_peak_tflops(double* tflops, void* stream) {
  if (!tflops) return fail(BGM_ERR_ARG, "bgm_fp32_peak_tflops: null output");
  cudaStream_t st = (cudaStream_t)stream;
  int dev = 0, sms = 0;
  BGM_CUDA_OK(cudaGetDevice(&dev));
  BGM_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  float* out = nullptr;
  BGM_CUDA_OK(cudaMalloc(&out, 4));
  cudaEvent_t e0, e1;
  BGM_CUDA_OK(cudaEventCreate(&e0));
  BGM_CUDA_OK(cudaEventCreate(&e1));
  const int iters = 4096, grid = sms * 8, block = 256;
  double best = 0.0;
  for (int rep = 0; rep < 6; ++rep) {
    BGM_CUDA_OK(cudaEventRecord(e0, st));
    bgm::fp32_peak_kernel<<<grid, block, 0, st>>>(out, iters, 0.999f, 1e-3f);
    BGM_CUDA_OK(cudaEventRecord(e1, st));
    BGM_CUDA_OK(cudaEventSynchronize(e1));
    float ms = 0.f;
    BGM_CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
    const double flop = 2.0 * 16 * 8 * (double)iters * grid * block;
    if (rep > 0) best = std::max(best, flop / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  *tflops = best;
  return 0;
}

// hmc_api.cu / train_api.cu / bnn.cu / host_rng.cu are separate translation units (bayesgm_b200/_build.py)
