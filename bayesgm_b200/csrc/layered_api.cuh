// Host side of the layered training engine (include/bgm_b200.h, bgm_lt_*): strings the kernels of
// layered.cuh into the training steps of CausalBGM -- train_disc_step / train_gen_step
// (causalbgm/base.py:305-377), update_g/h/f_net (:156-243), update_latent_variable_sgd (:246-302) and
// evaluate (:534-556) -- for Bayesian nets (DenseFlipout + batch-statistics BatchNormalization,
// networks/bnn.py) or deterministic nets, any batch size, any layer width.
// Included by train_api.cu (shares train.cuh's discriminator kernel).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <initializer_list>
#include <string>
#include <vector>

#include "layered.cuh"
#include "disc_big.cuh"

namespace bgm {
namespace lt {

constexpr int LT_MAXL = 8;

struct LNet {
  int L = 0;
  int dims[LT_MAXL + 1];
  int bayes = 0, net_id = 0;                   // bayes = bn_in && flip (BayesianFullyConnectedNet)
  int bn_in = 0, flip = 0;                      // input BatchNormalization / DenseFlipout layers
  int base = 0, n_params = 0;                  // slice of the group-0 flat vector
  int off_gamma = -1, off_beta = -1;           // absolute offsets (Bayesian nets)
  int off_w[LT_MAXL], off_rho[LT_MAXL], off_b[LT_MAXL];
};

struct Pass {                                   // what one forward call leaves behind for its backward
  const LNet* net = nullptr;
  int B = 0;
  float *mean = nullptr, *inv = nullptr, *xhat = nullptr;
  float* a[LT_MAXL + 1];                        // a[l] = input of layer l; a[L] = output
  int lda[LT_MAXL + 1];
  float* dW[LT_MAXL];
  signed char *sin[LT_MAXL], *sout[LT_MAXL];
  float* out() const { return a[net->L]; }
  int ldo() const { return net->dims[net->L]; }
};

struct DiscPass {
  float* pre[LT_MAXL];                          // Dense outputs
  float *xhat[LT_MAXL], *out[LT_MAXL], *mean[LT_MAXL], *inv[LT_MAXL];
  const float* in = nullptr;
  float* d = nullptr;                           // (B, 1)
};

struct Arena;
struct Ctx {                                    // what the generic passes need from a trainer
  Arena* ar;
  const float* th0;                             // parameter group 0
  float* g0;                                    // its gradient buffer
  int sm;
  uint64_t seed;
  const uint32_t* ctr_dev = nullptr;            // device-resident call counter (graph replay), see StepScalars
};

struct Arena {
  char* base = nullptr;
  size_t cap = 0, used = 0;
  size_t gen = 0;                               // bumped when the block moves: captured graphs hold its addresses
  bool overflow = false;
  template <class T> T* get(size_t n) {
    const size_t bytes = (n * sizeof(T) + 255) / 256 * 256;
    if (used + bytes > cap) { overflow = true; return reinterpret_cast<T*>(base); }
    T* p = reinterpret_cast<T*>(base + used);
    used += bytes;
    return p;
  }
};

}  // namespace lt
}  // namespace bgm

struct bgm_lt {
  bgm::lt::LNet g, e, f, h;
  bgm::tr::Disc dz;
  int z_dims[4];
  int zd = 0, p = 0, binary = 0, bayes = 0;
  float use_z_rec = 1.f;
  int n0 = 0, n1 = 0;
  float *theta[2] = {nullptr, nullptr}, *grad[2] = {nullptr, nullptr};
  float *m_pre[2] = {nullptr, nullptr}, *v_pre[2] = {nullptr, nullptr};
  long long step_pre[2] = {0, 0};
  float *m_it = nullptr, *v_it = nullptr;
  long long step_it[3] = {0, 0, 0}, step_z = 0;
  double lr = 0, b1 = 0.9, b2 = 0.99, lr_theta = 1e-4, lr_z = 1e-4;
  float s2v = -1.f, s2x = -1.f, s2y = -1.f;
  float kl_weight = 0.f;
  bgm::lt::Arena arena;
  float* scratch = nullptr;      // 64 floats of loss accumulators
  uint64_t seed = 0;
  uint32_t call_ctr = 0;
  int sm_count = 148, smem_disc = 0, wm_disc = 4, stage_disc = 0;
  // ---- CUDA-graph replay of a step (run_step) ----
  struct GraphEntry {
    int fn, bs, flag;
    const void* ptr[10];
    long long n;
    size_t arena_gen;
    cudaGraphExec_t exec;
  };
  std::vector<GraphEntry> graphs;
  cudaStream_t cap_stream = nullptr;       // capture only, never executes
  bgm::lt::StepScalars* sc_dev = nullptr;  // device copy of the step's changing scalars
  // mini-batch inputs (z, v, x, y rows or int32 row indices) are copied here and losses come back from here, so
  // that a replayed graph always reads and writes the same addresses whatever buffers the caller passes
  float *z_stage = nullptr, *v_stage = nullptr, *x_stage = nullptr, *y_stage = nullptr, *loss_stage = nullptr;
  int* idx_stage = nullptr;
  int stage_rows = 0;
  int use_graphs = 1;                      // BGM_LT_GRAPHS=0 in the environment: plain launches
};

// BGM flavour (bgm/base.py:145-291): generator = BaseVariationalNet (input BatchNormalization in training mode +
// Dense stack + [mean | variance] heads as ONE final layer of width 2 x_dim), encoder e_net, discriminators dz / dx.
// Same device parameter layout as the fused BGM trainer (bgm_bgmtrainer_create).
struct bgm_ltb {
  bgm::lt::LNet g, e;
  bgm::tr::Disc dz, dx;
  int dx_base = 0;
  int zd = 0, xd = 0, n_g = 0, n0 = 0, n1 = 0;
  float *theta[2] = {nullptr, nullptr}, *grad[2] = {nullptr, nullptr};
  float *m_pre[2] = {nullptr, nullptr}, *v_pre[2] = {nullptr, nullptr};
  long long step_pre[2] = {0, 0};
  float *m_it = nullptr, *v_it = nullptr;
  long long step_g = 0, step_z = 0;
  double lr = 0, b1 = 0.5, b2 = 0.9, lr_theta = 5e-3, lr_z = 5e-3;
  float alpha = 0.f, gamma = 0.f;
  float* moving = nullptr;       // [2 zd] moving mean | variance of the generator's input BatchNormalization
  bgm::lt::Arena arena;
  float* scratch = nullptr;
  int sm_count = 148;
};

namespace bgm {
namespace lt {

static inline int grid_for(long long total, int sm) {
  return (int)std::max<long long>(1, std::min<long long>((total + 255) / 256, (long long)sm * 8));
}

// tiled launches of the two dense kernels (layered.cuh, dense_tile_kernel)
static inline void dense_fwd(const float* A, int lda, const float* W, const float* dW, const signed char* s_in,
                             const signed char* s_out, const float* bias, int B, int K, int N, float* out, int ldo, int act,
                             cudaStream_t st) {
  DenseTileArgs a;
  a.X = A; a.ldx = lda; a.sX = s_in; a.W = W; a.dW = dW; a.sO = s_out; a.bias = bias; a.A_post = nullptr; a.lda = 0;
  a.O = out; a.ldo = ldo; a.B = B; a.K = K; a.N = N; a.act = act; a.accumulate = 0;
  dense_tile_kernel<false><<<dim3((N + 31) / 32, (B + 31) / 32), 256, 0, st>>>(a);
}
static inline void dense_bwd_input(const float* dY, int ldy, const float* W, const float* dW, const signed char* s_in,
                                   const signed char* s_out, const float* A_post, int lda, int B, int K, int N, float* dA,
                                   int ldd, int accumulate, cudaStream_t st) {
  DenseTileArgs a;
  a.X = dY; a.ldx = ldy; a.sX = s_out; a.W = W; a.dW = dW; a.sO = s_in; a.bias = nullptr; a.A_post = A_post; a.lda = lda;
  a.O = dA; a.ldo = ldd; a.B = B; a.K = K; a.N = N; a.act = 0; a.accumulate = accumulate;
  dense_tile_kernel<true><<<dim3((K + 31) / 32, (B + 31) / 32), 256, 0, st>>>(a);
}

static size_t pass_bytes(const LNet& n, long long B) {
  size_t b = 0;
  auto al = [](size_t x) { return (x + 255) / 256 * 256 + 256; };
  if (n.bn_in) b += 2 * al(4 * n.dims[0]) + 2 * al(4 * B * n.dims[0]) + 2 * al(4 * n.dims[0]);
  for (int l = 0; l < n.L; ++l) {
    const size_t K = n.dims[l], N = n.dims[l + 1];
    b += al(4 * B * N) + al(4 * B * K);                       // output, input gradient
    if (n.flip) b += al(4 * K * N) + al(B * K) + al(B * N);
  }
  b += al(4 * B * n.dims[n.L]);                              // output gradient
  return b;
}
static size_t disc_bytes(const tr::Disc& dz, long long B) {
  size_t b = 0;
  for (int l = 0; l <= dz.L; ++l) b += 8 * (size_t)(4 * B * std::max(dz.dims[l], dz.dims[l + 1]) + 512);
  return b;
}
static size_t step_bytes(const bgm_lt* t, long long B) {
  size_t b = 3 * pass_bytes(t->g, B) + 2 * pass_bytes(t->e, B) + 2 * pass_bytes(t->f, B) + 2 * pass_bytes(t->h, B);
  b += disc_bytes(t->dz, B) + 16 * (size_t)(4 * B * (t->p + 1 + t->zd + 8) + 512);
  if (B > 32) b += 4 * tr::disc_big_floats(t->dz, (int)B) + 1024;     // workspace of disc_grad_big_kernel
  return b + (1 << 16);
}
static int ensure_arena_bytes(Arena& ar, size_t need) {
  if (need > ar.cap) {
    BGM_CUDA_OK(cudaDeviceSynchronize());
    if (ar.base) cudaFree(ar.base);
    ar.base = nullptr;
    ar.cap = 0;
    BGM_CUDA_OK(cudaMalloc(&ar.base, need));
    ar.cap = need;
    ar.gen += 1;
  }
  ar.used = 0;
  ar.overflow = false;
  return 0;
}
static int ensure_arena(bgm_lt* t, long long B) { return ensure_arena_bytes(t->arena, step_bytes(t, B)); }
static Ctx ctx_of(bgm_lt* t) { return Ctx{&t->arena, t->theta[0], t->grad[0], t->sm_count, t->seed}; }
static int arena_ok(Arena& ar, const char* fn) {
  if (ar.overflow) return fail(BGM_ERR_NOMEM, std::string(fn) + ": workspace estimate too small (internal error)");
  BGM_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---- CUDA-graph replay.  A layered step is 15-250 small launches whose host cost (2-4 us each) rivals their GPU
// time at batch 32; its launch sequence depends only on (function, batch size, argument addresses), so it is
// captured once on a private stream and replayed.  What changes from step to step -- the noise call counter, the
// Adam bias corrections, the WGAN-GP epsilon -- lives in device memory (StepScalars, written by one 1-thread
// kernel in front of the graph); mini-batch rows / indices that arrive at changing addresses are copied to a
// fixed staging buffer first.  `body(stream, dev)` issues the step's launches; dev = true: read the scalars from
// t->sc_dev.
static inline uint32_t __float_as_uint_host(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static void drop_graphs(bgm_lt* t) {
  for (auto& g : t->graphs) cudaGraphExecDestroy(g.exec);
  t->graphs.clear();
}
static int ensure_stage(bgm_lt* t, int rows) {
  if (rows <= t->stage_rows) return 0;
  BGM_CUDA_OK(cudaDeviceSynchronize());
  drop_graphs(t);
  if (t->z_stage) cudaFree(t->z_stage);
  if (t->idx_stage) cudaFree(t->idx_stage);
  t->z_stage = nullptr; t->idx_stage = nullptr; t->stage_rows = 0;
  // one block: z (rows, zd) | v (rows, p) | x (rows) | y (rows) | losses (16)
  const size_t fl = (size_t)rows * (t->zd + t->p + 2) + 16;
  BGM_CUDA_OK(cudaMalloc(&t->z_stage, sizeof(float) * fl));
  BGM_CUDA_OK(cudaMalloc(&t->idx_stage, sizeof(int) * (size_t)rows));
  t->v_stage = t->z_stage + (size_t)rows * t->zd;
  t->x_stage = t->v_stage + (size_t)rows * t->p;
  t->y_stage = t->x_stage + rows;
  t->loss_stage = t->y_stage + rows;
  t->stage_rows = rows;
  return 0;
}
template <class Body>
static int run_step(bgm_lt* t, int fn, int bs, int flag, long long n, std::initializer_list<const void*> ptrs,
                    const StepScalars& sc, cudaStream_t st, Body body) {
  if (!t->use_graphs) return body(st, false);
  if (!t->sc_dev) BGM_CUDA_OK(cudaMalloc(&t->sc_dev, sizeof(StepScalars)));
  if (!t->cap_stream) BGM_CUDA_OK(cudaStreamCreateWithFlags(&t->cap_stream, cudaStreamNonBlocking));
  set_scalars_kernel<<<1, 1, 0, st>>>(t->sc_dev, sc);
  bgm_lt::GraphEntry key;
  memset(&key, 0, sizeof(key));
  key.fn = fn; key.bs = bs; key.flag = flag; key.n = n; key.arena_gen = t->arena.gen;
  int i = 0;
  for (const void* q : ptrs) key.ptr[i++] = q;
  for (auto& g : t->graphs)
    if (g.fn == key.fn && g.bs == key.bs && g.flag == key.flag && g.n == key.n && g.arena_gen == key.arena_gen &&
        memcmp(g.ptr, key.ptr, sizeof(key.ptr)) == 0) {
      t->arena.used = 0;
      BGM_CUDA_OK(cudaGraphLaunch(g.exec, st));
      return 0;
    }
  if (t->graphs.size() >= 48) drop_graphs(t);
  BGM_CUDA_OK(cudaStreamBeginCapture(t->cap_stream, cudaStreamCaptureModeThreadLocal));
  const int rc = body(t->cap_stream, true);
  cudaGraph_t graph = nullptr;
  const cudaError_t e = cudaStreamEndCapture(t->cap_stream, &graph);
  if (rc) {
    if (graph) cudaGraphDestroy(graph);
    return rc;
  }
  if (e != cudaSuccess) return fail(BGM_ERR_CUDA, std::string("layered step capture: ") + cudaGetErrorString(e));
  const cudaError_t e2 = cudaGraphInstantiate(&key.exec, graph, 0ull);
  cudaGraphDestroy(graph);
  if (e2 != cudaSuccess) return fail(BGM_ERR_CUDA, std::string("layered step graph: ") + cudaGetErrorString(e2));
  t->graphs.push_back(key);
  BGM_CUDA_OK(cudaGraphLaunch(key.exec, st));
  return 0;
}

// ---- one forward call of a net on B rows.  ext_mean / ext_inv: batch statistics computed elsewhere (chunked
// evaluation: the statistics are those of the WHOLE batch, the chunk only holds some of its rows). ----
static void net_fwd(const Ctx& C, const LNet& net, Pass& P, const float* X, int ldx, int B, uint32_t call, int64_t row0,
                    int const_col, const float* ext_mean, const float* ext_inv, cudaStream_t st) {
  Arena& ar = *C.ar;
  const float* th = C.th0;
  const int sm = C.sm;
  P.net = &net;
  P.B = B;
  const int K0 = net.dims[0];
  if (net.bn_in) {
    P.xhat = ar.get<float>((size_t)B * K0);
    P.a[0] = ar.get<float>((size_t)B * K0);
    P.lda[0] = K0;
    if (ext_mean) {
      P.mean = const_cast<float*>(ext_mean);
      P.inv = const_cast<float*>(ext_inv);
    } else {
      P.mean = ar.get<float>(K0);
      P.inv = ar.get<float>(K0);
      col_stats_kernel<<<K0, 256, 0, st>>>(X, ldx, B, K0, P.mean, P.inv);
    }
    bn_fwd_kernel<<<grid_for((long long)B * K0, sm), 256, 0, st>>>(X, ldx, P.mean, P.inv, th + net.off_gamma,
                                                                  th + net.off_beta, B, K0, P.xhat, P.a[0], 0, const_col);
  } else {
    P.a[0] = const_cast<float*>(X);
    P.lda[0] = ldx;
  }
  for (int l = 0; l < net.L; ++l) {
    const int K = net.dims[l], N = net.dims[l + 1];
    P.dW[l] = nullptr; P.sin[l] = nullptr; P.sout[l] = nullptr;
    if (net.flip) {
      P.dW[l] = ar.get<float>((size_t)K * N);
      P.sin[l] = ar.get<signed char>((size_t)B * K);
      P.sout[l] = ar.get<signed char>((size_t)B * N);
      const long long work = std::max<long long>((long long)K * N / 4, (long long)B * ((K + N + 31) / 32));
      flipout_noise_kernel<<<grid_for(work, sm), 256, 0, st>>>(th + net.off_rho[l], K, N, C.seed, 0, net.net_id, l, call,
                                                              row0, B, P.dW[l], P.sin[l], P.sout[l], C.ctr_dev);
    }
    P.a[l + 1] = ar.get<float>((size_t)B * N);
    P.lda[l + 1] = N;
    dense_fwd(P.a[l], P.lda[l], th + net.off_w[l], P.dW[l], P.sin[l],
                                                                    P.sout[l], th + net.off_b[l], B, K, N, P.a[l + 1], N,
                                                                    l < net.L - 1 ? 1 : 0, st);
  }
}

// ---- backward of one call: parameter gradients (+= into grad[0]) and / or the gradient w.r.t. the net's input ----
static void net_bwd(const Ctx& C, const LNet& net, const Pass& P, const float* dOut, bool param_grads, float* dX, int lddx,
                    bool accumulate_dx, cudaStream_t st) {
  Arena& ar = *C.ar;
  const float* th = C.th0;
  float* g = C.g0;
  const int sm = C.sm, B = P.B;
  const float* dY = dOut;
  for (int l = net.L - 1; l >= 0; --l) {
    const int K = net.dims[l], N = net.dims[l + 1];
    if (param_grads)
      dense_bwd_param_kernel<<<grid_for((long long)K * N, sm), 256, 0, st>>>(
          P.a[l], P.lda[l], dY, N, net.flip ? th + net.off_rho[l] : nullptr, P.dW[l], P.sin[l], P.sout[l], B, K, N,
          g + net.off_w[l], net.flip ? g + net.off_rho[l] : nullptr, g + net.off_b[l]);
    if (l > 0) {
      float* dA = ar.get<float>((size_t)B * K);
      dense_bwd_input(dY, N, th + net.off_w[l], P.dW[l], P.sin[l],
                                                                            P.sout[l], P.a[l], P.lda[l], B, K, N, dA, K, 0, st);
      dY = dA;
      continue;
    }
    if (net.bn_in) {
      if (!param_grads && !dX) break;
      float* dA0 = ar.get<float>((size_t)B * K);
      dense_bwd_input(dY, N, th + net.off_w[0], P.dW[0], P.sin[0],
                                                                            P.sout[0], nullptr, 0, B, K, N, dA0, K, 0, st);
      float* s1 = ar.get<float>(K);
      float* s2 = ar.get<float>(K);
      bn_bwd_sums_kernel<<<(K + 127) / 128, 128, 0, st>>>(dA0, nullptr, P.xhat, B, K, 0, s1, s2,
                                                          param_grads ? g + net.off_gamma : nullptr,
                                                          param_grads ? g + net.off_beta : nullptr);
      if (dX)
        bn_bwd_input_kernel<<<grid_for((long long)B * K, sm), 256, 0, st>>>(dA0, nullptr, P.xhat, P.inv, th + net.off_gamma,
                                                                           s1, s2, B, K, 0, dX, lddx, accumulate_dx ? 1 : 0);
    } else if (dX) {
      dense_bwd_input(dY, N, th + net.off_w[0], nullptr, nullptr,
                                                                            nullptr, nullptr, 0, B, K, N, dX, lddx,
                                                                            accumulate_dx ? 1 : 0, st);
    }
  }
}

// ---- Discriminator (networks/base.py:338-385): Dense -> BN(batch statistics) -> tanh blocks, Dense(1) ----
static void disc_fwd(const Ctx& C, const tr::Disc& dz, const float* th, DiscPass& D, const float* Z, int B, cudaStream_t st) {
  Arena& ar = *C.ar;
  const int sm = C.sm;
  D.in = Z;
  const float* a = Z;
  for (int l = 0; l < dz.L; ++l) {
    const int K = dz.dims[l], N = dz.dims[l + 1];
    D.pre[l] = ar.get<float>((size_t)B * N);
    D.xhat[l] = ar.get<float>((size_t)B * N);
    D.out[l] = ar.get<float>((size_t)B * N);
    D.mean[l] = ar.get<float>(N);
    D.inv[l] = ar.get<float>(N);
    dense_fwd(a, K, th + dz.w_off[l], nullptr, nullptr, nullptr,
                                                                    th + dz.b_off[l], B, K, N, D.pre[l], N, 0, st);
    col_stats_kernel<<<N, 256, 0, st>>>(D.pre[l], N, B, N, D.mean[l], D.inv[l]);
    bn_fwd_kernel<<<grid_for((long long)B * N, sm), 256, 0, st>>>(D.pre[l], N, D.mean[l], D.inv[l], th + dz.g_off[l],
                                                                 th + dz.be_off[l], B, N, D.xhat[l], D.out[l], 2, -1);
    a = D.out[l];
  }
  const int K = dz.dims[dz.L];
  D.d = ar.get<float>(B);
  dense_fwd(a, K, th + dz.w_off[dz.L], nullptr, nullptr, nullptr,
                                                   th + dz.b_off[dz.L], B, K, 1, D.d, 1, 0, st);
}
// Backward of one discriminator call from dd (B) = d loss / d D: parameter gradients (+= into g, Keras order of
// tr::Disc) when g != NULL, and the gradient w.r.t. the input added into dZ (B, dims[0]) when dZ != NULL.
static void disc_bwd(const Ctx& C, const tr::Disc& dz, const float* th, float* g, const DiscPass& D, const float* dd, int B,
                     float* dZ, cudaStream_t st) {
  Arena& ar = *C.ar;
  const int sm = C.sm;
  const float* dY = dd;
  int N = 1;
  for (int l = dz.L; l >= 0; --l) {
    const int K = dz.dims[l];
    if (l < dz.L) {
      // dY is the gradient w.r.t. out[l] (B, N = dims[l+1]): through tanh and BatchNorm to pre[l]
      float* s1 = ar.get<float>(N);
      float* s2 = ar.get<float>(N);
      float* dPre = ar.get<float>((size_t)B * N);
      bn_bwd_sums_kernel<<<(N + 127) / 128, 128, 0, st>>>(dY, D.out[l], D.xhat[l], B, N, 2, s1, s2,
                                                          g ? g + dz.g_off[l] : nullptr, g ? g + dz.be_off[l] : nullptr);
      bn_bwd_input_kernel<<<grid_for((long long)B * N, sm), 256, 0, st>>>(dY, D.out[l], D.xhat[l], D.inv[l], th + dz.g_off[l],
                                                                         s1, s2, B, N, 2, dPre, N, 0);
      dY = dPre;
    }
    const float* A_in = l == 0 ? D.in : D.out[l - 1];
    if (g)
      dense_bwd_param_kernel<<<grid_for((long long)K * N, sm), 256, 0, st>>>(A_in, K, dY, N, nullptr, nullptr, nullptr, nullptr, B,
                                                                            K, N, g + dz.w_off[l], nullptr, g + dz.b_off[l]);
    if (l == 0) {
      if (dZ)
        dense_bwd_input(dY, N, th + dz.w_off[0], nullptr, nullptr, nullptr,
                                                                              nullptr, 0, B, K, N, dZ, K, 1, st);
    } else {
      float* dA = ar.get<float>((size_t)B * K);
      dense_bwd_input(dY, N, th + dz.w_off[l], nullptr, nullptr, nullptr,
                                                                            nullptr, 0, B, K, N, dA, K, 0, st);
      dY = dA;
      N = K;
    }
  }
}

static int fill_net(const bgm_bnn_net_desc* d, LNet& n, int net_id, int bayes, int& off, int want_in, int want_out,
                    const char* name) {
  if (!d || !d->dims || !d->params || d->n_layers < 1 || d->n_layers > LT_MAXL)
    return fail(BGM_ERR_ARG, std::string(name) + ": need 1..8 layers");
  if (bayes && !d->bn) return fail(BGM_ERR_ARG, std::string(name) + ": Bayesian net without BatchNormalization parameters");
  if (d->dims[0] != want_in || d->dims[d->n_layers] != want_out)
    return fail(BGM_ERR_ARG, std::string(name) + ": input / output width does not match z_dims / v_dim");
  n.L = d->n_layers;
  n.bayes = bayes;
  n.bn_in = bayes;
  n.flip = bayes;
  n.net_id = net_id;
  n.base = off;
  for (int l = 0; l <= n.L; ++l) {
    n.dims[l] = d->dims[l];
    if (n.dims[l] < 1) return fail(BGM_ERR_ARG, std::string(name) + ": non-positive layer size");
  }
  if (bayes) {
    n.off_gamma = off; off += n.dims[0];
    n.off_beta = off; off += n.dims[0];
  }
  for (int l = 0; l < n.L; ++l) {
    const int K = n.dims[l], N = n.dims[l + 1];
    n.off_w[l] = off; off += K * N;
    n.off_rho[l] = -1;
    if (bayes) { n.off_rho[l] = off; off += K * N; }
    n.off_b[l] = off; off += N;
  }
  n.n_params = off - n.base;
  return 0;
}
static void pack_net(const bgm_bnn_net_desc* d, const LNet& n, std::vector<float>& host) {
  if (n.bayes) memcpy(host.data() + n.off_gamma, d->bn, sizeof(float) * 2 * n.dims[0]);
  // descriptor order == device order: per layer kernel (loc), [rho], bias
  const int body = n.n_params - (n.bayes ? 2 * n.dims[0] : 0);
  memcpy(host.data() + n.off_w[0], d->params, sizeof(float) * body);
}

static float lr_t_of(double lr, double b1, double b2, long long k) {
  return (float)(lr * std::sqrt(1.0 - std::pow(b2, (double)k)) / (1.0 - std::pow(b1, (double)k)));
}

// f input [z0, z1, x] and h input [z0, z2] from a (B, zd) latent block and x (B)
static void build_fh_inputs(bgm_lt* t, const float* Z, int ldz, const float* x, int B, float* fin, float* hin, cudaStream_t st) {
  const int d0 = t->z_dims[0], d1 = t->z_dims[1], d2 = t->z_dims[2], sm = t->sm_count;
  const int kf = d0 + d1 + 1, kh = d0 + d2;
  if (d0 + d1 > 0) gather_cols_kernel<<<grid_for((long long)B * (d0 + d1), sm), 256, 0, st>>>(Z, ldz, nullptr, 0, d0 + d1, B, fin, kf, 0);
  gather_cols_kernel<<<grid_for(B, sm), 256, 0, st>>>(x, 1, nullptr, 0, 1, B, fin, kf, d0 + d1);
  if (d0 > 0) gather_cols_kernel<<<grid_for((long long)B * d0, sm), 256, 0, st>>>(Z, ldz, nullptr, 0, d0, B, hin, kh, 0);
  if (d2 > 0) gather_cols_kernel<<<grid_for((long long)B * d2, sm), 256, 0, st>>>(Z, ldz, nullptr, d0 + d1, d2, B, hin, kh, d0);
}
// gradients w.r.t. the f / h inputs back onto the latent columns (the x column's gradient is dropped)
static void scatter_fh_grads(bgm_lt* t, const float* dFin, const float* dHin, int B, float* dZ, cudaStream_t st) {
  const int d0 = t->z_dims[0], d1 = t->z_dims[1], d2 = t->z_dims[2], sm = t->sm_count, zd = t->zd;
  const int kf = d0 + d1 + 1, kh = d0 + d2;
  if (dFin && d0 + d1 > 0) add_cols_kernel<<<grid_for((long long)B * (d0 + d1), sm), 256, 0, st>>>(dFin, kf, 0, d0 + d1, B, dZ, zd, 0);
  if (dHin && d0 > 0) add_cols_kernel<<<grid_for((long long)B * d0, sm), 256, 0, st>>>(dHin, kh, 0, d0, B, dZ, zd, 0);
  if (dHin && d2 > 0) add_cols_kernel<<<grid_for((long long)B * d2, sm), 256, 0, st>>>(dHin, kh, d0, d2, B, dZ, zd, d0 + d1);
}
static void zero(float* p, long long n, int sm, cudaStream_t st) {
  if (n > 0) fill_kernel<<<grid_for(n, sm), 256, 0, st>>>(p, n, 0.f);
}

}  // namespace lt
}  // namespace bgm

extern "C" {

void bgm_lt_destroy(bgm_lt* t) {
  if (!t) return;
  for (int g = 0; g < 2; ++g) {
    if (t->theta[g]) cudaFree(t->theta[g]);
    if (t->grad[g]) cudaFree(t->grad[g]);
    if (t->m_pre[g]) cudaFree(t->m_pre[g]);
    if (t->v_pre[g]) cudaFree(t->v_pre[g]);
  }
  if (t->m_it) cudaFree(t->m_it);
  if (t->v_it) cudaFree(t->v_it);
  if (t->scratch) cudaFree(t->scratch);
  if (t->arena.base) cudaFree(t->arena.base);
  for (auto& g : t->graphs) cudaGraphExecDestroy(g.exec);
  if (t->cap_stream) cudaStreamDestroy(t->cap_stream);
  if (t->sc_dev) cudaFree(t->sc_dev);
  if (t->z_stage) cudaFree(t->z_stage);
  if (t->idx_stage) cudaFree(t->idx_stage);
  delete t;
}

int bgm_lt_create(bgm_lt** out, const int z_dims[4], int v_dim, int binary_treatment, int use_z_rec, int bayes,
                  const bgm_bnn_net_desc* g_net, const bgm_bnn_net_desc* e_net, const bgm_bnn_net_desc* f_net,
                  const bgm_bnn_net_desc* h_net, const bgm_disc_desc* dz_net, float lr, float beta_1, float beta_2,
                  float kl_weight, uint64_t seed) {
  using namespace bgm;
  using namespace bgm::lt;
  if (!out || !z_dims || !dz_net) return fail(BGM_ERR_ARG, "bgm_lt_create: null argument");
  *out = nullptr;
  bgm_lt* t = new bgm_lt();
  for (int i = 0; i < 4; ++i) t->z_dims[i] = z_dims[i];
  const int d0 = z_dims[0], d1 = z_dims[1], d2 = z_dims[2];
  t->zd = z_dims[0] + z_dims[1] + z_dims[2] + z_dims[3];
  t->p = v_dim;
  t->binary = binary_treatment ? 1 : 0;
  t->use_z_rec = use_z_rec ? 1.f : 0.f;
  t->bayes = bayes ? 1 : 0;
  t->lr = lr; t->b1 = beta_1; t->b2 = beta_2;
  t->kl_weight = kl_weight;
  t->seed = seed;
  int off = 0, rc;
  if ((rc = fill_net(g_net, t->g, bnn::NET_G, t->bayes, off, t->zd, v_dim + 1, "g_net")) ||
      (rc = fill_net(e_net, t->e, bnn::NET_E, t->bayes, off, v_dim, t->zd, "e_net")) ||
      (rc = fill_net(f_net, t->f, bnn::NET_F, t->bayes, off, d0 + d1 + 1, 2, "f_net")) ||
      (rc = fill_net(h_net, t->h, bnn::NET_H, t->bayes, off, d0 + d2, 2, "h_net")) ||
      (rc = tr_fill_disc(dz_net, t->dz, t->zd, "dz_net"))) {
    delete t;
    return rc;
  }
  t->n0 = off;
  t->n1 = t->dz.n_params;
  t->wm_disc = (std::max(t->zd, 4) + 3) / 4 * 4;
  t->smem_disc = (2 * t->wm_disc * tr::LD + 3 * t->zd * tr::LD + tr::disc_smem_floats(t->dz, true)) * 4 + 64;
  t->stage_disc = t->smem_disc + 8 * t->dz.n_params <= 200 * 1024;     // DiscArgs.stage: parameters + gradients in shared memory
  if (t->stage_disc) t->smem_disc += 8 * t->dz.n_params;
  {
    const char* g = getenv("BGM_LT_GRAPHS");
    t->use_graphs = !(g && g[0] == '0');
  }
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&t->sm_count, cudaDevAttrMultiProcessorCount, dev);
  const int n[2] = {t->n0, t->n1};
  for (int gi = 0; gi < 2 && e == cudaSuccess; ++gi) {
    const size_t bytes = sizeof(float) * (size_t)n[gi];
    float** arrs[4] = {&t->theta[gi], &t->grad[gi], &t->m_pre[gi], &t->v_pre[gi]};
    for (float** a : arrs) {
      if (e == cudaSuccess) e = cudaMalloc(a, bytes);
      if (e == cudaSuccess) e = cudaMemset(*a, 0, bytes);
    }
  }
  if (e == cudaSuccess) e = cudaMalloc(&t->m_it, sizeof(float) * (size_t)t->n0);
  if (e == cudaSuccess) e = cudaMalloc(&t->v_it, sizeof(float) * (size_t)t->n0);
  if (e == cudaSuccess) e = cudaMemset(t->m_it, 0, sizeof(float) * (size_t)t->n0);
  if (e == cudaSuccess) e = cudaMemset(t->v_it, 0, sizeof(float) * (size_t)t->n0);
  if (e == cudaSuccess) e = cudaMalloc(&t->scratch, sizeof(float) * 64);
  if (e == cudaSuccess) {
    std::vector<float> host(t->n0, 0.f);
    pack_net(g_net, t->g, host);
    pack_net(e_net, t->e, host);
    pack_net(f_net, t->f, host);
    pack_net(h_net, t->h, host);
    e = cudaMemcpy(t->theta[0], host.data(), sizeof(float) * t->n0, cudaMemcpyHostToDevice);
  }
  if (e == cudaSuccess) e = cudaMemcpy(t->theta[1], dz_net->params, sizeof(float) * t->n1, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    bgm_lt_destroy(t);
    return fail(BGM_ERR_CUDA, std::string("bgm_lt_create: ") + cudaGetErrorString(e));
  }
  *out = t;
  return 0;
}

int bgm_lt_buffers(bgm_lt* t, int group, int* n_params, float** theta_dev, float** grad_dev) {
  if (!t || group < 0 || group > 1) return bgm::fail(BGM_ERR_ARG, "bgm_lt_buffers: bad trainer / group");
  if (n_params) *n_params = group == 0 ? t->n0 : t->n1;
  if (theta_dev) *theta_dev = t->theta[group];
  if (grad_dev) *grad_dev = t->grad[group];
  return 0;
}
int bgm_lt_get_params(bgm_lt* t, int group, float* host_out) {
  using namespace bgm;
  if (!t || group < 0 || group > 1 || !host_out) return fail(BGM_ERR_ARG, "bgm_lt_get_params: bad argument");
  BGM_CUDA_OK(cudaMemcpy(host_out, t->theta[group], sizeof(float) * (size_t)(group == 0 ? t->n0 : t->n1), cudaMemcpyDeviceToHost));
  return 0;
}
int bgm_lt_set_params(bgm_lt* t, int group, const float* host_in) {
  using namespace bgm;
  if (!t || group < 0 || group > 1 || !host_in) return fail(BGM_ERR_ARG, "bgm_lt_set_params: bad argument");
  BGM_CUDA_OK(cudaMemcpy(t->theta[group], host_in, sizeof(float) * (size_t)(group == 0 ? t->n0 : t->n1), cudaMemcpyHostToDevice));
  return 0;
}
int bgm_lt_set_call(bgm_lt* t, uint32_t call_counter) {
  if (!t) return bgm::fail(BGM_ERR_ARG, "bgm_lt_set_call: null trainer");
  t->call_ctr = call_counter;
  return 0;
}

int bgm_lt_adam(bgm_lt* t, int group, float grad_scale, void* stream) {
  using namespace bgm;
  using namespace bgm::lt;
  if (!t || group < 0 || group > 1) return fail(BGM_ERR_ARG, "bgm_lt_adam: bad trainer / group");
  const int n = group == 0 ? t->n0 : t->n1;
  t->step_pre[group] += 1;
  const float lr_t = lr_t_of(t->lr, t->b1, t->b2, t->step_pre[group]);
  lt::adam_kernel<<<grid_for(n, t->sm_count), 256, 0, (cudaStream_t)stream>>>(t->theta[group], t->grad[group], t->m_pre[group],
                                                                           t->v_pre[group], n, lr_t, (float)t->b1, (float)t->b2,
                                                                           1e-7f, grad_scale, nullptr);
  BGM_CUDA_OK(cudaGetLastError());
  return 0;
}

// train_disc_step gradients (causalbgm/base.py:305-323): z_ = e_net(v) on the layered engine (call id 16*ctr),
// then the three discriminator passes and the gradient penalty with its double backward in train.cuh's kernel.
int bgm_lt_disc_grad(bgm_lt* t, const float* z_dev, const float* v_dev, int bs, float epsilon, float gp_weight,
                     float* losses_dev, void* stream) {
  using namespace bgm;
  using namespace bgm::lt;
  if (!t || !z_dev || !v_dev || !losses_dev) return fail(BGM_ERR_ARG, "bgm_lt_disc_grad: null argument");
  if (bs < 2) return fail(BGM_ERR_ARG, "bgm_lt_disc_grad: batch size must be >= 2");
  cudaStream_t st0 = (cudaStream_t)stream;
  int rc = ensure_arena(t, bs);
  if (rc) return rc;
  const float* zsrc = z_dev;
  if (t->use_graphs) {
    if ((rc = ensure_stage(t, bs))) return rc;
    BGM_CUDA_OK(cudaMemcpyAsync(t->z_stage, z_dev, sizeof(float) * (size_t)bs * t->zd, cudaMemcpyDeviceToDevice, st0));
    BGM_CUDA_OK(cudaMemcpyAsync(t->v_stage, v_dev, sizeof(float) * (size_t)bs * t->p, cudaMemcpyDeviceToDevice, st0));
    zsrc = t->z_stage;
    v_dev = t->v_stage;
  }
  float* const losses_out = losses_dev;
  if (t->use_graphs) losses_dev = t->loss_stage;
  StepScalars sc;
  memset(&sc, 0, sizeof(sc));
  sc.ctr = t->call_ctr;
  sc.eps = epsilon;
  const uint32_t c0_host = t->call_ctr * 16u;
  t->call_ctr += 1;
  BGM_CUDA_OK(cudaFuncSetAttribute(tr::disc_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, t->smem_disc));
  rc = run_step(t, 1, bs, 0, 0, {(const void*)(uintptr_t)__float_as_uint_host(gp_weight)}, sc, st0, [&](cudaStream_t st, bool dev) -> int {
    t->arena.used = 0;
    Ctx C = ctx_of(t);
    if (dev) C.ctr_dev = &t->sc_dev->ctr;
    const uint32_t c0 = dev ? 0u : c0_host;
    Pass eA;
    net_fwd(C, t->e, eA, v_dev, t->p, bs, c0, 0, -1, nullptr, nullptr, st);
    tr::DiscArgs A;
    memset(&A, 0, sizeof(A));
    A.dz = t->dz; A.zd = t->zd; A.p = t->p; A.bs = bs;
    A.theta = nullptr; A.theta_d = t->theta[1]; A.grad_d = t->grad[1];
    A.z = zsrc; A.v = nullptr; A.zenc_in = eA.out();
    A.epsilon = epsilon; A.eps_dev = dev ? &t->sc_dev->eps : nullptr;
    A.gp_weight = gp_weight; A.losses = losses_dev; A.wm = t->wm_disc; A.stage = t->stage_disc;
    if (bs <= 32) {
      tr::disc_grad_kernel<<<1, tr::NTH, t->smem_disc, st>>>(A);
    } else {     // more than 32 rows: the same algorithm on a global workspace (disc_big.cuh)
      tr::BigDiscArgs G;
      memset(&G, 0, sizeof(G));
      G.dz = t->dz; G.B = bs; G.zd = t->zd; G.theta_d = t->theta[1]; G.grad_d = t->grad[1];
      G.z = zsrc; G.zenc = eA.out(); G.epsilon = epsilon; G.eps_dev = A.eps_dev; G.gp_weight = gp_weight;
      G.losses = losses_dev;
      G.ws = t->arena.get<float>(tr::disc_big_floats(t->dz, bs));
      tr::disc_grad_big_kernel<<<1, 512, 0, st>>>(G);
    }
    return arena_ok(t->arena, "bgm_lt_disc_grad");
  });
  if (rc == 0 && t->use_graphs)
    BGM_CUDA_OK(cudaMemcpyAsync(losses_out, t->loss_stage, sizeof(float) * 2, cudaMemcpyDeviceToDevice, st0));
  return rc;
}

// train_gen_step gradients (causalbgm/base.py:332-370).  Net calls and their noise ids (16*ctr + k): g(z) for v_
// (k=0) and again for its sigma head (k=1, a separate call in the reference, :336-337), e(v) (k=0), e(v_) (k=1),
// g(z_) (k=2), f twice (k=0,1), h twice (k=0,1).
int bgm_lt_gen_grad(bgm_lt* t, const float* z_dev, const float* v_dev, const float* x_dev, const float* y_dev, int bs,
                    float* losses_dev, void* stream) {
  using namespace bgm;
  using namespace bgm::lt;
  if (!t || !z_dev || !v_dev || !x_dev || !y_dev || !losses_dev) return fail(BGM_ERR_ARG, "bgm_lt_gen_grad: null argument");
  if (bs < 2) return fail(BGM_ERR_ARG, "bgm_lt_gen_grad: batch size must be >= 2");
  cudaStream_t st0 = (cudaStream_t)stream;
  int rc = ensure_arena(t, bs);
  if (rc) return rc;
  const float* zsrc = z_dev;
  if (t->use_graphs) {
    if ((rc = ensure_stage(t, bs))) return rc;
    BGM_CUDA_OK(cudaMemcpyAsync(t->z_stage, z_dev, sizeof(float) * (size_t)bs * t->zd, cudaMemcpyDeviceToDevice, st0));
    BGM_CUDA_OK(cudaMemcpyAsync(t->v_stage, v_dev, sizeof(float) * (size_t)bs * t->p, cudaMemcpyDeviceToDevice, st0));
    BGM_CUDA_OK(cudaMemcpyAsync(t->x_stage, x_dev, sizeof(float) * (size_t)bs, cudaMemcpyDeviceToDevice, st0));
    BGM_CUDA_OK(cudaMemcpyAsync(t->y_stage, y_dev, sizeof(float) * (size_t)bs, cudaMemcpyDeviceToDevice, st0));
    zsrc = t->z_stage;
    v_dev = t->v_stage; x_dev = t->x_stage; y_dev = t->y_stage;
  }
  float* const losses_out = losses_dev;
  if (t->use_graphs) losses_dev = t->loss_stage;
  StepScalars sc;
  memset(&sc, 0, sizeof(sc));
  sc.ctr = t->call_ctr;
  const uint32_t c0_host = t->call_ctr * 16u;
  t->call_ctr += 1;
  rc = run_step(t, 2, bs, 0, 0, {}, sc, st0, [&](cudaStream_t st, bool dev) -> int {
  t->arena.used = 0;
  const float* z_dev = zsrc;      // the staged copy inside a replayed graph
  Arena& ar = t->arena;
  const int B = bs, p = t->p, zd = t->zd, sm = t->sm_count;
  const int d0 = t->z_dims[0], d1 = t->z_dims[1], d2 = t->z_dims[2];
  const int kf = d0 + d1 + 1, kh = d0 + d2;
  const uint32_t c0 = dev ? 0u : c0_host;
  zero(t->grad[0], t->n0, sm, st);
  Ctx C = ctx_of(t);
  if (dev) C.ctr_dev = &t->sc_dev->ctr;
  Pass gA, gB, gC, eA, eB, fA, fB, hA, hB;
  net_fwd(C, t->g, gA, z_dev, zd, B, c0 + 0, 0, -1, nullptr, nullptr, st);
  net_fwd(C, t->g, gB, z_dev, zd, B, c0 + 1, 0, -1, nullptr, nullptr, st);
  net_fwd(C, t->e, eA, v_dev, p, B, c0 + 0, 0, -1, nullptr, nullptr, st);
  net_fwd(C, t->e, eB, gA.out(), p + 1, B, c0 + 1, 0, -1, nullptr, nullptr, st);       // e(v_): the first p columns of g(z)
  const float* zenc = eA.out();
  net_fwd(C, t->g, gC, zenc, zd, B, c0 + 2, 0, -1, nullptr, nullptr, st);
  DiscPass D;
  disc_fwd(C, t->dz, t->theta[1], D, zenc, B, st);
  float* fin = ar.get<float>((size_t)B * kf);
  float* hin = ar.get<float>((size_t)B * std::max(kh, 1));
  build_fh_inputs(t, zenc, zd, x_dev, B, fin, hin, st);
  net_fwd(C, t->f, fA, fin, kf, B, c0 + 0, 0, -1, nullptr, nullptr, st);
  net_fwd(C, t->f, fB, fin, kf, B, c0 + 1, 0, -1, nullptr, nullptr, st);
  net_fwd(C, t->h, hA, hin, kh, B, c0 + 0, 0, -1, nullptr, nullptr, st);
  net_fwd(C, t->h, hB, hin, kh, B, c0 + 1, 0, -1, nullptr, nullptr, st);
  GenLossArgs L;
  memset(&L, 0, sizeof(L));
  L.B = B; L.p = p; L.zd = zd; L.binary = t->binary; L.use_z_rec = t->use_z_rec;
  L.z = z_dev; L.v = v_dev; L.x = x_dev; L.y = y_dev;
  L.gB = gB.out(); L.gC = gC.out(); L.eB = eB.out(); L.d = D.d; L.fA = fA.out(); L.fB = fB.out(); L.hA = hA.out(); L.hB = hB.out();
  L.dgB = ar.get<float>((size_t)B * (p + 1)); L.dgC = ar.get<float>((size_t)B * (p + 1)); L.deB = ar.get<float>((size_t)B * zd);
  L.dd = ar.get<float>(B); L.dfA = ar.get<float>(B * 2); L.dfB = ar.get<float>(B * 2); L.dhA = ar.get<float>(B * 2);
  L.dhB = ar.get<float>(B * 2);
  L.losses = losses_dev;
  gen_loss_kernel<<<1, 256, 0, st>>>(L);
  float* dZ = ar.get<float>((size_t)B * zd);            // gradient w.r.t. z_ = e(v)
  float* dFin = ar.get<float>((size_t)B * kf);
  float* dHin = ar.get<float>((size_t)B * std::max(kh, 1));
  float* dgA = ar.get<float>((size_t)B * (p + 1));
  zero(dZ, (long long)B * zd, sm, st);
  zero(dgA, (long long)B * (p + 1), sm, st);
  disc_bwd(C, t->dz, t->theta[1], nullptr, D, L.dd, B, dZ, st);
  net_bwd(C, t->g, gC, L.dgC, true, dZ, zd, true, st);
  net_bwd(C, t->f, fA, L.dfA, true, dFin, kf, false, st);
  net_bwd(C, t->f, fB, L.dfB, true, dFin, kf, true, st);
  net_bwd(C, t->h, hA, L.dhA, true, dHin, kh, false, st);
  net_bwd(C, t->h, hB, L.dhB, true, dHin, kh, true, st);
  scatter_fh_grads(t, dFin, dHin, B, dZ, st);
  net_bwd(C, t->e, eB, L.deB, true, dgA, p + 1, false, st);   // d loss / d v_ lands in the first p columns of g(z)'s output gradient
  net_bwd(C, t->g, gA, dgA, true, nullptr, 0, false, st);
  net_bwd(C, t->g, gB, L.dgB, true, nullptr, 0, false, st);
  net_bwd(C, t->e, eA, dZ, true, nullptr, 0, false, st);
  return arena_ok(t->arena, "bgm_lt_gen_grad");
  });
  if (rc == 0 && t->use_graphs)
    BGM_CUDA_OK(cudaMemcpyAsync(losses_out, t->loss_stage, sizeof(float) * 6, cudaMemcpyDeviceToDevice, st0));
  return rc;
}

int bgm_lt_set_iter(bgm_lt* t, float lr_theta, float lr_z, float sigma_v, float sigma_x, float sigma_y) {
  using namespace bgm;
  if (!t) return fail(BGM_ERR_ARG, "bgm_lt_set_iter: null trainer");
  t->lr_theta = lr_theta; t->lr_z = lr_z;
  t->s2v = sigma_v >= 0.f ? sigma_v * sigma_v : -1.f;
  t->s2x = sigma_x >= 0.f ? sigma_x * sigma_x : -1.f;
  t->s2y = sigma_y >= 0.f ? sigma_y * sigma_y : -1.f;
  t->step_it[0] = t->step_it[1] = t->step_it[2] = 0;
  t->step_z = 0;
  BGM_CUDA_OK(cudaMemset(t->m_it, 0, sizeof(float) * (size_t)t->n0));
  BGM_CUDA_OK(cudaMemset(t->v_it, 0, sizeof(float) * (size_t)t->n0));
  return 0;
}

// update_g_net, update_h_net, update_f_net (causalbgm/base.py:156-243) on the rows idx_dev of the latent table and
// the data: one call of each net (noise id 16*ctr), Gaussian NLL (+ kl_weight * KL of the net's kernels for Bayesian
// nets, :171-173), gradients into grad[0] (apply 0 / 1) and one Keras-Adam step per net on its own optimizer state
// (apply 1; apply 2: only the updates, on grad_scale * the gradients already in the buffer).
int bgm_lt_iter_nets(bgm_lt* t, const float* zt_dev, const float* x_dev, const float* y_dev, const float* v_dev,
                     const int* idx_dev, int bs, int apply, float grad_scale, float* losses_dev, void* stream) {
  using namespace bgm;
  using namespace bgm::lt;
  if (!t || !zt_dev || !x_dev || !y_dev || !v_dev || !idx_dev || !losses_dev) return fail(BGM_ERR_ARG, "bgm_lt_iter_nets: null argument");
  if (bs < 1) return fail(BGM_ERR_ARG, "bgm_lt_iter_nets: batch size must be >= 1");
  cudaStream_t st0 = (cudaStream_t)stream;
  const int B = bs, p = t->p, zd = t->zd, sm = t->sm_count;
  const LNet* nets[3] = {&t->g, &t->h, &t->f};
  int rc = ensure_arena(t, bs);
  if (rc) return rc;
  const int* isrc = idx_dev;
  if (t->use_graphs) {
    if ((rc = ensure_stage(t, bs))) return rc;
    BGM_CUDA_OK(cudaMemcpyAsync(t->idx_stage, idx_dev, sizeof(int) * (size_t)bs, cudaMemcpyDeviceToDevice, st0));
    isrc = t->idx_stage;
  }
  StepScalars sc;
  memset(&sc, 0, sizeof(sc));
  sc.ctr = t->call_ctr;
  const uint32_t c0_host = t->call_ctr * 16u;
  if (apply != 2) t->call_ctr += 1;
  float lr_host[3] = {0.f, 0.f, 0.f};
  if (apply != 0)
    for (int i = 0; i < 3; ++i) {
      t->step_it[i] += 1;
      lr_host[i] = sc.lr[i] = lr_t_of(t->lr_theta, 0.9, 0.99, t->step_it[i]);
    }
  return run_step(t, 3, bs, apply, 0, {zt_dev, x_dev, y_dev, v_dev, losses_dev, (const void*)(uintptr_t)__float_as_uint_host(grad_scale)}, sc, st0,
                  [&](cudaStream_t st, bool dev) -> int {
  t->arena.used = 0;
  const int* idx_dev = isrc;
  if (apply != 2) {
    Arena& ar = t->arena;
    const int d0 = t->z_dims[0], d1 = t->z_dims[1], d2 = t->z_dims[2];
    const int kf = d0 + d1 + 1, kh = d0 + d2;
    const uint32_t c0 = dev ? 0u : c0_host;
    zero(t->grad[0], t->n0, sm, st);
    zero(losses_dev, 6, sm, st);
    float* zb = ar.get<float>((size_t)B * zd);
    float* vb = ar.get<float>((size_t)B * p);
    float* xb = ar.get<float>(B);
    float* yb = ar.get<float>(B);
    gather_cols_kernel<<<grid_for((long long)B * zd, sm), 256, 0, st>>>(zt_dev, zd, idx_dev, 0, zd, B, zb, zd, 0);
    gather_cols_kernel<<<grid_for((long long)B * p, sm), 256, 0, st>>>(v_dev, p, idx_dev, 0, p, B, vb, p, 0);
    gather_cols_kernel<<<grid_for(B, sm), 256, 0, st>>>(x_dev, 1, idx_dev, 0, 1, B, xb, 1, 0);
    gather_cols_kernel<<<grid_for(B, sm), 256, 0, st>>>(y_dev, 1, idx_dev, 0, 1, B, yb, 1, 0);
    float* fin = ar.get<float>((size_t)B * kf);
    float* hin = ar.get<float>((size_t)B * std::max(kh, 1));
    build_fh_inputs(t, zb, zd, xb, B, fin, hin, st);
    Ctx C = ctx_of(t);
    if (dev) C.ctr_dev = &t->sc_dev->ctr;
    Pass gP, hP, fP;
    net_fwd(C, t->g, gP, zb, zd, B, c0, 0, -1, nullptr, nullptr, st);
    net_fwd(C, t->h, hP, hin, kh, B, c0, 0, -1, nullptr, nullptr, st);
    net_fwd(C, t->f, fP, fin, kf, B, c0, 0, -1, nullptr, nullptr, st);
    const Pass* passes[3] = {&gP, &hP, &fP};
    const float* targets[3] = {vb, xb, yb};
    const int Ds[3] = {p, 1, 1};
    const float s2[3] = {t->s2v, t->s2x, t->s2y};
    for (int i = 0; i < 3; ++i) {
      const LNet& net = *nets[i];
      const int ldo = net.dims[net.L];
      float* dO = ar.get<float>((size_t)B * ldo);
      NllArgs N;
      memset(&N, 0, sizeof(N));
      N.B = B; N.D = Ds[i]; N.ldo = ldo; N.rcol = ldo - 1; N.binary = (i == 1 && t->binary) ? 1 : 0; N.s2_fixed = s2[i];
      N.target = targets[i]; N.MU = passes[i]->out(); N.RAW = passes[i]->out(); N.dMU = dO; N.dRAW = dO;
      N.losses = losses_dev + 2 * i;
      nll_loss_kernel<<<1, 256, 0, st>>>(N);
      if (net.bayes && t->kl_weight != 0.f)
        for (int l = 0; l < net.L; ++l) {
          const int cnt = net.dims[l] * net.dims[l + 1];
          kl_grad_kernel<<<grid_for(cnt, sm), 256, 0, st>>>(t->theta[0] + net.off_w[l], t->theta[0] + net.off_rho[l], cnt, t->kl_weight,
                                                           t->grad[0] + net.off_w[l], t->grad[0] + net.off_rho[l], losses_dev + 2 * i);
        }
      net_bwd(C, net, *passes[i], dO, true, nullptr, 0, false, st);
    }
    const int rc2 = arena_ok(t->arena, "bgm_lt_iter_nets");
    if (rc2) return rc2;
  }
  if (apply != 0) {
    for (int i = 0; i < 3; ++i) {
      const LNet& net = *nets[i];
      lt::adam_kernel<<<grid_for(net.n_params, sm), 256, 0, st>>>(t->theta[0] + net.base, t->grad[0] + net.base, t->m_it + net.base,
                                                               t->v_it + net.base, net.n_params, lr_host[i], 0.9f, 0.99f, 1e-7f, grad_scale,
                                                               dev ? &t->sc_dev->lr[i] : nullptr);
    }
    BGM_CUDA_OK(cudaGetLastError());
  }
  return 0;
  });
}


// update_latent_variable_sgd (causalbgm/base.py:246-302): g, h, f are each called TWICE (mean and variance heads
// come from separate calls, :259-286; noise ids 16*ctr and 16*ctr+1), gradient of loss_postrior_z w.r.t. the batch
// rows (through the batch statistics of the input BatchNormalization for Bayesian nets), then Keras Adam on the
// gathered variable = a dense sweep over the whole table.  gz_out_dev (bs, zd), optional: the gradient rows.
int bgm_lt_iter_latent(bgm_lt* t, float* zt_dev, float* m_dev, float* v_adam_dev, int* slot_dev, long long n,
                       const float* x_dev, const float* y_dev, const float* v_dev, const int* idx_dev, int bs,
                       float* loss_dev, float* gz_out_dev, void* stream) {
  using namespace bgm;
  using namespace bgm::lt;
  if (!t || !zt_dev || !m_dev || !v_adam_dev || !slot_dev || !x_dev || !y_dev || !v_dev || !idx_dev || !loss_dev)
    return fail(BGM_ERR_ARG, "bgm_lt_iter_latent: null argument");
  if (bs < 1 || n < bs) return fail(BGM_ERR_ARG, "bgm_lt_iter_latent: bad batch size / table size");
  cudaStream_t st0 = (cudaStream_t)stream;
  int rc = ensure_arena(t, bs);
  if (rc) return rc;
  const int* isrc = idx_dev;
  if (t->use_graphs) {
    if ((rc = ensure_stage(t, bs))) return rc;
    BGM_CUDA_OK(cudaMemcpyAsync(t->idx_stage, idx_dev, sizeof(int) * (size_t)bs, cudaMemcpyDeviceToDevice, st0));
    isrc = t->idx_stage;
  }
  StepScalars sc;
  memset(&sc, 0, sizeof(sc));
  sc.ctr = t->call_ctr;
  const uint32_t c0_host = t->call_ctr * 16u;
  t->call_ctr += 1;
  t->step_z += 1;
  const float lr_t = sc.lr[3] = lr_t_of(t->lr_z, 0.9, 0.99, t->step_z);
  return run_step(t, 4, bs, 0, n, {zt_dev, m_dev, v_adam_dev, slot_dev, x_dev, y_dev, v_dev, loss_dev, gz_out_dev}, sc, st0,
                  [&](cudaStream_t st, bool dev) -> int {
  t->arena.used = 0;
  const int* idx_dev = isrc;
  Arena& ar = t->arena;
  const int B = bs, p = t->p, zd = t->zd, sm = t->sm_count;
  const int d0 = t->z_dims[0], d1 = t->z_dims[1], d2 = t->z_dims[2];
  const int kf = d0 + d1 + 1, kh = d0 + d2;
  const uint32_t c0 = dev ? 0u : c0_host;
  float* losses = t->scratch;            // [0..1] v, [2..3] x, [4..5] y, [6] prior
  zero(losses, 8, sm, st);
  float* zb = ar.get<float>((size_t)B * zd);
  float* vb = ar.get<float>((size_t)B * p);
  float* xb = ar.get<float>(B);
  float* yb = ar.get<float>(B);
  gather_cols_kernel<<<grid_for((long long)B * zd, sm), 256, 0, st>>>(zt_dev, zd, idx_dev, 0, zd, B, zb, zd, 0);
  gather_cols_kernel<<<grid_for((long long)B * p, sm), 256, 0, st>>>(v_dev, p, idx_dev, 0, p, B, vb, p, 0);
  gather_cols_kernel<<<grid_for(B, sm), 256, 0, st>>>(x_dev, 1, idx_dev, 0, 1, B, xb, 1, 0);
  gather_cols_kernel<<<grid_for(B, sm), 256, 0, st>>>(y_dev, 1, idx_dev, 0, 1, B, yb, 1, 0);
  float* fin = ar.get<float>((size_t)B * kf);
  float* hin = ar.get<float>((size_t)B * std::max(kh, 1));
  build_fh_inputs(t, zb, zd, xb, B, fin, hin, st);
  Ctx C = ctx_of(t);
  if (dev) C.ctr_dev = &t->sc_dev->ctr;
  Pass gA, gB, hA, hB, fA, fB;
  net_fwd(C, t->g, gA, zb, zd, B, c0, 0, -1, nullptr, nullptr, st);
  net_fwd(C, t->g, gB, zb, zd, B, c0 + 1, 0, -1, nullptr, nullptr, st);
  net_fwd(C, t->h, hA, hin, kh, B, c0, 0, -1, nullptr, nullptr, st);
  net_fwd(C, t->h, hB, hin, kh, B, c0 + 1, 0, -1, nullptr, nullptr, st);
  net_fwd(C, t->f, fA, fin, kf, B, c0, 0, -1, nullptr, nullptr, st);
  net_fwd(C, t->f, fB, fin, kf, B, c0 + 1, 0, -1, nullptr, nullptr, st);
  const Pass* A[3] = {&gA, &hA, &fA};
  const Pass* R[3] = {&gB, &hB, &fB};
  const float* targets[3] = {vb, xb, yb};
  const int Ds[3] = {p, 1, 1};
  const float s2[3] = {t->s2v, t->s2x, t->s2y};
  float* dA[3];
  float* dR[3];
  for (int i = 0; i < 3; ++i) {
    const int ldo = A[i]->ldo();
    dA[i] = ar.get<float>((size_t)B * ldo);
    dR[i] = ar.get<float>((size_t)B * ldo);
    NllArgs N;
    memset(&N, 0, sizeof(N));
    N.B = B; N.D = Ds[i]; N.ldo = ldo; N.rcol = ldo - 1; N.binary = (i == 1 && t->binary) ? 1 : 0; N.s2_fixed = s2[i];
    N.target = targets[i]; N.MU = A[i]->out(); N.RAW = R[i]->out(); N.dMU = dA[i]; N.dRAW = dR[i];
    N.losses = losses + 2 * i;
    nll_loss_kernel<<<1, 256, 0, st>>>(N);
  }
  float* dZ = ar.get<float>((size_t)B * zd);
  float* dFin = ar.get<float>((size_t)B * kf);
  float* dHin = ar.get<float>((size_t)B * std::max(kh, 1));
  // prior (:291-292): mean_b |z_b|^2 / 2  ->  dZ = z / B
  gather_cols_kernel<<<grid_for((long long)B * zd, sm), 256, 0, st>>>(zb, zd, nullptr, 0, zd, B, dZ, zd, 0);
  prior_scale_kernel<<<1, 256, 0, st>>>(dZ, B * zd, 1.f / (float)B, losses + 6);
  net_bwd(C, t->g, gA, dA[0], false, dZ, zd, true, st);
  net_bwd(C, t->g, gB, dR[0], false, dZ, zd, true, st);
  net_bwd(C, t->h, hA, dA[1], false, dHin, kh, false, st);
  net_bwd(C, t->h, hB, dR[1], false, dHin, kh, true, st);
  net_bwd(C, t->f, fA, dA[2], false, dFin, kf, false, st);
  net_bwd(C, t->f, fB, dR[2], false, dFin, kf, true, st);
  scatter_fh_grads(t, dFin, dHin, B, dZ, st);
  sum_losses_kernel<<<1, 32, 0, st>>>(losses, loss_dev);
  if (gz_out_dev) BGM_CUDA_OK(cudaMemcpyAsync(gz_out_dev, dZ, sizeof(float) * (size_t)B * zd, cudaMemcpyDeviceToDevice, st));
  set_slots_kernel<<<grid_for(B, sm), 256, 0, st>>>(slot_dev, idx_dev, B, 1);
  tr::launch_latent_adam_sweep(zt_dev, m_dev, v_adam_dev, slot_dev, dZ, n, zd, lr_t, 0.9f, 0.99f, 1e-7f,
                               dev ? &t->sc_dev->lr[3] : nullptr, sm, st);
  set_slots_kernel<<<grid_for(B, sm), 256, 0, st>>>(slot_dev, idx_dev, B, 0);
  return arena_ok(t->arena, "bgm_lt_iter_latent");
  });
}

// CausalBGM.evaluate (causalbgm/base.py:534-556), the part that touches every row, in row chunks: the batch
// statistics of every net's input are those of ALL n rows (one call per net, noise id 16*ctr), the forward passes
// run chunk by chunk (signs keyed by the global row).  sums_dev[3] (float64) = sum (v - v^)^2, sum (x - x^)^2,
// sum (y - y^)^2; z = zt_dev rows or e_net(v) (then written to z_out_dev if given).
int bgm_lt_evaluate(bgm_lt* t, const float* zt_dev, const float* x_dev, const float* y_dev, const float* v_dev, int n,
                    double* sums_dev, float* z_out_dev, void* stream) {
  using namespace bgm;
  using namespace bgm::lt;
  if (!t || !x_dev || !y_dev || !v_dev || !sums_dev || n < 1) return fail(BGM_ERR_ARG, "bgm_lt_evaluate: null argument / n < 1");
  cudaStream_t st = (cudaStream_t)stream;
  const int p = t->p, zd = t->zd, sm = t->sm_count;
  const int d0 = t->z_dims[0], d1 = t->z_dims[1], d2 = t->z_dims[2];
  const int kf = d0 + d1 + 1, kh = d0 + d2;
  const int chunk = std::min(n, 16384);
  int rc = ensure_arena(t, chunk);
  if (rc) return rc;
  const uint32_t c0 = t->call_ctr * 16u;
  t->call_ctr += 1;
  const Ctx C = ctx_of(t);
  // whole-batch buffers (freed on return): z, f / h inputs, statistics
  float *zall = nullptr, *fin = nullptr, *hin = nullptr, *stats = nullptr;
  const int nstat = 2 * (p + zd + kf + std::max(kh, 1));
  cudaError_t e = cudaSuccess;
  if (!zt_dev) e = cudaMalloc(&zall, sizeof(float) * (size_t)n * zd);
  if (e == cudaSuccess) e = cudaMalloc(&fin, sizeof(float) * (size_t)n * kf);
  if (e == cudaSuccess) e = cudaMalloc(&hin, sizeof(float) * (size_t)n * std::max(kh, 1));
  if (e == cudaSuccess) e = cudaMalloc(&stats, sizeof(float) * nstat);
  auto cleanup = [&]() {
    cudaStreamSynchronize(st);
    if (zall) cudaFree(zall);
    if (fin) cudaFree(fin);
    if (hin) cudaFree(hin);
    if (stats) cudaFree(stats);
  };
  if (e != cudaSuccess) { cleanup(); return fail(BGM_ERR_CUDA, std::string("bgm_lt_evaluate: ") + cudaGetErrorString(e)); }
  float* st_v = stats;                 // mean | inv of each net's input
  float* st_z = st_v + 2 * p;
  float* st_f = st_z + 2 * zd;
  float* st_h = st_f + 2 * kf;
  zero(reinterpret_cast<float*>(sums_dev), 6, sm, st);
  const float* Z = zt_dev;
  if (!zt_dev) {
    if (t->bayes) col_stats_kernel<<<p, 256, 0, st>>>(v_dev, p, n, p, st_v, st_v + p);
    for (int r0 = 0; r0 < n; r0 += chunk) {
      const int B = std::min(chunk, n - r0);
      t->arena.used = 0;
      Pass eP;
      net_fwd(C, t->e, eP, v_dev + (size_t)r0 * p, p, B, c0, r0, -1, st_v, st_v + p, st);
      BGM_CUDA_OK(cudaMemcpyAsync(zall + (size_t)r0 * zd, eP.out(), sizeof(float) * (size_t)B * zd, cudaMemcpyDeviceToDevice, st));
    }
    Z = zall;
    if (z_out_dev) BGM_CUDA_OK(cudaMemcpyAsync(z_out_dev, zall, sizeof(float) * (size_t)n * zd, cudaMemcpyDeviceToDevice, st));
  }
  build_fh_inputs(t, Z, zd, x_dev, n, fin, hin, st);
  if (t->bayes) {
    col_stats_kernel<<<zd, 256, 0, st>>>(Z, zd, n, zd, st_z, st_z + zd);
    col_stats_kernel<<<kf, 256, 0, st>>>(fin, kf, n, kf, st_f, st_f + kf);
    if (kh > 0) col_stats_kernel<<<kh, 256, 0, st>>>(hin, kh, n, kh, st_h, st_h + kh);
  }
  for (int r0 = 0; r0 < n; r0 += chunk) {
    const int B = std::min(chunk, n - r0);
    t->arena.used = 0;
    Pass gP, fP, hP;
    net_fwd(C, t->g, gP, Z + (size_t)r0 * zd, zd, B, c0, r0, -1, st_z, st_z + zd, st);
    sq_err_kernel<<<grid_for((long long)B * p, sm), 256, 0, st>>>(v_dev + (size_t)r0 * p, p, gP.out(), p + 1, B, p, 0, sums_dev);
    t->arena.used = 0;
    net_fwd(C, t->h, hP, hin + (size_t)r0 * kh, kh, B, c0, r0, -1, st_h, st_h + kh, st);
    sq_err_kernel<<<grid_for(B, sm), 256, 0, st>>>(x_dev + r0, 1, hP.out(), 2, B, 1, t->binary, sums_dev + 1);
    t->arena.used = 0;
    net_fwd(C, t->f, fP, fin + (size_t)r0 * kf, kf, B, c0, r0, -1, st_f, st_f + kf, st);
    sq_err_kernel<<<grid_for(B, sm), 256, 0, st>>>(y_dev + r0, 1, fP.out(), 2, B, 1, 0, sums_dev + 2);
  }
  rc = arena_ok(t->arena, "bgm_lt_evaluate");
  cleanup();
  return rc;
}

// ------------------------------------------------------------------------------------------ BGM flavour ----
void bgm_ltb_destroy(bgm_ltb* t) {
  if (!t) return;
  for (int g = 0; g < 2; ++g) {
    if (t->theta[g]) cudaFree(t->theta[g]);
    if (t->grad[g]) cudaFree(t->grad[g]);
    if (t->m_pre[g]) cudaFree(t->m_pre[g]);
    if (t->v_pre[g]) cudaFree(t->v_pre[g]);
  }
  if (t->m_it) cudaFree(t->m_it);
  if (t->v_it) cudaFree(t->v_it);
  if (t->moving) cudaFree(t->moving);
  if (t->scratch) cudaFree(t->scratch);
  if (t->arena.base) cudaFree(t->arena.base);
  delete t;
}

int bgm_ltb_create(bgm_ltb** out, const bgm_varnet_desc* g, const bgm_net_desc* e_net, const bgm_disc_desc* dz_net,
                   const bgm_disc_desc* dx_net, float lr, float beta_1, float beta_2, float alpha, float gamma) {
  using namespace bgm;
  using namespace bgm::lt;
  if (!out || !g || !e_net || !dz_net || !dx_net) return fail(BGM_ERR_ARG, "bgm_ltb_create: null argument");
  *out = nullptr;
  if (gamma != 0.f)
    return fail(BGM_ERR_UNSUPPORTED, "bgm_ltb_create: the layered engine has no gradient-penalty double backward (gamma must be 0)");
  if (!g->units || !g->bn || !g->hidden_params || !g->mean_params || !g->var_params || g->n_hidden < 1 || g->n_hidden + 1 > LT_MAXL)
    return fail(BGM_ERR_ARG, "bgm_ltb_create: bad generator description");
  if (!e_net->dims || !e_net->params || e_net->n_layers < 1 || e_net->n_layers > LT_MAXL)
    return fail(BGM_ERR_ARG, "bgm_ltb_create: bad encoder description");
  const int zd = g->z_dim, xd = g->x_dim, nh = g->n_hidden;
  if (e_net->dims[0] != xd || e_net->dims[e_net->n_layers] != zd) return fail(BGM_ERR_ARG, "bgm_ltb_create: e_net must map x_dim -> z_dim");
  bgm_ltb* t = new bgm_ltb();
  t->zd = zd; t->xd = xd; t->alpha = alpha; t->gamma = gamma;
  t->lr = lr; t->b1 = beta_1; t->b2 = beta_2;
  int off = 0;
  LNet& G = t->g;
  G.L = nh + 1; G.bn_in = 1; G.flip = 0; G.bayes = 0; G.net_id = 0; G.base = 0;
  G.off_gamma = off; off += zd;
  G.off_beta = off; off += zd;
  G.dims[0] = zd;
  for (int l = 0; l < nh; ++l) G.dims[l + 1] = g->units[l];
  G.dims[nh + 1] = 2 * xd;
  for (int l = 0; l < G.L; ++l) {
    G.off_w[l] = off; off += G.dims[l] * G.dims[l + 1];
    G.off_rho[l] = -1;
    G.off_b[l] = off; off += G.dims[l + 1];
  }
  G.n_params = off;
  t->n_g = off;
  LNet& E = t->e;
  E.L = e_net->n_layers; E.base = off; E.net_id = 3;
  for (int l = 0; l <= E.L; ++l) E.dims[l] = e_net->dims[l];
  for (int l = 0; l < E.L; ++l) {
    E.off_w[l] = off; off += E.dims[l] * E.dims[l + 1];
    E.off_rho[l] = -1;
    E.off_b[l] = off; off += E.dims[l + 1];
  }
  E.n_params = off - E.base;
  t->n0 = off;
  int rc;
  if ((rc = tr_fill_disc(dz_net, t->dz, zd, "dz_net")) || (rc = tr_fill_disc(dx_net, t->dx, xd, "dx_net"))) { delete t; return rc; }
  t->dx_base = t->dz.n_params;
  t->n1 = t->dz.n_params + t->dx.n_params;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&t->sm_count, cudaDevAttrMultiProcessorCount, dev);
  const int n[2] = {t->n0, t->n1};
  for (int gi = 0; gi < 2 && e == cudaSuccess; ++gi) {
    const size_t bytes = sizeof(float) * (size_t)n[gi];
    float** arrs[4] = {&t->theta[gi], &t->grad[gi], &t->m_pre[gi], &t->v_pre[gi]};
    for (float** a : arrs) {
      if (e == cudaSuccess) e = cudaMalloc(a, bytes);
      if (e == cudaSuccess) e = cudaMemset(*a, 0, bytes);
    }
  }
  if (e == cudaSuccess) e = cudaMalloc(&t->m_it, sizeof(float) * (size_t)t->n_g);
  if (e == cudaSuccess) e = cudaMalloc(&t->v_it, sizeof(float) * (size_t)t->n_g);
  if (e == cudaSuccess) e = cudaMemset(t->m_it, 0, sizeof(float) * (size_t)t->n_g);
  if (e == cudaSuccess) e = cudaMemset(t->v_it, 0, sizeof(float) * (size_t)t->n_g);
  if (e == cudaSuccess) e = cudaMalloc(&t->moving, sizeof(float) * 2 * zd);
  if (e == cudaSuccess) e = cudaMalloc(&t->scratch, sizeof(float) * 64);
  if (e == cudaSuccess) {
    std::vector<float> host(t->n0, 0.f);
    memcpy(host.data() + G.off_gamma, g->bn, sizeof(float) * zd);
    memcpy(host.data() + G.off_beta, g->bn + zd, sizeof(float) * zd);
    const float* p = g->hidden_params;
    for (int l = 0; l < nh; ++l) {
      const int cnt = G.dims[l] * G.dims[l + 1] + G.dims[l + 1];
      memcpy(host.data() + G.off_w[l], p, sizeof(float) * cnt);
      p += cnt;
    }
    const int last = G.dims[nh];
    for (int k = 0; k < last; ++k)
      for (int c = 0; c < xd; ++c) {
        host[G.off_w[nh] + (size_t)k * 2 * xd + c] = g->mean_params[(size_t)k * xd + c];
        host[G.off_w[nh] + (size_t)k * 2 * xd + xd + c] = g->var_params[(size_t)k * xd + c];
      }
    for (int c = 0; c < xd; ++c) {
      host[G.off_b[nh] + c] = g->mean_params[(size_t)last * xd + c];
      host[G.off_b[nh] + xd + c] = g->var_params[(size_t)last * xd + c];
    }
    memcpy(host.data() + E.base, e_net->params, sizeof(float) * E.n_params);
    e = cudaMemcpy(t->theta[0], host.data(), sizeof(float) * t->n0, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(t->moving, g->bn + 2 * zd, sizeof(float) * 2 * zd, cudaMemcpyHostToDevice);
  }
  if (e == cudaSuccess) e = cudaMemcpy(t->theta[1], dz_net->params, sizeof(float) * t->dz.n_params, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(t->theta[1] + t->dx_base, dx_net->params, sizeof(float) * t->dx.n_params, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    bgm_ltb_destroy(t);
    return fail(BGM_ERR_CUDA, std::string("bgm_ltb_create: ") + cudaGetErrorString(e));
  }
  *out = t;
  return 0;
}

int bgm_ltb_buffers(bgm_ltb* t, int group, int* n_params, float** theta_dev, float** grad_dev) {
  if (!t || group < 0 || group > 1) return bgm::fail(BGM_ERR_ARG, "bgm_ltb_buffers: bad trainer / group");
  if (n_params) *n_params = group == 0 ? t->n0 : t->n1;
  if (theta_dev) *theta_dev = t->theta[group];
  if (grad_dev) *grad_dev = t->grad[group];
  return 0;
}
int bgm_ltb_get_params(bgm_ltb* t, int group, float* host_out) {
  using namespace bgm;
  if (!t || group < 0 || group > 1 || !host_out) return fail(BGM_ERR_ARG, "bgm_ltb_get_params: bad argument");
  BGM_CUDA_OK(cudaMemcpy(host_out, t->theta[group], sizeof(float) * (size_t)(group == 0 ? t->n0 : t->n1), cudaMemcpyDeviceToHost));
  return 0;
}
int bgm_ltb_bn_moving(bgm_ltb* t, float* host_inout, int set) {
  using namespace bgm;
  if (!t || !host_inout) return fail(BGM_ERR_ARG, "bgm_ltb_bn_moving: needs a trainer and a buffer");
  if (set) BGM_CUDA_OK(cudaMemcpy(t->moving, host_inout, sizeof(float) * 2 * t->zd, cudaMemcpyHostToDevice));
  else BGM_CUDA_OK(cudaMemcpy(host_inout, t->moving, sizeof(float) * 2 * t->zd, cudaMemcpyDeviceToHost));
  return 0;
}
int bgm_ltb_adam(bgm_ltb* t, int group, float grad_scale, void* stream) {
  using namespace bgm;
  using namespace bgm::lt;
  if (!t || group < 0 || group > 1) return fail(BGM_ERR_ARG, "bgm_ltb_adam: bad trainer / group");
  const int n = group == 0 ? t->n0 : t->n1;
  t->step_pre[group] += 1;
  const float lr_t = lr_t_of(t->lr, t->b1, t->b2, t->step_pre[group]);
  lt::adam_kernel<<<grid_for(n, t->sm_count), 256, 0, (cudaStream_t)stream>>>(t->theta[group], t->grad[group], t->m_pre[group],
                                                                           t->v_pre[group], n, lr_t, (float)t->b1, (float)t->b2,
                                                                           1e-7f, grad_scale, nullptr);
  BGM_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"

namespace bgm {
namespace lt {
static size_t ltb_step_bytes(const bgm_ltb* t, long long B) {
  size_t b = 2 * pass_bytes(t->g, B) + 2 * pass_bytes(t->e, B) + 2 * disc_bytes(t->dz, B) + 2 * disc_bytes(t->dx, B);
  b += 12 * (size_t)(4 * B * (2 * t->xd + t->zd + 8) + 512);
  return b + (1 << 16);
}
static Ctx ctx_of(bgm_ltb* t) { return Ctx{&t->arena, t->theta[0], t->grad[0], t->sm_count, 0}; }
// generator forward in TRAINING mode: batch statistics + the moving-statistics update Keras does on every such call
static void g_fwd_train(bgm_ltb* t, const Ctx& C, Pass& P, const float* Z, int ldz, int B, cudaStream_t st) {
  net_fwd(C, t->g, P, Z, ldz, B, 0, 0, -1, nullptr, nullptr, st);
  moving_update_kernel<<<1, 64, 0, st>>>(t->moving, P.mean, P.inv, t->zd);
}
}  // namespace lt
}  // namespace bgm

extern "C" {

// BGM.train_disc_step gradients (bgm/base.py:190-240, gamma == 0) into group 1 = [dz | dx]
int bgm_ltb_disc_grad(bgm_ltb* t, const float* z_dev, const float* x_dev, int bs, float eps_z, float eps_x,
                      const float* noise_dev, float* losses_dev, void* stream) {
  using namespace bgm;
  using namespace bgm::lt;
  (void)eps_z; (void)eps_x;     // only the gradient penalty (gamma != 0) reads the interpolation draws
  if (!t || !z_dev || !x_dev || !noise_dev || !losses_dev) return fail(BGM_ERR_ARG, "bgm_ltb_disc_grad: null argument");
  if (bs < 2) return fail(BGM_ERR_ARG, "bgm_ltb_disc_grad: batch size must be >= 2");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = ensure_arena_bytes(t->arena, ltb_step_bytes(t, bs));
  if (rc) return rc;
  Arena& ar = t->arena;
  const Ctx C = ctx_of(t);
  const int B = bs, xd = t->xd, zd = t->zd, sm = t->sm_count;
  zero(t->grad[1], t->n1, sm, st);
  Pass eA, gA;
  net_fwd(C, t->e, eA, x_dev, xd, B, 0, 0, -1, nullptr, nullptr, st);            // z_ = e(x) (:203)
  g_fwd_train(t, C, gA, z_dev, zd, B, st);                                        // (:207)
  float* xgen = ar.get<float>((size_t)B * xd);
  reparam_kernel<<<grid_for((long long)B * xd, sm), 256, 0, st>>>(gA.out(), noise_dev, B, xd, xgen);   // x_ (:208)
  const float* thz = t->theta[1];
  const float* thx = t->theta[1] + t->dx_base;
  DiscPass Dz, Dz_, Dx, Dx_;
  disc_fwd(C, t->dz, thz, Dz, z_dev, B, st);
  disc_fwd(C, t->dz, thz, Dz_, eA.out(), B, st);
  disc_fwd(C, t->dx, thx, Dx, x_dev, B, st);
  disc_fwd(C, t->dx, thx, Dx_, xgen, B, st);
  float* dd = ar.get<float>(4 * (size_t)B);
  bgm_disc_loss_kernel<<<1, 256, 0, st>>>(Dz.d, Dz_.d, Dx.d, Dx_.d, B, dd, dd + B, dd + 2 * B, dd + 3 * B, losses_dev);
  disc_bwd(C, t->dz, thz, t->grad[1], Dz, dd, B, nullptr, st);
  disc_bwd(C, t->dz, thz, t->grad[1], Dz_, dd + B, B, nullptr, st);
  disc_bwd(C, t->dx, thx, t->grad[1] + t->dx_base, Dx, dd + 2 * B, B, nullptr, st);
  disc_bwd(C, t->dx, thx, t->grad[1] + t->dx_base, Dx_, dd + 3 * B, B, nullptr, st);
  return arena_ok(t->arena, "bgm_ltb_disc_grad");
}

// BGM.train_gen_step gradients (bgm/base.py:247-285) into group 0 = [g | e]
int bgm_ltb_gen_grad(bgm_ltb* t, const float* z_dev, const float* x_dev, int bs, const float* noise1_dev,
                     const float* noise2_dev, float* losses_dev, void* stream) {
  using namespace bgm;
  using namespace bgm::lt;
  if (!t || !z_dev || !x_dev || !noise1_dev || !noise2_dev || !losses_dev) return fail(BGM_ERR_ARG, "bgm_ltb_gen_grad: null argument");
  if (bs < 2) return fail(BGM_ERR_ARG, "bgm_ltb_gen_grad: batch size must be >= 2");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = ensure_arena_bytes(t->arena, ltb_step_bytes(t, bs));
  if (rc) return rc;
  Arena& ar = t->arena;
  const Ctx C = ctx_of(t);
  const int B = bs, xd = t->xd, zd = t->zd, sm = t->sm_count;
  zero(t->grad[0], t->n0, sm, st);
  Pass gA, gB, eA, eB;
  g_fwd_train(t, C, gA, z_dev, zd, B, st);                                        // :258
  float* x1 = ar.get<float>((size_t)B * xd);
  reparam_kernel<<<grid_for((long long)B * xd, sm), 256, 0, st>>>(gA.out(), noise1_dev, B, xd, x1);    // x_ :259
  net_fwd(C, t->e, eA, x_dev, xd, B, 0, 0, -1, nullptr, nullptr, st);            // z_ :262
  net_fwd(C, t->e, eB, x1, xd, B, 0, 0, -1, nullptr, nullptr, st);               // z__ :264
  g_fwd_train(t, C, gB, eA.out(), zd, B, st);                                     // :266
  float* x2 = ar.get<float>((size_t)B * xd);
  reparam_kernel<<<grid_for((long long)B * xd, sm), 256, 0, st>>>(gB.out(), noise2_dev, B, xd, x2);    // x__ :267
  const float* thz = t->theta[1];
  const float* thx = t->theta[1] + t->dx_base;
  DiscPass Dx_, Dz_;
  disc_fwd(C, t->dx, thx, Dx_, x1, B, st);                                         // :269
  disc_fwd(C, t->dz, thz, Dz_, eA.out(), B, st);                                   // :270
  BgmGenLossArgs L;
  memset(&L, 0, sizeof(L));
  L.B = B; L.xd = xd; L.zd = zd; L.alpha = t->alpha;
  L.x = x_dev; L.z = z_dev; L.out1 = gA.out(); L.x2 = x2; L.z2 = eB.out(); L.dx_ = Dx_.d; L.dz_ = Dz_.d;
  L.dX2 = ar.get<float>((size_t)B * xd); L.dZ2 = ar.get<float>((size_t)B * zd); L.ddx = ar.get<float>(B); L.ddz = ar.get<float>(B);
  L.dOut1 = ar.get<float>((size_t)B * 2 * xd);
  L.losses = losses_dev;
  bgm_gen_loss_kernel<<<1, 256, 0, st>>>(L);
  float* dOut2 = ar.get<float>((size_t)B * 2 * xd);
  float* dZ_ = ar.get<float>((size_t)B * zd);          // gradient w.r.t. z_ = e(x)
  float* dX1 = ar.get<float>((size_t)B * xd);          // gradient w.r.t. x_
  zero(dZ_, (long long)B * zd, sm, st);
  zero(dX1, (long long)B * xd, sm, st);
  reparam_bwd_kernel<<<grid_for((long long)B * xd, sm), 256, 0, st>>>(L.dX2, gB.out(), noise2_dev, B, xd, dOut2, 0);
  net_bwd(C, t->g, gB, dOut2, true, dZ_, zd, true, st);
  disc_bwd(C, t->dz, thz, nullptr, Dz_, L.ddz, B, dZ_, st);
  net_bwd(C, t->e, eB, L.dZ2, true, dX1, xd, true, st);
  disc_bwd(C, t->dx, thx, nullptr, Dx_, L.ddx, B, dX1, st);
  reparam_bwd_kernel<<<grid_for((long long)B * xd, sm), 256, 0, st>>>(dX1, gA.out(), noise1_dev, B, xd, L.dOut1, 1);
  net_bwd(C, t->g, gA, L.dOut1, true, nullptr, 0, false, st);
  net_bwd(C, t->e, eA, dZ_, true, nullptr, 0, false, st);
  return arena_ok(t->arena, "bgm_ltb_gen_grad");
}

int bgm_ltb_set_iter(bgm_ltb* t, float lr_theta, float lr_z) {
  using namespace bgm;
  if (!t) return fail(BGM_ERR_ARG, "bgm_ltb_set_iter: null trainer");
  t->lr_theta = lr_theta; t->lr_z = lr_z;
  t->step_g = 0; t->step_z = 0;
  BGM_CUDA_OK(cudaMemset(t->m_it, 0, sizeof(float) * (size_t)t->n_g));
  BGM_CUDA_OK(cudaMemset(t->v_it, 0, sizeof(float) * (size_t)t->n_g));
  return 0;
}

// update_g_net (bgm/base.py:145-164) on the rows idx_dev
int bgm_ltb_iter_g(bgm_ltb* t, const float* zt_dev, const float* x_dev, const int* idx_dev, int bs, int apply,
                   float grad_scale, float* losses_dev, void* stream) {
  using namespace bgm;
  using namespace bgm::lt;
  if (!t || !zt_dev || !x_dev || !idx_dev || !losses_dev) return fail(BGM_ERR_ARG, "bgm_ltb_iter_g: null argument");
  if (bs < 1) return fail(BGM_ERR_ARG, "bgm_ltb_iter_g: batch size must be >= 1");
  cudaStream_t st = (cudaStream_t)stream;
  const int B = bs, xd = t->xd, zd = t->zd, sm = t->sm_count;
  if (apply != 2) {
    int rc = ensure_arena_bytes(t->arena, ltb_step_bytes(t, bs));
    if (rc) return rc;
    Arena& ar = t->arena;
    const Ctx C = ctx_of(t);
    zero(t->grad[0], t->n_g, sm, st);
    float* zb = ar.get<float>((size_t)B * zd);
    float* xb = ar.get<float>((size_t)B * xd);
    gather_cols_kernel<<<grid_for((long long)B * zd, sm), 256, 0, st>>>(zt_dev, zd, idx_dev, 0, zd, B, zb, zd, 0);
    gather_cols_kernel<<<grid_for((long long)B * xd, sm), 256, 0, st>>>(x_dev, xd, idx_dev, 0, xd, B, xb, xd, 0);
    Pass gP;
    g_fwd_train(t, C, gP, zb, zd, B, st);
    float* dOut = ar.get<float>((size_t)B * 2 * xd);
    bgm_nll_kernel<<<1, 256, 0, st>>>(xb, gP.out(), B, xd, dOut, losses_dev);
    net_bwd(C, t->g, gP, dOut, true, nullptr, 0, false, st);
    rc = arena_ok(t->arena, "bgm_ltb_iter_g");
    if (rc) return rc;
  }
  if (apply != 0) {
    t->step_g += 1;
    const float lr_t = lr_t_of(t->lr_theta, 0.9, 0.99, t->step_g);
    lt::adam_kernel<<<grid_for(t->n_g, sm), 256, 0, st>>>(t->theta[0], t->grad[0], t->m_it, t->v_it, t->n_g, lr_t, 0.9f, 0.99f, 1e-7f,
                                                       grad_scale, nullptr);
    BGM_CUDA_OK(cudaGetLastError());
  }
  return 0;
}

// update_latent_variable_sgd (bgm/base.py:167-187) + the write-back of :410-413
int bgm_ltb_iter_latent(bgm_ltb* t, float* zt_dev, const float* x_dev, const int* idx_dev, int bs, float* loss_dev,
                        float* gz_out_dev, void* stream) {
  using namespace bgm;
  using namespace bgm::lt;
  if (!t || !zt_dev || !x_dev || !idx_dev || !loss_dev) return fail(BGM_ERR_ARG, "bgm_ltb_iter_latent: null argument");
  if (bs < 1) return fail(BGM_ERR_ARG, "bgm_ltb_iter_latent: batch size must be >= 1");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = ensure_arena_bytes(t->arena, ltb_step_bytes(t, bs));
  if (rc) return rc;
  Arena& ar = t->arena;
  const Ctx C = ctx_of(t);
  const int B = bs, xd = t->xd, zd = t->zd, sm = t->sm_count;
  float* zb = ar.get<float>((size_t)B * zd);
  float* xb = ar.get<float>((size_t)B * xd);
  gather_cols_kernel<<<grid_for((long long)B * zd, sm), 256, 0, st>>>(zt_dev, zd, idx_dev, 0, zd, B, zb, zd, 0);
  gather_cols_kernel<<<grid_for((long long)B * xd, sm), 256, 0, st>>>(x_dev, xd, idx_dev, 0, xd, B, xb, xd, 0);
  Pass gP;
  g_fwd_train(t, C, gP, zb, zd, B, st);
  float* dOut = ar.get<float>((size_t)B * 2 * xd);
  float* losses = t->scratch;
  zero(losses, 8, sm, st);
  bgm_nll_kernel<<<1, 256, 0, st>>>(xb, gP.out(), B, xd, dOut, losses);
  float* dZ = ar.get<float>((size_t)B * zd);
  gather_cols_kernel<<<grid_for((long long)B * zd, sm), 256, 0, st>>>(zb, zd, nullptr, 0, zd, B, dZ, zd, 0);
  prior_scale_kernel<<<1, 256, 0, st>>>(dZ, B * zd, 1.f / (float)B, losses);        // loss_px_z + loss_prior_z (:179)
  net_bwd(C, t->g, gP, dOut, false, dZ, zd, true, st);
  BGM_CUDA_OK(cudaMemcpyAsync(loss_dev, losses, sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (gz_out_dev) BGM_CUDA_OK(cudaMemcpyAsync(gz_out_dev, dZ, sizeof(float) * (size_t)B * zd, cudaMemcpyDeviceToDevice, st));
  t->step_z += 1;
  const float lr_t = lr_t_of(t->lr_z, 0.9, 0.99, t->step_z);
  fresh_adam_rows_kernel<<<grid_for((long long)B * zd, sm), 256, 0, st>>>(zt_dev, idx_dev, dZ, B, zd, lr_t, 0.9f, 0.99f, 1e-7f);
  return arena_ok(t->arena, "bgm_ltb_iter_latent");
}

// evaluate(use_x_sd=False) (bgm/base.py:446-471): sum over rows and columns of (x - mu(z))^2, generator in inference mode
int bgm_ltb_evaluate(bgm_ltb* t, const float* zt_dev, const float* x_dev, int n, double* sum_dev, void* stream) {
  using namespace bgm;
  using namespace bgm::lt;
  if (!t || !zt_dev || !x_dev || !sum_dev || n < 1) return fail(BGM_ERR_ARG, "bgm_ltb_evaluate: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int chunk = std::min(n, 8192), xd = t->xd, zd = t->zd, sm = t->sm_count;
  int rc = ensure_arena_bytes(t->arena, ltb_step_bytes(t, chunk));
  if (rc) return rc;
  const Ctx C = ctx_of(t);
  zero(reinterpret_cast<float*>(sum_dev), 2, sm, st);
  float* stats = t->scratch + 8;         // zd <= 16 in practice; scratch holds 64 floats
  if (2 * zd > 56) return fail(BGM_ERR_UNSUPPORTED, "bgm_ltb_evaluate: z_dim <= 28");
  moving_to_stats_kernel<<<1, 64, 0, st>>>(t->moving, zd, stats, stats + zd);
  for (int r0 = 0; r0 < n; r0 += chunk) {
    const int B = std::min(chunk, n - r0);
    t->arena.used = 0;
    Pass gP;
    net_fwd(C, t->g, gP, zt_dev + (size_t)r0 * zd, zd, B, 0, r0, -1, stats, stats + zd, st);
    sq_err_kernel<<<grid_for((long long)B * xd, sm), 256, 0, st>>>(x_dev + (size_t)r0 * xd, xd, gP.out(), 2 * xd, B, xd, 0, sum_dev);
  }
  return arena_ok(t->arena, "bgm_ltb_evaluate");
}

// e_net(x) for all n rows -> z_out_dev (n, z_dim)  (`data_z_init = self.e_net(data)`, bgm/base.py:388)
int bgm_ltb_encode(bgm_ltb* t, const float* x_dev, int n, float* z_out_dev, void* stream) {
  using namespace bgm;
  using namespace bgm::lt;
  if (!t || !x_dev || !z_out_dev || n < 1) return fail(BGM_ERR_ARG, "bgm_ltb_encode: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int chunk = std::min(n, 8192), xd = t->xd, zd = t->zd;
  int rc = ensure_arena_bytes(t->arena, ltb_step_bytes(t, chunk));
  if (rc) return rc;
  const Ctx C = ctx_of(t);
  for (int r0 = 0; r0 < n; r0 += chunk) {
    const int B = std::min(chunk, n - r0);
    t->arena.used = 0;
    Pass eP;
    net_fwd(C, t->e, eP, x_dev + (size_t)r0 * xd, xd, B, 0, r0, -1, nullptr, nullptr, st);
    BGM_CUDA_OK(cudaMemcpyAsync(z_out_dev + (size_t)r0 * zd, eP.out(), sizeof(float) * (size_t)B * zd, cudaMemcpyDeviceToDevice, st));
  }
  return arena_ok(t->arena, "bgm_ltb_encode");
}

// Conditional prior rows for bgm_mh_args.prior_dev / bgm_causal_logpost_cond: prior_net evaluated on the
// n_segments one-hot inputs (a table of n_segments x (zd+1)), gathered by each row's segment.
int bgm_causal_prior_rows(const bgm_net_desc* prior_net, int n_segments, int zd, const int* seg_dev, int n,
                          float* prior_dev, int ldprior, void* stream) {
  using namespace bgm;
  using namespace bgm::lt;
  if (!prior_net || !seg_dev || !prior_dev || n < 1 || n_segments < 1 || zd < 1 || ldprior < zd + 1)
    return fail(BGM_ERR_ARG, "bgm_causal_prior_rows: bad argument");
  const int L = prior_net->n_layers;
  if (L < 1 || prior_net->dims[0] != n_segments || prior_net->dims[L] != zd + 1)
    return fail(BGM_ERR_ARG, "bgm_causal_prior_rows: prior_net must map n_segments -> zd + 1");
  cudaStream_t st = (cudaStream_t)stream;
  size_t n_par = 0;
  int wmax = n_segments;
  for (int l = 0; l < L; ++l) {
    n_par += (size_t)prior_net->dims[l] * prior_net->dims[l + 1] + prior_net->dims[l + 1];
    wmax = std::max(wmax, prior_net->dims[l + 1]);
  }
  float* buf = nullptr;
  const size_t act = (size_t)n_segments * wmax;
  BGM_CUDA_OK(cudaMalloc(&buf, sizeof(float) * (n_par + 2 * act)));
  cudaError_t e = cudaMemcpyAsync(buf, prior_net->params, sizeof(float) * n_par, cudaMemcpyHostToDevice, st);
  float* a = buf + n_par;
  float* b = a + act;
  if (e == cudaSuccess) {
    eye_kernel<<<4, 256, 0, st>>>(a, n_segments);
    const float* w = buf;
    for (int l = 0; l < L; ++l) {
      const int K = prior_net->dims[l], N = prior_net->dims[l + 1];
      dense_fwd(a, K, w, nullptr, nullptr, nullptr, w + (size_t)K * N,
                                                                            n_segments, K, N, b, N, l + 1 < L ? 1 : 0, st);
      w += (size_t)K * N + N;
      std::swap(a, b);
    }
    prior_rows_kernel<<<grid_for((long long)n * (zd + 1), 148), 256, 0, st>>>(a, zd, seg_dev, n_segments, n, prior_dev, ldprior);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(buf);
  if (e != cudaSuccess) return fail(BGM_ERR_CUDA, cudaGetErrorString(e));
  return 0;
}

}  // extern "C"
