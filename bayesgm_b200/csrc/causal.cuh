// CausalBGM hot path: packed g/f/h nets, log-posterior, persistent random-walk MH
// sampler and the effect (ITE / ADRF) kernel.
//
// Replaces causalbgm/base.py:765-817 (get_log_posterior), :820-904
// (metropolis_hastings_sampler) and :671-763 (infer_from_latent_posterior) of the
// reference.  One warp carries 32 observations (rows) through all T iterations; the
// weight image of the three nets sits in shared memory for the whole launch, the
// activations in the warp's private shared-memory slice, the chain state in
// registers (one row per lane); the covariate rows are re-read from L2 once per
// iteration.
#pragma once
#include <vector>

#include "common.cuh"

namespace bgm {

constexpr int MAX_OPS = 48;
constexpr int MAX_WARPS = 8;
constexpr int SCR_SLOTS = 2;  // reused per net: [sse | mu] and [sigma raw]
constexpr int ACT_ROWS = 64;  // widest hidden layer
constexpr int IMG_PAD = 64;   // floats after the image that the last prefetch may touch

struct CausalProgram {
  int n_ops;
  int g_end, f_end, h_end;     // op ranges: g [0,g_end) f [g_end,f_end) h [f_end,h_end)
  int f_img_begin, f_img_end;  // float range of the f-net tiles in the image
  int zd, kin, p, binary;
  int p_data;                  // columns of the data the SSE tiles read: p, or the projected width
  int proj;                    // 1: covariate likelihood through the QR projection (see bgm_b200.cu)
  float s2v, s2x, s2y;         // fixed variances, < 0 = learned head
  int image_floats;            // padded by IMG_PAD
  int per_warp_floats;
  TileOp ops[MAX_OPS];
};

// Shared-memory slice of one warp (floats): act[64][32] | zin[kin][32] | scr[2][32]
struct WarpSmem {
  float* act;
  float* zin;
  float* scr;
};
__device__ __forceinline__ WarpSmem warp_smem(float* base, const CausalProgram& P) {
  WarpSmem s;
  s.act = base;
  s.zin = s.act + ACT_ROWS * TILE_ROWS;
  s.scr = s.zin + P.kin * TILE_ROWS;
  return s;
}

// Runs one column tile: init accumulators (bias, or bias - data for the SSE
// epilogue), MAC over k, then the epilogue.
template <int C>
__device__ __forceinline__ void run_tile(const TileOp& op, const float* __restrict__ wimg,
                                         const WarpSmem& S, const float* __restrict__ v, int ldv,
                                         int p, int row0, int n, int rg, int cg,
                                         float (&sse)[RPT]) {
  const float* in = op.src ? S.act : S.zin;
  const float* w = wimg + op.w_off;
  const float* bias = wimg + op.b_off;
  float acc[RPT][C];
#pragma unroll
  for (int j = 0; j < C; ++j) {
    float bj = bias[ColMap<C>::col(cg, j)];
#pragma unroll
    for (int i = 0; i < RPT; ++i) acc[i][j] = bj;
  }
  if (op.epi == EPI_SSE) {
    // acc starts at bias - v so that after the MAC it holds mu - v.
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      int r = row0 + row_of(rg, i);
      r = r < n ? r : n - 1;
      const float* vr = v + (size_t)r * ldv + op.c0;
      if constexpr (C == 1) {
        int c = op.c0 + cg;
        float t = (c < p) ? __ldg(vr + cg) : 0.f;
        acc[i][0] -= t;
      } else {
#pragma unroll
        for (int h = 0; h < C / 4; ++h) {
          int cl = h * 32 + cg * 4;  // tile-local first column of this float4
          int c = op.c0 + cl;
          float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
          if (c < p) t = __ldg(reinterpret_cast<const float4*>(vr + cl));
          acc[i][h * 4 + 0] -= t.x;
          acc[i][h * 4 + 1] -= (c + 1 < p) ? t.y : 0.f;
          acc[i][h * 4 + 2] -= (c + 2 < p) ? t.z : 0.f;
          acc[i][h * 4 + 3] -= (c + 3 < p) ? t.w : 0.f;
        }
      }
    }
  }
  tile_mac<C>(in, w, op.kp, rg, cg, acc);
  if (op.epi == EPI_ACT || op.epi == EPI_LIN) {
    __syncwarp();  // every lane is done reading the buffer we overwrite
    const bool act = op.epi == EPI_ACT;
#pragma unroll
    for (int j = 0; j < C; ++j) {
      const int c = ColMap<C>::col(cg, j);
      const int chunk = (rg ^ act_swz(c)) << 2;
      float4 lo, hi;
      lo.x = act ? leaky(acc[0][j]) : acc[0][j];
      lo.y = act ? leaky(acc[1][j]) : acc[1][j];
      lo.z = act ? leaky(acc[2][j]) : acc[2][j];
      lo.w = act ? leaky(acc[3][j]) : acc[3][j];
      hi.x = act ? leaky(acc[4][j]) : acc[4][j];
      hi.y = act ? leaky(acc[5][j]) : acc[5][j];
      hi.z = act ? leaky(acc[6][j]) : acc[6][j];
      hi.w = act ? leaky(acc[7][j]) : acc[7][j];
      *reinterpret_cast<float4*>(S.act + c * TILE_ROWS + chunk) = lo;
      *reinterpret_cast<float4*>(S.act + c * TILE_ROWS + (chunk ^ 16)) = hi;
    }
    __syncwarp();
  } else if (op.epi == EPI_SSE) {
#pragma unroll
    for (int i = 0; i < RPT; ++i)
#pragma unroll
      for (int j = 0; j < C; ++j) sse[i] = fmaf(acc[i][j], acc[i][j], sse[i]);
  } else {  // EPI_OUT: raw outputs to scratch slots c0 + col
#pragma unroll
    for (int j = 0; j < C; ++j) {
      int c = ColMap<C>::col(cg, j);
      if (c < op.nvalid) {
#pragma unroll
        for (int i = 0; i < RPT; ++i) S.scr[(op.c0 + c) * TILE_ROWS + row_of(rg, i)] = acc[i][j];
      }
    }
  }
}

__device__ __forceinline__ void run_one(const TileOp& op, const float* __restrict__ wimg,
                                        const WarpSmem& S, const float* __restrict__ v, int ldv,
                                        int p, int row0, int n, int rg, int cg, float (&sse)[RPT]) {
  if (op.ctype == 0) run_tile<8>(op, wimg, S, v, ldv, p, row0, n, rg, cg, sse);
  else if (op.ctype == 1) run_tile<4>(op, wimg, S, v, ldv, p, row0, n, rg, cg, sse);
  else run_tile<1>(op, wimg, S, v, ldv, p, row0, n, rg, cg, sse);
}

// Log-posterior of the 32 rows whose z (and x) sit in S.zin; returns the value of
// this lane's row.  causalbgm/base.py:779-816.  ONE loop over the tile program (so
// each tile routine exists once in the instruction stream); the three losses are
// assembled as each net's last tile finishes, so two scratch rows suffice.
__device__ __forceinline__ float eval_logpost(const CausalProgram& P, const float* __restrict__ wimg,
                                              const WarpSmem& S, const float* __restrict__ v,
                                              int ldv, int row0, int n, int lane, float x_l,
                                              float y_l, float r0_l,
                                              const float* __restrict__ prior_row = nullptr) {
  const int rg = lane >> 3, cg = lane & 7;
  float sse[RPT];
#pragma unroll
  for (int i = 0; i < RPT; ++i) sse[i] = 0.f;
  float loss = 0.f;  // loss_pv, then + loss_py after the f-net; loss_px joins at the end
  float result = 0.f;
#pragma unroll 1
  for (int o = 0; o < P.n_ops; ++o) {
    const TileOp& op = P.ops[o];
    run_one(op, wimg, S, v, ldv, P.p_data, row0, n, rg, cg, sse);
    if (op.post == POST_NONE) continue;
    if (op.post == POST_G) {  // sum_j (v_j - mu_j)^2 and the sigma_v head
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        float s = sse[i];
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        if (cg == 0) S.scr[row_of(rg, i)] = s;
      }
      __syncwarp();
      const float sse_l = S.scr[lane] + r0_l;   // r0: part of v outside the last layer's row space
      const float s2v = P.s2v >= 0.f ? P.s2v : softplus_f(S.scr[TILE_ROWS + lane]) + 1e-6f;
      loss = sse_l / (2.f * s2v) + ((float)P.p * logf(s2v)) / 2.f;                 // :800-801
    } else if (op.post == POST_F) {  // outcome model
      __syncwarp();
      const float mu_y = S.scr[lane];
      const float s2y = P.s2y >= 0.f ? P.s2y : softplus_f(S.scr[TILE_ROWS + lane]) + 1e-6f;
      const float dy = y_l - mu_y;
      result = (dy * dy) / (2.f * s2y) + logf(s2y) / 2.f;                         // :809-810
    } else {  // POST_H: treatment model, prior, total
      __syncwarp();
      const float mu_x = S.scr[lane];
      float loss_px;
      if (P.binary) {                                                             // :804
        loss_px = fmaxf(mu_x, 0.f) - mu_x * x_l + log1pf(expf(-fabsf(mu_x)));
      } else {                                                                    // :806-807
        const float s2x = P.s2x >= 0.f ? P.s2x : softplus_f(S.scr[TILE_ROWS + lane]) + 1e-6f;
        const float d = x_l - mu_x;
        loss_px = (d * d) / (2.f * s2x) + logf(s2x) / 2.f;
      }
      float prior = 0.f;
      if (prior_row) {  // z | u ~ N(mu(u), sigma^2(u) I), causalbgm/identifiable.py:540-548
        for (int d = 0; d < P.zd; ++d) {
          const float dz = S.zin[act_idx(d, lane)] - prior_row[d];
          prior = fmaf(dz, dz, prior);
        }
        const float s2z = prior_row[P.zd];
        prior = prior / (2.f * s2z) + ((float)P.zd * logf(s2z)) / 2.f;
      } else {
        for (int d = 0; d < P.zd; ++d) {
          const float z = S.zin[act_idx(d, lane)];
          prior = fmaf(z, z, prior);
        }
        prior *= 0.5f;                                                            // :812
      }
      result = -(((loss + loss_px) + result) + prior);                            // :814-816
    }
    __syncwarp();  // scratch rows are reused by the next net
  }
  return result;
}

struct MhDev {
  bgm_mh_args a;
  int mode;     // 0: MH, 1: log-posterior of z_state only (out -> lp_state)
  int nchunks;  // iteration chunks per tile (dynamic scheduling granularity)
};

// ZMAX: compile-time bound on zd; the chain state z_cur[ZMAX] of the lane's row is
// held in registers.
template <int ZMAX>
__global__ void __launch_bounds__(MAX_WARPS * 32, 1)
causal_mh_kernel(const __grid_constant__ CausalProgram P, const float* __restrict__ image,
                 const __grid_constant__ MhDev D) {
  extern __shared__ __align__(128) float smem[];
  __shared__ uint64_t bar;
  bulk_load_to_smem(smem, image, (uint32_t)P.image_floats * 4u, &bar);
  const float* wimg = smem;

  const bgm_mh_args& A = D.a;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const WarpSmem S = warp_smem(smem + P.image_floats + warp * P.per_warp_floats, P);
  const int n = A.n, zd = P.zd;
  const int ntiles = (n + TILE_ROWS - 1) / TILE_ROWS;
  const double q_sd = A.q_sd_dev ? *A.q_sd_dev : 1.0;
  // the log-posterior of the initial state is evaluated as pseudo-iteration t_begin-1
  const bool need_init = !(A.init_mode == 0 && D.mode == 0);
  const int t_first = need_init ? A.t_begin - 1 : A.t_begin;
  const int t_last = D.mode == 1 ? A.t_begin : A.t_end;

  // Work units are (32-row tile, chunk of iterations), handed out by a global counter in
  // chunk-major order; chunk c of a tile waits for chunk c-1 of the same tile (progress
  // flag).  All CTAs are co-resident (grid <= #SMs) and units are claimed in increasing
  // order, so the unit a warp waits for is always owned by a running warp.
  const float* vdat = P.proj ? A.vproj_dev : A.v_dev;
  const int ldd = P.proj ? A.ldvproj : A.ldv;
  int* sched = A.sched_dev;
  const int n_iter = t_last - t_first;
  const int nchunks = D.nchunks;
  const int chunk_len = (n_iter + nchunks - 1) / nchunks;
  const long long total_units = (long long)ntiles * nchunks;
  for (;;) {
    long long u = 0;
    if (lane == 0) u = atomicAdd(reinterpret_cast<unsigned int*>(sched), 1u);
    u = __shfl_sync(0xffffffffu, u, 0);
    if (u >= total_units) break;
    const int chunk = (int)(u / ntiles);
    const int tile = (int)(u - (long long)chunk * ntiles);
    const int ta = t_first + chunk * chunk_len;
    const int tb = min(ta + chunk_len, t_last);
    if (chunk > 0) {
      if (lane == 0) {
        const volatile int* flag = sched + 1 + tile;
        while (*flag < chunk) __nanosleep(200);
        __threadfence();
      }
      __syncwarp();
    }
    const int row0 = tile * TILE_ROWS;
    const int row = row0 + lane;
    const bool valid = row < n;
    const int lrow = valid ? row : n - 1;
    const int nvalid_rows = min(TILE_ROWS, n - row0);
    const float x_l = A.x_dev[lrow], y_l = A.y_dev[lrow];
    const float r0_l = P.proj ? A.r0_dev[lrow] : 0.f;
    const int64_t grow = A.row_offset + lrow;
    const float* prior_row = A.prior_dev ? A.prior_dev + (size_t)lrow * A.ldprior : nullptr;
    float zc[ZMAX];
    // input buffer: rows [0,zd) proposal, row zd = x, remaining pad rows zero
    for (int k = zd; k < P.kin; ++k) S.zin[act_idx(k, lane)] = (k == zd) ? x_l : 0.f;
    // ---- initial state (:842) ----
    if (chunk == 0 && A.init_mode == 2) {
#pragma unroll
      for (int g = 0; g < ZMAX / 4; ++g) {
        if (g * 4 < zd) {
          float e[4];
          normal4(A.seed, grow, T_INIT, NOISE_PROPOSAL, g, e);
#pragma unroll
          for (int q = 0; q < 4; ++q) zc[g * 4 + q] = e[q];
        }
      }
    } else {
#pragma unroll
      for (int d = 0; d < ZMAX; ++d)
        if (d < zd) zc[d] = __ldcg(A.z_state_dev + (size_t)lrow * zd + d);
    }
    float lp_cur = (need_init && chunk == 0) ? 0.f : __ldcg(A.lp_state_dev + lrow);
    // ---- iterations (:860-898) ----
#pragma unroll 1
    for (int t = ta; t < tb; ++t) {
      const bool init_pass = t < A.t_begin;
      if (init_pass) {
#pragma unroll
        for (int d = 0; d < ZMAX; ++d)
          if (d < zd) S.zin[act_idx(d, lane)] = zc[d];
      } else if (A.eps_dev) {
        // proposal z' = z + float32(q_sd * eps) (:862): NumPy scales the unit normal by
        // q_sd in float64, .astype(float32) rounds once, then the float32 add.
        const float* e = A.eps_dev + ((size_t)t * n + lrow) * zd;
#pragma unroll
        for (int d = 0; d < ZMAX; ++d)
          if (d < zd) S.zin[act_idx(d, lane)] = __fadd_rn(zc[d], (float)(q_sd * (double)e[d]));
      } else {
#pragma unroll
        for (int g = 0; g < ZMAX / 4; ++g) {
          if (g * 4 < zd) {
            float e[4];
            normal4(A.seed, grow, (uint32_t)t, NOISE_PROPOSAL, g, e);
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (g * 4 + q < zd)
                S.zin[act_idx(g * 4 + q, lane)] = __fadd_rn(zc[g * 4 + q], (float)(q_sd * (double)e[q]));
          }
        }
      }
      __syncwarp();
      const float lp_prop = eval_logpost(P, wimg, S, vdat, ldd, row0, n, lane, x_l, y_l, r0_l, prior_row);
      if (init_pass) {
        lp_cur = lp_prop;
        continue;
      }
      // accept: u < exp(min(lp' - lp, 0))  (:868-870); a NaN ratio never accepts
      const float dlp = lp_prop - lp_cur;
      const float ratio = (dlp < 0.f) ? expf(dlp) : ((dlp >= 0.f) ? 1.f : __int_as_float(0x7fc00000));
      bool acc;
      if (A.u_dev) acc = A.u_dev[(size_t)t * n + lrow] < (double)ratio;
      else acc = uniform1(A.seed, grow, (uint32_t)t, NOISE_ACCEPT) < ratio;
      if (acc) {                                                                 // :871
#pragma unroll
        for (int d = 0; d < ZMAX; ++d)
          if (d < zd) zc[d] = S.zin[act_idx(d, lane)];
        lp_cur = lp_prop;
      }
      if (A.accept_mask_dev && valid) A.accept_mask_dev[(size_t)t * n + row] = acc ? 1 : 0;
      if (A.lp_trace_dev && valid) A.lp_trace_dev[(size_t)t * n + row] = lp_prop;
      if (A.accept_count_dev) {
        const unsigned b = __ballot_sync(0xffffffffu, acc && valid);
        if (lane == 0 && b) atomicAdd(A.accept_count_dev + t, __popc(b));
      }
      if (t >= A.burn_in && A.out_samples_dev) {                                 // :895-896
        // stage through the (idle) activation buffer so that the warp writes its
        // contiguous (rows x zd) block with full 128-byte lines
#pragma unroll
        for (int d = 0; d < ZMAX; ++d)
          if (d < zd) S.act[d * TILE_ROWS + lane] = zc[d];
        __syncwarp();
        float* dst = A.out_samples_dev + ((size_t)(t - A.burn_in) * n + row0) * zd;
        for (int idx = lane; idx < nvalid_rows * zd; idx += 32) {
          const int r = idx / zd, d = idx - r * zd;
          dst[idx] = S.act[d * TILE_ROWS + r];
        }
      }
      __syncwarp();
    }
    // ---- save state, publish the chunk ----
    if (valid) {
      if (D.mode == 0) {
#pragma unroll
        for (int d = 0; d < ZMAX; ++d)
          if (d < zd) A.z_state_dev[(size_t)row * zd + d] = zc[d];
      }
      A.lp_state_dev[row] = lp_cur;
    }
    __threadfence();
    __syncwarp();
    if (lane == 0) *reinterpret_cast<volatile int*>(sched + 1 + tile) = chunk + 1;
  }
}

// Effect kernel: f-net on kept states for each dose.  One warp tile = 32 rows of
// one kept sample.  infer_from_latent_posterior, causalbgm/base.py:671-763.
struct EffectDev {
  const float* z_samples;  // (n_keep, n, zd)
  int n_keep, n;
  const float* x_values;   // (n_x) ; binary: NULL -> {1, 0}
  int n_x;
  int sample_y;
  uint64_t seed;
  int64_t row_offset;
  const float* noise;      // optional injected N(0,1): (n_x, n_keep, n)
  double* adrf_sum;        // (n_x, n_keep)
  float* ite;              // (n_keep, n)
  float* heads;            // heads mode (bgm_causal_effect_heads): (n_keep*n, n_x, 2) = (mu_y, raw sigma head),
                           // no noise, no reduction
};

__global__ void __launch_bounds__(MAX_WARPS * 32, 1)
causal_effect_kernel(const __grid_constant__ CausalProgram P, const float* __restrict__ image,
                     const __grid_constant__ EffectDev E) {
  extern __shared__ __align__(128) float smem[];
  __shared__ uint64_t bar;
  const int img_floats = P.f_img_end - P.f_img_begin;
  bulk_load_to_smem(smem, image + P.f_img_begin, (uint32_t)img_floats * 4u, &bar);
  // tile ops carry offsets into the full image; rebase them onto the f-net copy
  const float* wimg = smem;
  const int rebase = P.f_img_begin;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warps = blockDim.x >> 5;
  const int rg = lane >> 3, cg = lane & 7;
  const WarpSmem S = warp_smem(smem + img_floats + IMG_PAD + warp * P.per_warp_floats, P);
  const int n = E.n, zd = P.zd;
  const int tiles_per_s = (n + TILE_ROWS - 1) / TILE_ROWS;
  const long long ntiles = (long long)tiles_per_s * E.n_keep;
  for (int k = zd; k < P.kin; ++k) S.zin[act_idx(k, lane)] = 0.f;
  for (long long tile = (long long)blockIdx.x * warps + warp; tile < ntiles;
       tile += (long long)gridDim.x * warps) {
    const int s = (int)(tile / tiles_per_s);
    const int row0 = (int)(tile - (long long)s * tiles_per_s) * TILE_ROWS;
    const int row = row0 + lane;
    const bool valid = row < n;
    const int lrow = valid ? row : n - 1;
    const int64_t grow = E.row_offset + lrow;
    const float* zs = E.z_samples + ((size_t)s * n + lrow) * zd;
    for (int d = 0; d < zd; ++d) S.zin[act_idx(d, lane)] = zs[d];
    float y_prev = 0.f;
    float nz4[4] = {0.f, 0.f, 0.f, 0.f};
    for (int j = 0; j < E.n_x; ++j) {
      const float xv = E.x_values ? E.x_values[j] : (j == 0 ? 1.f : 0.f);
      S.zin[act_idx(zd, lane)] = xv;
      __syncwarp();
      float sse[RPT] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
      for (int o = P.g_end; o < P.f_end; ++o) {
        TileOp op = P.ops[o];
        op.w_off -= rebase;
        op.b_off -= rebase;
        run_one(op, wimg, S, nullptr, 0, P.p, row0, n, rg, cg, sse);
      }
      __syncwarp();
      float y = S.scr[lane];                                                     // mu_y
      if (E.heads) {
        if (valid) {
          float2* hp = reinterpret_cast<float2*>(E.heads) + ((size_t)s * n + row) * E.n_x + j;
          *hp = make_float2(y, S.scr[TILE_ROWS + lane]);
        }
        __syncwarp();
        continue;
      }
      if (E.sample_y) {                                                          // :703-708
        const float s2 = P.s2y >= 0.f ? P.s2y : softplus_f(S.scr[TILE_ROWS + lane]) + 1e-6f;
        float e;
        if (E.noise) {
          e = E.noise[((size_t)j * E.n_keep + s) * n + lrow];
        } else {
          // doses 4g..4g+3 share one Philox block: component j & 3 of normal4(.., g)
          if ((j & 3) == 0) normal4(E.seed, grow, (uint32_t)s, NOISE_EFFECT, (uint32_t)(j >> 2), nz4);
          const int k4 = j & 3;
          e = k4 == 0 ? nz4[0] : (k4 == 1 ? nz4[1] : (k4 == 2 ? nz4[2] : nz4[3]));
        }
        y = fmaf(sqrtf(s2), e, y);
      }
      __syncwarp();
      if (P.binary) {                                                            // :731
        if (j == 0) y_prev = y;
        else if (valid) E.ite[(size_t)s * n + row] = y_prev - y;
      } else {                                                                   // :759
        float part = valid ? y : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (lane == 0) atomicAdd(E.adrf_sum + (size_t)j * E.n_keep + s, (double)part);
      }
    }
  }
}

// ---- memoised effect evaluation --------------------------------------------------------------
// A rejected proposal repeats the chain state, so at an acceptance rate a only a fraction ~a of the
// kept states of a row are distinct; f_net is deterministic, hence (mu_y, sigma_y) need to be
// evaluated once per DISTINCT state (bgm_causal_effect_heads on the compacted list) and every
// (kept state, row, dose) then only draws its own noise and accumulates (this kernel).  Results are
// identical to evaluating f_net at every kept state (infer_from_latent_posterior, :671-763).
// Index of the distinct states, all accesses coalesced (thread = row, sequential over the kept states):
// local[s*n + row] = number of distinct states among states 0..s of the row (>= 1; it changes exactly
// where state s differs bitwise from state s-1); rowtot[row] = local[(n_keep-1)*n + row].
template <int ZMAX>
__global__ void effect_rowscan_kernel(const float* __restrict__ z_samples, int n_keep, int n, int zd,
                                      int* __restrict__ local, int* __restrict__ rowtot) {
  for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < n; row += gridDim.x * blockDim.x) {
    float prev[ZMAX];
    int c = 0;
    for (int s = 0; s < n_keep; ++s) {
      const float* a = z_samples + ((size_t)s * n + row) * zd;
      int f = s == 0;
#pragma unroll
      for (int d = 0; d < ZMAX; ++d) {
        if (d < zd) {
          const float v = a[d];
          if (s > 0) f |= (__float_as_uint(v) != __float_as_uint(prev[d]));
          prev[d] = v;
        }
      }
      c += f;
      local[(size_t)s * n + row] = c;
    }
    rowtot[row] = c;
  }
}
// distinct states, compacted row by row: the states of `row` occupy zlist[rowend[row]-rowtot .. rowend[row])
// (rowend = inclusive scan of rowtot)
__global__ void effect_compact_kernel(const float* __restrict__ z_samples, int n_keep, int n, int zd,
                                      const int* __restrict__ local, const int* __restrict__ rowend,
                                      float* __restrict__ zlist) {
  for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < n; row += gridDim.x * blockDim.x) {
    const int base = rowend[row] - local[(size_t)(n_keep - 1) * n + row];
    int cprev = 0;
    for (int s = 0; s < n_keep; ++s) {
      const int c = local[(size_t)s * n + row];
      if (c != cprev) {
        const float* a = z_samples + ((size_t)s * n + row) * zd;
        float* o = zlist + (size_t)(base + c - 1) * zd;
        for (int d = 0; d < zd; ++d) o[d] = a[d];
        cprev = c;
      }
    }
  }
}
// Inclusive prefix sum of int32 (three small launches): per-block scans of 2048 elements, a
// single-block scan of the block totals, then the offsets are added.  scratch: ceil(n/2048) ints.
constexpr int SCAN_TILE = 2048;
__global__ void __launch_bounds__(256) scan_block_kernel(const int* __restrict__ in, int* __restrict__ out,
                                                         long long n, int* __restrict__ totals) {
  __shared__ int wsum[8];
  const long long base = (long long)blockIdx.x * SCAN_TILE + threadIdx.x * 8;
  int v[8], run = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    v[k] = (base + k < n) ? in[base + k] : 0;
    run += v[k];
    v[k] = run;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  int woff = 0;
  for (int w = 0; w < warp; ++w) woff += wsum[w];
  const int off = woff + inc - run;
#pragma unroll
  for (int k = 0; k < 8; ++k)
    if (base + k < n) out[base + k] = v[k] + off;
  if (threadIdx.x == 255) totals[blockIdx.x] = off + run;
}
__global__ void __launch_bounds__(1024) scan_totals_kernel(int* __restrict__ totals, int nblocks) {
  __shared__ int wsum[32];
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int b0 = 0; b0 < nblocks; b0 += 1024) {
    const int i = b0 + threadIdx.x;
    const int x = i < nblocks ? totals[i] : 0;
    int inc = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < warp; ++w) woff += wsum[w];
    const int carry = carry_s;
    if (i < nblocks) totals[i] = carry + woff + inc - x;   // exclusive offset of block i
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + woff + inc;
    __syncthreads();
  }
}
__global__ void __launch_bounds__(256) scan_add_kernel(int* __restrict__ out, long long n,
                                                       const int* __restrict__ totals) {
  const long long base = (long long)blockIdx.x * SCAN_TILE + threadIdx.x * 8;
  const int off = totals[blockIdx.x];
#pragma unroll
  for (int k = 0; k < 8; ++k)
    if (base + k < n) out[base + k] += off;
}
struct CombineDev {
  const float* heads;      // (n_distinct, n_x, 2)
  const int* local;        // (n_keep, n) per-row running count of distinct states
  const int* rowend;       // (n) inclusive scan of the per-row totals
  int n_keep, n, n_x, binary, sample_y;
  float s2y;               // fixed variance or < 0
  uint64_t seed;
  int64_t row_offset;
  const float* noise;      // optional injected N(0,1): (n_x, n_keep, n)
  double* adrf_sum;        // (n_x, n_keep)
  float* ite;              // (n_keep, n)
};
// one warp = 32 consecutive rows of one kept state
__global__ void __launch_bounds__(256) effect_combine_kernel(const __grid_constant__ CombineDev C) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int tiles_per_s = (C.n + 31) / 32;
  const long long ntiles = (long long)tiles_per_s * C.n_keep;
  for (long long tile = warp0; tile < ntiles; tile += nwarps) {
    const int s = (int)(tile / tiles_per_s);
    const int row = (int)(tile - (long long)s * tiles_per_s) * 32 + lane;
    const bool valid = row < C.n;
    const int lrow = valid ? row : C.n - 1;
    const int64_t grow = C.row_offset + lrow;
    const int ref = C.rowend[lrow] - C.local[(size_t)(C.n_keep - 1) * C.n + lrow] + C.local[(size_t)s * C.n + lrow] - 1;
    const float2* h = reinterpret_cast<const float2*>(C.heads) + (size_t)ref * C.n_x;
    float nz4[4] = {0.f, 0.f, 0.f, 0.f};
    float y_prev = 0.f;
    for (int j = 0; j < C.n_x; ++j) {
      const float2 mr = __ldg(h + j);
      float y = mr.x;
      if (C.sample_y) {
        const float s2 = C.s2y >= 0.f ? C.s2y : softplus_f(mr.y) + 1e-6f;
        float e;
        if (C.noise) {
          e = C.noise[((size_t)j * C.n_keep + s) * C.n + lrow];
        } else {
          if ((j & 3) == 0) normal4(C.seed, grow, (uint32_t)s, NOISE_EFFECT, (uint32_t)(j >> 2), nz4);
          const int k4 = j & 3;
          e = k4 == 0 ? nz4[0] : (k4 == 1 ? nz4[1] : (k4 == 2 ? nz4[2] : nz4[3]));
        }
        y = fmaf(sqrtf(s2), e, y);
      }
      if (C.binary) {
        if (j == 0) y_prev = y;
        else if (valid) C.ite[(size_t)s * C.n + row] = y_prev - y;
      } else {
        float part = valid ? y : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (lane == 0) atomicAdd(C.adrf_sum + (size_t)j * C.n_keep + s, (double)part);
      }
    }
  }
}

// The same reductions, row-major: thread = row, block = 128 rows x one group of 4 doses, the thread walks
// its row's kept states in order.  A rejected proposal repeats the state, so the (mu, sigma) of the group
// stay in registers until the running count `local` moves (each head is read once, 32 contiguous bytes),
// where the kernel above gathers 8 bytes per (kept state, row, dose) at a ~20 KB stride across the warp.
// Same Philox draws (normal4 of (row, s, group)), same float warp sums; per-warp partials of a segment of
// kept states are parked in shared memory and added to the float64 accumulators once per block in warp
// order (4x fewer atomics).
constexpr int COMBINE_SEG = 256;
__global__ void __launch_bounds__(128) effect_combine_rows_kernel(const __grid_constant__ CombineDev C) {
  __shared__ float part[4][COMBINE_SEG][4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_groups = (C.n_x + 3) >> 2;
  const int group = blockIdx.x % n_groups;
  const int row = (blockIdx.x / n_groups) * 128 + threadIdx.x;
  const bool valid = row < C.n;
  const int lrow = valid ? row : C.n - 1;
  const int64_t grow = C.row_offset + lrow;
  const int j0 = group * 4;
  const int nj = min(4, C.n_x - j0);
  const int base = C.rowend[lrow] - C.local[(size_t)(C.n_keep - 1) * C.n + lrow] - 1;
  int cur = -1;
  float mu[4] = {0.f, 0.f, 0.f, 0.f}, sd[4] = {0.f, 0.f, 0.f, 0.f};
  for (int s0 = 0; s0 < C.n_keep; s0 += COMBINE_SEG) {
    const int s1 = min(s0 + COMBINE_SEG, C.n_keep);
    for (int s = s0; s < s1; ++s) {
      const int loc = C.local[(size_t)s * C.n + lrow];
      if (loc != cur) {
        cur = loc;
        const float2* h = reinterpret_cast<const float2*>(C.heads) + (size_t)(base + loc) * C.n_x + j0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (k < nj) {
            const float2 mr = __ldg(h + k);
            mu[k] = mr.x;
            sd[k] = C.sample_y ? sqrtf(C.s2y >= 0.f ? C.s2y : softplus_f(mr.y) + 1e-6f) : 0.f;
          }
        }
      }
      float y[4] = {mu[0], mu[1], mu[2], mu[3]};
      if (C.sample_y) {
        float e[4];
        if (C.noise) {
#pragma unroll
          for (int k = 0; k < 4; ++k) e[k] = k < nj ? C.noise[((size_t)(j0 + k) * C.n_keep + s) * C.n + lrow] : 0.f;
        } else {
          normal4(C.seed, grow, (uint32_t)s, NOISE_EFFECT, (uint32_t)group, e);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) y[k] = fmaf(sd[k], e[k], y[k]);
      }
      if (C.binary) {
        if (valid) C.ite[(size_t)s * C.n + row] = y[0] - y[1];
        continue;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float v = valid ? y[k] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        y[k] = v;
      }
      if (lane == 0) *reinterpret_cast<float4*>(part[warp][s - s0]) = make_float4(y[0], y[1], y[2], y[3]);
    }
    if (C.binary) continue;
    __syncthreads();
    for (int i = threadIdx.x; i < (s1 - s0) * 4; i += 128) {
      const int ss = i >> 2, k = i & 3;
      if (k < nj) {
        const double tot = (((double)part[0][ss][k] + (double)part[1][ss][k]) + (double)part[2][ss][k]) + (double)part[3][ss][k];
        atomicAdd(C.adrf_sum + (size_t)(j0 + k) * C.n_keep + s0 + ss, tot);
      }
    }
    __syncthreads();
  }
}

// 1-thread kernel: the q_sd adaptation rule, causalbgm/base.py:880-890.
__global__ void mh_adapt_qsd_kernel(const int* __restrict__ accept_count, int t, int window,
                                    long long n_total, double target, double tol, double* q_sd) {
  int lo = t - window + 1;
  if (lo < 0) lo = 0;
  long long s = 0;
  for (int i = lo; i <= t; ++i) s += accept_count[i];
  const double rate = (double)s / ((double)(t - lo + 1) * (double)n_total);
  double q = *q_sd;
  if (rate < target - tol) q *= 0.9;
  else if (rate > target + tol) q *= 1.1;
  *q_sd = q;
}

// Projection of the covariates onto the row space of g_net's last layer (see
// bgm_b200.cu): t = (v - b) U  (n x H) and r0 = |v - b|^2 - |t|^2.  One warp per row,
// U (p x HP) and b in shared memory.
__global__ void __launch_bounds__(256)
causal_project_kernel(const float* __restrict__ v, int ldv, int n, int p, int HP,
                      const float* __restrict__ Ub, float* __restrict__ t, int ldt,
                      float* __restrict__ r0, int stage) {
  extern __shared__ __align__(16) float psm[];
  // U [p][HP] | b [p]: staged in shared memory when it fits (stage != 0), else read through L1/L2
  // (v_dim >~ 890 at HP = 64 exceeds the 227 KB opt-in limit)
  const float* U = Ub;
  const float* b = Ub + p * HP;
  if (stage) {
    for (int i = threadIdx.x; i < p * HP + p; i += blockDim.x) psm[i] = Ub[i];
    __syncthreads();
    U = psm;
    b = psm + p * HP;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  for (int row = blockIdx.x * warps + warp; row < n; row += gridDim.x * warps) {
    const float* vr = v + (size_t)row * ldv;
    float t0 = 0.f, t1 = 0.f;
    double nv = 0.0;
    for (int k0 = 0; k0 < p; k0 += 32) {
      const int kk = k0 + lane;
      const float mine = kk < p ? vr[kk] - b[kk] : 0.f;
      nv += (double)mine * (double)mine;
      const int lim = min(32, p - k0);
      for (int j = 0; j < lim; ++j) {
        const float vk = __shfl_sync(0xffffffffu, mine, j);
        const float* u = U + (k0 + j) * HP;
        t0 = fmaf(vk, u[lane], t0);
        if (HP > 32) t1 = fmaf(vk, u[32 + lane], t1);
      }
    }
    double nt = (double)t0 * (double)t0 + (double)t1 * (double)t1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      nv += __shfl_xor_sync(0xffffffffu, nv, o);
      nt += __shfl_xor_sync(0xffffffffu, nt, o);
    }
    float* tr = t + (size_t)row * ldt;
    if (lane < HP) tr[lane] = t0;
    if (32 + lane < HP) tr[32 + lane] = t1;
    if (lane == 0) r0[row] = (float)fmax(nv - nt, 0.0);
  }
}

__global__ void mh_noise_kernel(uint64_t seed, int64_t row_offset, int n, int zd, int t_begin,
                                int t_end, float* z0, float* eps, double* u) {
  const int T = t_end - t_begin;
  const long long total = (long long)(T + 1) * n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ti = (int)(i / n);
    const int row = (int)(i - (long long)ti * n);
    const int64_t grow = row_offset + row;
    const bool init = ti == T;
    const uint32_t t = init ? T_INIT : (uint32_t)(t_begin + ti);
    float* dst = init ? (z0 ? z0 + (size_t)row * zd : nullptr)
                      : (eps ? eps + ((size_t)ti * n + row) * zd : nullptr);
    if (dst) {
      for (int g = 0; g * 4 < zd; ++g) {
        float e[4];
        normal4(seed, grow, t, NOISE_PROPOSAL, g, e);
        for (int q = 0; q < 4; ++q)
          if (g * 4 + q < zd) dst[g * 4 + q] = e[q];
      }
    }
    if (!init && u) u[(size_t)ti * n + row] = (double)uniform1(seed, grow, t, NOISE_ACCEPT);
  }
}

}  // namespace bgm
