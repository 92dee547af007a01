// BGM hot path: Hamiltonian Monte Carlo over the latent z of a mean/variance
// generator (BaseVariationalNet), with missing observations.
//
// Replaces bgm/base.py:665-705 (get_log_posterior), :709-830 (tfp_mcmc_sampler; the
// integrator / accept / step-size rule of TFP 0.18 restated in SURVEY A.5) and
// :511-525 (predict_on_posteriors) of the reference.
//
// Engine: a CTA is `ncons` warps of 32 observations each.  NO weight is resident: every
// weight tile of the per-leapfrog program (forward layers, fused mean/variance head
// tiles with their transposes, backward layers) is streamed through a 2-stage
// shared-memory ring with cp.async.bulk (TMA bulk copy) completing on mbarriers; every
// warp uses each staged tile for its own rows, so the L2->SM stream is shared by
// ncons*32 rows.  There is no producer warp: the LAST warp to finish a stage (an
// arrival counter in shared memory) re-arms it with the tile two ops ahead.  The
// gradient d logp / d z is hand-written: LeakyReLU sign bits of the forward pass
// stay in the registers of the thread that needs them in the backward pass.
#pragma once
#include "common.cuh"

namespace bgm {

constexpr int HMC_MAX_OPS = 176;
constexpr int HMC_MAXL = 6;                          // hidden layers of g_net
constexpr int HMC_MAXZ = 16;                         // z_dim
constexpr int HMC_STAGE_FLOATS = 64 * 64 + 64 + 64;  // tile + bias + prefetch pad
constexpr int HMC_STAGES = 2;
constexpr int HMC_BUF = 64 * TILE_ROWS;              // one activation-format buffer

enum : unsigned char { HK_FWD = 0, HK_HEADF = 1, HK_HEADB = 2, HK_BWD = 3, HK_DZ = 4 };

struct HmcOp {      // 16 bytes
  int g_off;        // float offset of the tile [kp][NT] | bias[NT] in the global image
  int bytes;        // bytes the producer copies (multiple of 16)
  short kp;         // reduction length, multiple of 4
  unsigned char kind;
  unsigned char layer;  // HK_FWD / HK_BWD: hidden layer whose sign bits are written / applied
  short c0;         // HK_HEAD*: first data column; HK_DZ: first z index
  short flags;      // bit0: first head tile, bit1: last head tile
};

struct HmcProgram {
  int n_ops;
  int zd, kin, x_dim, nh;
  float bn_mean[HMC_MAXZ], bn_inv[HMC_MAXZ], bn_beta[HMC_MAXZ];  // BN(z) = (z-mean)*inv+beta
  HmcOp ops[HMC_MAX_OPS];
};

enum { HMC_RUN = 0, HMC_EVAL = 1, HMC_PREDICT = 2 };

struct HmcDev {
  bgm_hmc_args a;
  int mode;
  int ncons;
  // HMC_PREDICT: rows are (sample, row) pairs of z_samples (n_keep*n, zd)
  const float* z_in;      // EVAL: (n,zd) ; PREDICT: (n_rows, zd)
  float* out_grad;        // EVAL: (n,zd)
  float* out_x;           // PREDICT: (n_rows, x_dim) draws
  const float* noise_x;   // PREDICT: optional injected N(0,1) (n_rows, x_dim)
  float* out_var;         // PREDICT: if non-NULL, out_x receives mu and out_var sigma^2 (no draw)
  int n_per_sample;       // PREDICT: n (rows per sample), for the Philox key
  int sample0;            // PREDICT: index of the first sample in this call
};

// 1 / x for x in a safe range (no zero / denormal / inf): MUFU.RCP + one Newton step, within 1 ulp.  A plain division
// compiles to a range test with a slow-path call per use.
__device__ __forceinline__ float rcp_newton(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return fmaf(r, fmaf(-x, r, 1.f), r);
}

// ------------------------------------------------------------- mbarriers -----
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Stores an 8x8 register block in activation format (act[col][row], swizzled).
__device__ __forceinline__ void store_block(float* buf, const float (&v)[RPT][8], int rg, int cg) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = ColMap<8>::col(cg, j);
    const int chunk = (rg ^ act_swz(c)) << 2;
    *reinterpret_cast<float4*>(buf + c * TILE_ROWS + chunk) = make_float4(v[0][j], v[1][j], v[2][j], v[3][j]);
    *reinterpret_cast<float4*>(buf + c * TILE_ROWS + (chunk ^ 16)) = make_float4(v[4][j], v[5][j], v[6][j], v[7][j]);
  }
}

struct HmcWarp {
  float *zin, *act, *gbuf, *zs, *ps, *gs;
};

// Ring position: identical in every warp of the CTA by construction (all warps run the
// same op sequence).
struct RingPos {
  uint32_t it;        // ops consumed so far by this warp
  uint32_t total;     // ops this CTA will consume in the whole launch
  __device__ __forceinline__ int stage() const { return it & (HMC_STAGES - 1); }
  __device__ __forceinline__ uint32_t parity() const { return (it / HMC_STAGES) & 1; }
};
struct Ring {
  float* buf;
  uint64_t* full;
  int* done;          // per stage: warps that finished the tile currently staged
  const float* image;
  int ncons;
};
__device__ __forceinline__ void ring_issue(const Ring& R, const HmcOp& op, int st) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  mbar_expect_tx(R.full + st, (uint32_t)op.bytes);
  bulk_g2s(R.buf + st * HMC_STAGE_FLOATS, R.image + op.g_off, (uint32_t)op.bytes, R.full + st);
}
// Called by every warp after its last read of stage `st`; the last arriver refills it.
__device__ __forceinline__ void ring_release(const HmcProgram& P, const Ring& R, const RingPos& rp, int o,
                                             int st, int lane) {
  __syncwarp();
  if (lane == 0) {
    __threadfence_block();
    if (atomicAdd(R.done + st, 1) == R.ncons - 1) {
      atomicExch(R.done + st, 0);
      if (rp.it + HMC_STAGES < rp.total) {
        int o2 = o + HMC_STAGES;
        while (o2 >= P.n_ops) o2 -= P.n_ops;
        ring_issue(R, P.ops[o2], st);
      }
    }
  }
}

// One pass over the tile program for the 32 rows whose BN(z) sits in W.zin.
//   want_lp : also accumulate the likelihood terms (needed for log p only)
//   predict : forward only; heads produce draws x = mu + sqrt(s2) * N(0,1)
// Returns the likelihood loss of this lane's row (0 if !want_lp); the gradient of the
// LOSS w.r.t. BN(z) is left in W.gbuf rows [0, zd) (activation format, unswizzled
// scratch: gbuf[d*32 + row]).
__device__ __forceinline__ float hmc_program(const HmcProgram& P, const HmcDev& D, const HmcWarp& W,
                                             const Ring& R, RingPos& rp, int row0, int n_rows, int lane,
                                             bool want_lp) {
  const int rg = lane >> 3, cg = lane & 7;
  const bool predict = D.mode == HMC_PREDICT;
  unsigned long long sg[HMC_MAXL];
#pragma unroll
  for (int l = 0; l < HMC_MAXL; ++l) sg[l] = 0ull;
  float loss8[RPT];
#pragma unroll
  for (int i = 0; i < RPT; ++i) loss8[i] = 0.f;
  float gh[RPT][8];
#pragma unroll 1
  for (int o = 0; o < P.n_ops; ++o) {
    const HmcOp op = P.ops[o];
    const int st = rp.stage();
    mbar_wait(R.full + st, rp.parity());
    const float* w = R.buf + st * HMC_STAGE_FLOATS;
    if (op.kind == HK_DZ) {
      float acc[RPT][1];
#pragma unroll
      for (int i = 0; i < RPT; ++i) acc[i][0] = 0.f;
      tile_mac<1>(W.act, w, op.kp, rg, cg, acc);
      __syncwarp();
#pragma unroll
      for (int i = 0; i < RPT; ++i) W.gbuf[(op.c0 + cg) * TILE_ROWS + row_of(rg, i)] = acc[i][0];
    } else if (op.kind == HK_HEADB) {
      if (op.flags & 1) {
#pragma unroll
        for (int i = 0; i < RPT; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) gh[i][j] = 0.f;
      }
      tile_mac<8>(W.gbuf, w, op.kp, rg, cg, gh);
      __syncwarp();
      if (op.flags & 2) {  // d loss / d h_last complete: through the last LeakyReLU, to act
        float v[RPT][8];
        unsigned long long bits = 0ull;
#pragma unroll
        for (int l = 0; l < HMC_MAXL; ++l)
          if (l == op.layer) bits = sg[l];
#pragma unroll
        for (int i = 0; i < RPT; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) v[i][j] = ((bits >> (i * 8 + j)) & 1ull) ? gh[i][j] : 0.2f * gh[i][j];
        store_block(W.act, v, rg, cg);
      }
    } else {
      const float* bias = w + op.kp * 64;
      float acc[RPT][8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float bj = bias[ColMap<8>::col(cg, j)];
#pragma unroll
        for (int i = 0; i < RPT; ++i) acc[i][j] = bj;
      }
      const float* in = (op.kind == HK_FWD && op.layer == 0) ? W.zin : W.act;
      tile_mac<8>(in, w, op.kp, rg, cg, acc);
      __syncwarp();
      if (op.kind == HK_FWD) {
        unsigned long long bits = 0ull;
#pragma unroll
        for (int i = 0; i < RPT; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (acc[i][j] > 0.f) bits |= 1ull << (i * 8 + j);
            else acc[i][j] *= 0.2f;
          }
#pragma unroll
        for (int l = 0; l < HMC_MAXL; ++l)
          if (l == op.layer) sg[l] = bits;
        store_block(W.act, acc, rg, cg);
      } else if (op.kind == HK_BWD) {  // gradient w.r.t. the output of hidden layer op.layer
        unsigned long long bits = 0ull;
#pragma unroll
        for (int l = 0; l < HMC_MAXL; ++l)
          if (l == op.layer) bits = sg[l];
#pragma unroll
        for (int i = 0; i < RPT; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (!((bits >> (i * 8 + j)) & 1ull)) acc[i][j] *= 0.2f;
        store_block(W.act, acc, rg, cg);
      } else {  // HK_HEADF: acc[i][0..3] = mu, acc[i][4..7] = raw variance head of 4 data columns
        const int c = op.c0 + cg * 4;
        const bool cols_in = c < D.a.ldx;
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          int r = row0 + row_of(rg, i);
          const bool rvalid = r < n_rows;
          r = rvalid ? r : n_rows - 1;
          if (predict) {
            // bgm/base.py:517-521: x = mu + sqrt(softplus(raw)+1e-6) * N(0,1)
            float e4[4];
            if (D.out_var) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                e4[q] = 0.f;
                if (rvalid && c + q < P.x_dim)
                  D.out_var[(size_t)r * P.x_dim + c + q] = softplus_f(acc[i][4 + q]) + 1e-6f;
              }
            } else if (D.noise_x) {
#pragma unroll
              for (int q = 0; q < 4; ++q) e4[q] = (c + q < P.x_dim) ? D.noise_x[(size_t)r * P.x_dim + c + q] : 0.f;
            } else {
              const int s = D.sample0 + r / D.n_per_sample;
              const int64_t grow = D.a.row_offset + (r - (r / D.n_per_sample) * D.n_per_sample);
              normal4(D.a.seed, grow, (uint32_t)s, NOISE_PREDICT, (uint32_t)(c >> 2), e4);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float s2 = softplus_f(acc[i][4 + q]) + 1e-6f;
              const float xv = fmaf(e4[q], sqrtf(s2), acc[i][q]);
              if (rvalid && c + q < P.x_dim) D.out_x[(size_t)r * P.x_dim + c + q] = xv;
            }
          } else {
            float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (cols_in) xv = __ldg(reinterpret_cast<const float4*>(D.a.x_dev + (size_t)r * D.a.ldx + c));
            const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const bool obs = cols_in && (c + q < P.x_dim) && (xs[q] == xs[q]);   // NaN = missing
              const float raw = acc[i][4 + q];
              const float e = expf(-fabsf(raw));
              const float s2 = (fmaxf(raw, 0.f) + log1pf(e)) + 1e-6f;                // :  softplus + eps
              const float inv = rcp_newton(s2);
              const float d = obs ? xs[q] - acc[i][q] : 0.f;
              const float r1 = rcp_newton(1.f + e);
              const float sig = raw >= 0.f ? r1 : e * r1;
              if (want_lp && obs) loss8[i] += (d * d) * (0.5f * inv) + 0.5f * logf(s2);   // bgm/base.py:683-684
              acc[i][q] = obs ? -d * inv : 0.f;                                           // d loss / d mu
              acc[i][4 + q] = obs ? (0.5f * inv - (0.5f * d * d) * (inv * inv)) * sig : 0.f;  // d loss / d raw
            }
          }
        }
        if (!predict) store_block(W.gbuf, acc, rg, cg);
      }
    }
    ring_release(P, R, rp, o, st, lane);
    ++rp.it;
  }
  float loss = 0.f;
  if (want_lp) {
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      float s = loss8[i];
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      if (cg == 0) W.gbuf[HMC_MAXZ * TILE_ROWS + row_of(rg, i)] = s;
    }
    __syncwarp();
    loss = W.gbuf[HMC_MAXZ * TILE_ROWS + lane];
  }
  __syncwarp();
  return loss;
}

// lane = row helpers on the per-warp state arrays ([d][32])
__device__ __forceinline__ void write_zin(const HmcProgram& P, const HmcWarp& W, int lane) {
  for (int d = 0; d < P.zd; ++d)
    W.zin[act_idx(d, lane)] = (W.zs[d * TILE_ROWS + lane] - P.bn_mean[d]) * P.bn_inv[d] + P.bn_beta[d];
  __syncwarp();
}
// log posterior and its gradient from the program's outputs (bgm/base.py:702-704)
__device__ __forceinline__ float finish_grad(const HmcProgram& P, const HmcWarp& W, int lane, float loss) {
  float prior = 0.f;
  for (int d = 0; d < P.zd; ++d) {
    const float z = W.zs[d * TILE_ROWS + lane];
    prior = fmaf(z, z, prior);
    W.gs[d * TILE_ROWS + lane] = -(W.gbuf[d * TILE_ROWS + lane] * P.bn_inv[d] + z);
  }
  __syncwarp();
  return -(0.5f * prior + loss);
}

__global__ void __launch_bounds__(256, 1)
hmc_kernel(const __grid_constant__ HmcProgram P, const float* __restrict__ image,
           const __grid_constant__ HmcDev D) {
  extern __shared__ __align__(128) float smem[];
  __shared__ uint64_t full[HMC_STAGES];
  __shared__ int done[HMC_STAGES];
  const bgm_hmc_args& A = D.a;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ncons = D.ncons;

  const int n_rows = A.n;
  const int ntiles = (n_rows + TILE_ROWS - 1) / TILE_ROWS;
  const int nblocks = (ntiles + ncons - 1) / ncons;
  const int L = A.num_leapfrog;
  const bool need_init = D.mode != HMC_RUN || A.init_mode != 0;
  const int steps = D.mode == HMC_RUN ? A.t_end - A.t_begin : 0;
  const int runs_per_block = (need_init ? 1 : 0) + steps * L;
  int my_blocks = 0;
  for (int b = blockIdx.x; b < nblocks; b += gridDim.x) ++my_blocks;
  Ring R;
  R.buf = smem;
  R.full = full;
  R.done = done;
  R.image = image;
  R.ncons = ncons;
  RingPos rp;
  rp.it = 0;
  rp.total = (uint32_t)my_blocks * (uint32_t)runs_per_block * (uint32_t)P.n_ops;
  if (threadIdx.x == 0) {
    for (int s = 0; s < HMC_STAGES; ++s) {
      mbar_init(full + s, 1);
      done[s] = 0;
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int s = 0; s < HMC_STAGES; ++s)
      if ((uint32_t)s < rp.total) ring_issue(R, P.ops[s % P.n_ops], s);
  }
  __syncthreads();
  // ------------------------------------------------------------------- consumers
  const int zd = P.zd;
  const int per_warp = P.kin * TILE_ROWS + 2 * HMC_BUF + 3 * zd * TILE_ROWS;
  HmcWarp W;
  W.zin = smem + HMC_STAGES * HMC_STAGE_FLOATS + warp * per_warp;
  W.act = W.zin + P.kin * TILE_ROWS;
  W.gbuf = W.act + HMC_BUF;
  W.zs = W.gbuf + HMC_BUF;
  W.ps = W.zs + zd * TILE_ROWS;
  W.gs = W.ps + zd * TILE_ROWS;
  for (int k = zd; k < P.kin; ++k) W.zin[act_idx(k, lane)] = 0.f;
  const float eps = (D.mode == HMC_RUN && A.step_dev) ? *A.step_dev : 0.f;

  for (int b = blockIdx.x; b < nblocks; b += gridDim.x) {
    const int row0 = (b * ncons + warp) * TILE_ROWS;
    const int row = row0 + lane;
    const bool valid = row < n_rows;
    const int lrow = valid ? row : n_rows - 1;
    const int64_t grow = A.row_offset + lrow;

    if (D.mode != HMC_RUN) {  // EVAL: log p and gradient at z_in ; PREDICT: draws at z_in
      for (int d = 0; d < zd; ++d) W.zs[d * TILE_ROWS + lane] = D.z_in[(size_t)lrow * zd + d];
      write_zin(P, W, lane);
      const float loss = hmc_program(P, D, W, R, rp, row0, n_rows, lane, D.mode == HMC_EVAL);
      if (D.mode == HMC_EVAL) {
        const float lp = finish_grad(P, W, lane, loss);
        if (valid) {
          A.lp_state_dev[row] = lp;
          if (D.out_grad)
            for (int d = 0; d < zd; ++d) D.out_grad[(size_t)row * zd + d] = W.gs[d * TILE_ROWS + lane];
        }
      }
      __syncwarp();
      continue;
    }

    // ---- chain state (bgm/base.py:778) ----
    if (A.init_mode == 2) {
      for (int g = 0; g * 4 < zd; ++g) {
        float e[4];
        normal4(A.seed, grow, T_INIT, NOISE_MOMENTUM, g, e);
        for (int q = 0; q < 4; ++q)
          if (g * 4 + q < zd) W.zs[(g * 4 + q) * TILE_ROWS + lane] = e[q];
      }
    } else {
      for (int d = 0; d < zd; ++d) W.zs[d * TILE_ROWS + lane] = A.z_state_dev[(size_t)lrow * zd + d];
    }
    float lp_cur;
    if (need_init) {
      write_zin(P, W, lane);
      const float loss = hmc_program(P, D, W, R, rp, row0, n_rows, lane, true);
      lp_cur = finish_grad(P, W, lane, loss);
      if (valid) {
        for (int d = 0; d < zd; ++d) {
          A.z_state_dev[(size_t)row * zd + d] = W.zs[d * TILE_ROWS + lane];
          A.g_state_dev[(size_t)row * zd + d] = W.gs[d * TILE_ROWS + lane];
        }
      }
    } else {
      lp_cur = A.lp_state_dev[lrow];
      for (int d = 0; d < zd; ++d) W.gs[d * TILE_ROWS + lane] = A.g_state_dev[(size_t)lrow * zd + d];
    }

    // ---- HMC steps (TFP 0.18 HamiltonianMonteCarlo, unit mass) ----
#pragma unroll 1
    for (int t = A.t_begin; t < A.t_end; ++t) {
      float ke0 = 0.f;
      for (int g = 0; g * 4 < zd; ++g) {
        float e[4];
        if (A.mom_dev) {
          for (int q = 0; q < 4; ++q)
            e[q] = (g * 4 + q < zd) ? A.mom_dev[((size_t)t * A.n + lrow) * zd + g * 4 + q] : 0.f;
        } else {
          normal4(A.seed, grow, (uint32_t)t, NOISE_MOMENTUM, g, e);
        }
        for (int q = 0; q < 4; ++q) {
          const int d = g * 4 + q;
          if (d < zd) {
            ke0 = fmaf(e[q], e[q], ke0);
            W.ps[d * TILE_ROWS + lane] = e[q] + (0.5f * eps) * W.gs[d * TILE_ROWS + lane];
          }
        }
      }
      ke0 *= 0.5f;
      float lp_new = lp_cur;
#pragma unroll 1
      for (int l = 0; l < L; ++l) {
        for (int d = 0; d < zd; ++d) W.zs[d * TILE_ROWS + lane] += eps * W.ps[d * TILE_ROWS + lane];
        write_zin(P, W, lane);
        const bool last = l == L - 1;
        const float loss = hmc_program(P, D, W, R, rp, row0, n_rows, lane, last);
        const float lp = finish_grad(P, W, lane, loss);
        if (last) lp_new = lp;
        for (int d = 0; d < zd; ++d) W.ps[d * TILE_ROWS + lane] += eps * W.gs[d * TILE_ROWS + lane];
      }
      float ke1 = 0.f;
      for (int d = 0; d < zd; ++d) {
        const float p = W.ps[d * TILE_ROWS + lane] - (0.5f * eps) * W.gs[d * TILE_ROWS + lane];
        ke1 = fmaf(p, p, ke1);
      }
      ke1 *= 0.5f;
      float log_accept = (lp_new - lp_cur) + (ke0 - ke1);
      if (!(fabsf(log_accept) <= 3.0e38f)) log_accept = -INFINITY;   // NaN / inf energy: reject
      float logu;
      if (A.logu_dev) logu = A.logu_dev[(size_t)t * A.n + lrow];
      else logu = logf(u01_open1(noise_block(A.seed, grow, (uint32_t)t, NOISE_ACCEPT, 0).x));
      const bool acc = logu < log_accept;
      if (acc) {
        lp_cur = lp_new;
        if (valid)
          for (int d = 0; d < zd; ++d) {
            A.z_state_dev[(size_t)row * zd + d] = W.zs[d * TILE_ROWS + lane];
            A.g_state_dev[(size_t)row * zd + d] = W.gs[d * TILE_ROWS + lane];
          }
      } else {
        for (int d = 0; d < zd; ++d) {
          W.zs[d * TILE_ROWS + lane] = A.z_state_dev[(size_t)lrow * zd + d];
          W.gs[d * TILE_ROWS + lane] = A.g_state_dev[(size_t)lrow * zd + d];
        }
      }
      if (A.accept_mask_dev && valid) A.accept_mask_dev[(size_t)t * A.n + row] = acc ? 1 : 0;
      if (A.log_accept_dev && valid) A.log_accept_dev[(size_t)t * A.n + row] = log_accept;
      if (A.accept_count_dev) {
        const unsigned bal = __ballot_sync(0xffffffffu, acc && valid);
        if (lane == 0 && bal) atomicAdd(A.accept_count_dev + t, __popc(bal));
      }
      if (A.accept_stat_dev) {   // SimpleStepSizeAdaptation: mean over ALL chains of exp(min(log_accept, 0))
        float a = valid ? expf(fminf(log_accept, 0.f)) : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) atomicAdd(A.accept_stat_dev + t, (double)a);
      }
      if (t >= A.burn_in && A.out_samples_dev && valid) {
        float* dst = A.out_samples_dev + ((size_t)(t - A.burn_in) * A.n + row) * zd;
        for (int d = 0; d < zd; ++d) dst[d] = W.zs[d * TILE_ROWS + lane];
      }
      __syncwarp();
    }
    if (valid) A.lp_state_dev[row] = lp_cur;
    __syncwarp();
  }
}

// 1-thread kernel: TFP SimpleStepSizeAdaptation (adaptation_rate 0.01, float32 step):
// step *= 1.01 if mean_chains exp(min(log_accept,0)) > target else step /= 1.01.
__global__ void hmc_adapt_kernel(const double* __restrict__ stat, int t, long long n_total, float target,
                                 float rate, float* step) {
  const double mean = stat[t] / (double)n_total;
  const float s = *step;
  *step = (mean > (double)target) ? s * (1.f + rate) : s / (1.f + rate);
}

__global__ void hmc_noise_kernel(uint64_t seed, int64_t row_offset, int n, int zd, int t_begin, int t_end,
                                 float* z0, float* mom, float* logu) {
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
                      : (mom ? mom + ((size_t)ti * n + row) * zd : nullptr);
    if (dst) {
      for (int g = 0; g * 4 < zd; ++g) {
        float e[4];
        normal4(seed, grow, t, NOISE_MOMENTUM, g, e);
        for (int q = 0; q < 4; ++q)
          if (g * 4 + q < zd) dst[g * 4 + q] = e[q];
      }
    }
    if (!init && logu) logu[(size_t)ti * n + row] = logf(u01_open1(noise_block(seed, grow, t, NOISE_ACCEPT, 0).x));
  }
}

}  // namespace bgm
