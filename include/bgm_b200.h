/*
 * bgm_b200.h -- C ABI of the B200-native bayesgm hot path (libbgm_b200.so).
 *
 * The reference (liuq-lab/bayesgm @ 85350a1) has no FFI / plugin layer: its seam is
 * the Python method surface of `CausalBGM` / `BGM`.  Each
 * entry point below replaces the device-side work of one of those methods; the
 * Python host (bayesgm_b200/causalbgm.py, bgm.py) keeps the reference's method
 * names and kwargs and binds these symbols with ctypes.
 *
 * Conventions: plain pointers and sizes, no torch types.  Pointers named *_dev are
 * caller-owned DEVICE memory on the current CUDA device; `stream` is a cudaStream_t
 * passed as void* (NULL = default stream).  Every function returns 0 on success or
 * a negative bgm_status; bgm_last_error() gives the message (thread-local).  No
 * hidden device allocation happens outside the *_create functions.  All arithmetic
 * is IEEE fp32 (no fast-math, no plain-TF32 rounding): on the CUDA cores, or -- the
 * sampler's tensor engine, see bgm_causal_set_sampler -- as error-compensated 3xTF32
 * products with fp32 accumulation on the tensor cores.  Row-major layouts.
 */
#ifndef BGM_B200_H
#define BGM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  BGM_OK = 0,
  BGM_ERR_ARG = -1,      /* bad argument (message says which) */
  BGM_ERR_UNSUPPORTED = -2, /* layer width / dimension outside what the kernels cover */
  BGM_ERR_CUDA = -3,     /* CUDA runtime error (message carries cudaGetErrorString) */
  BGM_ERR_NOMEM = -4     /* packed model does not fit in shared memory */
} bgm_status;

const char* bgm_last_error(void);
int bgm_version(void);
/* SM count, opt-in shared memory per block, SM clock (kHz) of the current device. */
int bgm_device_info(int* sm_count, int* smem_optin_bytes, int* clock_khz);

/* ---------------------------------------------------------------- networks -- */
/* A Dense stack as the reference builds it (networks/base.py:17-26): n_layers
 * layers, dims[n_layers+1] = [in, units..., out]; `params` (HOST memory) is the
 * concatenation, layer by layer, of the Keras arrays kernel[in][out] (row-major)
 * then bias[out].  LeakyReLU(0.2) after every layer but the last (:38-51). */
typedef struct {
  int n_layers;
  const int* dims;
  const float* params;
} bgm_net_desc;

/* ------------------------------------------------------- CausalBGM sampler -- */
typedef struct bgm_causal bgm_causal; /* opaque: g/f/h nets packed for the kernels */

/* Replaces the construction-time state get_log_posterior reads
 * (causalbgm/base.py:74-81, :765-798).  z_dims = [z0,z1,z2,z3]; sigma_* < 0 means
 * "learned head" (softplus(last output)+1e-6), >= 0 is the fixed `sigma_*` of the
 * params dict (variance = sigma^2, :781-798).  g: sum(z_dims) -> v_dim+1,
 * f: z0+z1+1 -> 2, h: z0+z2 -> 2.  Hidden widths up to 64 are supported. */
int bgm_causal_create(bgm_causal** out, const int z_dims[4], int v_dim, int binary_treatment,
                      float sigma_v, float sigma_x, float sigma_y,
                      const bgm_net_desc* g_net, const bgm_net_desc* f_net,
                      const bgm_net_desc* h_net);
void bgm_causal_destroy(bgm_causal* m);
/* Packed-model facts: shared-memory bytes per CTA, warps per CTA, tile ops per
 * log-posterior evaluation, algorithmic FMAs per row per evaluation (the reference's
 * formula, unpadded), FMAs the kernel issues per row per evaluation, and proj_dim:
 * 0, or the width H of the projected covariates (see bgm_causal_project). */
int bgm_causal_info(const bgm_causal* m, int* smem_bytes, int* warps_per_cta, int* n_ops,
                    long long* macs_per_row, long long* issued_macs_per_row, int* proj_dim);

/* Sampler engines.  Two execution plans of the SAME algorithm (same arguments, same
 * Philox streams) serve bgm_causal_logpost / bgm_causal_mh:
 *   1  SIMT   : every layer on the fp32 FMA pipe (causal.cuh), any net shape;
 *   2  tensor : the 64x64 layers of g_net on the 5th-gen tensor cores (tcgen05, accumulators
 *               and activations in TMEM) as error-compensated 3xTF32 products (fp32-level
 *               error, no plain-TF32 rounding anywhere), everything else on the FMA pipe
 *               (causal_tc.cuh).  Needs g hidden layers of 64 units (>= 2 of them), v_dim > 72
 *               (projected likelihood) and f / h units [64, 32, 8].
 * kind 0 = auto (tensor when available).  bgm_causal_sampler_info reports the engine the
 * next launch will use, whether the tensor engine exists for this model, its shared-memory
 * bytes and the multiply-adds it issues per row per evaluation (tensor + FMA pipe). */
int bgm_causal_set_sampler(bgm_causal* m, int kind);
int bgm_causal_sampler_info(const bgm_causal* m, int* active_kind, int* tensor_available,
                            int* tensor_smem_bytes, long long* tensor_issued_macs_per_row);
/* Name of the __global__ function the next bgm_causal_mh / bgm_causal_logpost launch runs (as it
 * appears in an ncu launch list), e.g. "causal_mh_tc16_kernel<8, true>" (second argument: first layers of f / h on
 * the tensor cores as well, z_dim <= 6). */
int bgm_causal_kernel_name(const bgm_causal* m, char* buf, int len);

/* Covariate projection.  With mu_v = h M + b (M: H x v_dim, the last layer of g_net,
 * H <= 64 < v_dim) and M^T = U R (thin QR, computed at create time in float64),
 *   sum_j (v_j - mu_v,j)^2 = | (v-b) U - h R^T |^2 + ( |v-b|^2 - |(v-b) U|^2 ),
 * an exact identity: the likelihood term of causalbgm/base.py:800 only sees h through
 * the H-dimensional row space of M.  The data-dependent parts t = (v-b) U (n,H) and
 * r0 = |v-b|^2 - |t|^2 (n) do not depend on z, so they are computed ONCE per data set
 * by this call and every log-posterior evaluation then costs H*H instead of H*v_dim
 * FMAs for that layer.  Models with proj_dim == 0 (v_dim <= H + 8) do not use it.
 * vproj_dev: (n, ldvproj), ldvproj >= proj_dim, ldvproj % 4 == 0; r0_dev: (n). */
int bgm_causal_project(const bgm_causal* m, const float* v_dev, int ldv, int n, float* vproj_dev,
                       int ldvproj, float* r0_dev, void* stream);

/* CausalBGM.get_log_posterior (causalbgm/base.py:765-817).
 * x_dev,y_dev: (n) ; v_dev: (n, ldv) with ldv >= v_dim, ldv % 4 == 0, 16-byte
 * aligned base; z_dev: (n, sum z_dims); out_logp_dev: (n).  Models with proj_dim > 0
 * read vproj_dev / r0_dev (from bgm_causal_project) instead of v_dev, which may then
 * be NULL.  sched_dev: see bgm_mh_args. */
int bgm_causal_logpost(const bgm_causal* m, const float* x_dev, const float* y_dev,
                       const float* v_dev, int ldv, const float* vproj_dev, int ldvproj,
                       const float* r0_dev, const float* z_dev, int n, float* out_logp_dev,
                       int* sched_dev, void* stream);
/* The same with the conditional prior rows of bgm_mh_args.prior_dev
 * (IdentifiableCausalBGM.get_log_posterior, causalbgm/identifiable.py:505-556). */
int bgm_causal_logpost_cond(const bgm_causal* m, const float* x_dev, const float* y_dev,
                            const float* v_dev, int ldv, const float* vproj_dev, int ldvproj,
                            const float* r0_dev, const float* z_dev, int n, const float* prior_dev,
                            int ldprior, float* out_logp_dev, int* sched_dev, void* stream);

/* CausalBGM.metropolis_hastings_sampler (causalbgm/base.py:820-904): iterations
 * [t_begin, t_end) of n independent random-walk MH chains in ONE persistent launch.
 * The chain state lives in z_state_dev / lp_state_dev between launches so that the
 * adaptive-q_sd path (:880-892) can run as a sequence of launches with
 * bgm_mh_adapt_qsd in between, all on `stream`, without host synchronisation. */
typedef struct {
  const float* x_dev;       /* (n)    treatment                                   */
  const float* y_dev;       /* (n)    outcome                                     */
  const float* v_dev;       /* (n,ldv) covariates (unused if the model projects)  */
  int ldv;
  int n;
  const float* vproj_dev;   /* (n,ldvproj) projected covariates, bgm_causal_project */
  const float* r0_dev;      /* (n)                                                */
  int ldvproj;
  int* sched_dev;           /* 4*(ceil(n/32)+1) bytes of scratch for the in-kernel
                               work scheduler; zeroed by the call                 */
  float* z_state_dev;       /* (n,zd) current state: in (init_mode 0/1) and out   */
  float* lp_state_dev;      /* (n)    cached log-posterior of the current state   */
  int init_mode;            /* 0: continue (lp_state valid); 1: z_state given,
                               evaluate its log-posterior first; 2: draw
                               z0 ~ N(0,1) from the Philox stream (:842), then 1  */
  int t_begin, t_end;       /* iteration range of this launch                     */
  int burn_in;              /* samples of iterations t >= burn_in are kept (:895) */
  const double* q_sd_dev;   /* (1) proposal sd (float64 like the reference's Python
                               float), read once at launch                        */
  /* noise: injected (both non-NULL; test / exact-parity mode) or Philox4x32-10  */
  const float* eps_dev;     /* (T,n,zd) UNIT normals; proposal step = float32(q_sd*eps)
                               computed in float64 like normal(0,q_sd).astype(f32)  */
  const double* u_dev;      /* (T,n)  uniforms, compared as float64 like :870     */
  uint64_t seed;            /* Philox key                                         */
  int64_t row_offset;       /* global index of row 0 (multi-GPU shards)           */
  /* outputs, each may be NULL */
  float* out_samples_dev;   /* (t_end_total-burn_in, n, zd) kept states           */
  int* accept_count_dev;    /* (T) accepted proposals per iteration (atomicAdd)   */
  uint8_t* accept_mask_dev; /* (T,n) per-row accept decisions (trace)             */
  float* lp_trace_dev;      /* (T,n) proposed log-posteriors (trace)              */
  /* conditional prior z | u ~ N(mu(u), sigma^2(u) I) of IdentifiableCausalBGM.get_log_posterior
   * (causalbgm/identifiable.py:540-548): row i holds mu_z[zd] then sigma^2; NULL = the N(0,I)
   * prior of causalbgm/base.py:812.  Runs on the SIMT engine.                                    */
  const float* prior_dev;   /* (n,ldprior), ldprior >= zd+1                        */
  int ldprior;
} bgm_mh_args;

/* Rows for bgm_mh_args.prior_dev: prior_net (n_segments -> zd+1, BaseFullyConnectedNet) evaluated on
 * the one-hot auxiliary variable u of every row (causalbgm/identifiable.py:540-543, :566-570):
 * prior_dev[r] = (mu_z[zd], softplus(last output)+1e-6) of segment seg_dev[r].  Synchronises the stream. */
int bgm_causal_prior_rows(const bgm_net_desc* prior_net, int n_segments, int zd, const int* seg_dev, int n,
                          float* prior_dev, int ldprior, void* stream);

int bgm_causal_mh(const bgm_causal* m, const bgm_mh_args* args, void* stream);

/* The windowed acceptance-rate rule of causalbgm/base.py:880-890, evaluated on the
 * device after iteration `t`: rate over iterations (t-window, t] (clipped at 0) of
 * accept_count / (len * n_total); q_sd *= 0.9 / 1.1 outside target +- tolerance. */
int bgm_mh_adapt_qsd(const int* accept_count_dev, int t, int window, long long n_total,
                     double target, double tolerance, double* q_sd_dev, void* stream);

/* Writes exactly the noise bgm_causal_mh would draw from Philox for rows
 * [0,n) + row_offset: z0 (n,zd), eps (t_end-t_begin,n,zd) unit normals, u (.,n)
 * as float64.  Used by the parity tests to replay Philox runs through the oracle. */
int bgm_mh_noise(uint64_t seed, int64_t row_offset, int n, int zd, int t_begin, int t_end,
                 float* z0_dev, float* eps_dev, double* u_dev, void* stream);

/* CausalBGM.infer_from_latent_posterior (causalbgm/base.py:671-763) on kept states.
 * z_samples_dev: (n_keep, n, zd).  Continuous: for each x_values[j] and sample s,
 * adrf_sum_dev[j*n_keep+s] += sum over rows of y(s,row,j) (float64 accumulators,
 * caller zeroes them and divides by n; shards add into the same layout).
 * Binary: ite_dev[s*n+row] = y(z,1) - y(z,0).  sample_y != 0 adds
 * sqrt(sigma_y^2)*N(0,1) from Philox (seed,row_offset); noise_dev, if non-NULL,
 * replaces those draws (continuous: (n_x,n_keep,n); binary: (2,n_keep,n)). */
int bgm_causal_effect(const bgm_causal* m, const float* z_samples_dev, int n_keep, int n,
                      const float* x_values_dev, int n_x, int sample_y, uint64_t seed,
                      int64_t row_offset, const float* noise_dev, double* adrf_sum_dev,
                      float* ite_dev, void* stream);

/* Memoised form of the same computation.  A rejected MH proposal repeats the chain state, so only a
 * fraction ~ acceptance rate of a row's kept states are distinct, and f_net is deterministic: (mu_y,
 * sigma_y head) are evaluated once per DISTINCT state and every (kept state, row, dose) then only
 * draws its noise and accumulates.  Identical results (same Philox draws, same f_net arithmetic).
 *   bgm_causal_effect_index   local_dev[s*n+row] = number of distinct states among kept states 0..s of the
 *                             row (it changes exactly where state s differs bitwise from state s-1);
 *                             rowtot_dev[row] = the row's total; rowend_dev = inclusive prefix sum of
 *                             rowtot_dev (rowend_dev[n-1] = number of distinct states); scratch_dev:
 *                             ceil(n/2048) ints.  n*n_keep < 2^31, zd <= 32.
 *   bgm_causal_effect_compact zlist_dev (n_distinct, zd): the distinct states, row by row (the states of
 *                             `row` occupy [rowend[row]-rowtot[row], rowend[row])).
 *   bgm_causal_effect_heads   heads_dev (n_states, n_x, 2) = (mu_y, raw sigma_y head) of f_net at every
 *                             dose (binary: n_x = 2, doses {1, 0}, x_values_dev NULL) for a list of states.
 *   bgm_causal_effect_combine the reductions of bgm_causal_effect from heads_dev, local_dev, rowend_dev. */
int bgm_causal_effect_index(const float* z_samples_dev, int n_keep, int n, int zd, int* local_dev, int* rowtot_dev,
                            int* rowend_dev, int* scratch_dev, void* stream);
int bgm_causal_effect_compact(const float* z_samples_dev, int n_keep, int n, int zd, const int* local_dev,
                              const int* rowend_dev, float* zlist_dev, void* stream);
int bgm_causal_effect_heads(const bgm_causal* m, const float* states_dev, int n_states, const float* x_values_dev,
                            int n_x, float* heads_dev, void* stream);
int bgm_causal_effect_combine(const bgm_causal* m, const float* heads_dev, const int* local_dev, const int* rowend_dev,
                              int n_keep, int n, int n_x, int sample_y, uint64_t seed, int64_t row_offset,
                              const float* noise_dev, double* adrf_sum_dev, float* ite_dev, void* stream);

/* ------------------------------------------- CausalBGM on Bayesian networks -- */
/* `use_bnn: True` (the default of every shipped CausalBGM config, causalbgm/base.py:64-72):
 * g / f / h are `BayesianFullyConnectedNet`s (networks/bnn.py:4-38): a BatchNormalization on
 * the input that uses the statistics of the CALL's batch (Keras hands the outer call's
 * `training=True` default down to the sub-layer; eps 1e-3, biased variance), then
 * tfp.layers.DenseFlipout layers with LeakyReLU(0.2) between them.  A DenseFlipout call draws
 * ONE kernel perturbation dW = sigma * N(0,1), sigma = finfo(float32).eps + softplus(rho), shared
 * by the rows of the batch, and per-row Rademacher vectors: y = x loc + ((x o s_in) dW) o s_out + b.
 * Descriptor arrays are HOST memory: bn = gamma[in] | beta[in]; params = per layer
 * loc[in][out], rho[in][out] (`kernel_posterior_untransformed_scale`), bias[out].
 * Widths: layer inputs <= 64 (so hidden widths <= 64), sum(z_dims) <= 32. */
typedef struct {
  int n_layers;
  const int* dims;
  const float* bn;
  const float* params;
} bgm_bnn_net_desc;

typedef struct bgm_bnn bgm_bnn; /* opaque: loc / sigma / bias images of g, f, h */

int bgm_bnn_create(bgm_bnn** out, const int z_dims[4], int v_dim, int binary_treatment, float sigma_v,
                   float sigma_x, float sigma_y, const bgm_bnn_net_desc* g_net,
                   const bgm_bnn_net_desc* f_net, const bgm_bnn_net_desc* h_net);
void bgm_bnn_destroy(bgm_bnn* m);
/* shared-memory bytes per CTA, rows per CTA, multiply-adds per row per log-posterior evaluation
 * (2 x the deterministic count: loc and perturbation products). */
int bgm_bnn_info(const bgm_bnn* m, int* smem_bytes, int* rows_per_cta, long long* macs_per_eval);
/* Execution plan of bgm_bnn_logpost / bgm_bnn_mh: 1 = one thread per row (csrc/bnn.cuh), 2 (default) = two threads
 * per row over a table-driven chunk program with double-buffered weight chunks (csrc/bnn2.cuh).  Same noise
 * streams, same per-accumulator arithmetic. */
int bgm_bnn_set_plan(bgm_bnn* m, int plan);
/* Number of doubles of scratch (per-CTA partial sums of the batch statistics) the calls below
 * need for n rows; -1 on a bad argument. */
long long bgm_bnn_scratch_doubles(const bgm_bnn* m, int n);

/* CausalBGM.get_log_posterior (causalbgm/base.py:765-817) with Bayesian nets: ONE call of each
 * net on the batch of n rows -- batch statistics over these n rows, Flipout noise from the Philox
 * streams (seed, slice, call) (csrc/bnn.cuh; restated in oracle/bnn.py).  v_dev: (n, ldv),
 * ldv % 4 == 0, 16-byte aligned; z_dev: (n, zd). */
int bgm_bnn_logpost(const bgm_bnn* m, const float* x_dev, const float* y_dev, const float* v_dev, int ldv,
                    const float* z_dev, int n, uint64_t seed, int slice, int64_t row_offset, uint32_t call,
                    double* scratch_dev, float* out_logp_dev, void* stream);

/* CausalBGM.metropolis_hastings_sampler (:820-904) with Bayesian nets: iterations
 * [t_begin, t_end) over the n rows of ONE `bs` slice (they share batch statistics): one launch per
 * iteration; proposal and current state are BOTH evaluated every iteration with fresh network
 * noise (call ids 2t, 2t+1) like :865-866.  Uses the fields of bgm_mh_args except vproj_dev,
 * r0_dev, sched_dev and lp_state_dev; init_mode 0/1: z_state_dev given, 2: z0 ~ N(0,1).
 * lp_trace_dev receives the proposals' log-posteriors, lp_cur_trace_dev (T,n; may be NULL) the
 * current states'.  `slice` keys the kernel-perturbation stream (slices are independent runs). */
int bgm_bnn_mh(const bgm_bnn* m, const bgm_mh_args* args, int slice, double* scratch_dev,
               float* lp_cur_trace_dev, void* stream);

/* CausalBGM.infer_from_latent_posterior (:671-763) with a Bayesian f_net: one f_net call per
 * (kept state s, dose j) on the n rows of state s (call id s*n_x + j; batch statistics of z0, z1
 * over those rows; the tiled dose column has batch variance 0 and normalises to beta exactly as in
 * the reference).  Arguments as bgm_causal_effect; stats_scratch_dev: n_keep * 2 * (z0+z1) floats. */
int bgm_bnn_effect(const bgm_bnn* m, const float* z_samples_dev, int n_keep, int n, const float* x_values_dev,
                   int n_x, int sample_y, uint64_t seed, int64_t row_offset, const float* noise_dev,
                   float* stats_scratch_dev, double* adrf_sum_dev, float* ite_dev, void* stream);

/* Test hook: the Flipout noise of (net 0 g / 1 f / 2 h, layer, call) exactly as the kernels draw
 * it: eps_dev (K, N) unit normals, sign_in_dev (rows, K), sign_out_dev (rows, N) as +-1 int8 for
 * global rows row_offset + [0, rows); any pointer may be NULL. */
int bgm_bnn_noise(const bgm_bnn* m, uint64_t seed, int slice, int net, int layer, uint32_t call,
                  int64_t row_offset, int rows, float* eps_dev, signed char* sign_in_dev,
                  signed char* sign_out_dev, void* stream);

/* --------------------------------------------------------------- BGM / HMC -- */
/* The generator of BGM, `BaseVariationalNet` (networks/base.py:53-117), in
 * inference mode as bgm/base.py:679 calls it: BatchNormalization on z with its
 * moving statistics, n_hidden Dense+LeakyReLU(0.2) layers, a mean head and a
 * softplus(+1e-6) variance head.  All arrays are HOST memory in Keras layout:
 *   bn            4*z_dim floats: gamma, beta, moving_mean, moving_variance (eps 1e-3)
 *   hidden_params per layer kernel[in][out] then bias[out], concatenated
 *   mean_params   kernel[units[-1]][x_dim] then bias[x_dim]; var_params likewise.
 * Supported: z_dim <= 16, hidden widths <= 64, n_hidden <= 6, any x_dim <= 2600. */
typedef struct {
  int z_dim, x_dim, n_hidden;
  const int* units;
  const float* bn;
  const float* hidden_params;
  const float* mean_params;
  const float* var_params;
} bgm_varnet_desc;

typedef struct bgm_hmc bgm_hmc; /* opaque: g_net tiles (forward + transposed) in device memory */

int bgm_hmc_create(bgm_hmc** out, const bgm_varnet_desc* g_net);
void bgm_hmc_destroy(bgm_hmc* m);
/* shared-memory bytes per CTA at 8 consumer warps, tile ops per gradient evaluation,
 * algorithmic and issued FMAs per row per gradient evaluation (forward + d/dz). */
int bgm_hmc_info(const bgm_hmc* m, int* smem_bytes, int* n_ops, long long* macs_per_grad,
                 long long* issued_macs_per_grad);

/* HMC engines, two execution plans of the SAME algorithm (same arguments, same Philox streams) behind
 * bgm_hmc_logpost_grad / bgm_hmc_run:
 *   1  SIMT   : streamed fp32 tiles on the FMA pipe (hmc.cuh), any hidden widths <= 64;
 *   2  tensor : every 64-wide product of a gradient evaluation on the 5th-gen tensor cores (tcgen05, operands
 *               and accumulators in TMEM) as error-compensated 3xTF32 (hmc_tc.cuh).  Needs >= 2 hidden layers,
 *               all 64 wide.
 * kind 0 = auto (tensor when available); bgm_hmc_predict / bgm_hmc_heads follow the same choice (forward-only pass). */
int bgm_hmc_set_engine(bgm_hmc* m, int kind);
int bgm_hmc_engine_info(const bgm_hmc* m, int* active_kind, int* tensor_available, int* tensor_smem_bytes,
                        long long* tensor_issued_macs_per_grad);

/* BGM.get_log_posterior (bgm/base.py:665-705) and its gradient w.r.t. z (what TFP's
 * HMC obtains by autodiff).  x_dev: (n, ldx), ldx % 4 == 0, 16-byte aligned; a NaN
 * entry is a MISSING observation (the input convention of BGM.predict, :527-545) and
 * contributes nothing -- the dense equivalent of the reference's ind_x1 / obs_mask
 * gather (:689-700).  z_dev: (n,z_dim); out_logp_dev: (n); out_grad_dev: (n,z_dim) or NULL. */
int bgm_hmc_logpost_grad(const bgm_hmc* m, const float* x_dev, int ldx, const float* z_dev, int n,
                         float* out_logp_dev, float* out_grad_dev, void* stream);

/* BGM.tfp_mcmc_sampler (bgm/base.py:709-830): HMC steps [t_begin, t_end) of n chains
 * that share ONE step size (read from step_dev at launch), num_leapfrog leapfrog steps
 * each, in one persistent launch.  While TFP's SimpleStepSizeAdaptation is active (the
 * first int(0.8*burn_in) steps, :805-809) the host launches one step at a time with
 * bgm_hmc_adapt in between on the same stream; afterwards a single launch runs the rest. */
typedef struct {
  const float* x_dev;        /* (n,ldx) data, NaN = missing                          */
  int ldx;
  int n;
  float* z_state_dev;        /* (n,z_dim) current state, in/out                      */
  float* g_state_dev;        /* (n,z_dim) gradient of log p at the current state     */
  float* lp_state_dev;       /* (n)       log p at the current state                 */
  int init_mode;             /* 0: continue; 1: z_state given, evaluate log p / grad
                                first; 2: z0 ~ N(0,1) from Philox (:778), then 1     */
  int t_begin, t_end;
  int burn_in;               /* states of steps t >= burn_in are kept                */
  int num_leapfrog;
  const float* step_dev;     /* (1) shared step size (float32 like TFP's state dtype)*/
  const float* mom_dev;      /* (T,n,z_dim) injected momenta N(0,1), or NULL: Philox */
  const float* logu_dev;     /* (T,n) injected log-uniforms, or NULL: Philox         */
  uint64_t seed;
  int64_t row_offset;
  float* out_samples_dev;    /* (n_mcmc,n,z_dim) or NULL                             */
  double* accept_stat_dev;   /* (T) += sum over rows of exp(min(log_accept,0)), or NULL */
  int* accept_count_dev;     /* (T) accepted chains per step, or NULL                */
  uint8_t* accept_mask_dev;  /* (T,n) trace, or NULL                                 */
  float* log_accept_dev;     /* (T,n) trace, or NULL                                 */
} bgm_hmc_args;

int bgm_hmc_run(const bgm_hmc* m, const bgm_hmc_args* args, void* stream);

/* TFP SimpleStepSizeAdaptation after step t (adaptation_rate `rate`, float32):
 * step *= (1+rate) if accept_stat[t]/n_total > target else step /= (1+rate).
 * Multi-GPU: all-reduce accept_stat_dev[t] over the shards before this call. */
int bgm_hmc_adapt(const double* accept_stat_dev, int t, long long n_total, float target, float rate,
                  float* step_dev, void* stream);

/* The Philox noise bgm_hmc_run draws: z0 (n,z_dim), momenta (t_end-t_begin,n,z_dim),
 * log-uniforms (.,n); any pointer may be NULL.  Test hook (oracle replay). */
int bgm_hmc_noise(uint64_t seed, int64_t row_offset, int n, int z_dim, int t_begin, int t_end,
                  float* z0_dev, float* mom_dev, float* logu_dev, void* stream);

/* BGM.predict_on_posteriors (bgm/base.py:511-525): x = mu(z) + sqrt(sigma^2(z)) * N(0,1)
 * for z_samples_dev (n_keep, n, z_dim) -> out_x_dev (n_keep, n, x_dim).  The N(0,1)
 * draws come from Philox keyed by (seed, row_offset+row, sample0+s, column) or from
 * noise_dev (n_keep, n, x_dim) if non-NULL. */
int bgm_hmc_predict(const bgm_hmc* m, const float* z_samples_dev, int n_keep, int n, int sample0,
                    uint64_t seed, int64_t row_offset, const float* noise_dev, float* out_x_dev,
                    void* stream);

/* Reduction of the posterior-predictive draws in BGM.predict (bgm/base.py:640-660): per column of
 * draws_dev (S, M) the mean over S and np.quantile at q_lo / q_hi (linear interpolation), in one
 * pass (thread = column keeps the few smallest / largest values in registers).  lo_dev / hi_dev may
 * both be NULL (mean only).  BGM_ERR_UNSUPPORTED when a quantile needs more than 16 order statistics
 * from one end of the sample (the caller sorts instead). */
int bgm_column_quantiles(const float* draws_dev, int S, long long M, double q_lo, double q_hi,
                         float* mean_dev, float* lo_dev, float* hi_dev, void* stream);

/* The generator's heads (bgm/base.py:483-509 `generate`, networks/base.py:98-111): mu(z) and
 * sigma^2(z) = softplus(raw) + 1e-6 for z_dev (n, z_dim) -> out_mu_dev, out_var_dev (n, x_dim). */
int bgm_hmc_heads(const bgm_hmc* m, const float* z_dev, int n, float* out_mu_dev, float* out_var_dev,
                  void* stream);

/* ------------------------------------------------------- EGM training steps -- */
/* The Discriminator of networks/base.py:338-385: n_hidden blocks of
 * Dense -> BatchNormalization (batch statistics in every call, eps 1e-3) -> tanh, then
 * Dense(1).  dims[n_hidden+2] = [in, units..., 1]; `params` (HOST) in the order of
 * Keras' trainable_variables: per block kernel[in][out], bias, gamma, beta; then the
 * output kernel, bias.  (The BN moving statistics are never read by the reference's
 * training path and are not kept.) */
typedef struct {
  int n_hidden;
  const int* dims;
  const float* params;
} bgm_disc_desc;

typedef struct bgm_trainer bgm_trainer; /* opaque: parameters, gradients, Adam moments on the device */

/* State of CausalBGM's EGM phase (causalbgm/base.py:74-87): nets g,e,f,h (parameter
 * group 0, flat in the order g|e|f|h, each net kernel,bias per layer = the order
 * train_gen_step concatenates trainable_variables, :370) and dz_net (group 1), each
 * group with its own Keras Adam(lr, beta_1, beta_2, eps=1e-7) (:86-87). */
int bgm_trainer_create(bgm_trainer** out, const int z_dims[4], int v_dim, int binary_treatment, int use_z_rec,
                       const bgm_net_desc* g_net, const bgm_net_desc* e_net, const bgm_net_desc* f_net,
                       const bgm_net_desc* h_net, const bgm_disc_desc* dz_net, float lr, float beta_1,
                       float beta_2);
void bgm_trainer_destroy(bgm_trainer* t);
/* Device pointers of a group's flat parameter / gradient buffers (for the data-parallel
 * gradient all-reduce, the only collective of the training path) and their length. */
int bgm_trainer_buffers(bgm_trainer* t, int group, int* n_params, float** theta_dev, float** grad_dev);
int bgm_trainer_get_params(bgm_trainer* t, int group, float* host_out);
int bgm_trainer_set_params(bgm_trainer* t, int group, const float* host_in);

/* Gradients of train_disc_step (causalbgm/base.py:305-323) w.r.t. dz_net into the
 * group-1 gradient buffer: e_net forward, three discriminator passes, WGAN-GP term
 * gp_weight * mean((|d D(z_hat)/d z_hat| - 1)^2) with its double backward.  z_dev: (bs,zd)
 * prior draws, v_dev: (bs,v_dim) covariate rows, 2 <= bs <= 32; epsilon: the U(0,1)
 * draw of :307; losses_dev[2] = dz_loss, d_loss. */
int bgm_train_disc_grad(bgm_trainer* t, const float* z_dev, const float* v_dev, int bs, float epsilon,
                        float gp_weight, float* losses_dev, void* stream);
/* Gradients of train_gen_step (causalbgm/base.py:332-370) w.r.t. g,e,f,h into the
 * group-0 gradient buffer; losses_dev[6] = e_loss_adv, l2_loss_v, l2_loss_z, l2_loss_x,
 * l2_loss_y, g_e_loss (:377).  x_dev, y_dev: (bs). */
int bgm_train_gen_grad(bgm_trainer* t, const float* z_dev, const float* v_dev, const float* x_dev,
                       const float* y_dev, int bs, float* losses_dev, void* stream);
/* ---- BGM flavour of the EGM steps (bgm/base.py:190-291) ---- */
/* Generator = BaseVariationalNet in TRAINING mode (input BatchNormalization on batch
 * statistics, moving statistics updated with momentum .99 on every call), encoder e_net,
 * discriminators dz_net (latent) and dx_net (data); LSGAN targets .9/.1; reg weight `alpha`
 * (:279), gradient-penalty weight `gamma` (:236; 0 skips that path).  Parameter group 0 on
 * the device: g = [gamma | beta | hidden kernel,bias ... | [W_mean|W_var] | [b_mean|b_var]],
 * then e; group 1 = [dz | dx].  Adam of the reference: beta = (.5, .9) (bgm/base.py:83-85). */
int bgm_bgmtrainer_create(bgm_trainer** out, const bgm_varnet_desc* g_net, const bgm_net_desc* e_net,
                          const bgm_disc_desc* dz_net, const bgm_disc_desc* dx_net, float lr, float beta_1,
                          float beta_2, float alpha, float gamma);
/* BN moving mean | variance (2*z_dim floats, HOST): read (set == 0) or overwrite. */
int bgm_trainer_bn_moving(bgm_trainer* t, float* host_inout, int set);
/* BGM.train_disc_step gradients (:190-240) into group 1.  noise_dev: (bs,x_dim) N(0,1)
 * draws of reparameterize (:208); eps_z / eps_x: the U(0,1) draws of :199-200;
 * losses_dev[3] = dz_loss, dx_loss, d_loss. */
int bgm_bgm_train_disc_grad(bgm_trainer* t, const float* z_dev, const float* x_dev, int bs, float eps_z, float eps_x,
                            const float* noise_dev, float* losses_dev, void* stream);
/* BGM.train_gen_step gradients (:247-285) into group 0.  noise1/2_dev: the N(0,1) draws of the
 * two reparameterize calls (:259, :267); losses_dev[6] = g_loss_adv, e_loss_adv, l2_loss_z,
 * l2_loss_x, reg_loss, g_e_loss (:291). */
int bgm_bgm_train_gen_grad(bgm_trainer* t, const float* z_dev, const float* x_dev, int bs, const float* noise1_dev,
                           const float* noise2_dev, float* losses_dev, void* stream);

/* One Keras-Adam step of the group on (grad_scale * gradient buffer); grad_scale is
 * 1/world_size after a summing all-reduce, 1 on a single GPU. */
int bgm_train_adam(bgm_trainer* t, int group, float grad_scale, void* stream);
/* ---- iterative phase of CausalBGM.fit (causalbgm/base.py:156-302, :488-514) ---- */
/* Learning rates of g/h/f_optimizer (lr_theta) and posterior_optimizer (lr_z), all Adam
 * beta=(.9,.99) (:89-92); sigma_* >= 0 are the fixed `sigma_*` keys of the params dict,
 * < 0 = learned softplus head.  Resets the Adam state of this phase. */
int bgm_trainer_set_iter(bgm_trainer* t, float lr_theta, float lr_z, float sigma_v, float sigma_x, float sigma_y);
/* update_g_net, update_h_net, update_f_net (:156-243) on the mini-batch rows idx_dev[0..bs)
 * of the device-resident data and latent table zt_dev (n, zd): gradients (apply 0 or 1) and
 * the three Adam updates (apply 1; apply 2 = only the update, after an all-reduce of the
 * group-0 gradient buffer, with grad_scale = 1/world).  losses_dev[6] = loss_v, loss_mse_v,
 * loss_x, loss_mse_x, loss_y, loss_mse_y.  1 <= bs <= 32. */
int bgm_train_iter_nets(bgm_trainer* t, const float* zt_dev, const float* x_dev, const float* y_dev,
                        const float* v_dev, const int* idx_dev, int bs, int apply, float grad_scale,
                        float* losses_dev, void* stream);
/* update_latent_variable_sgd (:246-302): gradient of loss_postrior_z w.r.t. the batch rows
 * of the latent table, then Keras Adam on a gathered variable = a DENSE sweep (m, v decay
 * and every one of the n rows moves).  m_dev, v_dev_adam: (n, zd) Adam moments;
 * slot_dev: (n) int32 scratch initialised to -1.  loss_dev[1] = loss_postrior_z. */
int bgm_train_iter_latent(bgm_trainer* t, float* zt_dev, float* m_dev, float* v_dev_adam, int* slot_dev,
                          long long n, const float* x_dev, const float* y_dev, const float* v_dev,
                          const int* idx_dev, int bs, float* loss_dev, void* stream);
/* CausalBGM.evaluate (:534-556), the part that touches every row: sums_dev[3] = sum (v-v^)^2,
 * sum (x-x^)^2, sum (y-y^)^2 (float64) with z = zt_dev rows, or z = e_net(v) if zt_dev is
 * NULL; z_out_dev (n, zd), if non-NULL, receives the z used (the `data_z_init = e_net(data_v)`
 * of :479). */
int bgm_causal_evaluate(bgm_trainer* t, const float* zt_dev, const float* x_dev, const float* y_dev,
                        const float* v_dev, int n, double* sums_dev, float* z_out_dev, void* stream);

/* ---- iterative phase of BGM.fit (bgm/base.py:145-187, loop :397-415) on a BGM trainer ----
 * bgm_bgmtrainer_set_iter: g_optimizer / posterior_optimizer (:87-88, Adam beta = (0.9, 0.99)) state.
 * bgm_bgm_iter_g: update_g_net on the rows idx_dev of (zt_dev (n,z_dim), x_dev (n,x_dim)): generator in
 *   TRAINING mode (input BatchNormalization with batch statistics, moving statistics updated),
 *   losses_dev[2] = loss_x, loss_mse_x; apply 0: gradients only, 1: + Adam, 2: Adam on the gradients
 *   already in the buffer (after an all-reduce).
 * bgm_bgm_iter_latent: update_latent_variable_sgd: gradient of mean_r(-log p(x_r|z_r) + |z_r|^2/2) w.r.t.
 *   the batch rows through the training-mode BatchNormalization, then Adam on a FRESH variable per
 *   batch (the reference wraps every batch in a new tf.Variable: zero slots, shared step count) written
 *   back into zt_dev; loss_dev[1]; gz_out_dev (bs,z_dim) optionally receives the gradient rows.
 * bgm_bgm_evaluate: sum_dev[1] (float64) = sum over rows and columns of (x - mu(z))^2, generator in
 *   inference mode (evaluate with use_x_sd=False, :446-471). */
int bgm_bgmtrainer_set_iter(bgm_trainer* t, float lr_theta, float lr_z);
int bgm_bgm_iter_g(bgm_trainer* t, const float* zt_dev, const float* x_dev, const int* idx_dev, int bs,
                   int apply, float grad_scale, float* losses_dev, void* stream);
int bgm_bgm_iter_latent(bgm_trainer* t, float* zt_dev, const float* x_dev, const int* idx_dev, int bs,
                        float* loss_dev, float* gz_out_dev, void* stream);
int bgm_bgm_evaluate(bgm_trainer* t, const float* zt_dev, const float* x_dev, int n, double* sum_dev,
                     void* stream);

/* ---- layered training engine: the same training steps for BAYESIAN nets, any batch size, any width ----
 * The single-CTA kernels above fuse a whole step but take deterministic nets, batches <= 32 and widths that fit
 * one SM's shared memory.  bgm_lt_* runs the steps of CausalBGM (train_disc_step / train_gen_step
 * causalbgm/base.py:305-377, update_g/h/f_net :156-243, update_latent_variable_sgd :246-302, evaluate :534-556)
 * as a sequence of per-layer kernels (csrc/layered.cuh): Dense / DenseFlipout forward, input- and parameter-
 * backward, BatchNormalization on batch statistics forward / backward, the losses, Keras Adam.  Nets are
 * described by bgm_bnn_net_desc; bayes == 0: deterministic Dense stacks (bn == NULL, params = kernel, bias per
 * layer).  Parameter group 0 = g | e | f | h, each in Keras trainable_variables order (gamma, beta, then loc, rho,
 * bias per DenseFlipout layer); group 1 = dz_net.  Network noise: Philox (seed; call id 16*counter + k, see
 * bgm_lt_set_call; signs keyed by the row's position in the batch), restated in oracle/train_bnn.py.
 * The gradient-penalty step (bgm_lt_disc_grad) reuses the fused discriminator kernel and keeps its batch <= 32
 * limit; every other step takes any batch size.  The workspace grows (cudaMalloc) when a larger batch arrives. */
typedef struct bgm_lt bgm_lt;
int bgm_lt_create(bgm_lt** out, const int z_dims[4], int v_dim, int binary_treatment, int use_z_rec, int bayes,
                  const bgm_bnn_net_desc* g_net, const bgm_bnn_net_desc* e_net, const bgm_bnn_net_desc* f_net,
                  const bgm_bnn_net_desc* h_net, const bgm_disc_desc* dz_net, float lr, float beta_1, float beta_2,
                  float kl_weight, uint64_t seed);
void bgm_lt_destroy(bgm_lt* t);
int bgm_lt_buffers(bgm_lt* t, int group, int* n_params, float** theta_dev, float** grad_dev);
int bgm_lt_get_params(bgm_lt* t, int group, float* host_out);
int bgm_lt_set_params(bgm_lt* t, int group, const float* host_in);
/* Sets the step counter that keys the network noise of the next step (every step increments it). */
int bgm_lt_set_call(bgm_lt* t, uint32_t call_counter);
int bgm_lt_disc_grad(bgm_lt* t, const float* z_dev, const float* v_dev, int bs, float epsilon, float gp_weight,
                     float* losses_dev, void* stream);
int bgm_lt_gen_grad(bgm_lt* t, const float* z_dev, const float* v_dev, const float* x_dev, const float* y_dev, int bs,
                    float* losses_dev, void* stream);
int bgm_lt_adam(bgm_lt* t, int group, float grad_scale, void* stream);
int bgm_lt_set_iter(bgm_lt* t, float lr_theta, float lr_z, float sigma_v, float sigma_x, float sigma_y);
int bgm_lt_iter_nets(bgm_lt* t, const float* zt_dev, const float* x_dev, const float* y_dev, const float* v_dev,
                     const int* idx_dev, int bs, int apply, float grad_scale, float* losses_dev, void* stream);
int bgm_lt_iter_latent(bgm_lt* t, float* zt_dev, float* m_dev, float* v_adam_dev, int* slot_dev, long long n,
                       const float* x_dev, const float* y_dev, const float* v_dev, const int* idx_dev, int bs,
                       float* loss_dev, float* gz_out_dev, void* stream);
int bgm_lt_evaluate(bgm_lt* t, const float* zt_dev, const float* x_dev, const float* y_dev, const float* v_dev, int n,
                    double* sums_dev, float* z_out_dev, void* stream);

/* BGM flavour of the layered engine (bgm/base.py:145-291): the entry points of bgm_bgmtrainer_create /
 * bgm_bgm_train_disc_grad / bgm_bgm_train_gen_grad / bgm_train_adam / bgm_trainer_bn_moving /
 * bgm_bgmtrainer_set_iter / bgm_bgm_iter_g / bgm_bgm_iter_latent / bgm_bgm_evaluate with the same arguments,
 * the same device parameter layout and the same semantics, for any x_dim and any batch size (the fused kernels
 * stop at x_dim ~ 110 and 32 rows).  gamma (gradient-penalty weight, :236) must be 0: the layered engine has no
 * double backward.  bgm_ltb_encode: e_net(x) for n rows (`data_z_init = self.e_net(data)`, :388). */
typedef struct bgm_ltb bgm_ltb;
int bgm_ltb_create(bgm_ltb** out, const bgm_varnet_desc* g_net, const bgm_net_desc* e_net, const bgm_disc_desc* dz_net,
                   const bgm_disc_desc* dx_net, float lr, float beta_1, float beta_2, float alpha, float gamma);
void bgm_ltb_destroy(bgm_ltb* t);
int bgm_ltb_buffers(bgm_ltb* t, int group, int* n_params, float** theta_dev, float** grad_dev);
int bgm_ltb_get_params(bgm_ltb* t, int group, float* host_out);
int bgm_ltb_bn_moving(bgm_ltb* t, float* host_inout, int set);
int bgm_ltb_adam(bgm_ltb* t, int group, float grad_scale, void* stream);
int bgm_ltb_disc_grad(bgm_ltb* t, const float* z_dev, const float* x_dev, int bs, float eps_z, float eps_x,
                      const float* noise_dev, float* losses_dev, void* stream);
int bgm_ltb_gen_grad(bgm_ltb* t, const float* z_dev, const float* x_dev, int bs, const float* noise1_dev,
                     const float* noise2_dev, float* losses_dev, void* stream);
int bgm_ltb_set_iter(bgm_ltb* t, float lr_theta, float lr_z);
int bgm_ltb_iter_g(bgm_ltb* t, const float* zt_dev, const float* x_dev, const int* idx_dev, int bs, int apply,
                   float grad_scale, float* losses_dev, void* stream);
int bgm_ltb_iter_latent(bgm_ltb* t, float* zt_dev, const float* x_dev, const int* idx_dev, int bs, float* loss_dev,
                        float* gz_out_dev, void* stream);
int bgm_ltb_evaluate(bgm_ltb* t, const float* zt_dev, const float* x_dev, int n, double* sum_dev, void* stream);
int bgm_ltb_encode(bgm_ltb* t, const float* x_dev, int n, float* z_out_dev, void* stream);

/* ---- host-side index / prior streams (no device work) ----
 * NumPy's LEGACY generator (MT19937 `RandomState`) restated natively so that the mini-batch index
 * and prior streams of the training loops are produced bit-exactly off the Python thread:
 * `np.random.choice(n, bs, replace=False)` (causalbgm/base.py:406,:413; a full permutation of n per
 * call), `np.random.normal` (`Gaussian_sampler.get_batch`, prior_samplers.py:46-59) and
 * `np.random.rand`.  The state is exactly `np.random.get_state()`: key[624], pos, has_gauss,
 * cached_gaussian, and is written back with `np.random.set_state()` by the caller. */
typedef struct {
  uint32_t key[624];
  int pos;
  int has_gauss;
  double gauss;
} bgm_mt19937_state;
/* out[size] = np.random.choice(n, size, replace=False); work: n int32 of scratch. */
int bgm_host_choice(bgm_mt19937_state* st, int n, int size, int32_t* out, int32_t* work);
/* out[count] = np.random.normal(loc, scale, count).astype(float32) */
int bgm_host_normal(bgm_mt19937_state* st, double loc, double scale, long long count, float* out);
/* out[count] = np.random.rand(count) */
int bgm_host_rand(bgm_mt19937_state* st, long long count, double* out);
/* `iters` iterations of egm_init's draws in the reference's order (:405-413): g_d_freq x [choice,
 * get_batch], then [get_batch, choice].  idx_out: (iters, g_d_freq+1, bs) int32,
 * z_out: (iters, g_d_freq+1, bs, zd) float32; work: n int32. */
int bgm_host_egm_stream(bgm_mt19937_state* st, int n, int bs, int zd, int g_d_freq, int iters, int32_t* idx_out,
                        float* z_out, int32_t* work);

/* dst[r][:] = src[idx[r]][:dim] -- mini-batch gather from device-resident data (:406-416). */
int bgm_gather_rows(const float* src_dev, int ld, const int* idx_dev, int bs, int dim, float* dst_dev,
                    void* stream);

/* Dependent-FFMA micro-benchmark: measured fp32 FMA peak of the device in TFLOP/s
 * (the roofline denominator for the SIMT kernels; MEASURED_PEAKS.json has none). */
int bgm_fp32_peak_tflops(double* tflops, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BGM_B200_H */
