"""Oracle restatement of the reference's BAYESIAN networks and of the CausalBGM posterior path
running on them (`use_bnn=True`, the default of every shipped CausalBGM config).  TEST
INFRASTRUCTURE -- see oracle/__init__.py.

Follows `src/bayesgm/models/networks/bnn.py:4-38` (`BayesianFullyConnectedNet`): an input
`BatchNormalization` followed by `tfp.layers.DenseFlipout` layers with LeakyReLU(0.2) between
them, and `src/bayesgm/models/causalbgm/base.py:765-817 / :820-904 / :671-763` evaluated on such
nets.  Third-party semantics restated (TF 2.10 / TFP 0.18, not vendored, parity unpinned):

* `DenseFlipout.call` (tfp/python/layers/dense_variational.py): one draw per CALL of the kernel
  perturbation dW = sigma * eps, eps ~ N(0,1) [in,out], sigma = finfo(float32).eps + softplus(rho)
  (`default_mean_field_normal_fn`), shared by all rows of the batch, and per-ROW Rademacher sign
  vectors s_in (in) and s_out (out):  y = x @ loc + ((x * s_in) @ dW) * s_out + bias.  The bias
  posterior is deterministic (`is_singular=True`).  Initialisers: loc, bias ~ N(0, 0.1^2),
  rho ~ N(-3, 0.1^2).  KL(q || N(0,1)) of every kernel is added to `losses`.
* `BatchNormalization` (eps 1e-3, momentum .99) is called as `self.norm_layer(inputs)` inside
  `call(self, inputs, training=True)` (bnn.py:25-27): Keras hands the outer call's training value
  (the signature default, True) down to the sub-layer, so the input is normalised with BATCH
  statistics (biased variance) in EVERY call -- inside get_log_posterior, predict and evaluate too.
  Consequence restated faithfully: a constant input column (the dose column that
  infer_from_latent_posterior / evaluate tile, :736-741, :563-566) normalises to exactly beta.

Noise: a `provider` supplies, per (net, layer, call), eps [K,N], s_in [rows,K], s_out [rows,N].
`PhiloxFlipout` reproduces the streams of the CUDA kernels (csrc/bnn.cuh) so that kernel and oracle
can be compared value for value; `NumpyFlipout` draws from a NumPy generator (CPU baseline).
"""
import numpy as np

from .nets import leaky_relu, softplus, sigmoid, BN_EPS
from .philox import philox4x32_10, philox_normal4

SCALE_EPS = np.float32(np.finfo(np.float32).eps)
NET_ID = dict(g=0, f=1, h=2, e=3)
NOISE_BNN_W, NOISE_BNN_SIGN = 5, 6


def init_bnn(rs, dims, bn_random=False):
    """Parameters of a BayesianFullyConnectedNet: dict(bn=dict(gamma, beta, mean, var),
    layers=[(loc[in,out], rho[in,out], bias[out]), ...]) with TFP's default initialisers."""
    k = dims[0]
    if bn_random:
        bn = dict(gamma=(1 + 0.2 * rs.standard_normal(k)).astype(np.float32),
                  beta=(0.2 * rs.standard_normal(k)).astype(np.float32))
    else:
        bn = dict(gamma=np.ones(k, np.float32), beta=np.zeros(k, np.float32))
    bn['mean'] = np.zeros(k, np.float32)
    bn['var'] = np.ones(k, np.float32)
    layers = []
    for i in range(len(dims) - 1):
        loc = (0.1 * rs.standard_normal((dims[i], dims[i + 1]))).astype(np.float32)
        rho = (-3.0 + 0.1 * rs.standard_normal((dims[i], dims[i + 1]))).astype(np.float32)
        bias = (0.1 * rs.standard_normal(dims[i + 1])).astype(np.float32)
        layers.append((loc, rho, bias))
    return dict(bn=bn, layers=layers)


def kernel_scale(rho):
    return (SCALE_EPS + softplus(rho)).astype(np.float32)


def kl_divergence(net):
    """sum over the DenseFlipout kernels of KL(N(loc, sigma) || N(0, 1)) (float64)."""
    tot = 0.0
    for loc, rho, _ in net['layers']:
        s = kernel_scale(rho).astype(np.float64)
        tot += float(np.sum(-np.log(s) + 0.5 * (s * s + loc.astype(np.float64) ** 2) - 0.5))
    return tot


class NumpyFlipout(object):
    def __init__(self, rs):
        self.rs = rs

    def flipout(self, net, layer, call, rows, K, N):
        eps = self.rs.standard_normal((K, N)).astype(np.float32)
        s_in = (1.0 - 2.0 * self.rs.randint(0, 2, size=(rows, K))).astype(np.float32)
        s_out = (1.0 - 2.0 * self.rs.randint(0, 2, size=(rows, N))).astype(np.float32)
        return eps, s_in, s_out


class PhiloxFlipout(object):
    """The Philox4x32-10 streams of csrc/bnn.cuh.
    eps: element (k, c) of layer (net, layer) is normal (k*N4 + c) % 4 of
         normal4(seed, row = slice<<44 | net<<40 | layer<<36 | (k*N4 + c)/4, t = call, NOISE_BNN_W, 0),
         N4 = N rounded up to a multiple of 4.
    signs: bit i of the 128-bit blocks noise_block(seed, global row, t = call, NOISE_BNN_SIGN,
         j = net<<8 | layer<<4 | blk) (word i/32 within the block, bit i%32; blocks concatenated):
         s_in[k] = bit k, s_out[c] = bit K + c; a set bit is -1."""

    def __init__(self, seed, slice_id=0, row_offset=0, rows_of=None):
        self.seed, self.slice_id, self.row_offset = int(seed), int(slice_id), int(row_offset)
        self.rows_of = rows_of           # optional: call -> explicit global row ids

    def flipout(self, net, layer, call, rows, K, N):
        net_id = NET_ID[net] if isinstance(net, str) else int(net)
        N4 = (N + 3) // 4 * 4
        groups = np.arange(K * N4 // 4, dtype=np.int64)
        base = (self.slice_id << 44) | (net_id << 40) | (layer << 36)
        eps = philox_normal4(self.seed, base + groups, call, NOISE_BNN_W, 0).reshape(K, N4)[:, :N]
        grow = (np.arange(rows, dtype=np.int64) + self.row_offset) if self.rows_of is None \
            else np.asarray(self.rows_of(call), np.int64)
        nblk = (K + N + 127) // 128
        words = []
        key = np.array([self.seed & 0xFFFFFFFF, (self.seed >> 32) & 0xFFFFFFFF], np.uint64)
        for b in range(nblk):
            j = (net_id << 8) | (layer << 4) | b
            c = np.stack([np.full(grow.shape, call & 0xFFFFFFFF, np.uint64), (grow & 0xFFFFFFFF).astype(np.uint64),
                          ((grow >> 32) & 0xFFFFFFFF).astype(np.uint64),
                          np.full(grow.shape, ((NOISE_BNN_SIGN << 24) | j) & 0xFFFFFFFF, np.uint64)],
                         axis=-1).astype(np.uint32)
            words.append(philox4x32_10(c, key))
        w = np.concatenate(words, axis=-1)                                  # (rows, 4*nblk) uint32
        bits = ((w[:, :, None] >> np.arange(32, dtype=np.uint32)[None, None, :]) & 1).reshape(len(grow), -1)
        sg = (1.0 - 2.0 * bits[:, :K + N]).astype(np.float32)
        return eps.astype(np.float32), sg[:, :K], sg[:, K:K + N]


def batch_norm_train(bn, x):
    """Keras BatchNormalization, training mode: batch mean / biased variance, eps 1e-3."""
    x = np.asarray(x, np.float32)
    mean = x.astype(np.float64).mean(axis=0)
    var = ((x.astype(np.float64) - mean) ** 2).mean(axis=0)
    inv = (np.float32(1.0) / np.sqrt(var.astype(np.float32) + BN_EPS)).astype(np.float32)
    return ((x - mean.astype(np.float32)) * inv * bn['gamma'] + bn['beta']).astype(np.float32)


def bnn_forward(net, x, provider, net_name, call, constant_cols=()):
    """BayesianFullyConnectedNet.call (bnn.py:25-38).  `constant_cols`: input columns known to be
    constant over the batch (a tiled dose): their batch variance is exactly 0 and the normalised
    value exactly 0, so BN returns beta (fp32 rounding of TF's batch mean aside)."""
    x = np.asarray(x, np.float32)
    h = batch_norm_train(net['bn'], x)
    for c in constant_cols:
        h[:, c] = net['bn']['beta'][c]
    L = len(net['layers'])
    rows = x.shape[0]
    for l, (loc, rho, bias) in enumerate(net['layers']):
        K, N = loc.shape
        eps, s_in, s_out = provider.flipout(net_name, l, call, rows, K, N)
        dW = (kernel_scale(rho) * eps).astype(np.float32)
        out = (h @ loc + ((h * s_in) @ dW) * s_out + bias).astype(np.float32)
        h = leaky_relu(out) if l < L - 1 else out
    return h


def log_posterior(params, nets, data_x, data_y, data_v, data_z, provider, call=0, eps=1e-6):
    """causalbgm/base.py:765-817 on Bayesian nets: one fresh Flipout draw and one set of batch
    statistics per net per call."""
    d0, d1, d2, _ = params['z_dims']
    p = params['v_dim']
    f32 = np.float32
    z = np.asarray(data_z, f32)
    x = np.asarray(data_x, f32).reshape(-1, 1)
    y = np.asarray(data_y, f32).reshape(-1, 1)
    v = np.asarray(data_v, f32)
    z0, z1, z2 = z[:, :d0], z[:, d0:d0 + d1], z[:, d0 + d1:d0 + d1 + d2]
    g_out = bnn_forward(nets['g'], z, provider, 'g', call)
    mu_v = g_out[:, :p]
    s2v = f32(params['sigma_v'] ** 2) if 'sigma_v' in params else softplus(g_out[:, -1]) + f32(eps)
    h_out = bnn_forward(nets['h'], np.concatenate([z0, z2], axis=-1), provider, 'h', call)
    mu_x = h_out[:, :1]
    s2x = f32(params['sigma_x'] ** 2) if 'sigma_x' in params else softplus(h_out[:, -1]) + f32(eps)
    f_out = bnn_forward(nets['f'], np.concatenate([z0, z1, x], axis=-1), provider, 'f', call)
    mu_y = f_out[:, :1]
    s2y = f32(params['sigma_y'] ** 2) if 'sigma_y' in params else softplus(f_out[:, -1]) + f32(eps)
    loss_pv = ((v - mu_v) ** 2).sum(axis=1) / (2 * s2v) + f32(p) * np.log(s2v) / 2
    if params['binary_treatment']:
        l = mu_x[:, 0]
        loss_px = np.maximum(l, 0) - l * x[:, 0] + np.log1p(np.exp(-np.abs(l)))
    else:
        loss_px = ((x - mu_x) ** 2).sum(axis=1) / (2 * s2x) + np.log(s2x) / 2
    loss_py = ((y - mu_y) ** 2).sum(axis=1) / (2 * s2y) + np.log(s2y) / 2
    loss_prior = (z ** 2).sum(axis=1) / 2
    return (-(loss_pv + loss_px + loss_py + loss_prior)).astype(f32)


def mh_sampler(params, nets, data, provider, initial_q_sd=1.0, q_sd=None, burn_in=5000, n_keep=3000,
               target_acceptance_rate=0.25, tolerance=0.05, adjustment_interval=50, adaptive_sd=None,
               window_size=100, noise=None, return_trace=False):
    """causalbgm/base.py:820-904 on Bayesian nets: BOTH log-posteriors are re-evaluated every
    iteration (:865-866) with fresh network noise -- call ids 2t (proposal) and 2t+1 (current)."""
    from .causal import NumpyGlobalNoise
    data_x, data_y, data_v = data
    n = len(data_x)
    zd = sum(params['z_dims'])
    noise = noise if noise is not None else NumpyGlobalNoise()
    cur = noise.initial(n, zd)
    samples, t, recent = [], 0, []
    if adaptive_sd is None:
        adaptive_sd = (q_sd is None or q_sd <= 0)
    if adaptive_sd:
        q_sd = initial_q_sd
    trace = dict(accept=[], lp_prop=[], lp_cur=[], q_sd=[])
    while len(samples) < n_keep:
        prop = cur + noise.proposal(q_sd, n, zd)
        lp_p = log_posterior(params, nets, data_x, data_y, data_v, prop, provider, call=2 * t)
        lp_c = log_posterior(params, nets, data_x, data_y, data_v, cur, provider, call=2 * t + 1)
        ratio = np.exp(np.minimum(lp_p - lp_c, 0))
        idx = noise.uniform(n) < ratio
        cur[idx] = prop[idx]
        if return_trace:
            trace['accept'].append(idx.copy())
            trace['lp_prop'].append(lp_p)
            trace['lp_cur'].append(lp_c)
            trace['q_sd'].append(float(q_sd))
        recent.append(idx)
        if len(recent) > window_size:
            recent = recent[-window_size:]
        if adaptive_sd and t < burn_in and t % adjustment_interval == 0 and t > 0:
            rate = np.sum(recent) / (len(recent) * n)
            if rate < target_acceptance_rate - tolerance:
                q_sd *= 0.9
            elif rate > target_acceptance_rate + tolerance:
                q_sd *= 1.1
        if t >= burn_in:
            samples.append(cur.copy())
        t += 1
    out = np.array(samples)
    if return_trace:
        trace['q_sd_final'] = float(q_sd)
        return out, trace
    return out


def infer_from_latent_posterior(params, nets, data_posterior_z, provider, x_values=None, sample_y=True,
                                eps=1e-6, normal_fn=None):
    """causalbgm/base.py:671-763 on a Bayesian f_net: one f_net CALL per (kept state s, dose j)
    (the bodies of the nested tf.map_fn), call id s * n_x + j; binary: doses (1, 0)."""
    zs = np.asarray(data_posterior_z, np.float32)
    n_keep, n, _ = zs.shape
    d0, d1, _, _ = params['z_dims']
    f32 = np.float32
    binary = bool(params['binary_treatment'])
    xs = np.array([1.0, 0.0]) if binary else np.atleast_1d(np.asarray(x_values, dtype=float))
    n_x = len(xs)
    mu = np.empty((n_x, n_keep, n), f32)
    s2 = np.empty((n_x, n_keep, n), f32)
    for s in range(n_keep):
        for j, xv in enumerate(xs):
            inp = np.concatenate([zs[s][:, :d0], zs[s][:, d0:d0 + d1], np.full((n, 1), f32(xv), f32)], axis=-1)
            out = bnn_forward(nets['f'], inp, provider, 'f', s * n_x + j, constant_cols=(d0 + d1,))
            mu[j, s] = out[:, 0]
            s2[j, s] = f32(params['sigma_y'] ** 2) if 'sigma_y' in params else softplus(out[:, 1]) + f32(eps)
    if sample_y:
        y = (mu + np.sqrt(s2) * normal_fn(mu.shape).astype(f32)).astype(f32)
    else:
        y = mu
    if binary:
        return (y[0] - y[1]).astype(f32)
    return y.mean(axis=2)


def evaluate(params, nets, data, provider, data_z=None, nb_intervals=200):
    """causalbgm/base.py:534-570 on Bayesian nets.  Call ids: e 0, g 0, f 0 (fit), h 0, then the
    dose / ITE calls 1 + j on f."""
    from .causal import percentile_nearest
    data_x, data_y, data_v = [np.asarray(a, np.float32) for a in data]
    data_x, data_y = data_x.reshape(-1, 1), data_y.reshape(-1, 1)
    d0, d1, d2, _ = params['z_dims']
    p = params['v_dim']
    n = len(data_x)
    z = bnn_forward(nets['e'], data_v, provider, 'e', 0) if data_z is None else np.asarray(data_z, np.float32)
    z0, z1, z2 = z[:, :d0], z[:, d0:d0 + d1], z[:, d0 + d1:d0 + d1 + d2]
    v_pred = bnn_forward(nets['g'], z, provider, 'g', 0)[:, :p]
    y_pred = bnn_forward(nets['f'], np.concatenate([z0, z1, data_x], axis=-1), provider, 'f', 0)[:, :1]
    x_pred = bnn_forward(nets['h'], np.concatenate([z0, z2], axis=-1), provider, 'h', 0)[:, :1]
    if params['binary_treatment']:
        x_pred = sigmoid(x_pred)
    mse_v, mse_x, mse_y = np.mean((data_v - v_pred) ** 2), np.mean((data_x - x_pred) ** 2), np.mean((data_y - y_pred) ** 2)

    def f_at(xv, call):
        inp = np.concatenate([z0, z1, np.full((n, 1), np.float32(xv), np.float32)], axis=-1)
        return bnn_forward(nets['f'], inp, provider, 'f', call, constant_cols=(d0 + d1,))[:, :1]
    if params['binary_treatment']:
        return f_at(1.0, 1) - f_at(0.0, 2), mse_x, mse_y, mse_v
    x_min, x_max = percentile_nearest(data_x, 5.0), percentile_nearest(data_x, 95.0)
    xs = np.linspace(x_min, x_max, nb_intervals).astype(np.float32)
    dose = np.array([f_at(xv, 1 + j).mean() for j, xv in enumerate(xs)], np.float32)
    return dose, mse_x, mse_y, mse_v
