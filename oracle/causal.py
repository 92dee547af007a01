"""Oracle restatement of CausalBGM's posterior-sampling path (test infrastructure).

Follows `src/bayesgm/models/causalbgm/base.py`:
  * get_log_posterior            :765-817 -> `log_posterior`
  * metropolis_hastings_sampler  :820-904 -> `mh_sampler`
  * infer_from_latent_posterior  :671-763 -> `infer_from_latent_posterior`
  * predict                      :573-668 -> `predict`
  * evaluate                     :534-570 -> `evaluate`
`nets` is a dict {'g','e','f','h'} of Keras-layout Dense stacks (oracle.nets).
`params` is the reference's config dict (src/configs/*.yaml).
"""
import numpy as np

from .nets import mlp_forward, softplus, sigmoid


def _split(params):
    d0, d1, d2, d3 = params['z_dims']
    return d0, d1, d2, d3


def log_posterior(params, nets, data_x, data_y, data_v, data_z, eps=1e-6, prior=None):
    """causalbgm/base.py:765-817, float32 throughout.  `prior=(mu_z (n,zd), sigma2_z (n,))` replaces
    the N(0,I) prior by the conditional one of IdentifiableCausalBGM.get_log_posterior
    (causalbgm/identifiable.py:505-556, prior term :540-548)."""
    d0, d1, d2, _ = _split(params)
    p = params['v_dim']
    f32 = np.float32
    z = np.asarray(data_z, f32)
    x = np.asarray(data_x, f32)
    y = np.asarray(data_y, f32)
    v = np.asarray(data_v, f32)
    z0, z1, z2 = z[:, :d0], z[:, d0:d0 + d1], z[:, d0 + d1:d0 + d1 + d2]

    g_out = mlp_forward(nets['g'], z)                                   # :779
    mu_v = g_out[:, :p]
    if 'sigma_v' in params:                                             # :781-784
        s2v = f32(params['sigma_v'] ** 2)
    else:
        s2v = softplus(g_out[:, -1]) + f32(eps)

    h_out = mlp_forward(nets['h'], np.concatenate([z0, z2], axis=-1))   # :786
    mu_x = h_out[:, :1]
    if 'sigma_x' in params:
        s2x = f32(params['sigma_x'] ** 2)
    else:
        s2x = softplus(h_out[:, -1]) + f32(eps)

    f_out = mlp_forward(nets['f'], np.concatenate([z0, z1, x], axis=-1))  # :793
    mu_y = f_out[:, :1]
    if 'sigma_y' in params:
        s2y = f32(params['sigma_y'] ** 2)
    else:
        s2y = softplus(f_out[:, -1]) + f32(eps)

    loss_pv = ((v - mu_v) ** 2).sum(axis=1) / (2 * s2v) + f32(p) * np.log(s2v) / 2  # :800
    if params['binary_treatment']:                                      # :803-804
        l = mu_x[:, 0]
        loss_px = np.maximum(l, 0) - l * x[:, 0] + np.log1p(np.exp(-np.abs(l)))
    else:                                                               # :806
        loss_px = ((x - mu_x) ** 2).sum(axis=1) / (2 * s2x) + np.log(s2x) / 2
    loss_py = ((y - mu_y) ** 2).sum(axis=1) / (2 * s2y) + np.log(s2y) / 2   # :809
    if prior is None:
        loss_prior = (z ** 2).sum(axis=1) / 2                           # :812
    else:                                                               # identifiable.py:540-548
        mu_z, s2z = np.asarray(prior[0], f32), np.asarray(prior[1], f32).reshape(-1)
        term1 = ((z - mu_z) ** 2).sum(axis=1) / (f32(2.0) * s2z)
        term2 = f32(z.shape[1]) * np.log(s2z) / f32(2.0)
        loss_prior = term1 + term2
    return (-(loss_pv + loss_px + loss_py + loss_prior)).astype(f32)    # :814-816


class InjectedNoise(object):
    """Stands in for NumPy's global generator: replays pre-drawn noise.

    `z0` (n,zd) float32; `eps` (T,n,zd) float32 UNIT normals (scaled by q_sd at
    use, like np.random.normal(0, q_sd)); `u` (T,n) float64 uniforms.
    """

    def __init__(self, z0, eps, u):
        self.z0, self.eps, self.u, self.t = z0, eps, u, 0

    def initial(self, n, zd):
        return self.z0.astype(np.float32).copy()

    def proposal(self, q_sd, n, zd):
        # normal(0, q_sd) scales the unit draw in float64; .astype('float32') rounds once (:862)
        return (np.float64(q_sd) * self.eps[self.t].astype(np.float64)).astype(np.float32)

    def uniform(self, n):
        u = self.u[self.t]
        self.t += 1
        return u


class NumpyGlobalNoise(object):
    """The reference's RNG call order on NumPy's legacy generator:
    normal(0,1,(n,zd)) once (:842); per iteration normal(0,q_sd,(n,zd)) (:862) then
    rand(n) (:870); normals are float64 then .astype('float32')."""

    def __init__(self, rs=None):
        self.rs = rs if rs is not None else np.random

    def initial(self, n, zd):
        return self.rs.normal(0, 1, size=(n, zd)).astype('float32')

    def proposal(self, q_sd, n, zd):
        return self.rs.normal(0, q_sd, size=(n, zd)).astype('float32')

    def uniform(self, n):
        return self.rs.rand(n)


def mh_sampler(params, nets, data, initial_q_sd=1.0, q_sd=None, burn_in=5000, n_keep=3000,
               target_acceptance_rate=0.25, tolerance=0.05, adjustment_interval=50,
               adaptive_sd=None, window_size=100, noise=None, recompute_current=True,
               return_trace=False, prior=None, nets_at=None):
    """causalbgm/base.py:820-904.

    `prior`: conditional prior rows of IdentifiableCausalBGM (identifiable.py:559-616 is this loop
    with `data_u` passed on).  `nets_at(t)`: the nets of iteration t -- FullMCMCCausalBGM draws one
    posterior weight sample per iteration and evaluates BOTH states with it
    (causalbgm/fullmcmc.py:438-449), so the current state's value cannot be cached.

    `recompute_current=True` evaluates the current state's log-posterior every
    iteration exactly like :866; False caches it (what the CUDA kernel does --
    identical values, half the work).  `return_trace` additionally returns the
    per-iteration accept masks, q_sd history and log-posteriors for state-for-state
    comparison.
    """
    data_x, data_y, data_v = data
    n = len(data_x)
    zd = sum(params['z_dims'])
    noise = noise if noise is not None else NumpyGlobalNoise()

    current_state = noise.initial(n, zd)                                 # :842
    samples, counter = [], 0
    recent = []
    if adaptive_sd is None:                                              # :852-853
        adaptive_sd = (q_sd is None or q_sd <= 0)
    if adaptive_sd:
        q_sd = initial_q_sd
    trace = dict(accept=[], q_sd=[], lp_prop=[], lp_cur=[])
    cur_lp = None
    while len(samples) < n_keep:                                         # :860
        proposed_state = current_state + noise.proposal(q_sd, n, zd)     # :862
        if nets_at is not None:                                          # fullmcmc.py:441-444
            nets = nets_at(counter)
        prop_lp = log_posterior(params, nets, data_x, data_y, data_v, proposed_state, prior=prior)  # :865
        if recompute_current or cur_lp is None or nets_at is not None:
            cur_lp = log_posterior(params, nets, data_x, data_y, data_v, current_state, prior=prior)  # :866
        ratio = np.exp(np.minimum(prop_lp - cur_lp, 0))                  # :868 (float32)
        indices = noise.uniform(n) < ratio                               # :870 (f64 < f32)
        current_state[indices] = proposed_state[indices]                 # :871
        if not recompute_current:
            cur_lp = np.where(indices, prop_lp, cur_lp)
        if return_trace:
            trace['accept'].append(indices.copy())
            trace['q_sd'].append(float(q_sd))
            trace['lp_prop'].append(prop_lp)
            trace['lp_cur'].append(np.asarray(cur_lp).copy())
        recent.append(indices)                                           # :874-877
        if len(recent) > window_size:
            recent = recent[-window_size:]
        if adaptive_sd and counter < burn_in and counter % adjustment_interval == 0 and counter > 0:
            rate = np.sum(recent) / (len(recent) * n)                    # :882
            if rate < target_acceptance_rate - tolerance:                # :887-890
                q_sd *= 0.9
            elif rate > target_acceptance_rate + tolerance:
                q_sd *= 1.1
        if counter >= burn_in:                                           # :895-896
            samples.append(current_state.copy())
        counter += 1
    out = np.array(samples)                                              # :904
    if return_trace:
        trace['final_acceptance_rate'] = np.sum(recent) / (len(recent) * n)  # :901
        trace['q_sd_final'] = float(q_sd)
        return out, trace
    return out


def f_net_on(params, nets, z, xcol):
    d0, d1, _, _ = _split(params)
    inp = np.concatenate([z[:, :d0], z[:, d0:d0 + d1], xcol], axis=-1).astype(np.float32)
    return mlp_forward(nets['f'], inp)


def conditional_prior(params, prior_net, segments, eps=1e-6):
    """IdentifiableCausalBGM's p(z|u) (causalbgm/identifiable.py:540-543): prior_net on the one-hot
    rows of `segments` (:566-570) -> (mu_z (n,zd), sigma2_z (n,)), float32."""
    zd = sum(params['z_dims'])
    u = np.eye(int(params.get('n_segments', 10)), dtype=np.float32)[np.asarray(segments)]
    out = mlp_forward(prior_net, u)
    return out[:, :zd], softplus(out[:, -1]) + np.float32(eps)


def infer_from_latent_posterior(params, nets, data_posterior_z, x_values=None, sample_y=True,
                                eps=1e-6, normal_fn=None, nets_at=None):
    """causalbgm/base.py:671-763.  `nets_at(s)`: the nets paired with kept state s
    (FullMCMCCausalBGM.infer_from_latent_posterior, causalbgm/fullmcmc.py:285-342; same outputs,
    this function keeps the base class's (len(x_values), n_keep) orientation).

    `normal_fn(shape)` supplies the N(0,1) draws of tf.random.normal (:704,:725,
    :753); required when sample_y=True (TF's stream is not reproducible outside TF).
    Binary: returns ITE (n_keep, n).  Continuous: ADRF draws (len(x_values), n_keep).
    """
    zs = np.asarray(data_posterior_z, np.float32)
    n_keep, n, _ = zs.shape
    f32 = np.float32

    def draw(mu, s2):
        if not sample_y:
            return mu
        return (mu + np.sqrt(s2) * normal_fn(mu.shape).astype(f32)).astype(f32)

    def head(out):
        mu = out[:, 0]
        s2 = f32(params['sigma_y'] ** 2) if 'sigma_y' in params else softplus(out[:, 1]) + f32(eps)
        return mu, s2

    if params['binary_treatment']:
        pos = np.empty((n_keep, n), f32)
        neg = np.empty((n_keep, n), f32)
        mus, s2s = [], []
        for xv in (1.0, 0.0):                                            # :690-729
            mu_all = np.empty((n_keep, n), f32)
            s2_all = np.empty((n_keep, n), f32)
            for s in range(n_keep):
                mu, s2 = head(f_net_on(params, nets_at(s) if nets_at else nets, zs[s], np.full((n, 1), xv, f32)))
                mu_all[s] = mu
                s2_all[s] = s2
            mus.append(mu_all)
            s2s.append(s2_all)
        pos = draw(mus[0], s2s[0])
        neg = draw(mus[1], s2s[1])
        return (pos - neg).astype(f32)                                   # :731
    x_values = np.atleast_1d(np.asarray(x_values, dtype=float))
    out = np.empty((len(x_values), n_keep), f32)
    for j, xv in enumerate(x_values):                                    # :736-761
        mu_all = np.empty((n_keep, n), f32)
        s2_all = np.empty((n_keep, n), f32)
        for s in range(n_keep):
            mu, s2 = head(f_net_on(params, nets_at(s) if nets_at else nets, zs[s], np.full((n, 1), f32(xv), f32)))
            mu_all[s] = mu
            s2_all[s] = s2
        out[j] = draw(mu_all, s2_all).mean(axis=1)                       # :759
    return out


def predict(params, nets, data, alpha=0.01, n_mcmc=3000, burn_in=5000, x_values=None, q_sd=1.0,
            sample_y=True, bs=10000, noise_factory=None, normal_fn=None):
    """causalbgm/base.py:573-668.  `noise_factory(start, end)` returns the noise
    source for one `bs` slice (default: NumPy global generator, consumed sequentially
    across slices as in the reference)."""
    assert 0 < alpha < 1
    if not params['binary_treatment'] and x_values is None:
        raise ValueError("For continuous treatment, 'x_values' must not be None. "
                         "Provide a list or a single treatment value.")
    if x_values is not None:
        x_values = np.array([x_values], dtype=float) if np.isscalar(x_values) \
            else np.array(x_values, dtype=float)
    data_x, data_y, data_v = data
    n_test = len(data_x)
    bs = max(1, int(bs))
    if params['binary_treatment']:
        ite_mean = np.zeros(n_test, np.float32)
        up = np.zeros(n_test, np.float32)
        lo = np.zeros(n_test, np.float32)
        for start in range(0, n_test, bs):
            end = min(start + bs, n_test)
            batch = (data_x[start:end], data_y[start:end], data_v[start:end])
            noise = noise_factory(start, end) if noise_factory else None
            zs = mh_sampler(params, nets, batch, burn_in=burn_in, n_keep=n_mcmc, q_sd=q_sd, noise=noise)
            ce = infer_from_latent_posterior(params, nets, zs, x_values, sample_y, normal_fn=normal_fn)
            ite_mean[start:end] = np.mean(ce, axis=0)                    # :640-642
            up[start:end] = np.quantile(ce, 1 - alpha / 2, axis=0)
            lo[start:end] = np.quantile(ce, alpha / 2, axis=0)
        return ite_mean, np.stack([lo, up], axis=1)
    sums = np.zeros((len(x_values), n_mcmc), np.float32)
    n_seen = 0
    for start in range(0, n_test, bs):
        end = min(start + bs, n_test)
        batch = (data_x[start:end], data_y[start:end], data_v[start:end])
        noise = noise_factory(start, end) if noise_factory else None
        zs = mh_sampler(params, nets, batch, burn_in=burn_in, n_keep=n_mcmc, q_sd=q_sd, noise=noise)
        be = infer_from_latent_posterior(params, nets, zs, x_values, sample_y, normal_fn=normal_fn)
        sums += be * (end - start)                                       # :660
        n_seen += end - start
    ce = sums / float(n_seen)                                            # :663-667
    adrf = np.mean(ce, axis=1)
    up = np.quantile(ce, 1 - alpha / 2, axis=1)
    lo = np.quantile(ce, alpha / 2, axis=1)
    return adrf, np.stack([lo, up], axis=1)


def percentile_nearest(a, q):
    """tfp.stats.percentile(x, q) with the default interpolation='nearest' over the
    flattened input (TFP 0.18, not vendored): index round((n-1)*q/100) of the sort."""
    s = np.sort(np.asarray(a).ravel())
    idx = int(np.round((len(s) - 1) * q / 100.0))
    return s[idx]


def evaluate(params, nets, data, data_z=None, nb_intervals=200):
    """causalbgm/base.py:534-570."""
    data_x, data_y, data_v = [np.asarray(a, np.float32) for a in data]
    d0, d1, d2, _ = _split(params)
    p = params['v_dim']
    if data_z is None:
        data_z = mlp_forward(nets['e'], data_v)
    z = np.asarray(data_z, np.float32)
    z0, z1, z2 = z[:, :d0], z[:, d0:d0 + d1], z[:, d0 + d1:d0 + d1 + d2]
    v_pred = mlp_forward(nets['g'], z)[:, :p]
    y_pred = mlp_forward(nets['f'], np.concatenate([z0, z1, data_x], axis=-1))[:, :1]
    x_pred = mlp_forward(nets['h'], np.concatenate([z0, z2], axis=-1))[:, :1]
    if params['binary_treatment']:
        x_pred = sigmoid(x_pred)
    mse_v = np.mean((data_v - v_pred) ** 2)
    mse_x = np.mean((data_x - x_pred) ** 2)
    mse_y = np.mean((data_y - y_pred) ** 2)
    n = len(data_x)
    if params['binary_treatment']:
        pos = f_net_on(params, nets, z, np.ones((n, 1), np.float32))[:, :1]
        neg = f_net_on(params, nets, z, np.zeros((n, 1), np.float32))[:, :1]
        return pos - neg, mse_x, mse_y, mse_v
    x_min = percentile_nearest(data_x, 5.0)
    x_max = percentile_nearest(data_x, 95.0)
    xs = np.linspace(x_min, x_max, nb_intervals).astype(np.float32)
    dose = np.array([f_net_on(params, nets, z, np.full((n, 1), xv, np.float32))[:, :1].mean()
                     for xv in xs], np.float32)
    return dose, mse_x, mse_y, mse_v
