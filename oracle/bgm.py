"""Oracle restatement of BGM's HMC posterior path (test infrastructure).

Follows `src/bayesgm/models/bgm/base.py`:
  * get_log_posterior      :665-705 -> `log_posterior` (+ hand-written d/dz)
  * tfp_mcmc_sampler       :709-830 -> `pad_index_lists`, `hmc_sampler`
  * predict_on_posteriors  :511-525 -> `predict_on_posteriors`
  * predict                :527-663 -> `predict`
The integrator / accept / step-size adaptation arithmetic lives in
tensorflow-probability==0.18.0 (src/setup.py:14-16), which is NOT under
/root/reference; it is restated from its published algorithm:
  HamiltonianMonteCarlo (unit mass): p ~ N(0,I); p += eps/2 * grad; L x { z += eps*p;
  grad = d logp(z); p += eps*grad }; p -= eps/2 * grad;
  log_accept = logp(z') - logp(z) + (|p|^2/2 - |p'|^2/2); accept iff log(U) < log_accept.
  SimpleStepSizeAdaptation(adaptation_rate=0.01, target 0.75): for the first
  int(0.8*burn_in) steps, with a = log-mean-exp over ALL chains of min(log_accept,0),
  eps <- eps*1.01 if a > log(0.75) else eps/1.01 (applies from the next step on).
"""
import numpy as np

from .nets import variational_forward, bn_inference, leaky_relu, LEAKY_SLOPE, BN_EPS


def pad_index_lists(ind_x1, n, x_dim=None):
    """bgm/base.py:741-775: list-of-lists -> (n,K_max) int32 indices + float mask."""
    if isinstance(ind_x1, (list, tuple)) and len(ind_x1) > 0 and isinstance(ind_x1[0], (list, tuple)):
        assert len(ind_x1) == n
        max_len = max(len(r) for r in ind_x1) if n > 0 else 0
        assert max_len > 0, "No observed features"
        ind = np.zeros((n, max_len), np.int32)
        mask = np.zeros((n, max_len), np.float32)
        for i, row in enumerate(ind_x1):
            L = len(row)
            if L > 0:
                ind[i, :L] = np.array(row, np.int32)
                mask[i, :L] = 1.0
        return ind, mask
    ind = np.asarray(ind_x1, np.int32)
    if ind.ndim == 1:
        ind = np.broadcast_to(ind[None, :], (n, ind.shape[0])).copy()
    elif ind.ndim != 2:
        raise ValueError("ind_x1 must be rank 1 or 2 if tensor-like.")
    return ind, np.ones(ind.shape, np.float32)


def dense_mask_from_indices(ind, mask, x_dim):
    """(n,x_dim) 0/1 weights equivalent to the gather formulation (:689-700): each
    padded slot contributes `mask` times the term of column `ind` (duplicates add)."""
    n = ind.shape[0]
    w = np.zeros((n, x_dim), np.float32)
    np.add.at(w, (np.repeat(np.arange(n), ind.shape[1]), ind.ravel()), mask.ravel())
    return w


def log_posterior(p, data_z, data_x, ind_x1=None, obs_mask=None):
    """bgm/base.py:665-705 (g_net in inference mode, :679)."""
    z = np.asarray(data_z, np.float32)
    x = np.asarray(data_x, np.float32)
    mu, s2 = variational_forward(p, z)
    if ind_x1 is None:                                                    # :682-685
        loss = ((x - mu) ** 2 / (2 * s2) + 0.5 * np.log(s2)).sum(axis=1)
    else:                                                                 # :689-700
        xc = np.take_along_axis(x, ind_x1, axis=1)
        mc = np.take_along_axis(mu, ind_x1, axis=1)
        sc = np.take_along_axis(s2, ind_x1, axis=1)
        ll = (xc - mc) ** 2 / (2 * sc) + 0.5 * np.log(sc)
        if obs_mask is not None:
            ll = ll * obs_mask
        loss = ll.sum(axis=1)
    prior = (z ** 2).sum(axis=1) / 2                                      # :702
    return (-(prior + loss)).astype(np.float32)


def log_posterior_and_grad(p, data_z, data_x, w=None, eps=1e-6):
    """log posterior and its gradient w.r.t. z (what TFP obtains by autodiff through
    :665-705), written out by hand; `w` is the dense (n,x_dim) observation weight
    (None = all observed).  Checked against torch autograd in tests/test_oracle.py."""
    f32 = np.float32
    z = np.asarray(data_z, f32)
    x = np.asarray(data_x, f32)
    h = bn_inference(p['bn'], z)
    pres, acts = [], [h]
    for W, b in p['hidden']:
        a = (h @ W + b).astype(f32)
        pres.append(a)
        h = leaky_relu(a)
        acts.append(h)
    mu = (h @ p['mean'][0] + p['mean'][1]).astype(f32)
    raw = (h @ p['var'][0] + p['var'][1]).astype(f32)
    sp = (np.maximum(raw, 0) + np.log1p(np.exp(-np.abs(raw)))).astype(f32)
    s2 = sp + f32(eps)
    d = x - mu
    term = d * d / (2 * s2) + 0.5 * np.log(s2)
    if w is not None:
        term = term * w
    lp = -(term.sum(axis=1) + (z ** 2).sum(axis=1) / 2)
    # backward of loss = sum(term) (+ prior)
    ww = f32(1.0) if w is None else w
    g_mu = (-d / s2) * ww
    sig = (1.0 / (1.0 + np.exp(-raw))).astype(f32)
    g_raw = ((-(d * d) / (2 * s2 * s2) + 0.5 / s2) * sig) * ww
    gh = (g_mu @ p['mean'][0].T + g_raw @ p['var'][0].T).astype(f32)
    for (W, b), a in zip(reversed(p['hidden']), reversed(pres)):
        ga = gh * np.where(a > 0, f32(1.0), LEAKY_SLOPE)
        gh = (ga @ W.T).astype(f32)
    inv = (p['bn']['gamma'] / np.sqrt(p['bn']['var'] + BN_EPS)).astype(f32)
    g_z_loss = gh * inv + z
    return lp.astype(f32), (-g_z_loss).astype(f32)


def hmc_sampler(p, data, w=None, z0=None, n_mcmc=3000, burn_in=5000, step_size=0.01,
                num_leapfrog_steps=10, momentum=None, log_u=None, rs=None,
                adaptation_rate=0.01, target_accept=0.75, return_trace=False):
    """bgm/base.py:778-821 + TFP 0.18 semantics (module docstring).

    Injected-noise mode: `z0` (n,zd), `momentum` (T,n,zd) N(0,1) draws and `log_u`
    (T,n) = log(uniform) draws, T = burn_in + n_mcmc.  Otherwise draws come from
    `rs` (a RandomState) -- TF's own stream is not reproducible outside TF.
    All rows share ONE scalar step size (bgm/base.py:805-809).
    """
    f32 = np.float32
    x = np.asarray(data, f32)
    n = x.shape[0]
    zd = p['hidden'][0][0].shape[0]
    rs = rs if rs is not None else np.random.RandomState(0)
    z = (z0 if z0 is not None else rs.standard_normal((n, zd))).astype(f32).copy()
    n_adapt = int(burn_in * 0.8)                                          # :807
    eps = f32(step_size)
    lp, g = log_posterior_and_grad(p, z, x, w)
    samples = []
    trace = dict(accept=[], step=[], log_accept=[])
    T = burn_in + n_mcmc
    for t in range(T):
        mom = (momentum[t] if momentum is not None else rs.standard_normal((n, zd))).astype(f32)
        zt, pt, gt = z.copy(), mom.copy(), g
        pt = pt + f32(0.5) * eps * gt
        lpt = lp
        for _ in range(num_leapfrog_steps):
            zt = (zt + eps * pt).astype(f32)
            lpt, gt = log_posterior_and_grad(p, zt, x, w)
            pt = (pt + eps * gt).astype(f32)
        pt = (pt - f32(0.5) * eps * gt).astype(f32)
        ke0 = f32(0.5) * (mom * mom).sum(axis=1)
        ke1 = f32(0.5) * (pt * pt).sum(axis=1)
        log_accept = (lpt - lp + (ke0 - ke1)).astype(f32)
        log_accept = np.where(np.isfinite(log_accept), log_accept, f32(-np.inf))
        lu = (log_u[t] if log_u is not None else np.log(rs.uniform(size=n))).astype(f32)
        acc = lu < log_accept
        z = np.where(acc[:, None], zt, z)
        lp = np.where(acc, lpt, lp)
        g = np.where(acc[:, None], gt, g)
        if return_trace:
            trace['accept'].append(acc.copy())
            trace['step'].append(float(eps))
            trace['log_accept'].append(log_accept.copy())
        if t < n_adapt:
            la = np.minimum(log_accept, 0).astype(np.float64)
            m = la.max()
            log_mean = m + np.log(np.mean(np.exp(la - m))) if np.isfinite(m) else -np.inf
            if log_mean > np.log(target_accept):
                eps = f32(eps * f32(1.0 + adaptation_rate))
            else:
                eps = f32(eps / f32(1.0 + adaptation_rate))
        if t >= burn_in:
            samples.append(z.copy())
    out = np.array(samples)
    if return_trace:
        trace['step_final'] = float(eps)
        return out, trace
    return out


def predict_on_posteriors(p, zs, noise):
    """bgm/base.py:511-525 with the reparameterisation draws passed in
    (`noise` (n_mcmc, n, x_dim))."""
    n_mcmc, n, zd = zs.shape
    mu, s2 = variational_forward(p, zs.reshape(-1, zd))
    xd = mu.shape[1]
    return (noise.reshape(-1, xd) * np.sqrt(s2) + mu).reshape(n_mcmc, n, xd).astype(np.float32)


def impute_from_samples(data_np, pred_all, alpha):
    """bgm/base.py:618-663: intervals on missing dims + posterior-mean imputation."""
    miss = np.isnan(data_np)
    obs_mask = 1.0 - miss.astype(np.float32)
    data_obs = np.nan_to_num(data_np, nan=0.0)
    n = data_np.shape[0]
    same = np.all(miss == miss[0])
    if same:
        idx = np.where(miss[0])[0]
        if idx.size == 0:
            interval = np.zeros((n, 0, 2), np.float32)
        else:
            dim = pred_all[:, :, idx]
            interval = np.stack([np.quantile(dim, alpha / 2.0, axis=0),
                                 np.quantile(dim, 1.0 - alpha / 2.0, axis=0)], axis=-1)
    else:
        interval = []
        for i in range(n):
            idx = np.where(miss[i])[0]
            if idx.size == 0:
                interval.append(np.zeros((0, 2), np.float32))
                continue
            dim = pred_all[:, i, idx]
            interval.append(np.stack([np.quantile(dim, alpha / 2.0, axis=0),
                                      np.quantile(dim, 1.0 - alpha / 2.0, axis=0)], axis=-1))
    imputed = np.mean(pred_all, axis=0)
    imputed = miss.astype(np.float32) * imputed + obs_mask * data_obs
    return imputed, interval
