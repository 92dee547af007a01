"""GPU parity tests of the BGM / HMC hot path: CUDA (through the C ABI, via the
reference-shaped Python methods) vs the CPU oracle on the same seeded inputs.

Tolerances (fp32):
  * log-posterior |d| <= 1e-4 * max(1,|logp|); gradient |d| <= 2e-4 * max(1, max|grad| of the row);
  * injected-noise HMC: shared step-size history identical, accept masks identical except at
    rounding-level ties (|log u - log_accept| < 1e-3), kept states of agreeing chains within
    2e-3 absolute (10-leapfrog trajectories amplify the 1e-6 relative gradient differences);
  * posterior-predictive draws with injected noise: rtol 1e-4.
"""
import numpy as np
import pytest

from oracle import bgm
from helpers import bgm_params, bgm_oracle_net, bgm_product_model, hmc_noise, philox_normal4

pytestmark = pytest.mark.gpu

CASES = [
    # n, x_dim, z_dim, g_units, missing fraction
    (100, 10, 3, (64,) * 5, 0.0),        # cfg-1 shape
    (77, 70, 10, (64,) * 5, 0.3),        # MCAR mask, ragged tile
    (40, 500, 10, (64,) * 5, 0.3),       # cfg-5 shape (16 head tiles)
    (33, 33, 4, (16, 16), 0.5),          # narrow nets, x_dim not a multiple of 4
    (1, 8, 1, (8,), 0.0),                # single row, one hidden layer
    (64, 20, 16, (64, 32, 48), 0.2),     # largest z_dim, mixed widths
]


def make_case(n, x_dim, z_dim, units, miss, seed=5):
    params = bgm_params(x_dim, z_dim, units)
    p = bgm_oracle_net(params)
    rs = np.random.RandomState(seed)
    x = rs.standard_normal((n, x_dim)).astype(np.float32)
    w = (rs.uniform(size=(n, x_dim)) >= miss).astype(np.float32)
    w[:, 0] = 1.0
    xn = x.copy()
    xn[w == 0] = np.nan
    return params, p, x, w, xn


def engine_model(params, p, engine):
    """engine: 'simt' | 'tensor' (skips when the net shape has no tensor engine) -- BGM.set_hmc_engine."""
    m = bgm_product_model(params, p)
    if engine == 'tensor' and not m.hmc_engine_info()['tensor_available']:
        pytest.skip("no tensor engine for this net shape (needs >= 2 hidden layers, all 64 wide)")
    m.set_hmc_engine(engine)
    assert m.hmc_engine_info()['engine'] == engine
    return m


@pytest.mark.parametrize("engine", ["simt", "tensor"])
@pytest.mark.parametrize("n,x_dim,z_dim,units,miss", CASES + [(300, 97, 7, (64, 64), 0.5), (129, 32, 12, (64,) * 6, 0.1), (70, 33, 16, (64, 64, 64), 0.2),
                                                          (5, 3, 1, (64, 64), 0.0)])
def test_log_posterior_and_gradient_parity(n, x_dim, z_dim, units, miss, engine):
    params, p, x, w, xn = make_case(n, x_dim, z_dim, units, miss)
    z = np.random.RandomState(1).standard_normal((n, z_dim)).astype(np.float32)
    want_lp, want_g = bgm.log_posterior_and_grad(p, z, x, w)
    m = engine_model(params, p, engine)
    lp, g = m.get_log_posterior(z, xn, return_grad=True)
    assert lp.shape == (n,) and g.shape == (n, z_dim)
    err = np.abs(lp - want_lp) / np.maximum(1, np.abs(want_lp))
    assert err.max() <= 1e-4, err.max()
    scale = np.maximum(1, np.abs(want_g).max(axis=1, keepdims=True))
    assert (np.abs(g - want_g) / scale).max() <= 2e-4
    # the reference's padded-gather form (ind_x1, obs_mask) gives the same numbers
    lists = [np.where(w[i] > 0)[0].tolist() for i in range(n)]
    ind, mask = bgm.pad_index_lists(lists, n)
    lp2 = m.get_log_posterior(z, x, ind, mask)
    np.testing.assert_array_equal(lp2, lp)


def compare_hmc(sg, trg, so, tro, log_u, burn_in):
    acc_o = np.array(tro['accept'])
    T, n = acc_o.shape
    np.testing.assert_allclose(trg['step'], np.array(tro['step'], np.float32), rtol=1e-6)
    mism = trg['accept'] != acc_o
    first = np.where(mism.any(axis=0), mism.argmax(axis=0), T)
    clean = first == T
    assert (~clean).mean() <= 0.02 + 1.0 / n
    la_o = np.array(tro['log_accept'])
    for r in np.where(~clean)[0]:
        t = first[r]
        assert abs(log_u[t, r] - la_o[t, r]) < 1e-3, "chain %d diverged at t=%d without a tie" % (r, t)
    np.testing.assert_allclose(sg[:, clean], so[:, clean], rtol=0, atol=2e-3)
    fin = np.isfinite(la_o) & (np.arange(T)[:, None] <= first[None, :])
    assert np.abs(trg['log_accept'] - la_o)[fin].max() < 2e-2
    return clean.mean()


@pytest.mark.parametrize("engine", ["simt", "tensor"])
@pytest.mark.parametrize("n,x_dim,z_dim,units,miss", CASES[:4] + [(300, 97, 7, (64, 64), 0.5)])
def test_hmc_injected_noise_matches_oracle(n, x_dim, z_dim, units, miss, engine):
    params, p, x, w, xn = make_case(n, x_dim, z_dim, units, miss)
    burn_in, n_mcmc, L, step = 10, 6, 5, 0.02
    nz = hmc_noise(n, z_dim, burn_in + n_mcmc)
    so, tro = bgm.hmc_sampler(p, x, w, z0=nz['z0'], n_mcmc=n_mcmc, burn_in=burn_in, step_size=step,
                              num_leapfrog_steps=L, momentum=nz['momentum'], log_u=nz['log_u'],
                              return_trace=True)
    m = engine_model(params, p, engine)
    sg, trg = m.tfp_mcmc_sampler(xn, n_mcmc=n_mcmc, burn_in=burn_in, step_size=step, num_leapfrog_steps=L,
                                 noise=nz, return_trace=True, verbose=0)
    assert sg.shape == so.shape == (n_mcmc, n, z_dim) and sg.dtype == np.float32
    compare_hmc(sg, trg, so, tro, nz['log_u'], burn_in)
    assert len(set(tro['step'])) > 2                           # the shared step size adapted
    np.testing.assert_array_equal(trg['accept_count'], trg['accept'].sum(axis=1))
    np.testing.assert_array_equal(trg['z_final'], sg[-1])
    # cached log p / gradient of the final state are those of the final state
    lp, g = m.get_log_posterior(trg['z_final'], xn, return_grad=True)
    np.testing.assert_allclose(lp, trg['lp_final'], rtol=1e-6, atol=1e-5)
    np.testing.assert_allclose(g, trg['g_final'], rtol=1e-5, atol=1e-5)


def test_hmc_index_list_form_equals_nan_form():
    params, p, x, w, xn = make_case(50, 12, 3, (32, 32), 0.4)
    m = bgm_product_model(params, p)
    lists = [np.where(w[i] > 0)[0].tolist() for i in range(50)]
    a = m.tfp_mcmc_sampler(x, ind_x1=lists, n_mcmc=4, burn_in=5, seed=3, verbose=0)
    b = m.tfp_mcmc_sampler(xn, n_mcmc=4, burn_in=5, seed=3, verbose=0)
    np.testing.assert_array_equal(a, b)
    shared = [0, 3, 7]
    c = m.tfp_mcmc_sampler(x, ind_x1=shared, n_mcmc=4, burn_in=5, seed=3, verbose=0)
    x2 = np.full_like(x, np.nan)
    x2[:, shared] = x[:, shared]
    d = m.tfp_mcmc_sampler(x2, n_mcmc=4, burn_in=5, seed=3, verbose=0)
    np.testing.assert_array_equal(c, d)


def test_duplicate_indices_in_ind_x1_add_like_the_reference_gather():
    """bgm/base.py:689-700 gathers (x, mu, sigma^2) at ind_x1 and sums: a feature listed m times counts m times."""
    params, p, x, w, xn = make_case(41, 12, 3, (32, 32), 0.0)
    m = bgm_product_model(params, p)
    rs = np.random.RandomState(4)
    z = rs.standard_normal((41, 3)).astype(np.float32)
    lists = [rs.randint(0, 12, size=rs.randint(1, 9)).tolist() for _ in range(41)]     # with repeats, ragged
    assert any(len(set(r)) < len(r) for r in lists)
    ind, mask = bgm.pad_index_lists(lists, 41)
    want = bgm.log_posterior(p, z, x, ind, mask)
    got, g = m.get_log_posterior(z, x, ind, mask, return_grad=True)
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-3)
    # weights = multiplicities in the oracle's dense gradient form
    cnt = np.zeros((41, 12), np.float32)
    for i, r in enumerate(lists):
        np.add.at(cnt[i], r, 1)
    wl, wg = bgm.log_posterior_and_grad(p, z, x, cnt)
    np.testing.assert_allclose(got, wl, rtol=1e-4, atol=1e-3)
    np.testing.assert_allclose(g, wg, rtol=2e-4, atol=2e-4 * max(1.0, np.abs(wg).max()))
    # the sampler accepts the same lists; a fixed (K,) list with a repeat too
    a = m.tfp_mcmc_sampler(x, ind_x1=lists, n_mcmc=3, burn_in=4, seed=3, verbose=0)
    assert a.shape == (3, 41, 3) and np.isfinite(a).all()
    b = m.tfp_mcmc_sampler(x, ind_x1=[0, 3, 3, 7], n_mcmc=3, burn_in=4, seed=3, verbose=0)
    c = m.tfp_mcmc_sampler(x, ind_x1=[0, 3, 7], n_mcmc=3, burn_in=4, seed=3, verbose=0)
    assert np.isfinite(b).all() and not np.array_equal(b, c)


@pytest.mark.parametrize("engine", ["simt", "tensor"])
def test_hmc_philox_run_replayed_through_oracle(engine):
    params, p, x, w, xn = make_case(200, 10, 3, (64,) * 5, 0.0)
    m = engine_model(params, p, engine)
    burn_in, n_mcmc, L, seed = 10, 10, 10, 42
    sg, trg = m.tfp_mcmc_sampler(xn, n_mcmc=n_mcmc, burn_in=burn_in, step_size=0.01, num_leapfrog_steps=L,
                                 seed=seed, return_trace=True, verbose=0)
    nz = m.philox_noise(seed, 200, burn_in + n_mcmc)
    so, tro = bgm.hmc_sampler(p, x, None, z0=nz['z0'], n_mcmc=n_mcmc, burn_in=burn_in, step_size=0.01,
                              num_leapfrog_steps=L, momentum=nz['momentum'], log_u=nz['log_u'],
                              return_trace=True)
    compare_hmc(sg, trg, so, tro, nz['log_u'], burn_in)
    e = nz['momentum'].ravel()
    assert abs(e.mean()) < 0.03 and abs(e.std() - 1) < 0.03
    u = np.exp(nz['log_u'])
    assert abs(u.mean() - 0.5) < 0.03
    # the dumped stream is the documented Philox construction (kind 3 = momentum)
    rows = np.arange(200)
    want = philox_normal4(seed, rows, 4, 3, 0)[:, :3]
    np.testing.assert_allclose(nz['momentum'][4], want, rtol=2e-5, atol=2e-6)


def test_predict_on_posteriors_parity():
    params, p, x, w, xn = make_case(30, 45, 5, (64, 64), 0.0)
    rs = np.random.RandomState(8)
    zs = rs.standard_normal((7, 30, 5)).astype(np.float32)
    noise = rs.standard_normal((7, 30, 45)).astype(np.float32)
    want = bgm.predict_on_posteriors(p, zs, noise)
    m = bgm_product_model(params, p)
    got = m.predict_on_posteriors(zs, noise=noise)
    assert got.shape == (7, 30, 45)
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-4)
    # Philox draws: keyed by (row, sample, column), N(0,1) through the same heads
    mu = m.predict_on_posteriors(zs, noise=np.zeros_like(noise))
    sd = m.predict_on_posteriors(zs, noise=np.ones_like(noise)) - mu
    d = (m.predict_on_posteriors(zs, seed=77) - mu) / sd
    assert abs(d.mean()) < 0.05 and abs(d.std() - 1) < 0.05
    want_e = philox_normal4(77, np.arange(30), 2, 4, 3)          # sample 2, columns 12..15
    np.testing.assert_allclose(d[2, :, 12:16], want_e, rtol=1e-3, atol=2e-3)


def test_predict_imputation_shapes_and_semantics():
    params, p, x, w, xn = make_case(60, 12, 3, (32, 32), 0.3)
    m = bgm_product_model(params, p)
    imp, interval = m.predict(xn, alpha=0.1, bs=25, n_mcmc=40, burn_in=10, seed=1, verbose=0)
    miss = np.isnan(xn)
    assert imp.shape == xn.shape and np.isfinite(imp).all()
    np.testing.assert_array_equal(imp[~miss], xn[~miss])            # observed entries intact (:662)
    assert isinstance(interval, list) and len(interval) == 60        # ragged pattern -> per-row list (:637)
    for i in range(60):
        k = int(miss[i].sum())
        assert interval[i].shape == (k, 2)
        if k:
            assert (interval[i][:, 0] <= interval[i][:, 1]).all()
    # shared missing pattern -> (n, n_missing, 2) array (:625-635); samples on request
    x2 = x.copy()
    x2[:, [2, 5]] = np.nan
    smp, iv = m.predict(x2, alpha=0.1, return_samples=True, bs=32, n_mcmc=30, burn_in=10, seed=1, verbose=0)
    assert smp.shape == (30, 60, 12) and iv.shape == (60, 2, 2)
    lo = np.quantile(smp[:, :, [2, 5]], 0.05, axis=0)
    np.testing.assert_allclose(iv[..., 0], lo, rtol=1e-4, atol=1e-4)
    # no missing values at all
    _, iv0 = m.predict(x, n_mcmc=5, burn_in=5, verbose=0)
    assert iv0.shape == (60, 0, 2)
    with pytest.raises(AssertionError):
        m.predict(x, alpha=0.0, verbose=0)


@pytest.mark.parametrize("S,alpha", [(1, 0.5), (2, 0.1), (25, 0.2), (100, 0.05), (100, 0.1), (200, 0.1), (333, 0.05), (7, 1.0)])
def test_column_quantiles_kernel_against_numpy(S, alpha):
    """bgm_column_quantiles (mean and np.quantile at alpha/2, 1 - alpha/2 of every column in one pass)."""
    import torch
    from bayesgm_b200 import _lib
    rs = np.random.RandomState(S)
    M = 1000 + S
    draws = rs.standard_normal((S, M)).astype(np.float32)
    draws[:, :7] = np.round(draws[:, :7])                     # ties
    d = torch.from_numpy(draws).cuda()
    mean = torch.empty(M, dtype=torch.float32, device='cuda')
    lo, hi = torch.empty_like(mean), torch.empty_like(mean)
    _lib.call("bgm_column_quantiles", _lib.ptr(d), S, M, alpha / 2, 1 - alpha / 2, _lib.ptr(mean), _lib.ptr(lo), _lib.ptr(hi),
              _lib.stream_ptr())
    np.testing.assert_allclose(mean.cpu().numpy(), draws.mean(axis=0), rtol=1e-5, atol=1e-6)
    d64 = draws.astype(np.float64)     # (np.quantile on float32 input interpolates in float32 and is off by up to ~5e-6)
    np.testing.assert_allclose(lo.cpu().numpy(), np.quantile(d64, alpha / 2, axis=0), rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(hi.cpu().numpy(), np.quantile(d64, 1 - alpha / 2, axis=0), rtol=1e-6, atol=1e-6)
    if S == 333:      # 16 + order statistics from an end: the entry point refuses, BGM.predict sorts instead
        with pytest.raises(_lib.BgmError):
            _lib.call("bgm_column_quantiles", _lib.ptr(d), S, M, 0.2, 0.8, _lib.ptr(mean), _lib.ptr(lo), _lib.ptr(hi), _lib.stream_ptr())


def test_imputation_deep_quantiles_fall_back_to_the_sort():
    params, p, x, w, xn = make_case(30, 9, 2, (16,), 0.3)
    m = bgm_product_model(params, p)
    smp, iv = m.predict(xn, alpha=0.5, return_samples=True, bs=16, n_mcmc=120, burn_in=8, seed=4, verbose=0)
    want_imp, want_iv = bgm.impute_from_samples(xn, smp, 0.5)
    for a, b in zip(iv, want_iv):
        np.testing.assert_allclose(a, b, rtol=1e-4, atol=1e-4)


def test_imputation_matches_oracle_reduction():
    """predict()'s device-side reductions equal the oracle's host reductions on the same draws."""
    params, p, x, w, xn = make_case(40, 9, 2, (16,), 0.3)
    m = bgm_product_model(params, p)
    smp, iv = m.predict(xn, alpha=0.2, return_samples=True, bs=16, n_mcmc=25, burn_in=8, seed=4, verbose=0)
    imp, iv2 = m.predict(xn, alpha=0.2, return_samples=False, bs=16, n_mcmc=25, burn_in=8, seed=4, verbose=0)
    want_imp, want_iv = bgm.impute_from_samples(xn, smp, 0.2)
    np.testing.assert_allclose(imp, want_imp, rtol=1e-5, atol=1e-5)
    for a, b in zip(iv2, want_iv):
        np.testing.assert_allclose(a, b, rtol=1e-4, atol=1e-4)


def test_golden_bgm():
    import os
    from test_golden import load, bgm_setup
    for name in ("bgm_x10_z3", "bgm_x70_z10_mcar"):
        g = load(name)
        p, mom, log_u = bgm_setup(g)
        params = bgm_params(int(g['x_dim']), int(g['z_dim']), [int(u) for u in g['units']])
        m = bgm_product_model(params, p)
        xn = g['x'].copy()
        xn[g['w'] == 0] = np.nan
        lp, gr = m.get_log_posterior(g['z0'], xn, return_grad=True)
        assert (np.abs(lp - g['logp']) / np.maximum(1, np.abs(g['logp']))).max() <= 1e-4
        np.testing.assert_allclose(gr, g['grad'], rtol=1e-3, atol=2e-4 * max(1, np.abs(g['grad']).max()))
        s, tr = m.tfp_mcmc_sampler(xn, n_mcmc=int(g['n_mcmc']), burn_in=int(g['burn_in']), step_size=float(g['step']),
                                   num_leapfrog_steps=int(g['L']), noise=dict(z0=g['z0'], momentum=mom, log_u=log_u),
                                   return_trace=True, verbose=0)
        np.testing.assert_allclose(tr['step'], g['steps'], rtol=1e-6)
        clean = (tr['accept'] == g['accept']).all(axis=0)
        assert clean.mean() >= 0.95
        np.testing.assert_allclose(s[:, clean], g['samples'][:, clean], rtol=0, atol=2e-3)


def test_full_size_properties():
    """cfg-5 per-GPU shape at reduced row count (n=37888 = one full wave, x_dim=500, z_dim=10, 30% MCAR)."""
    n, xd, zd = 37888, 500, 10
    params = bgm_params(xd, zd)
    p = bgm_oracle_net(params, bn_random=False)
    rs = np.random.RandomState(0)
    x = rs.standard_normal((n, xd)).astype(np.float32)
    x[rs.uniform(size=(n, xd)) < 0.3] = np.nan
    m = bgm_product_model(params, p)
    s1, tr = m.tfp_mcmc_sampler(x, n_mcmc=2, burn_in=3, step_size=0.01, num_leapfrog_steps=3, seed=11,
                                return_trace=True, verbose=0)
    assert s1.shape == (2, n, zd) and np.isfinite(s1).all()
    s2 = m.tfp_mcmc_sampler(x, n_mcmc=2, burn_in=3, step_size=0.01, num_leapfrog_steps=3, seed=11, verbose=0)
    np.testing.assert_array_equal(s1, s2)                        # deterministic
    np.testing.assert_array_equal(tr['accept_count'], tr['accept'].sum(axis=1))
    changed = (s1[1:] != s1[:-1]).any(axis=2)
    np.testing.assert_array_equal(changed, tr['accept'][4:])
    idx = rs.choice(n, 128, replace=False)
    xo = np.nan_to_num(x[idx], nan=0.0)
    wo = (~np.isnan(x[idx])).astype(np.float32)
    want_lp, want_g = bgm.log_posterior_and_grad(p, tr['z_final'][idx], xo, wo)
    err = np.abs(tr['lp_final'][idx] - want_lp) / np.maximum(1, np.abs(want_lp))
    assert err.max() <= 1e-4
    assert np.abs(tr['g_final'][idx] - want_g).max() <= 2e-4 * max(1, np.abs(want_g).max())
