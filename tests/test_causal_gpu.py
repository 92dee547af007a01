"""GPU parity tests of the CausalBGM hot path: CUDA (through the C ABI, via the
reference-shaped Python methods) vs the CPU oracle on the same seeded inputs.

Tolerances (fp32):
  * log-posterior: |d| <= 1e-4 * max(1, |logp|) per row;
  * injected-noise MH: accept masks identical and states BIT-identical for every
    chain up to its first mask mismatch; mismatching chains <= 1% over the run and
    each first mismatch explained by |u - ratio| < 1e-3 (a rounding-level tie);
  * effects with sample_y=False: rtol 1e-4.
"""
import numpy as np
import pytest

from oracle import causal
from helpers import causal_params, causal_nets, causal_data, product_model, injected_noise

pytestmark = pytest.mark.gpu

LP_TOL = 1e-4


def lp_close(a, b):
    err = np.abs(a - b) / np.maximum(1.0, np.abs(b))
    assert err.max() <= LP_TOL, "log-posterior mismatch: max rel err %.3g" % err.max()


CASES = [
    # n, v_dim, z_dims, binary, extra
    (100, 200, [1, 1, 1, 2], False, {}),            # cfg-3 shape (zd=5)
    (97, 200, [1, 1, 1, 7], False, {}),             # shipped Sim_Hirano_Imbens z_dims, ragged tile
    (64, 100, [3, 6, 3, 6], True, {}),              # cfg-2 shape, binary treatment
    (33, 177, [3, 6, 3, 6], True, {}),              # Semi_acic v_dim (not a multiple of 4)
    (1, 10, [1, 1, 1, 1], False, {}),               # single row
    (31, 10, [1, 1, 1, 0], False, dict(g_units=[8, 8], f_units=[8, 8], h_units=[8, 8])),  # R-test nets
    (40, 50, [2, 1, 1, 1], False, dict(g_units=[40, 10, 64], f_units=[20], h_units=[33, 5])),
    (40, 30, [1, 1, 1, 2], False, dict(sigma_v=0.8, sigma_x=1.1, sigma_y=0.9)),  # fixed-variance keys
]


# every case on the fp32 FMA-pipe engine; the standard net shapes (CASES[:4]) also on the
# tensor-core engine (3xTF32) -- same tolerances for both
ENGINE_CASES = ([c + ('simt',) for c in CASES] + [c + ('tensor',) for c in CASES[:4]] +
                [(1, 200, [1, 1, 1, 2], False, {}, 'tensor'), (129, 200, [2, 3, 2, 5], False, {}, 'tensor'),
                 (257, 100, [4, 4, 4, 4], True, dict(g_units=[64, 64]), 'tensor')])   # single row; zd=12; zd=16, 2-layer g


@pytest.mark.parametrize("n,v_dim,z_dims,binary,extra,engine", ENGINE_CASES)
def test_log_posterior_parity(n, v_dim, z_dims, binary, extra, engine):
    params = causal_params(v_dim, z_dims, binary, **extra)
    nets = causal_nets(params)
    x, y, v = causal_data(n, v_dim, binary)
    z = np.random.RandomState(11).standard_normal((n, sum(z_dims))).astype(np.float32)
    want = causal.log_posterior(params, nets, x, y, v, z)
    m = product_model(params, nets, engine)
    got = m.get_log_posterior(x, y, v, z)
    assert m.sampler_info()['engine'] == engine
    assert got.shape == (n,) and got.dtype == np.float32
    lp_close(got, want)


def test_auto_engine_selection():
    std = product_model(causal_params(200, [1, 1, 1, 2]), causal_nets(causal_params(200, [1, 1, 1, 2])))
    assert std.sampler_info()['engine'] == 'tensor' and std.sampler_info()['tensor_available']
    p = causal_params(10, [1, 1, 1, 0], g_units=[8, 8], f_units=[8, 8], h_units=[8, 8])
    small = product_model(p, causal_nets(p))
    assert small.sampler_info()['engine'] == 'simt' and not small.sampler_info()['tensor_available']
    with pytest.raises(Exception):
        small.set_sampler_engine('tensor')


def compare_chains(samples_g, tr_g, samples_o, tr_o, u, burn_in):
    """State-for-state comparison of a GPU run and an oracle run on the same noise."""
    acc_g, acc_o = tr_g['accept'], np.array(tr_o['accept'])
    T, n = acc_o.shape
    mism = acc_g != acc_o
    first = np.where(mism.any(axis=0), mism.argmax(axis=0), T)
    clean = first == T
    # chains whose decisions all agree are bit-identical
    np.testing.assert_array_equal(samples_g[:, clean], samples_o[:, clean])
    # proposed log-posteriors agree up to each chain's first mismatch
    lp_o = np.array(tr_o['lp_prop'])
    tt = np.arange(T)[:, None]
    before = tt <= first[None, :]
    err = np.abs(tr_g['lp_prop'] - lp_o) / np.maximum(1.0, np.abs(lp_o))
    assert err[before].max() <= LP_TOL
    # mismatches are rare and are rounding-level ties
    assert (~clean).mean() <= 0.01 + 1.0 / n, "%.3f of chains diverged" % (~clean).mean()
    ratio_o = np.exp(np.minimum(lp_o - np.array(tr_o['lp_cur']), 0))
    for r in np.where(~clean)[0]:
        t = first[r]
        assert abs(u[t, r] - ratio_o[t, r]) < 1e-3, "chain %d diverged at t=%d without a tie" % (r, t)
        # states agree up to the mismatch
        k = t - burn_in
        if k > 0:
            np.testing.assert_array_equal(samples_g[:k, r], samples_o[:k, r])
    return clean.mean()


MH_CASES = ([c + ('simt',) for c in CASES[:4] + CASES[5:]] + [c + ('tensor',) for c in CASES[:4]] +
            [(300, 200, [1, 1, 1, 2], False, {}, 'tensor')])   # 3 row tiles of 128, ragged last one


@pytest.mark.parametrize("n,v_dim,z_dims,binary,extra,engine", MH_CASES)
def test_mh_injected_noise_state_for_state(n, v_dim, z_dims, binary, extra, engine):
    params = causal_params(v_dim, z_dims, binary, **extra)
    nets = causal_nets(params)
    data = causal_data(n, v_dim, binary)
    burn_in, n_keep = 15, 25
    nz = injected_noise(n, sum(z_dims), burn_in + n_keep)
    so, tro = causal.mh_sampler(params, nets, data, q_sd=0.3, burn_in=burn_in, n_keep=n_keep,
                                noise=causal.InjectedNoise(**nz), return_trace=True)
    m = product_model(params, nets, engine)
    sg, trg = m.metropolis_hastings_sampler(data, q_sd=0.3, burn_in=burn_in, n_keep=n_keep, noise=nz,
                                            return_trace=True, verbose=0)
    assert sg.shape == so.shape == (n_keep, n, sum(z_dims)) and sg.dtype == np.float32
    compare_chains(sg, trg, so, tro, nz['u'], burn_in)
    # bookkeeping: per-iteration counts match the masks, final state is the last sample
    np.testing.assert_array_equal(trg['accept_count'], trg['accept'].sum(axis=1))
    np.testing.assert_array_equal(trg['z_final'], sg[-1])
    assert 0.0 < trg['accept'].mean() < 1.0


@pytest.mark.parametrize("engine", ['simt', 'tensor'])
def test_mh_philox_run_replayed_through_oracle(engine):
    """The production noise path: sample with the in-kernel Philox stream, dump the
    very same stream, replay it through the oracle."""
    params = causal_params(200, [1, 1, 1, 2])
    nets = causal_nets(params)
    data = causal_data(200, 200)
    burn_in, n_keep, seed = 20, 20, 20261017
    m = product_model(params, nets, engine)
    sg, trg = m.metropolis_hastings_sampler(data, q_sd=0.5, burn_in=burn_in, n_keep=n_keep, seed=seed,
                                            return_trace=True, verbose=0)
    nz = m.philox_noise(seed, 200, burn_in + n_keep)
    so, tro = causal.mh_sampler(params, nets, data, q_sd=0.5, burn_in=burn_in, n_keep=n_keep,
                                noise=causal.InjectedNoise(**nz), return_trace=True)
    compare_chains(sg, trg, so, tro, nz['u'], burn_in)
    # and the stream itself is N(0,1) / U(0,1)
    e = nz['eps'].ravel()
    assert abs(e.mean()) < 0.02 and abs(e.std() - 1) < 0.02
    assert abs(np.mean(e ** 4) - 3) < 0.2
    assert abs(nz['u'].mean() - 0.5) < 0.02 and nz['u'].min() >= 0 and nz['u'].max() < 1


@pytest.mark.parametrize("v_dim,n,engine", [(20, 64, 'simt'), (200, 200, 'tensor'), (200, 200, 'simt')])
def test_mh_adaptive_q_sd_matches_oracle(v_dim, n, engine):
    """Adaptive proposal scale (:880-890): one launch per 50-iteration stretch with the chain state
    handed over through z_state / lp_state, the rule evaluated on the device in between."""
    params = causal_params(v_dim, [1, 1, 1, 2])
    nets = causal_nets(params)
    data = causal_data(n, v_dim)
    burn_in, n_keep = 230, 10
    nz = injected_noise(n, 5, burn_in + n_keep)
    so, tro = causal.mh_sampler(params, nets, data, q_sd=None, initial_q_sd=3.0, burn_in=burn_in,
                                n_keep=n_keep, noise=causal.InjectedNoise(**nz), return_trace=True)
    m = product_model(params, nets, engine)
    sg, trg = m.metropolis_hastings_sampler(data, q_sd=None, initial_q_sd=3.0, burn_in=burn_in,
                                            n_keep=n_keep, noise=nz, return_trace=True, verbose=0)
    assert len(set(tro['q_sd'])) >= 3                       # the rule fired
    frac_clean = (trg['accept'] == np.array(tro['accept'])).all(axis=0).mean()
    if frac_clean == 1.0:
        assert abs(trg['q_sd_final'] - tro['q_sd_final']) < 1e-6 * tro['q_sd_final']
        np.testing.assert_array_equal(sg, so)
    else:   # a rounding tie changed one chain: q_sd history still follows the same rule
        assert abs(trg['q_sd_final'] / tro['q_sd_final'] - 1) < 0.25


@pytest.mark.parametrize("binary", [False, True])
def test_effect_parity(binary):
    z_dims = [2, 2, 1, 3]
    params = causal_params(16, z_dims, binary)
    nets = causal_nets(params)
    n, n_keep = 77, 9
    rs = np.random.RandomState(2)
    zs = rs.standard_normal((n_keep, n, 8)).astype(np.float32)
    m = product_model(params, nets)
    xv = None if binary else [0.0, 0.7, 3.0]
    want = causal.infer_from_latent_posterior(params, nets, zs, xv, sample_y=False)
    got = m.infer_from_latent_posterior(zs, x_values=xv, sample_y=False)
    assert got.shape == want.shape
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-5)
    # sample_y=True with the N(0,1) draws injected on both sides
    k = 2 if binary else 3
    noise = rs.standard_normal((k, n_keep, n)).astype(np.float32)
    it = iter(noise)
    want = causal.infer_from_latent_posterior(params, nets, zs, xv, sample_y=True,
                                              normal_fn=lambda shape: next(it))
    got = m.infer_from_latent_posterior(zs, x_values=xv, sample_y=True, noise=noise)
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("binary,engine", [(False, 'tensor'), (True, 'tensor'), (False, 'simt'), (True, 'simt')])
def test_effect_parity_standard_nets(binary, engine):
    """The effect kernel on the standard net shape (tensor-core engine available): 3 row tiles of
    128 per kept state with a ragged last one, 1..5 doses (pipeline fill / drain), both engines."""
    z_dims = [3, 6, 3, 6] if binary else [1, 1, 1, 2]
    v_dim = 100 if binary else 200
    params = causal_params(v_dim, z_dims, binary)
    nets = causal_nets(params)
    n, n_keep = 300, 3
    rs = np.random.RandomState(4)
    zs = rs.standard_normal((n_keep, n, sum(z_dims))).astype(np.float32)
    m = product_model(params, nets, engine)
    assert m.sampler_info()['engine'] == engine
    for xv in ([None] if binary else [[0.3], [0.0, 1.5], [0.0, 0.7, 3.0, 1.1, 2.2]]):
        want = causal.infer_from_latent_posterior(params, nets, zs, xv, sample_y=False)
        got = m.infer_from_latent_posterior(zs, x_values=xv, sample_y=False)
        assert got.shape == want.shape
        np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-5)
        k = 2 if binary else len(xv)
        noise = rs.standard_normal((k, n_keep, n)).astype(np.float32)
        it = iter(noise)
        want = causal.infer_from_latent_posterior(params, nets, zs, xv, sample_y=True,
                                                  normal_fn=lambda shape: next(it))
        got = m.infer_from_latent_posterior(zs, x_values=xv, sample_y=True, noise=noise)
        np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-4)
    if not binary:   # the Philox draws are keyed by (row, kept state, dose): engines agree
        a = m.infer_from_latent_posterior(zs, x_values=[0.5, 2.0], sample_y=True, seed=11)
        m2 = product_model(params, nets, 'simt' if engine == 'tensor' else 'tensor')
        b = m2.infer_from_latent_posterior(zs, x_values=[0.5, 2.0], sample_y=True, seed=11)
        np.testing.assert_allclose(a, b, rtol=1e-4, atol=1e-5)


def test_effect_sample_y_philox_is_distributionally_right():
    params = causal_params(16, [1, 1, 1, 1])
    nets = causal_nets(params)
    n, n_keep = 4096, 4
    zs = np.random.RandomState(3).standard_normal((n_keep, n, 4)).astype(np.float32)
    m = product_model(params, nets)
    mean_only = m.infer_from_latent_posterior(zs, x_values=[1.0], sample_y=False)
    drawn = m.infer_from_latent_posterior(zs, x_values=[1.0], sample_y=True, seed=5)
    # sigma_y^2 = softplus(.)+1e-6 is O(1): the row-mean of n draws moves by ~ 1/sqrt(n)
    assert np.all(np.abs(drawn - mean_only) < 6.0 / np.sqrt(n))
    assert np.any(drawn != mean_only)


def test_predict_continuous_matches_oracle_on_replayed_noise():
    params = causal_params(40, [1, 1, 1, 2])
    nets = causal_nets(params)
    n = 150
    data = causal_data(n, 40)
    m = product_model(params, nets)
    seed, burn_in, n_mcmc, bs = 99, 10, 12, 64
    xv = np.linspace(0, 3, 5)
    adrf, interval = m.predict(data, alpha=0.1, n_mcmc=n_mcmc, burn_in=burn_in, x_values=xv, q_sd=0.4,
                               sample_y=False, bs=bs, seed=seed, verbose=0)
    assert adrf.shape == (5,) and interval.shape == (5, 2)

    def factory(start, end):
        return causal.InjectedNoise(**m.philox_noise(seed, end - start, burn_in + n_mcmc, row_offset=start))
    want, want_int = causal.predict(params, nets, data, alpha=0.1, n_mcmc=n_mcmc, burn_in=burn_in,
                                    x_values=xv, q_sd=0.4, sample_y=False, bs=bs, noise_factory=factory)
    np.testing.assert_allclose(adrf, want, rtol=2e-4, atol=2e-4)
    np.testing.assert_allclose(interval, want_int, rtol=2e-4, atol=2e-4)
    # bs only slices the work: Philox noise is keyed by the global row, so any bs gives the same answer
    adrf2, _ = m.predict(data, alpha=0.1, n_mcmc=n_mcmc, burn_in=burn_in, x_values=xv, q_sd=0.4,
                         sample_y=False, bs=n, seed=seed, verbose=0)
    np.testing.assert_allclose(adrf2, adrf, rtol=1e-6, atol=1e-6)


def test_predict_binary_matches_oracle_on_replayed_noise():
    params = causal_params(24, [2, 2, 2, 2], binary=True)
    nets = causal_nets(params)
    n = 70
    data = causal_data(n, 24, binary=True)
    m = product_model(params, nets)
    seed, burn_in, n_mcmc = 5, 8, 16
    ite, interval = m.predict(data, alpha=0.2, n_mcmc=n_mcmc, burn_in=burn_in, q_sd=0.4, sample_y=False,
                              bs=32, seed=seed, verbose=0)
    assert ite.shape == (n,) and interval.shape == (n, 2)

    def factory(start, end):
        return causal.InjectedNoise(**m.philox_noise(seed, end - start, burn_in + n_mcmc, row_offset=start))
    want, want_int = causal.predict(params, nets, data, alpha=0.2, n_mcmc=n_mcmc, burn_in=burn_in, q_sd=0.4,
                                    sample_y=False, bs=32, noise_factory=factory)
    np.testing.assert_allclose(ite, want, rtol=1e-3, atol=2e-4)
    np.testing.assert_allclose(interval, want_int, rtol=1e-3, atol=2e-4)


def test_api_errors_match_reference():
    params = causal_params(8, [1, 1, 1, 1])
    m = product_model(params, causal_nets(params))
    data = causal_data(10, 8)
    with pytest.raises(ValueError):       # causalbgm/base.py:610-611
        m.predict(data, x_values=None, verbose=0)
    with pytest.raises(AssertionError):   # :606
        m.predict(data, alpha=1.5, x_values=[1.0], verbose=0)


def test_full_size_properties():
    """BASELINE cfg-3 size (n=100k, p=200, zd=5), short T: properties that need no oracle."""
    n, p = 100000, 200
    params = causal_params(p, [1, 1, 1, 2])
    nets = causal_nets(params)
    rs = np.random.RandomState(0)
    v = rs.standard_normal((n, p)).astype(np.float32)
    x = rs.exponential(size=(n, 1)).astype(np.float32)
    y = (x + rs.standard_normal((n, 1))).astype(np.float32)
    m = product_model(params, nets)
    burn_in, n_keep, seed = 4, 6, 1234
    s1, tr1 = m.metropolis_hastings_sampler((x, y, v), q_sd=0.2, burn_in=burn_in, n_keep=n_keep, seed=seed,
                                            return_trace=True, verbose=0)
    assert s1.shape == (n_keep, n, 5) and np.isfinite(s1).all()
    # determinism
    s2 = m.metropolis_hastings_sampler((x, y, v), q_sd=0.2, burn_in=burn_in, n_keep=n_keep, seed=seed, verbose=0)
    np.testing.assert_array_equal(s1, s2)
    # counts vs masks; a kept state changes exactly where a proposal was accepted
    np.testing.assert_array_equal(tr1['accept_count'], tr1['accept'].sum(axis=1))
    changed = (s1[1:] != s1[:-1]).any(axis=2)
    np.testing.assert_array_equal(changed, tr1['accept'][burn_in + 1:])
    # the cached log-posterior of the final state is the log-posterior of the final state
    lp = m.get_log_posterior(x, y, v, tr1['z_final'])
    np.testing.assert_array_equal(lp, tr1['lp_final'])
    # spot-check 256 random rows of the last iteration against the oracle
    idx = rs.choice(n, 256, replace=False)
    want = causal.log_posterior(params, nets, x[idx], y[idx], v[idx], tr1['z_final'][idx])
    lp_close(lp[idx], want)


@pytest.mark.parametrize("binary,engine,n,n_keep", [(False, 'tensor', 333, 40), (True, 'tensor', 333, 40), (False, 'simt', 333, 40),
                                                     (False, 'tensor', 4321, 300), (True, 'tensor', 4100, 17)])
def test_memoised_effect_equals_direct_evaluation(binary, engine, n, n_keep):
    """Kept states with long runs of repeats (what a 25 % acceptance rate produces): evaluating f_net
    once per distinct state and combining gives exactly what evaluating it at every kept state gives."""
    z_dims = [3, 6, 3, 6] if binary else [1, 1, 1, 2]
    v_dim = 100 if binary else 200
    params = causal_params(v_dim, z_dims, binary)
    m = product_model(params, causal_nets(params), engine)
    rs = np.random.RandomState(12)
    zd = sum(z_dims)                                 # n >= 4096: the row-major combine kernel (segments of 256 kept states)
    zs = np.empty((n_keep, n, zd), np.float32)
    zs[0] = rs.standard_normal((n, zd))
    for s in range(1, n_keep):                       # each row moves with probability 0.25
        move = rs.uniform(size=n) < 0.25
        zs[s] = np.where(move[:, None], zs[s - 1] + 0.3 * rs.standard_normal((n, zd)).astype(np.float32), zs[s - 1])
    import torch
    zd_ = torch.from_numpy(zs).cuda()
    xv = None if binary else np.linspace(0, 3, 7)
    nz = rs.standard_normal((2 if binary else 7, n_keep, n)).astype(np.float32)
    for sample_y, noise in ((False, None), (True, None), (True, nz)):
        a = m._effect_device(zd_, n_keep, n, xv, sample_y, 5, 1000, noise=noise, memoise=True).cpu().numpy()
        frac = m.last_distinct_fraction
        b = m._effect_device(zd_, n_keep, n, xv, sample_y, 5, 1000, noise=noise, memoise=False).cpu().numpy()
        if binary:
            np.testing.assert_array_equal(a, b)      # per-subject values: bit-identical
        else:
            np.testing.assert_allclose(a, b, rtol=1e-12, atol=1e-9)   # float64 sums, atomic order only
    assert 0.2 < frac < 0.35


def test_first_layer_product_variant_is_selected_and_agrees_with_the_fma_variant(monkeypatch):
    """z_dim <= 6: the first layers of f_net / h_net run as one tensor-core product (causal_mh_tc16_kernel<8, true>,
    DESIGN.md 4.1b); z_dim 7..8 and BGM_TC16_NO_L1 keep them on the FMA pipe.  Same Philox streams, same accept
    rule: the two variants must produce the same chains up to rounding-level ties."""
    params = causal_params(200, [1, 1, 1, 2])
    nets = causal_nets(params)
    data = causal_data(300, 200)
    m1 = product_model(params, nets, 'tensor')
    assert m1.sampler_info()['kernel'] == 'causal_mh_tc16_kernel<8, true>'
    p8 = causal_params(200, [2, 2, 2, 2])
    assert product_model(p8, causal_nets(p8), 'tensor').sampler_info()['kernel'] == 'causal_mh_tc16_kernel<8, false>'
    s1, t1 = m1.metropolis_hastings_sampler(data, q_sd=0.5, burn_in=30, n_keep=30, seed=5, return_trace=True, verbose=0)
    monkeypatch.setenv("BGM_TC16_NO_L1", "1")        # read when the device model is built
    m0 = product_model(params, nets, 'tensor')
    assert m0.sampler_info()['kernel'] == 'causal_mh_tc16_kernel<8, false>'
    s0, t0 = m0.metropolis_hastings_sampler(data, q_sd=0.5, burn_in=30, n_keep=30, seed=5, return_trace=True, verbose=0)
    same = np.all(t1['accept'] == t0['accept'], axis=0)
    assert same.mean() >= 0.99, "chains that took a different decision: %d of %d" % ((~same).sum(), same.size)
    np.testing.assert_array_equal(s1[:, same], s0[:, same])
    lp_close(t1['lp_prop'][:, same], t0['lp_prop'][:, same])


def test_predict_asks_the_driver_for_free_memory_once(monkeypatch):
    """cudaMemGetInfo blocks for up to ~85 ms on a shared host (profiles/r02Y_predict_jitter_before.log): predict()
    may call it when the process first sizes its buffers, never per call."""
    import torch
    params = causal_params(200, [1, 1, 1, 2])
    nets = causal_nets(params)
    data = causal_data(500, 200)
    m = product_model(params, nets)
    kw = dict(alpha=0.01, n_mcmc=20, burn_in=20, x_values=np.linspace(0, 3, 5), q_sd=1.0, sample_y=True, bs=500, verbose=0)
    m.predict(data, **kw)
    calls = []
    real = torch.cuda.mem_get_info
    monkeypatch.setattr(torch.cuda, "mem_get_info", lambda *a, **k: (calls.append(1), real(*a, **k))[1])
    m.predict(data, **kw)
    m.predict(data, **kw)
    assert len(calls) == 0
