"""IdentifiableCausalBGM (conditional prior, causalbgm/identifiable.py:505-616) and
FullMCMCCausalBGM (per-iteration weight sample, causalbgm/fullmcmc.py:285-487) as flags on the sampler /
effect kernels, against the oracle on the same injected noise."""
import numpy as np
import pytest

from helpers import causal_params, causal_nets, causal_data, injected_noise
from oracle import causal as oc
from oracle import nets as onets

pytestmark = pytest.mark.gpu

FLAT = lambda layers: [a for W, b in layers for a in (W, b)]
FLATV = lambda layers: np.concatenate([np.concatenate([W.ravel(), b.ravel()]) for W, b in layers]).astype(np.float32)


def _ident(params, nets, prior_net):
    from bayesgm_b200 import IdentifiableCausalBGM
    m = IdentifiableCausalBGM(params=params)
    m.set_weights(g=FLAT(nets['g']), e=FLAT(nets['e']), f=FLAT(nets['f']), h=FLAT(nets['h']))
    m.prior_net.set_weights(FLAT(prior_net))
    return m


@pytest.mark.parametrize("v_dim,z_dims,binary,n_seg", [(200, [1, 1, 1, 2], False, 10), (30, [2, 2, 2, 4], True, 4),
                                                        (100, [3, 3, 6, 6], False, 7)])
def test_identifiable_log_posterior(v_dim, z_dims, binary, n_seg):
    params = causal_params(v_dim, z_dims, binary=binary, n_segments=n_seg)
    nets = causal_nets(params)
    zd = sum(z_dims)
    prior_net = onets.init_mlp(np.random.RandomState(3), [n_seg, 64, zd + 1], 0.3)
    n = 333
    x, y, v = causal_data(n, v_dim, binary=binary)
    rs = np.random.RandomState(1)
    z = rs.standard_normal((n, zd)).astype(np.float32)
    seg = rs.randint(0, n_seg, size=n)
    u = np.eye(n_seg, dtype=np.float32)[seg]
    m = _ident(params, nets, prior_net)
    got = m.get_log_posterior(x, y, v, z, u)
    want = oc.log_posterior(params, nets, x, y, v, z, prior=oc.conditional_prior(params, prior_net, seg))
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-3)
    # a prior net that outputs mu = 0, softplus(s) + 1e-6 = 1 is the N(0, I) prior of the base class
    flat0 = [np.zeros_like(a) for a in FLAT(prior_net)]
    flat0[-1][-1] = np.log(np.expm1(1.0 - 1e-6))
    m.prior_net.set_weights(flat0)
    base = oc.log_posterior(params, nets, x, y, v, z)
    np.testing.assert_allclose(m.get_log_posterior(x, y, v, z, u), base, rtol=1e-4, atol=1e-3)
    with pytest.raises(NotImplementedError):
        m.get_log_posterior(x, y, v, z, 0.5 * u)


def test_identifiable_sampler_state_for_state():
    params = causal_params(200, [1, 1, 1, 2], n_segments=5)
    nets = causal_nets(params)
    prior_net = onets.init_mlp(np.random.RandomState(4), [5, 64, 6], 0.5)
    n, T = 257, 60
    x, y, v = causal_data(n, 200)
    noise = injected_noise(n, 5, T)
    seg = np.random.RandomState(2).randint(0, 5, size=n)
    m = _ident(params, nets, prior_net)
    got, u, tr = m.metropolis_hastings_sampler((x, y, v), q_sd=0.7, burn_in=20, n_keep=40, noise=noise, segments=seg,
                                               return_trace=True, verbose=0)
    assert u.shape == (n, 5) and np.array_equal(u.argmax(1), seg)
    want, wtr = oc.mh_sampler(params, nets, (x, y, v), q_sd=0.7, burn_in=20, n_keep=40,
                              noise=oc.InjectedNoise(noise['z0'], noise['eps'], noise['u']), return_trace=True,
                              prior=oc.conditional_prior(params, prior_net, seg))
    agree = (tr['accept'] == np.array(wtr['accept'])).mean()
    assert agree > 0.999, agree
    same = np.all(got == want, axis=(0, 2))
    assert same.mean() > 0.98, same.mean()
    np.testing.assert_allclose(tr['lp_prop'][0], wtr['lp_prop'][0], rtol=1e-4, atol=1e-3)
    # the prior matters: the same noise under N(0, I) gives other chains
    base = oc.mh_sampler(params, nets, (x, y, v), q_sd=0.7, burn_in=20, n_keep=40,
                         noise=oc.InjectedNoise(noise['z0'], noise['eps'], noise['u']))
    assert not np.array_equal(base, want)


def test_identifiable_predict_runs_and_matches_oracle_reduction():
    params = causal_params(50, [1, 1, 1, 2], n_segments=3)
    nets = causal_nets(params)
    prior_net = onets.init_mlp(np.random.RandomState(4), [3, 64, 6], 0.5)
    n = 200
    x, y, v = causal_data(n, 50)
    m = _ident(params, nets, prior_net)
    np.random.seed(5)
    adrf, interval = m.predict((x, y, v), alpha=0.1, n_mcmc=50, burn_in=30, x_values=[0.5, 1.5], q_sd=0.5,
                               sample_y=False, seed=11, verbose=0)
    np.random.seed(5)
    seg = np.random.randint(0, 3, size=n)
    nz = m.philox_noise(11, n, 80)
    zs = oc.mh_sampler(params, nets, (x, y, v), q_sd=0.5, burn_in=30, n_keep=50,
                       noise=oc.InjectedNoise(nz['z0'], nz['eps'], nz['u']),
                       prior=oc.conditional_prior(params, prior_net, seg))
    ce = oc.infer_from_latent_posterior(params, nets, zs, x_values=[0.5, 1.5], sample_y=False)
    np.testing.assert_allclose(adrf, ce.mean(axis=1), rtol=2e-3, atol=2e-3)
    assert interval.shape == (2, 2) and np.all(interval[:, 0] <= interval[:, 1])


def _bank(params, S, seed=50):
    banks = [causal_nets(params, seed=seed + i) for i in range(S)]
    return banks, [np.stack([FLATV(b[k]) for b in banks]) for k in ('g', 'h', 'f')]


@pytest.mark.parametrize("binary", [False, True])
def test_fullmcmc_sampler_and_effect(binary):
    from bayesgm_b200 import FullMCMCCausalBGM
    params = causal_params(200, [1, 1, 1, 2], binary=binary)
    S, n, T = 4, 193, 50
    banks, (gs, hs, fs) = _bank(params, S)
    x, y, v = causal_data(n, 200, binary=binary)
    noise = injected_noise(n, 5, T)
    widx = np.random.RandomState(8).randint(0, S, size=T)
    m = FullMCMCCausalBGM(params=params)
    got, tr = m.metropolis_hastings_sampler((x, y, v), gs, hs, fs, q_sd=0.6, burn_in=20, n_keep=30, noise=noise,
                                            weight_idx=widx, return_trace=True, verbose=0)
    want, wtr = oc.mh_sampler(params, None, (x, y, v), q_sd=0.6, burn_in=20, n_keep=30,
                              noise=oc.InjectedNoise(noise['z0'], noise['eps'], noise['u']), return_trace=True,
                              nets_at=lambda t: banks[widx[t]])
    agree = (tr['accept'] == np.array(wtr['accept'])).mean()
    assert agree > 0.999, agree
    assert np.all(got == want, axis=(0, 2)).mean() > 0.97
    # log-posterior for one explicit set of weights (fullmcmc.py:344-394)
    z = noise['z0']
    lp = m.get_log_posterior(x, y, v, z, gs[2], hs[2], fs[2])
    np.testing.assert_allclose(lp, oc.log_posterior(params, banks[2], x, y, v, z), rtol=1e-4, atol=1e-3)
    # effect: kept state s with f_net sample pair[s] (:243-247, :285-342)
    pair = np.random.RandomState(9).randint(0, S, size=30)
    n_x = 2 if binary else 3
    xv = None if binary else [0.2, 1.0, 2.5]
    nz = np.random.RandomState(10).standard_normal((n_x, 30, n)).astype(np.float32)
    eff = m.infer_from_latent_posterior(want, f_net_weights=fs[pair], x_values=xv, sample_y=True, noise=nz)
    it = iter(nz) if binary else iter(nz)
    ref = oc.infer_from_latent_posterior(params, None, want, x_values=xv, sample_y=True,
                                         normal_fn=lambda shape: next(it), nets_at=lambda s: banks[pair[s]])
    if binary:
        np.testing.assert_allclose(eff, ref, rtol=1e-4, atol=2e-4)
    else:
        np.testing.assert_allclose(eff, ref.T, rtol=1e-4, atol=2e-4)


def test_fullmcmc_predict_with_one_weight_sample_is_the_base_model():
    """With a single weight sample every iteration uses the same nets: the bank path (one launch per
    iteration, current state re-evaluated) must reproduce the base sampler's chains on the same seed."""
    from bayesgm_b200 import FullMCMCCausalBGM
    from helpers import product_model
    params = causal_params(200, [1, 1, 1, 2])
    banks, (gs, hs, fs) = _bank(params, 1)
    n = 300
    x, y, v = causal_data(n, 200)
    m = FullMCMCCausalBGM(params=params)
    m.set_weight_samples(gs, hs, fs)
    got = m.metropolis_hastings_sampler((x, y, v), burn_in=60, n_keep=20, seed=5, verbose=0)     # adaptive q_sd
    base = product_model(params, banks[0])
    want = base.metropolis_hastings_sampler((x, y, v), burn_in=60, n_keep=20, seed=5, verbose=0)
    assert np.all(got == want, axis=(0, 2)).mean() > 0.97
    assert abs(m.last_q_sd - base.last_q_sd) < 1e-12
    adrf, interval = m.predict((x, y, v), alpha=0.1, n_mcmc=30, burn_in=30, x_values=[0.5, 1.0], q_sd=0.5, seed=3, verbose=0)
    assert adrf.shape == (2,) and interval.shape == (2, 2) and np.all(np.isfinite(adrf))
    with pytest.raises(NotImplementedError):
        m.run_mcmc_training((x, y, v))
