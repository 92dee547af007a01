"""CPU checks that pin the oracle as far as it can be pinned without TensorFlow
(the reference's own tests hold no golden vectors for this path)."""
import numpy as np
import pytest
import torch

from oracle import nets as onets, causal, bgm
from helpers import causal_params, causal_nets, causal_data, injected_noise, philox4x32_10


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for c, k, want in kat:
        got = philox4x32_10(np.array(c, np.uint32), np.array(k, np.uint32))
        assert tuple(int(g) for g in got) == want


def torch_mlp(layers, x):
    h = x
    for W, b in layers[:-1]:
        h = torch.nn.functional.leaky_relu(h @ torch.from_numpy(W).double() + torch.from_numpy(b).double(), 0.2)
    W, b = layers[-1]
    return h @ torch.from_numpy(W).double() + torch.from_numpy(b).double()


@pytest.mark.parametrize("binary", [False, True])
def test_causal_log_posterior_matches_independent_float64(binary):
    params = causal_params(20, [1, 2, 1, 3], binary=binary)
    nets = causal_nets(params)
    x, y, v = causal_data(50, 20, binary)
    z = np.random.RandomState(1).standard_normal((50, 7)).astype(np.float32)
    lp = causal.log_posterior(params, nets, x, y, v, z)
    # independent restatement in float64 torch straight from causalbgm/base.py:775-816
    zt, xt, yt, vt = [torch.from_numpy(a).double() for a in (z, x, y, v)]
    g = torch_mlp(nets['g'], zt)
    h = torch_mlp(nets['h'], torch.cat([zt[:, :1], zt[:, 3:4]], 1))
    f = torch_mlp(nets['f'], torch.cat([zt[:, :1], zt[:, 1:3], xt], 1))
    sp = torch.nn.functional.softplus
    s2v, s2x, s2y = sp(g[:, -1]) + 1e-6, sp(h[:, -1]) + 1e-6, sp(f[:, -1]) + 1e-6
    lv = ((vt - g[:, :20]) ** 2).sum(1) / (2 * s2v) + 20 * torch.log(s2v) / 2
    if binary:
        lx = torch.nn.functional.binary_cross_entropy_with_logits(h[:, 0], xt[:, 0], reduction='none')
    else:
        lx = ((xt - h[:, :1]) ** 2).sum(1) / (2 * s2x) + torch.log(s2x) / 2
    ly = ((yt - f[:, :1]) ** 2).sum(1) / (2 * s2y) + torch.log(s2y) / 2
    want = -(lv + lx + ly + (zt ** 2).sum(1) / 2)
    np.testing.assert_allclose(lp, want.numpy(), rtol=2e-5, atol=2e-4)


def test_fixed_sigma_keys_are_honoured():
    params = causal_params(8, [1, 1, 1, 1], sigma_v=0.7, sigma_y=1.3)
    nets = causal_nets(params)
    x, y, v = causal_data(10, 8)
    z = np.zeros((10, 4), np.float32)
    a = causal.log_posterior(params, nets, x, y, v, z)
    params2 = dict(params, sigma_v=0.9)
    b = causal.log_posterior(params2, nets, x, y, v, z)
    assert not np.allclose(a, b)


def test_mh_cached_current_equals_recomputed_current():
    params = causal_params(12, [1, 1, 1, 2])
    nets = causal_nets(params)
    data = causal_data(40, 12)
    nz = injected_noise(40, 5, 30)
    a = causal.mh_sampler(params, nets, data, q_sd=1.0, burn_in=10, n_keep=20,
                          noise=causal.InjectedNoise(**nz), recompute_current=True)
    b = causal.mh_sampler(params, nets, data, q_sd=1.0, burn_in=10, n_keep=20,
                          noise=causal.InjectedNoise(**nz), recompute_current=False)
    assert a.shape == (20, 40, 5)
    np.testing.assert_array_equal(a, b)


def test_mh_numpy_global_stream_call_order():
    """normal(n,zd) once, then per iteration normal(n,zd), rand(n)."""
    params = causal_params(6, [1, 1, 1, 1])
    nets = causal_nets(params)
    data = causal_data(8, 6)
    np.random.seed(5)
    a = causal.mh_sampler(params, nets, data, q_sd=0.5, burn_in=3, n_keep=4)
    rs = np.random.RandomState(5)
    z0 = rs.normal(0, 1, (8, 4)).astype('float32')
    eps, u = [], []
    for _ in range(7):
        eps.append(rs.normal(0, 0.5, (8, 4)).astype('float32') / np.float32(0.5))
        u.append(rs.rand(8))
    b = causal.mh_sampler(params, nets, data, q_sd=0.5, burn_in=3, n_keep=4,
                          noise=causal.InjectedNoise(z0, np.array(eps), np.array(u)))
    np.testing.assert_allclose(a, b, rtol=0, atol=1e-6)


def test_mh_adaptive_q_sd_moves_towards_target():
    params = causal_params(6, [1, 1, 1, 1])
    nets = causal_nets(params)
    data = causal_data(64, 6)
    np.random.seed(0)
    _, tr = causal.mh_sampler(params, nets, data, q_sd=None, initial_q_sd=5.0, burn_in=400, n_keep=5,
                              return_trace=True)
    assert tr['q_sd_final'] < 5.0
    assert len(set(tr['q_sd'])) > 1


def test_bgm_hand_gradient_matches_autograd():
    rs = np.random.RandomState(3)
    p = onets.init_variational(rs, 3, 10, [16, 16], bias_scale=0.1, bn_random=True)
    x = rs.standard_normal((20, 10)).astype(np.float32)
    w = (rs.uniform(size=(20, 10)) < 0.7).astype(np.float32)
    z = rs.standard_normal((20, 3)).astype(np.float32)
    lp, g = bgm.log_posterior_and_grad(p, z, x, w)
    zt = torch.from_numpy(z).double().requires_grad_(True)
    T = lambda a: torch.from_numpy(np.asarray(a)).double()
    h = (zt - T(p['bn']['mean'])) * T(p['bn']['gamma']) / torch.sqrt(T(p['bn']['var']) + 1e-3) + T(p['bn']['beta'])
    for W, b in p['hidden']:
        h = torch.nn.functional.leaky_relu(h @ T(W) + T(b), 0.2)
    mu = h @ T(p['mean'][0]) + T(p['mean'][1])
    s2 = torch.nn.functional.softplus(h @ T(p['var'][0]) + T(p['var'][1])) + 1e-6
    lpt = -(((T(x) - mu) ** 2 / (2 * s2) + 0.5 * torch.log(s2)) * T(w)).sum(1) - (zt ** 2).sum(1) / 2
    lpt.sum().backward()
    np.testing.assert_allclose(lp, lpt.detach().numpy(), rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(g, zt.grad.numpy(), rtol=1e-4, atol=1e-4)


def test_bgm_gather_and_dense_mask_formulations_agree():
    rs = np.random.RandomState(4)
    p = onets.init_variational(rs, 3, 10, [16, 16], bias_scale=0.1)
    x = rs.standard_normal((6, 10)).astype(np.float32)
    lists = [sorted(rs.choice(10, size=rs.randint(1, 10), replace=False).tolist()) for _ in range(6)]
    ind, mask = bgm.pad_index_lists(lists, 6)
    z = rs.standard_normal((6, 3)).astype(np.float32)
    a = bgm.log_posterior(p, z, x, ind, mask)
    w = bgm.dense_mask_from_indices(ind, mask, 10)
    b, _ = bgm.log_posterior_and_grad(p, z, x, w)
    np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-5)


def test_hmc_small_step_accepts_and_adapts():
    rs = np.random.RandomState(5)
    p = onets.init_variational(rs, 3, 10, [16, 16], bias_scale=0.1)
    x = rs.standard_normal((32, 10)).astype(np.float32)
    s, tr = bgm.hmc_sampler(p, x, n_mcmc=5, burn_in=20, step_size=0.01, num_leapfrog_steps=5,
                            rs=np.random.RandomState(1), return_trace=True)
    assert s.shape == (5, 32, 3)
    assert np.mean(tr['accept']) > 0.9           # tiny steps conserve energy
    assert tr['step'][16] > tr['step'][0]        # so the shared step size grows (16 adaptation steps)
    assert tr['step'][-1] == tr['step'][17]      # and freezes after int(0.8*burn_in)


def test_torch_cpu_restatement_matches_numpy_oracle():
    """oracle/causal_torch.py (the multi-threaded CPU timing arm of bench.py) is the same algorithm as
    oracle/causal.py: same log-posterior, and on the same NumPy RNG stream the same chains."""
    import torch
    from oracle import causal_torch
    for binary in (False, True):
        params = causal_params(12, [1, 2, 1, 2], binary=binary)
        nets = causal_nets(params)
        x, y, v = causal_data(64, 12, binary=binary)
        z = np.random.RandomState(3).standard_normal((64, 6)).astype(np.float32)
        want = causal.log_posterior(params, nets, x, y, v, z)
        tn = causal_torch.to_torch_nets(nets)
        got = causal_torch.log_posterior(params, tn, *[torch.from_numpy(a) for a in (x, y, v, z)]).numpy()
        np.testing.assert_allclose(got, want, rtol=2e-5, atol=2e-5)
        a = causal_torch.mh_sampler(params, nets, (x, y, v), q_sd=0.4, burn_in=2, n_keep=4, rs=np.random.RandomState(9))
        b = causal.mh_sampler(params, nets, (x, y, v), q_sd=0.4, burn_in=2, n_keep=4,
                              noise=causal.NumpyGlobalNoise(np.random.RandomState(9)))
        same = (a == b).all(axis=(0, 2))
        assert same.mean() > 0.9


def test_conditional_prior_and_per_iteration_nets_reduce_to_the_base_algorithm():
    """oracle.causal: identifiable.py:540-548 with mu = 0, sigma^2 = 1 is the N(0, I) prior of base.py:812; a
    constant `nets_at` (fullmcmc.py:441-449 with one weight sample) is the base sampler."""
    from helpers import causal_params, causal_nets, causal_data, injected_noise
    from oracle import causal as oc
    from oracle import nets as onets
    params = causal_params(20, [1, 1, 1, 2], n_segments=4)
    nets = causal_nets(params)
    n, T = 40, 12
    x, y, v = causal_data(n, 20)
    z = np.random.RandomState(0).standard_normal((n, 5)).astype(np.float32)
    base = oc.log_posterior(params, nets, x, y, v, z)
    unit = (np.zeros((n, 5), np.float32), np.ones(n, np.float32))
    np.testing.assert_allclose(oc.log_posterior(params, nets, x, y, v, z, prior=unit), base, rtol=1e-6, atol=1e-5)
    # by hand for one row: -(sum (z - mu)^2 / (2 s2) + zd log(s2) / 2) replaces -sum z^2 / 2
    prior_net = onets.init_mlp(np.random.RandomState(3), [4, 64, 6], 0.5)
    seg = np.random.RandomState(1).randint(0, 4, size=n)
    mu, s2 = oc.conditional_prior(params, prior_net, seg)
    assert mu.shape == (n, 5) and s2.shape == (n,) and np.all(s2 > 0)
    assert np.array_equal(mu[seg == seg[0]], np.repeat(mu[:1], (seg == seg[0]).sum(), axis=0))   # a table over segments
    got = oc.log_posterior(params, nets, x, y, v, z, prior=(mu, s2))
    want = base.astype(np.float64) + (z.astype(np.float64) ** 2).sum(1) / 2 - (
        ((z - mu).astype(np.float64) ** 2).sum(1) / (2 * s2) + 5 * np.log(s2.astype(np.float64)) / 2)
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-3)
    nz = injected_noise(n, 5, T)
    mk = lambda: oc.InjectedNoise(nz['z0'], nz['eps'], nz['u'])
    a = oc.mh_sampler(params, nets, (x, y, v), q_sd=0.5, burn_in=4, n_keep=8, noise=mk())
    b = oc.mh_sampler(params, None, (x, y, v), q_sd=0.5, burn_in=4, n_keep=8, noise=mk(), nets_at=lambda t: nets)
    assert np.array_equal(a, b)
    c = oc.infer_from_latent_posterior(params, nets, a, x_values=[0.5], sample_y=False)
    d = oc.infer_from_latent_posterior(params, None, a, x_values=[0.5], sample_y=False, nets_at=lambda s: nets)
    assert np.array_equal(c, d)
