"""GPU parity of the Bayesian-network path (`use_bnn=True`, networks/bnn.py under
causalbgm/base.py:765-817, :820-904, :671-763) against oracle/bnn.py on the SAME noise: the oracle
regenerates the kernels' Philox streams (PhiloxFlipout), so values are compared one for one.
Tolerances: log-posterior 2e-4 * max(1, |lp|) (the kernel accumulates each Dense product in index
order with fp32 FMAs, NumPy's matmul blocks differently)."""
import numpy as np
import pytest

from oracle import bnn as obnn
from oracle import causal
from helpers import causal_params, causal_data, injected_noise

pytestmark = pytest.mark.gpu


def bnn_case(v_dim, z_dims, binary=False, g_units=(64,) * 5, f_units=(64, 32, 8), h_units=(64, 32, 8), seed=5, **extra):
    params = causal_params(v_dim, z_dims, binary=binary, g_units=g_units, f_units=f_units, h_units=h_units, **extra)
    params['use_bnn'] = True
    rs = np.random.RandomState(seed)
    zd = sum(z_dims)
    d0, d1, d2, _ = z_dims
    nets = dict(g=obnn.init_bnn(rs, [zd] + list(g_units) + [v_dim + 1], bn_random=True),
                e=obnn.init_bnn(rs, [v_dim] + [64] * 5 + [zd], bn_random=True),
                f=obnn.init_bnn(rs, [d0 + d1 + 1] + list(f_units) + [2], bn_random=True),
                h=obnn.init_bnn(rs, [d0 + d2] + list(h_units) + [2], bn_random=True))
    return params, nets


def product(params, nets, plan=None):
    from bayesgm_b200 import CausalBGM
    m = CausalBGM(params=params, random_seed=None)
    if plan:
        m.bnn_plan = plan
    kw = {}
    for k in 'gefh':
        w = [nets[k]['bn'][q] for q in ('gamma', 'beta', 'mean', 'var')]
        for layer in nets[k]['layers']:
            w += list(layer)
        kw[k] = w
    m.set_weights(**kw)
    return m


CASES = [
    dict(v_dim=200, z_dims=[1, 1, 1, 2]),
    dict(v_dim=200, z_dims=[1, 1, 1, 7]),
    dict(v_dim=100, z_dims=[3, 6, 3, 6], binary=True),
    dict(v_dim=37, z_dims=[2, 1, 2, 4], g_units=(20, 12), f_units=(9, 5), h_units=(33, 8)),
    dict(v_dim=64, z_dims=[1, 1, 1, 2], sigma_v=0.8, sigma_x=1.1, sigma_y=0.9),
]


def test_noise_streams_match_oracle():
    import ctypes as C
    import torch
    from bayesgm_b200 import _lib
    params, nets = bnn_case(200, [1, 1, 1, 2])
    m = product(params, nets)
    h = m._device_model()
    for net, name, layer in ((0, 'g', 0), (0, 'g', 5), (1, 'f', 1), (2, 'h', 3)):
        loc = nets[name]['layers'][layer][0]
        K, N = loc.shape
        rows, call, seed, sl, off = 70, 11, 12345678901, 3, 1000
        eps = torch.empty((K, N), dtype=torch.float32, device='cuda')
        si = torch.empty((rows, K), dtype=torch.int8, device='cuda')
        so = torch.empty((rows, N), dtype=torch.int8, device='cuda')
        _lib.call("bgm_bnn_noise", h, seed, sl, net, layer, call, off, rows, _lib.ptr(eps), _lib.ptr(si), _lib.ptr(so),
                  _lib.stream_ptr())
        we, wi, wo = obnn.PhiloxFlipout(seed, slice_id=sl, row_offset=off).flipout(name, layer, call, rows, K, N)
        np.testing.assert_allclose(eps.cpu().numpy(), we, rtol=0, atol=2e-6)
        np.testing.assert_array_equal(si.cpu().numpy().astype(np.float32), wi)
        np.testing.assert_array_equal(so.cpu().numpy().astype(np.float32), wo)
        assert abs(wi.mean()) < 0.2 and abs(we.mean()) < 0.2


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("n", [300, 1])
@pytest.mark.parametrize("plan", [1, 2])
def test_log_posterior_matches_oracle(case, n, plan):
    params, nets = bnn_case(**case)
    if n == 1 and case['v_dim'] != 200:
        pytest.skip("single-row batch checked on the standard shapes")
    x, y, v = causal_data(n, params['v_dim'], binary=params['binary_treatment'])
    z = np.random.RandomState(2).standard_normal((n, sum(params['z_dims']))).astype(np.float32)
    m = product(params, nets, plan)
    for call in (0, 7):
        got = m.get_log_posterior(x, y, v, z, seed=99, call=call)
        want = obnn.log_posterior(params, nets, x, y, v, z, obnn.PhiloxFlipout(99), call=call)
        err = np.abs(got - want) / np.maximum(1.0, np.abs(want))
        assert err.max() < 2e-4, (call, err.max())
    # a different call id is a different noise draw
    other = m.get_log_posterior(x, y, v, z, seed=99, call=8)
    assert np.abs(other - got).max() > 1e-3


@pytest.mark.parametrize("case", CASES[:4])
@pytest.mark.parametrize("mode", ["injected", "philox"])
@pytest.mark.parametrize("plan", [1, 2])
def test_mh_trace_matches_oracle(case, mode, plan):
    params, nets = bnn_case(**case)
    n, burn_in, n_keep, seed = 300, 4, 6, 4242
    T = burn_in + n_keep
    zd = sum(params['z_dims'])
    data = causal_data(n, params['v_dim'], binary=params['binary_treatment'])
    m = product(params, nets, plan)
    if mode == "injected":
        nz = injected_noise(n, zd, T)
        sg, tr = m.metropolis_hastings_sampler(data, q_sd=0.3, burn_in=burn_in, n_keep=n_keep, seed=seed, noise=nz,
                                               return_trace=True, verbose=0)
    else:
        sg, tr = m.metropolis_hastings_sampler(data, q_sd=0.3, burn_in=burn_in, n_keep=n_keep, seed=seed,
                                               return_trace=True, verbose=0)
        nz = m.philox_noise(seed, n, T)
    so, tro = obnn.mh_sampler(params, nets, data, obnn.PhiloxFlipout(seed), q_sd=0.3, burn_in=burn_in, n_keep=n_keep,
                              noise=causal.InjectedNoise(**nz), return_trace=True)
    acc_o = np.array(tro['accept'])
    # the rows of a slice are coupled through the batch statistics: compare up to the first iteration in
    # which any accept decision differs (a rounding-level tie), everything before must agree
    same_t = (tr['accept'] == acc_o).all(axis=1)
    first_bad = T if same_t.all() else int(np.argmin(same_t))
    assert first_bad >= T - 2, "accept decisions diverge at iteration %d of %d" % (first_bad, T)
    for t in range(min(first_bad + 1, T)):
        for name, got, want in (("lp_prop", tr['lp_prop'][t], tro['lp_prop'][t]), ("lp_cur", tr['lp_cur'][t], tro['lp_cur'][t])):
            err = np.abs(got - want) / np.maximum(1.0, np.abs(want))
            assert err.max() < 3e-4, (name, t, err.max())
    if first_bad == T:
        np.testing.assert_array_equal(sg, so)
    assert 0.02 < tr['accept'].mean() < 0.98


@pytest.mark.parametrize("binary", [False, True])
@pytest.mark.parametrize("sample_y", [False, True])
def test_effect_matches_oracle(binary, sample_y):
    params, nets = bnn_case(100 if binary else 200, [3, 6, 3, 6] if binary else [1, 1, 1, 2], binary=binary)
    n, n_keep = 300, 4
    zd = sum(params['z_dims'])
    rs = np.random.RandomState(3)
    zs = rs.standard_normal((n_keep, n, zd)).astype(np.float32)
    xs = None if binary else np.array([0.0, 0.7, 2.5])
    n_x = 2 if binary else 3
    noise = rs.standard_normal((n_x, n_keep, n)).astype(np.float32)
    m = product(params, nets)
    got = m.infer_from_latent_posterior(zs, x_values=xs, sample_y=sample_y, seed=31, noise=noise if sample_y else None)
    want = obnn.infer_from_latent_posterior(params, nets, zs, obnn.PhiloxFlipout(31), x_values=xs, sample_y=sample_y,
                                            normal_fn=(lambda shape: noise) if sample_y else None)
    assert got.shape == want.shape
    np.testing.assert_allclose(got, want, rtol=2e-4, atol=2e-4)


def test_predict_runs_on_shipped_config_verbatim():
    """src/configs/Sim_Hirano_Imbens.yaml as shipped (use_bnn: True) constructs and predicts."""
    from bayesgm_b200 import CausalBGM
    params = dict(dataset='Sim_Hirano_Imbens', output_dir='/tmp/bgm_b200_test', save_res=False, save_model=False,
                  binary_treatment=False, use_bnn=True, z_dims=[1, 1, 1, 7], v_dim=200, lr_theta=0.0001, lr_z=0.0001,
                  g_units=[64] * 5, f_units=[64, 32, 8], h_units=[64, 32, 8], kl_weight=0.0001, lr=0.0002, g_d_freq=5,
                  use_z_rec=True, e_units=[64] * 5, dz_units=[64, 32, 8])
    m = CausalBGM(params=params, random_seed=1)
    x, y, v = causal_data(700, 200)
    adrf, interval = m.predict(data=(x, y, v), alpha=0.05, n_mcmc=20, burn_in=20, x_values=[0.5, 1.5, 2.5], q_sd=1.0,
                               bs=300, verbose=0)
    assert adrf.shape == (3,) and interval.shape == (3, 2)
    assert np.isfinite(adrf).all() and (interval[:, 0] <= interval[:, 1]).all()
    assert 0.0 < m.last_acceptance_rate < 1.0
    # adaptive proposal scale
    s = m.metropolis_hastings_sampler((x[:256], y[:256], v[:256]), q_sd=None, burn_in=120, n_keep=5, seed=3, verbose=0)
    assert s.shape == (5, 256, 10) and np.isfinite(s).all()
