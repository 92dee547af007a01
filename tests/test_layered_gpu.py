"""GPU parity of the layered training engine (csrc/layered.cuh): CausalBGM's training steps on Bayesian nets
(DenseFlipout + batch-statistics BatchNorm) against torch autograd over the same Philox noise
(oracle/train_bnn.py), and on deterministic nets (any batch size) against oracle/train.py."""
import numpy as np
import pytest

from oracle import train as otrain
from oracle import train_bnn as obt
from helpers import causal_params, causal_nets, causal_data, product_model
from test_bnn_gpu import bnn_case, product as bnn_product

pytestmark = pytest.mark.gpu


def flat_bnn(grads, order='gefh'):
    return np.concatenate([a.ravel() for k in order for a in grads[k]])


def close(got, want, rtol=2e-3, what=""):
    scale = max(np.abs(want).max(), 1e-6)
    err = np.abs(got - want).max() / scale
    assert err < rtol, (what, err, np.abs(want).max())


def per_tensor_close(flat_got, grad_lists, order, rtol=3e-3):
    o = 0
    for k in order:
        for i, w in enumerate(grad_lists[k]):
            g = flat_got[o:o + w.size].reshape(w.shape)
            o += w.size
            scale = max(np.abs(w).max(), 1e-5)
            err = np.abs(g - w).max() / scale
            assert err < rtol, (k, i, err, np.abs(w).max())
    assert o == flat_got.size


BNN_CASES = [dict(v_dim=200, z_dims=[1, 1, 1, 7]), dict(v_dim=100, z_dims=[3, 6, 3, 6], binary=True),
             dict(v_dim=37, z_dims=[2, 1, 2, 4], g_units=(20, 12), f_units=(9, 5), h_units=(33, 8))]


def make_model(case, seed=77):
    params, nets = bnn_case(**case)
    m = bnn_product(params, nets)
    m._lt_seed = seed
    return params, nets, m


@pytest.mark.parametrize("case", BNN_CASES)
@pytest.mark.parametrize("bs", [32, 7, 80])
def test_bnn_gen_step_gradients(case, bs):
    params, nets, m = make_model(case)
    x, y, v = causal_data(bs, params['v_dim'], binary=params['binary_treatment'])
    z = np.random.RandomState(1).standard_normal((bs, sum(params['z_dims']))).astype(np.float32)
    m.set_noise_counter(5)
    losses, flat = m.gradients('gen', z, v, x, y)
    wl, wg = obt.gen_step(params, nets, m.dz_net.as_oracle_params(), z, v, x, y, seed=77, ctr=5)
    np.testing.assert_allclose(losses, wl, rtol=3e-4, atol=1e-5)
    per_tensor_close(flat, wg, 'gefh')


@pytest.mark.parametrize("case", BNN_CASES[:2])
def test_bnn_disc_step_gradients(case):
    params, nets, m = make_model(case)
    bs = 32
    x, y, v = causal_data(bs, params['v_dim'], binary=params['binary_treatment'])
    z = np.random.RandomState(2).standard_normal((bs, sum(params['z_dims']))).astype(np.float32)
    m.set_noise_counter(9)
    losses, flat = m.gradients('disc', z, v, epsilon=0.37)
    dzl, dl, grads = obt.disc_step(params, nets, m.dz_net.as_oracle_params(), z, v, 0.37, seed=77, ctr=9)
    np.testing.assert_allclose(losses, [dzl, dl], rtol=5e-4, atol=1e-5)
    want = np.concatenate([g.ravel() for g in grads])
    close(flat, want, 3e-3, "disc gradient")


@pytest.mark.parametrize("bs", [33, 96, 257])
@pytest.mark.parametrize("bnn", [False, True])
def test_disc_step_gradients_beyond_32_rows(bs, bnn):
    """train_disc_step (causalbgm/base.py:305-323) on mini-batches the fused kernel cannot hold: the three
    discriminator passes and the gradient-penalty double backward of csrc/disc_big.cuh against torch autograd."""
    if bnn:
        params, nets, m = make_model(BNN_CASES[0])
    else:
        params = causal_params(200, [1, 1, 1, 2])
        nets = causal_nets(params)
        m = product_model(params, nets)
        m._set_layered(True)
    x, y, v = causal_data(bs, params['v_dim'], binary=params['binary_treatment'])
    z = np.random.RandomState(2).standard_normal((bs, sum(params['z_dims']))).astype(np.float32)
    if bnn:
        m.set_noise_counter(9)
        losses, flat = m.gradients('disc', z, v, epsilon=0.37)
        dzl, dl, grads = obt.disc_step(params, nets, m.dz_net.as_oracle_params(), z, v, 0.37, seed=77, ctr=9)
    else:
        losses, flat = m.gradients('disc', z, v, epsilon=0.37)
        dzl, dl, grads = otrain.disc_step(params, nets, m.dz_net.as_oracle_params(), z, v, 0.37)
    np.testing.assert_allclose(losses, [dzl, dl], rtol=5e-4, atol=1e-5)
    want = np.concatenate([np.asarray(g).ravel() for g in grads])
    close(flat, want, 3e-3, "disc gradient")


def test_fit_with_batch_64_runs_egm_and_iterative_phase():
    from bayesgm_b200 import CausalBGM
    from bayesgm_b200.datasets import Sim_Hirano_Imbens_sampler
    params = dict(dataset='Sim_Hirano_Imbens', output_dir='/tmp/bgm_b200_test', save_res=False, save_model=False,
                  binary_treatment=False, use_bnn=False, z_dims=[1, 1, 1, 2], v_dim=40, lr_theta=0.001, lr_z=0.001,
                  g_units=[64] * 3, f_units=[64, 32, 8], h_units=[64, 32, 8], kl_weight=0.0001, lr=0.0002, g_d_freq=5,
                  use_z_rec=True, e_units=[64] * 3, dz_units=[64, 32, 8])
    x, y, v = Sim_Hirano_Imbens_sampler(N=256, v_dim=40).load_all()
    m = CausalBGM(params=params, random_seed=3)
    w0 = m.get_weights()
    m.fit(data=(x, y, v), batch_size=64, epochs=1, epochs_per_eval=10, use_egm_init=True, egm_n_iter=8, egm_batches_per_eval=100, verbose=0)
    w1 = m.get_weights()
    for k in ('g', 'e', 'f', 'h', 'dz'):
        assert any(np.abs(a - b).max() > 0 for a, b in zip(w0[k], w1[k])), k
        assert all(np.isfinite(a).all() for a in w1[k])


@pytest.mark.parametrize("case", BNN_CASES)
def test_bnn_iterative_steps(case):
    params, nets, m = make_model(case)
    n, bs = 300, 32
    zd = sum(params['z_dims'])
    x, y, v = causal_data(n, params['v_dim'], binary=params['binary_treatment'])
    rs = np.random.RandomState(4)
    zt = rs.standard_normal((n, zd)).astype(np.float32)
    idx = rs.choice(n, bs, replace=False)
    m.set_noise_counter(3)
    losses, flat = m.iter_gradients(zt, (x, y, v), idx)
    wl, wg = obt.iter_nets_step(params, nets, zt[idx], x[idx], y[idx], v[idx], seed=77, ctr=3)
    np.testing.assert_allclose(losses, wl, rtol=5e-4, atol=1e-5)
    # group 0 = g | e | f | h: e has no gradient in this phase
    wg['e'] = [np.zeros(tuple(a.shape), np.float32) for a in obt.param_list(obt.net_to_t(nets['e'], False))]
    per_tensor_close(flat, wg, 'gefh')
    # latent step
    m.set_noise_counter(4)
    loss, gz, z_new = m.latent_step(zt, (x, y, v), idx)
    wloss, wgz = obt.latent_step(params, nets, zt[idx], x[idx], y[idx], v[idx], seed=77, ctr=4)
    assert abs(loss - wloss) < 5e-4 * max(1.0, abs(wloss))
    close(gz, wgz, 3e-3, "latent gradient")
    # Keras Adam on the gathered variable: first step moves the batch rows by lr * sign(g), the others not at all
    lr = params['lr_z']
    moved = z_new - zt
    others = np.setdiff1d(np.arange(n), idx)
    assert np.abs(moved[others]).max() == 0.0
    big = np.abs(wgz) > 1e-4
    np.testing.assert_allclose(moved[idx][big], (-lr * np.sign(wgz))[big], rtol=2e-2)


def test_bnn_evaluate_matches_oracle_across_chunks():
    params, nets, m = make_model(BNN_CASES[0])
    n = 17000                                   # > one 16384-row chunk
    x, y, v = causal_data(n, params['v_dim'])
    m.set_noise_counter(11)
    causal_pre, mse_x, mse_y, mse_v = m.evaluate((x, y, v), nb_intervals=5)
    wx, wy, wv, wz = obt.evaluate_mse(params, nets, (x, y, v), seed=77, ctr=11)
    np.testing.assert_allclose([mse_x, mse_y, mse_v], [wx, wy, wv], rtol=5e-4)
    assert causal_pre.shape == (5,) and np.isfinite(causal_pre).all()
    zt = np.random.RandomState(0).standard_normal((n, sum(params['z_dims']))).astype(np.float32)
    m.set_noise_counter(12)
    _, mse_x, mse_y, mse_v = m.evaluate((x, y, v), data_z=zt, nb_intervals=3)
    wx, wy, wv, _ = obt.evaluate_mse(params, nets, (x, y, v), seed=77, ctr=12, data_z=zt)
    np.testing.assert_allclose([mse_x, mse_y, mse_v], [wx, wy, wv], rtol=5e-4)


@pytest.mark.parametrize("bs", [32, 100])
def test_deterministic_nets_on_the_layered_engine(bs):
    """use_bnn=False through the same kernels (no Flipout, no input BN): any batch size."""
    params = causal_params(200, [1, 1, 1, 2])
    nets = causal_nets(params)
    m = product_model(params, nets)
    m._set_layered(True)
    x, y, v = causal_data(bs, 200)
    z = np.random.RandomState(1).standard_normal((bs, 5)).astype(np.float32)
    losses, flat = m.gradients('gen', z, v, x, y)
    wl, wg = otrain.gen_step(params, nets, m.dz_net.as_oracle_params(), z, v, x, y)
    np.testing.assert_allclose(losses, wl, rtol=3e-4, atol=1e-6)
    per_tensor_close(flat, wg, 'gefh')


def test_bnn_fit_runs_end_to_end_on_shipped_config():
    """src/configs/Sim_Hirano_Imbens.yaml verbatim (use_bnn: True): egm_init + iterative phase + predict."""
    from bayesgm_b200 import CausalBGM
    from bayesgm_b200.datasets import Sim_Hirano_Imbens_sampler
    params = dict(dataset='Sim_Hirano_Imbens', output_dir='/tmp/bgm_b200_test', save_res=False, save_model=False,
                  binary_treatment=False, use_bnn=True, z_dims=[1, 1, 1, 7], v_dim=200, lr_theta=0.0001, lr_z=0.0001,
                  g_units=[64] * 5, f_units=[64, 32, 8], h_units=[64, 32, 8], kl_weight=0.0001, lr=0.0002, g_d_freq=5,
                  use_z_rec=True, e_units=[64] * 5, dz_units=[64, 32, 8])
    x, y, v = Sim_Hirano_Imbens_sampler(N=640, v_dim=200).load_all()
    m = CausalBGM(params=params, random_seed=3)
    w0 = m.get_weights()
    m.fit(data=(x, y, v), epochs=1, epochs_per_eval=1, use_egm_init=True, egm_n_iter=30, egm_batches_per_eval=10, verbose=0)
    w1 = m.get_weights()
    for k in ('g', 'e', 'f', 'h', 'dz'):
        assert any(np.abs(a - b).max() > 0 for a, b in zip(w0[k], w1[k])), k
        assert all(np.isfinite(a).all() for a in w1[k])
    assert m.data_z.shape == (640, 10) and np.isfinite(m.data_z).all()
    assert len(m.egm_history) == 4 and all(np.isfinite(h[1:]).all() for h in m.egm_history)
    adrf, interval = m.predict(data=(x, y, v), alpha=0.05, n_mcmc=10, burn_in=10, x_values=[0.5, 1.5], q_sd=1.0, bs=320, verbose=0)
    assert np.isfinite(adrf).all() and interval.shape == (2, 2)


@pytest.mark.parametrize("bnn", [True, False])
def test_graph_replay_is_bit_identical_to_plain_launches(bnn, monkeypatch):
    """The layered steps are captured once as CUDA graphs and replayed with their changing scalars (noise call
    counter, Adam bias corrections, WGAN-GP epsilon) in device memory: the same fit with BGM_LT_GRAPHS=0 (plain
    launches, scalars by value) must give the same parameters and latent table, bit for bit."""
    from bayesgm_b200 import CausalBGM
    from bayesgm_b200.datasets import Sim_Hirano_Imbens_sampler
    params = dict(dataset='Sim_Hirano_Imbens', output_dir='/tmp/bgm_b200_test', save_res=False, save_model=False,
                  binary_treatment=False, use_bnn=bnn, z_dims=[1, 1, 1, 2], v_dim=40, lr_theta=0.001, lr_z=0.001,
                  g_units=[64] * 3, f_units=[64, 32, 8], h_units=[64, 32, 8], kl_weight=0.0001, lr=0.0002, g_d_freq=5,
                  use_z_rec=True, e_units=[64] * 3, dz_units=[64, 32, 8])
    x, y, v = Sim_Hirano_Imbens_sampler(N=192, v_dim=40).load_all()
    out = {}
    for flag in ('0', '1'):
        monkeypatch.setenv('BGM_LT_GRAPHS', flag)
        np.random.seed(11)
        m = CausalBGM(params=dict(params), random_seed=3)
        if not bnn:
            m._set_layered(True)
        m.fit(data=(x, y, v), epochs=1, epochs_per_eval=10, use_egm_init=True, egm_n_iter=12, egm_batches_per_eval=100, verbose=0)
        w = m.get_weights()
        out[flag] = [a for k in ('g', 'e', 'f', 'h', 'dz') for a in w[k]] + [m.data_z]
    assert len(out['0']) == len(out['1'])
    for a, b in zip(out['0'], out['1']):
        np.testing.assert_array_equal(a, b)


# ------------------------------------------------------------------ BGM flavour (a14) ----
from oracle import train_bgm                                   # noqa: E402
from oracle import train                                        # noqa: E402
from test_train_gpu import make_bgm, bgm_gen_device_layout, check_grads, split_like   # noqa: E402

BGM_WIDE = [dict(x_dim=500, z_dim=10, bs=32), dict(x_dim=500, z_dim=10, bs=64, alpha=0.1), dict(x_dim=100, z_dim=10, bs=96),
            dict(x_dim=33, z_dim=5, bs=17, g_units=(16, 16), e_units=[16, 16], alpha=0.5)]


@pytest.mark.parametrize("kw", BGM_WIDE)
def test_bgm_layered_gen_and_disc_gradients(kw):
    """BASELINE cfg-5 width (x_dim = 500) and batches > 32: beyond the fused single-CTA kernels."""
    params, g, e, dz, dx, m, z, x, n1, n2 = make_bgm(**kw)
    m._set_layered(True)
    want_losses, want, stats = train_bgm.gen_step(params, g, e, dz, dx, z, x, n1, n2)
    losses, flat = m.gradients('gen', z, x, noise=n1, noise2=n2)
    np.testing.assert_allclose(losses, want_losses, rtol=3e-4, atol=1e-6)
    want_dev = bgm_gen_device_layout(want, g, params['x_dim'])
    check_grads(split_like(flat, want_dev), want_dev)
    want_losses, want, stats = train_bgm.disc_step(params, g, e, dz, dx, z, x, 0.3, 0.7, n1)
    losses, flat = m.gradients('disc', z, x, eps_z=0.3, eps_x=0.7, noise=n1)
    np.testing.assert_allclose(losses, want_losses, rtol=3e-4, atol=2e-6)
    check_grads(split_like(flat, want), want)


def test_bgm_wide_model_falls_back_to_the_layered_engine_and_trains():
    """x_dim = 500 does not fit the fused kernels (BGM_ERR_NOMEM): the model picks the layered engine by itself;
    steps track the oracle, the moving statistics follow Keras' update, fit() runs end to end."""
    params, g, e, dz, dx, m, z, x, n1, n2 = make_bgm(x_dim=500, z_dim=10, bs=32, lr=1e-3)
    import copy
    g0 = copy.deepcopy(g)
    gen_opt, d_opt = train.Adam(1e-3, 0.5, 0.9), train.Adam(1e-3, 0.5, 0.9)
    for k in range(2):
        _, grads, stats = train_bgm.disc_step(params, g, e, dz, dx, z, x, 0.3, 0.7, n1)
        d_opt.apply(train.disc_flat_params(dz) + train.disc_flat_params(dx), grads)
        train_bgm.update_moving(g, stats)
        m.train_disc_step(z, x, eps_z=0.3, eps_x=0.7, noise=n1)
        _, grads, stats = train_bgm.gen_step(params, g, e, dz, dx, z, x, n1, n2)
        gen_opt.apply(train_bgm.g_flat_params(g) + train.flat_params(e), grads)
        train_bgm.update_moving(g, stats)
        m.train_gen_step(z, x, noise1=n1, noise2=n2)
    assert m._layered
    w = m.get_weights()
    np.testing.assert_allclose(w['g'][2], g['bn']['mean'], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(w['g'][3], g['bn']['var'], rtol=1e-4, atol=1e-5)
    assert not np.allclose(w['g'][2], g0['bn']['mean'])
    want_g = [g['bn']['gamma'], g['bn']['beta']] + [a for W, b in g['hidden'] for a in (W, b)] + list(g['mean']) + list(g['var'])
    for a, b in zip(w['g'][:2] + w['g'][4:], want_g):
        assert np.median(np.abs(a - b)) <= 2e-5
    # iterative phase on the same engine: gradient rows and updated rows vs the oracle
    rs = np.random.RandomState(8)
    n = 200
    data = rs.standard_normal((n, 500)).astype(np.float32)
    zt = rs.standard_normal((n, 10)).astype(np.float32)
    idx = rs.choice(n, 48, replace=False)
    gnow = dict(bn=dict(gamma=w['g'][0], beta=w['g'][1], mean=w['g'][2], var=w['g'][3]),
                hidden=[(w['g'][4 + 2 * i], w['g'][5 + 2 * i]) for i in range(len(g['hidden']))],
                mean=(w['g'][-4], w['g'][-3]), var=(w['g'][-2], w['g'][-1]))
    lx, lm, gg, _ = train_bgm.iter_g_grads(gnow, zt[idx], data[idx])
    (gl, gm), gzl, ggz, cur = m.iter_step(zt, data, idx, apply=False)
    assert abs(gl - lx) < 3e-4 * max(1, abs(lx)) and abs(gm - lm) < 3e-4 * max(1, abs(lm))
    wl, wgz, _ = train_bgm.iter_latent_grad(gnow, zt[idx], data[idx])
    assert abs(gzl - wl) < 3e-4 * max(1, abs(wl))
    assert np.abs(ggz - wgz).max() < 3e-3 * np.abs(wgz).max()
    assert np.abs(cur[np.setdiff1d(np.arange(n), idx)] - zt[np.setdiff1d(np.arange(n), idx)]).max() == 0
    # evaluate / encode / fit
    mse = m.evaluate(data, data_z=zt, use_x_sd=False)
    assert np.isfinite(mse)
    m2_params, g2, *_ = make_bgm(x_dim=500, z_dim=10, bs=32)
    from helpers import bgm_product_model
    m2 = bgm_product_model(dict(m2_params, save_res=False, save_model=False), g2)
    m2.fit(data, batch_size=64, epochs=1, epochs_per_eval=1, use_egm_init=True, egm_n_iter=6, egm_batches_per_eval=3, verbose=0)
    assert m2._layered and np.isfinite(m2.history_loss).all()
    zenc = m2._encode_host(data)
    want = data
    L = len(m2.e_net.layers)
    for i, (W, b) in enumerate(m2.e_net.layers):
        want = want @ W + b
        if i < L - 1:
            want = np.where(want > 0, want, 0.2 * want)
    np.testing.assert_allclose(zenc, want, rtol=2e-4, atol=2e-4)
