"""GPU parity tests of the EGM training steps: CUDA (hand-derived backward and double
backward, through the C ABI) vs the torch-autograd oracle on the same batches.

Tolerances (fp32): losses rtol 2e-4; every parameter-gradient tensor
|d| <= 2e-3 * max|grad of that tensor| + 1e-7 (different summation orders through ~10
layers and a double backward; + 2e-6 * the largest gradient of the step); after k optimizer steps parameters within 5e-4 absolute
(Adam normalises the update to ~lr per step regardless of the gradient scale, so a
sign-level disagreement on a near-zero gradient moves a parameter by at most k*lr).
"""
import copy

import numpy as np
import pytest

from oracle import train, nets as onets
from helpers import causal_params, causal_nets, causal_data, product_model

pytestmark = pytest.mark.gpu


def make(v_dim=200, z_dims=(1, 1, 1, 2), binary=False, bs=32, seed=0, **extra):
    params = causal_params(v_dim, list(z_dims), binary, **extra)
    nets = causal_nets(params, seed=11)
    rs = np.random.RandomState(seed)
    zd = sum(z_dims)
    dz = onets.init_discriminator(rs, zd, params['dz_units'])
    for bn in dz['bns']:
        bn['gamma'] = (1 + 0.1 * rs.standard_normal(bn['gamma'].shape)).astype(np.float32)
        bn['beta'] = (0.1 * rs.standard_normal(bn['beta'].shape)).astype(np.float32)
    for i, (W, b) in enumerate(dz['layers']):
        dz['layers'][i] = (W, (0.1 * rs.standard_normal(b.shape)).astype(np.float32))
    x, y, v = causal_data(bs, v_dim, binary, seed=seed + 1)
    z = rs.standard_normal((bs, zd)).astype(np.float32)
    m = product_model(params, nets)
    m.set_weights(dz=train.disc_flat_params(dz))
    return params, nets, dz, m, z, v, x, y


def split_like(flat, arrays):
    out, o = [], 0
    for a in arrays:
        out.append(flat[o:o + a.size].reshape(a.shape))
        o += a.size
    assert o == flat.size
    return out


def check_grads(got_list, want_list, names=None):
    # a Dense bias in front of a batch-statistics BatchNorm has a mathematically ZERO gradient:
    # both sides then hold rounding noise, so the absolute floor scales with the largest gradient
    gmax = max(float(np.abs(w).max()) for w in want_list)
    for i, (g, w) in enumerate(zip(got_list, want_list)):
        tol = 2e-3 * np.abs(w).max() + 2e-6 * gmax + 1e-7
        err = np.abs(g - w).max()
        assert err <= tol, "tensor %d (%s): err %.3g > tol %.3g (max |g| %.3g)" % (
            i, names[i] if names else "?", err, tol, np.abs(w).max())


GEN_CASES = [
    dict(),                                                       # cfg-3 shape, batch 32
    dict(v_dim=100, z_dims=(3, 6, 3, 6), binary=True),            # cfg-2 shape, binary treatment
    dict(v_dim=177, z_dims=(2, 1, 1, 1), bs=20),                  # ragged batch, odd v_dim
    dict(v_dim=10, z_dims=(1, 1, 1, 0), g_units=[8, 8], e_units=[8, 8], f_units=[8, 8], h_units=[8, 8],
         dz_units=[8, 8]),                                        # the R tests' tiny nets
    dict(use_z_rec=False),
]


@pytest.mark.parametrize("kw", GEN_CASES)
def test_gen_step_gradients(kw):
    params, nets, dz, m, z, v, x, y = make(**kw)
    want_losses, want = train.gen_step(params, nets, dz, z, v, x, y)
    losses, flat = m.gradients('gen', z, v, x, y)
    np.testing.assert_allclose(losses, want_losses, rtol=2e-4, atol=1e-6)
    want_list = want['g'] + want['e'] + want['f'] + want['h']
    check_grads(split_like(flat, want_list), want_list)


DISC_CASES = [dict(), dict(v_dim=100, z_dims=(3, 6, 3, 6)), dict(v_dim=40, z_dims=(1, 1, 1, 1), bs=9),
              dict(v_dim=12, z_dims=(1, 1, 1, 2), dz_units=[16, 16, 16, 4])]


@pytest.mark.parametrize("kw", DISC_CASES)
def test_disc_step_gradients_including_the_gradient_penalty(kw):
    params, nets, dz, m, z, v, x, y = make(**kw)
    for eps in (0.3, 0.9):
        dz_loss, d_loss, want = train.disc_step(params, nets, dz, z, v, eps)
        losses, flat = m.gradients('disc', z, v, epsilon=eps)
        np.testing.assert_allclose(losses, [dz_loss, d_loss], rtol=2e-4, atol=2e-6)
        check_grads(split_like(flat, want), want)


def test_optimizer_steps_track_the_oracle():
    params, nets, dz, m, z, v, x, y = make(v_dim=60, z_dims=(1, 2, 1, 2), lr=1e-3)
    tr = train.EgmTrainer(params, copy.deepcopy(nets), copy.deepcopy(dz))
    rs = np.random.RandomState(5)
    for k in range(4):
        zz = rs.standard_normal(z.shape).astype(np.float32)
        eps = float(rs.uniform())
        a = tr.train_disc_step(zz, v, eps)
        b = m.train_disc_step(zz, v, epsilon=eps)
        np.testing.assert_allclose(b, a, rtol=1e-3, atol=1e-5)
        a = tr.train_gen_step(zz, v, x, y)
        b = m.train_gen_step(zz, v, x, y)
        np.testing.assert_allclose(b, a, rtol=1e-3, atol=1e-5)
    w = m.get_weights()
    for name in ('g', 'e', 'f', 'h'):
        for got, want in zip(w[name], train.flat_params(tr.nets[name])):
            assert np.abs(got - want).max() <= 5e-4
    for i, (got, want) in enumerate(zip(w['dz'], train.disc_flat_params(tr.dz))):
        if i % 4 == 1 and i < len(w['dz']) - 2:
            # Dense bias in front of a batch-statistics BN: zero true gradient, Adam turns the
            # rounding noise into +-lr steps on both sides; the value never reaches the output
            assert np.abs(got - want).max() <= 2 * 4 * 1e-3 + 1e-6
            continue
        assert np.abs(got - want).max() <= 5e-4, i
    # ... and the discriminators agree as functions
    zt = rs.standard_normal(z.shape).astype(np.float32)
    dz_got = dict(layers=[(w['dz'][4 * b], w['dz'][4 * b + 1]) for b in range(3)] + [(w['dz'][-2], w['dz'][-1])],
                  bns=[dict(gamma=w['dz'][4 * b + 2], beta=w['dz'][4 * b + 3]) for b in range(3)])
    np.testing.assert_allclose(onets.discriminator_forward(dz_got, zt), onets.discriminator_forward(tr.dz, zt),
                               rtol=0, atol=5e-3)
    # the sampler picks up the trained weights
    from oracle import causal
    lp = m.get_log_posterior(x, y, v, zz)
    want_lp = causal.log_posterior(params, {k: [(W, b) for W, b in zip(w[k][0::2], w[k][1::2])] for k in 'gefh'},
                                   x, y, v, zz)
    assert (np.abs(lp - want_lp) / np.maximum(1, np.abs(want_lp))).max() <= 1e-4


def test_egm_init_follows_the_reference_index_stream():
    """Same NumPy seed on both sides: mini-batch indices and prior draws are bit-identical
    (host RNG, reference call order), so after a few iterations the parameters agree."""
    params = causal_params(30, [1, 1, 1, 2], lr=1e-3, g_d_freq=2)
    nets = causal_nets(params, seed=11)
    data = causal_data(500, 30)
    m = product_model(params, nets)
    dz0 = m.dz_net.as_oracle_params()
    tr = train.EgmTrainer(params, copy.deepcopy(nets), copy.deepcopy(dz0))
    eps_stream = np.random.RandomState(77).uniform(size=100).astype(np.float32)
    it = iter(eps_stream)
    np.random.seed(123)
    tr.egm_init(data, 3, 16, m.z_sampler, lambda: float(next(it)))
    after_oracle = np.random.get_state()[1][:5].copy()

    class Eps(object):
        def __init__(self):
            self.i = 0

        def uniform(self):
            self.i += 1
            return eps_stream[self.i - 1]
    m._eps_rng = Eps()
    np.random.seed(123)
    m.egm_init(data, egm_n_iter=3, batch_size=16, egm_batches_per_eval=2, verbose=0)
    np.testing.assert_array_equal(np.random.get_state()[1][:5], after_oracle)   # same number of draws
    w = m.get_weights()
    for name in ('g', 'e', 'f', 'h'):
        for got, want in zip(w[name], train.flat_params(tr.nets[name])):
            assert np.abs(got - want).max() <= 1e-3


# ------------------------------------------------------------------ BGM (a14) --
from oracle import train_bgm
from helpers import bgm_params, bgm_oracle_net, bgm_product_model


def make_bgm(x_dim=100, z_dim=10, bs=32, seed=0, **extra):
    params = bgm_params(x_dim, z_dim, **extra)
    g = bgm_oracle_net(params, seed=31, bn_random=True)
    rs = np.random.RandomState(seed)
    e = onets.init_mlp(rs, [x_dim] + params['e_units'] + [z_dim], 0.1)
    dz = onets.init_discriminator(rs, z_dim, params['dz_units'])
    dx = onets.init_discriminator(rs, x_dim, params['dx_units'])
    for d in (dz, dx):
        for bn in d['bns']:
            bn['gamma'] = (1 + 0.1 * rs.standard_normal(bn['gamma'].shape)).astype(np.float32)
            bn['beta'] = (0.1 * rs.standard_normal(bn['beta'].shape)).astype(np.float32)
    m = bgm_product_model(params, g)
    m.set_weights(e=[a for W, b in e for a in (W, b)], dz=train.disc_flat_params(dz), dx=train.disc_flat_params(dx))
    z = rs.standard_normal((bs, z_dim)).astype(np.float32)
    x = rs.standard_normal((bs, x_dim)).astype(np.float32)
    n1 = rs.standard_normal((bs, x_dim)).astype(np.float32)
    n2 = rs.standard_normal((bs, x_dim)).astype(np.float32)
    return params, g, e, dz, dx, m, z, x, n1, n2


def bgm_gen_device_layout(grads, g, x_dim):
    """oracle (Keras order) gradient list -> tensors in the device layout of group 0."""
    nh = len(g['hidden'])
    gl = grads[:2 + 2 * nh + 4]
    el = grads[2 + 2 * nh + 4:]
    out = gl[:2 + 2 * nh]
    wm, bm, wv, bv = gl[2 + 2 * nh:]
    out += [np.concatenate([wm, wv], axis=1), np.concatenate([bm, bv])]
    return out + el


@pytest.mark.parametrize("kw", [dict(), dict(x_dim=10, z_dim=3, bs=17), dict(x_dim=4, z_dim=2, alpha=0.1),
                                dict(x_dim=33, z_dim=5, g_units=(16, 16), e_units=[16, 16], alpha=0.5)])
def test_bgm_gen_step_gradients(kw):
    params, g, e, dz, dx, m, z, x, n1, n2 = make_bgm(**kw)
    want_losses, want, stats = train_bgm.gen_step(params, g, e, dz, dx, z, x, n1, n2)
    losses, flat = m.gradients('gen', z, x, noise=n1, noise2=n2)
    np.testing.assert_allclose(losses, want_losses, rtol=2e-4, atol=1e-6)
    want_dev = bgm_gen_device_layout(want, g, params['x_dim'])
    check_grads(split_like(flat, want_dev), want_dev)


@pytest.mark.parametrize("kw", [dict(), dict(x_dim=10, z_dim=3, bs=17), dict(x_dim=20, z_dim=4, gamma=1.0),
                                dict(x_dim=4, z_dim=2, gamma=0.5, bs=12)])
def test_bgm_disc_step_gradients(kw):
    params, g, e, dz, dx, m, z, x, n1, n2 = make_bgm(**kw)
    want_losses, want, stats = train_bgm.disc_step(params, g, e, dz, dx, z, x, 0.3, 0.7, n1)
    losses, flat = m.gradients('disc', z, x, eps_z=0.3, eps_x=0.7, noise=n1)
    np.testing.assert_allclose(losses, want_losses, rtol=2e-4, atol=2e-6)
    check_grads(split_like(flat, want), want)


def test_bgm_steps_update_weights_and_moving_statistics():
    params, g, e, dz, dx, m, z, x, n1, n2 = make_bgm(x_dim=12, z_dim=3, lr=1e-3)
    import copy
    g0 = copy.deepcopy(g)
    gen_opt, d_opt = train.Adam(1e-3, 0.5, 0.9), train.Adam(1e-3, 0.5, 0.9)
    for k in range(3):
        _, grads, stats = train_bgm.disc_step(params, g, e, dz, dx, z, x, 0.3, 0.7, n1)
        d_opt.apply(train.disc_flat_params(dz) + train.disc_flat_params(dx), grads)
        train_bgm.update_moving(g, stats)
        m.train_disc_step(z, x, eps_z=0.3, eps_x=0.7, noise=n1)
        _, grads, stats = train_bgm.gen_step(params, g, e, dz, dx, z, x, n1, n2)
        gen_opt.apply(train_bgm.g_flat_params(g) + train.flat_params(e), grads)
        train_bgm.update_moving(g, stats)
        m.train_gen_step(z, x, noise1=n1, noise2=n2)
    w = m.get_weights()
    # Keras order of g: gamma, beta, moving_mean, moving_var, hidden..., mean, var
    np.testing.assert_allclose(w['g'][2], g['bn']['mean'], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(w['g'][3], g['bn']['var'], rtol=1e-4, atol=1e-5)
    assert not np.allclose(w['g'][2], g0['bn']['mean'])
    want_g = [g['bn']['gamma'], g['bn']['beta']] + [a for W, b in g['hidden'] for a in (W, b)] + \
        list(g['mean']) + list(g['var'])
    got_g = w['g'][:2] + w['g'][4:]
    for a, b in zip(got_g, want_g):
        assert np.abs(a - b).max() <= 6 * 1e-3 * 0.5     # k*lr bound; most entries agree to 1e-5
        assert np.median(np.abs(a - b)) <= 2e-5
    for a, b in zip(w['e'], train.flat_params(e)):
        assert np.median(np.abs(a - b)) <= 2e-5


def test_bgm_egm_init_runs_and_keeps_the_reference_batch_stream():
    params = bgm_params(8, 2, lr=1e-3)
    g = bgm_oracle_net(params, seed=31)
    m = bgm_product_model(params, g)
    data = np.random.RandomState(3).standard_normal((200, 8)).astype(np.float32)
    from bayesgm_b200.datasets import Base_sampler
    np.random.seed(5)
    dl, gl = m.egm_init(data, egm_n_iter=4, batch_size=16, egm_batches_per_eval=2, verbose=0)
    after = np.random.get_state()[1][:4].copy()
    assert np.isfinite(dl).all() and np.isfinite(gl).all()
    # the same host-RNG consumption as the reference loop: Base_sampler ctor, then per iteration
    # (g_d_freq + 1) x (next_batch, get_batch)
    np.random.seed(5)
    s = Base_sampler(x=data, y=data, v=data, batch_size=16, normalize=False)
    for it in range(5):
        for _ in range(params['g_d_freq'] + 1):
            s.next_batch()
            m.z_sampler.get_batch(16)
    np.testing.assert_array_equal(np.random.get_state()[1][:4], after)


# --------------------------------------------------- iterative phase of fit (N2) --
def test_fit_iterative_phase_tracks_the_oracle():
    """fit(use_egm_init=False): same NumPy seed on both sides -> same random latent table and
    epoch permutations (bit-exact host streams); g/h/f updates, the latent gradient and the
    DENSE Keras-Adam sweep over the whole table then agree with the oracle loop."""
    from oracle import causal
    params = causal_params(20, [1, 2, 1, 1], lr_theta=1e-3, lr_z=1e-3)
    nets = causal_nets(params, seed=11)
    n, bs = 70, 16                                   # 70 = 4 batches of 16 + a ragged 6
    data = causal_data(n, 20)
    m = product_model(params, nets)
    np.random.seed(42)
    m.fit(data, epochs=1, epochs_per_eval=1, batch_size=bs, use_egm_init=False, verbose=0)
    np.random.seed(42)
    z0 = np.random.normal(0, 1, size=(n, 5)).astype('float32')
    tr = train.IterTrainer(params, copy.deepcopy(nets), z0)
    for epoch in range(2):
        last = tr.epoch(data, bs)
    # every row of the table moved (dense Adam), and it moved like the oracle's
    assert (np.abs(m.data_z - z0) > 0).all()
    np.testing.assert_allclose(m.data_z, tr.data_z, rtol=0, atol=2e-4)
    w = m.get_weights()
    for name in ('g', 'f', 'h'):
        for got, want in zip(w[name], train.flat_params(tr.nets[name])):
            assert np.median(np.abs(got - want)) <= 2e-5 and np.abs(got - want).max() <= 3e-3
    for got, want in zip(w['e'], train.flat_params(nets['e'])):
        np.testing.assert_array_equal(got, want)                 # e_net is not trained in this phase
    np.testing.assert_allclose(m.last_iter_losses[:6], last[0], rtol=2e-3, atol=1e-5)
    np.testing.assert_allclose(m.last_iter_losses[6], last[1], rtol=2e-3)
    # evaluate (:534-570) on the trained state
    got = m.evaluate(data, data_z=m.data_z)
    onets_now = {k: [(W, b) for W, b in zip(w[k][0::2], w[k][1::2])] for k in 'gefh'}
    want = causal.evaluate(params, onets_now, data, data_z=m.data_z)
    np.testing.assert_allclose(got[0], want[0], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(got[1:], want[1:], rtol=1e-4)
    assert m.best_causal_pre.shape == (200,) and m.best_epoch in (0, 1)
    got2 = m.evaluate(data)                                      # data_z=None -> z = e_net(v)
    want2 = causal.evaluate(params, onets_now, data)
    np.testing.assert_allclose(got2[0], want2[0], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(got2[1:], want2[1:], rtol=1e-4)


def test_fit_with_egm_init_binary_end_to_end():
    params = causal_params(12, [1, 1, 1, 1], binary=True, lr=1e-3, lr_theta=1e-3, lr_z=1e-3, g_d_freq=1)
    nets = causal_nets(params, seed=11)
    data = causal_data(64, 12, binary=True)
    m = product_model(params, nets)
    np.random.seed(1)
    m.fit(data, epochs=0, epochs_per_eval=1, batch_size=32, use_egm_init=True, egm_n_iter=3, verbose=0)
    assert m.data_z.shape == (64, 4) and np.isfinite(m.data_z).all()
    assert m.best_causal_pre.shape == (64, 1)                    # ITE per subject (:559-563)
    ite, interval = m.predict(data, n_mcmc=10, burn_in=10, q_sd=0.5, verbose=0)
    assert ite.shape == (64,) and np.isfinite(ite).all()


# ------------------------------------------ BGM: iterative phase of fit (bgm/base.py:145-187) --
@pytest.mark.parametrize("kw", [dict(x_dim=20, z_dim=4), dict(x_dim=100, z_dim=10, bs=32),
                                dict(x_dim=7, z_dim=3, bs=11, g_units=(16, 16))])
def test_bgm_iter_step_tracks_the_oracle(kw):
    """update_g_net + update_latent_variable_sgd, three consecutive mini-batches: losses, the
    gradient w.r.t. the latent rows (through the training-mode BatchNormalization), the scattered
    latent table (fresh-variable Adam), the generator parameters and the BN moving statistics."""
    import copy
    bs = kw.pop('bs', 16)
    params, g, e, dz, dx, m, _, _, _, _ = make_bgm(bs=bs, lr_theta=1e-3, lr_z=1e-2, **kw)
    x_dim, z_dim = params['x_dim'], params['z_dim']
    rs = np.random.RandomState(8)
    n = 80
    data = rs.standard_normal((n, x_dim)).astype(np.float32)
    table = rs.standard_normal((n, z_dim)).astype(np.float32)
    ot = train_bgm.BgmIterTrainer(params, copy.deepcopy(g), table)
    cur = table.copy()
    for k in range(3):
        idx = rs.choice(n, bs, replace=False)
        (wl, wm), wz, wgz = ot.step(data, idx)
        (gl, gm), gzl, ggz, cur = m.iter_step(cur, data, idx)
        np.testing.assert_allclose([gl, gm, gzl], [wl, wm, wz], rtol=3e-4, atol=1e-5)
        np.testing.assert_allclose(ggz, wgz, rtol=2e-3, atol=2e-5 * max(1.0, np.abs(wgz).max()))
        # the Adam step on a fresh variable is ~ lr_t * sign(g): rows agree unless a gradient is ~0
        np.testing.assert_allclose(cur, ot.data_z, rtol=0, atol=2.5e-3)
        assert np.median(np.abs(cur - ot.data_z)) < 1e-6
        untouched = np.setdiff1d(np.arange(n), idx)
        np.testing.assert_array_equal(cur[untouched], table[untouched] if k == 0 else prev[untouched])
        prev = cur.copy()
    w = m.get_weights()
    np.testing.assert_allclose(w['g'][2], ot.g['bn']['mean'], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(w['g'][3], ot.g['bn']['var'], rtol=1e-4, atol=1e-5)
    want_g = train_bgm.g_flat_params(ot.g)
    got_g = w['g'][:2] + w['g'][4:]
    for a, b in zip(got_g, want_g):
        assert np.abs(a - b).max() <= 3 * 1e-3 * 1.01      # k * lr_theta bound
        assert np.median(np.abs(a - b)) <= 2e-5


def test_bgm_fit_end_to_end_tracks_the_oracle_loop():
    """BGM.fit(use_egm_init=False): the reference's NumPy streams (N(0,1) table, epoch permutations),
    two epochs of mini-batches with the incomplete last batch skipped, evaluate(use_x_sd=False)."""
    import copy
    params = bgm_params(12, 3, lr_theta=1e-3, lr_z=1e-3)
    g = bgm_oracle_net(params, seed=31)
    m = bgm_product_model(params, g)
    n, bs = 70, 16
    data = np.random.RandomState(4).standard_normal((n, 12)).astype(np.float32)
    np.random.seed(11)
    m.fit(data, batch_size=bs, epochs=1, epochs_per_eval=1, use_egm_init=False, verbose=0)
    np.random.seed(11)
    z0 = np.random.normal(0, 1, size=(n, 3)).astype('float32')
    ot = train_bgm.BgmIterTrainer(params, copy.deepcopy(g), z0)
    for _ in range(2):
        ot.epoch(data, bs)
    got_z = m.data_z.cpu().numpy()
    assert np.median(np.abs(got_z - ot.data_z)) < 1e-5
    assert np.abs(got_z - ot.data_z).max() < 2 * 8 * 1e-3      # epochs * batches * lr_z bound
    assert len(m.history_loss) == 2 and all(np.isfinite(m.history_loss))
    assert abs(m.evaluate(data, m.data_z, use_x_sd=False) - ot.mse(data)) < 1e-3 * max(1.0, ot.mse(data))
    # use_x_sd=True adds sigma^2 on average
    assert m.evaluate(data, m.data_z, use_x_sd=True) > m.evaluate(data, m.data_z, use_x_sd=False)


def test_save_model_and_save_res_paths(tmp_path):
    """save_model / save_res of the params dict: weights and results are written where the reference
    writes its checkpoints / result files, and load_weights restores a model that predicts identically."""
    import os
    # BGM
    params = bgm_params(8, 2, lr=1e-3, lr_theta=1e-3, lr_z=1e-3, output_dir=str(tmp_path), save_model=True, save_res=True,
                        dataset='bgm_case')       # the timestamped directories have 1 s resolution: keep the two models apart
    g = bgm_oracle_net(params, seed=31)
    m = bgm_product_model(params, g)
    data = np.random.RandomState(3).standard_normal((64, 8)).astype(np.float32)
    np.random.seed(2)
    m.fit(data, batch_size=16, epochs=1, epochs_per_eval=1, use_egm_init=True, egm_n_iter=2, egm_batches_per_eval=2,
          verbose=0)
    assert os.path.exists(m.checkpoint_path + "/weights_at_1.npz")
    assert os.path.exists(m.checkpoint_path + "/weights_at_egm_init_2.npz")
    assert os.path.exists(m.save_dir + "/data_gen_at_1.npz") and os.path.exists(m.save_dir + "/params.txt")
    m2 = bgm_product_model(params, bgm_oracle_net(params, seed=99))
    m2.load_weights(m.checkpoint_path + "/weights_at_1.npz")
    z = np.random.RandomState(1).standard_normal((1, 10, 2)).astype(np.float32)
    zero = np.zeros((1, 10, 8), np.float32)
    np.testing.assert_array_equal(m.predict_on_posteriors(z, noise=zero), m2.predict_on_posteriors(z, noise=zero))
    # CausalBGM
    cp = causal_params(12, [1, 1, 1, 1], output_dir=str(tmp_path), save_model=True, save_res=True, dataset='causal_case')
    cm = product_model(cp, causal_nets(cp))
    cdata = causal_data(48, 12)
    np.random.seed(3)
    cm.fit(cdata, epochs=1, epochs_per_eval=1, batch_size=16, use_egm_init=True, egm_n_iter=2, egm_batches_per_eval=2,
           verbose=0)
    saved = [f for f in os.listdir(cm.checkpoint_path) if f.startswith("weights_at_")]
    assert saved and os.path.exists(cm.save_dir + "/causal_pre_egm_init_iter-2.txt")
    cm2 = product_model(cp, causal_nets(cp, seed=5))
    cm2.load_weights(os.path.join(cm.checkpoint_path, saved[0]))
    zz = np.random.RandomState(2).standard_normal((48, 4)).astype(np.float32)
    if saved[0] == "weights_at_%d.npz" % cm.best_epoch and cm.best_epoch == 1:
        np.testing.assert_array_equal(cm.get_log_posterior(*cdata, zz), cm2.get_log_posterior(*cdata, zz))
