"""CPU checks of the training oracle (oracle/train.py): Keras-Adam arithmetic, loss
bookkeeping, and the autograd gradient-penalty path against float64 finite differences."""
import numpy as np
import torch

from oracle import train, nets as onets
from helpers import causal_params, causal_nets


def small_case(binary=False, bs=8):
    params = causal_params(6, [1, 2, 1, 1], binary, g_units=[8, 8], e_units=[8, 8], f_units=[8, 4],
                           h_units=[8, 4], dz_units=[8, 4], lr=1e-2)
    nets = causal_nets(params, seed=3)
    rs = np.random.RandomState(4)
    dz = onets.init_discriminator(rs, 5, [8, 4])
    for bn in dz['bns']:
        bn['gamma'] = (1 + 0.1 * rs.standard_normal(bn['gamma'].shape)).astype(np.float32)
        bn['beta'] = (0.1 * rs.standard_normal(bn['beta'].shape)).astype(np.float32)
    z = rs.standard_normal((bs, 5)).astype(np.float32)
    v = rs.standard_normal((bs, 6)).astype(np.float32)
    x = (rs.uniform(size=(bs, 1)) < 0.5).astype(np.float32) if binary else rs.exponential(size=(bs, 1)).astype(np.float32)
    y = rs.standard_normal((bs, 1)).astype(np.float32)
    return params, nets, dz, z, v, x, y


def test_keras_adam_first_steps():
    opt = train.Adam(0.1, 0.9, 0.99)
    p = [np.array([1.0, -2.0], np.float32)]
    g = [np.array([0.5, -0.25], np.float32)]
    opt.apply(p, g)
    # t=1: m = .1 g, v = .01 g^2, lr_t = lr*sqrt(1-.99)/(1-.9) = lr -> step = lr * sign(g) (up to eps)
    np.testing.assert_allclose(p[0], [1.0 - 0.1, -2.0 + 0.1], rtol=1e-5)
    opt.apply(p, g)
    m = 0.1 * 0.5 * (1 + 0.9)
    v = 0.01 * 0.25 * (1 + 0.99)
    lr_t = 0.1 * np.sqrt(1 - 0.99 ** 2) / (1 - 0.9 ** 2)
    np.testing.assert_allclose(p[0][0], 0.9 - lr_t * m / (np.sqrt(v) + 1e-7), rtol=1e-5)


def test_gen_step_total_is_the_sum_of_its_parts():
    params, nets, dz, z, v, x, y = small_case()
    (e_adv, l2v, l2z, l2x, l2y, total), grads = train.gen_step(params, nets, dz, z, v, x, y)
    assert set(grads) == {'g', 'e', 'f', 'h'} and len(grads['g']) == 2 * len(nets['g'])
    g_out = onets.mlp_forward(nets['g'], z)
    zenc = onets.mlp_forward(nets['e'], v)
    f_out = onets.mlp_forward(nets['f'], np.concatenate([zenc[:, :1], zenc[:, 1:3], x], 1))
    h_out = onets.mlp_forward(nets['h'], np.concatenate([zenc[:, :1], zenc[:, 3:4]], 1))
    sig = (g_out[:, -1] ** 2).mean() + (f_out[:, -1] ** 2).mean() + (h_out[:, -1] ** 2).mean()
    np.testing.assert_allclose(total, e_adv + l2v + l2z + l2x + l2y + 0.001 * sig, rtol=1e-5)
    np.testing.assert_allclose(e_adv, -onets.discriminator_forward(dz, zenc).mean(), rtol=1e-5)


def test_disc_gradient_matches_float64_finite_differences():
    params, nets, dz, z, v, x, y = small_case()
    dz_loss, d_loss, grads = train.disc_step(params, nets, dz, z, v, 0.3)

    def loss64(dz64):
        e = [(torch.tensor(W, dtype=torch.float64), torch.tensor(b, dtype=torch.float64)) for W, b in nets['e']]
        pt = dict(layers=[(torch.tensor(W, dtype=torch.float64), torch.tensor(b, dtype=torch.float64))
                          for W, b in dz64['layers']],
                  bn=[(torch.tensor(bn['gamma'], dtype=torch.float64), torch.tensor(bn['beta'], dtype=torch.float64))
                      for bn in dz64['bns']])
        zt, vt = torch.tensor(z, dtype=torch.float64), torch.tensor(v, dtype=torch.float64)
        z_ = train.mlp(e, vt)
        zh = (zt * 0.3 + z_ * 0.7).requires_grad_(True)
        gz = torch.autograd.grad(train.disc(pt, zh).sum(), zh)[0]
        gp = ((torch.sqrt((gz ** 2).sum(1)) - 1) ** 2).mean()
        return float(-train.disc(pt, zt).mean() + train.disc(pt, z_).mean() + 10 * gp)
    import copy
    base = copy.deepcopy(dz)
    for key in ('layers', 'bns'):
        pass
    # perturb a few entries of each parameter kind
    flat = train.disc_flat_params(dz)
    checks = [(0, (1, 2)), (1, (3,)), (2, (0,)), (3, (5,)), (4, (2, 1)), (6, (1,)), (8, (3, 0))]
    h = 1e-5
    for idx, pos in checks:
        d1, d2 = copy.deepcopy(base), copy.deepcopy(base)
        train.disc_flat_params(d1)[idx][pos] += h
        train.disc_flat_params(d2)[idx][pos] -= h
        # deepcopy keeps float32 storage; rebuild as float64 perturbations through the loss
        for d, sgn in ((d1, +1), (d2, -1)):
            arr = train.disc_flat_params(d)[idx].astype(np.float64)
            arr[pos] = float(train.disc_flat_params(base)[idx][pos]) + sgn * h
            target = train.disc_flat_params(d)
            # write back as float64 array
            if idx in (0, 1, 4, 5, 8, 9):
                li = {0: 0, 1: 0, 4: 1, 5: 1, 8: 2, 9: 2}[idx]
                W, b = d['layers'][li]
                d['layers'][li] = (arr, b) if idx % 4 == 0 else (W, arr)
            else:
                bi = {2: 0, 3: 0, 6: 1, 7: 1}[idx]
                d['bns'][bi]['gamma' if idx % 4 == 2 else 'beta'] = arr
        fd = (loss64(d1) - loss64(d2)) / (2 * h)
        np.testing.assert_allclose(grads[idx][pos], fd, rtol=2e-3, atol=2e-5)
    assert d_loss >= dz_loss          # the penalty is non-negative


# ---- BGM iterative phase (oracle/train_bgm.py): the autograd gradients against float64 finite differences ----
def _bgm_loss64(g, z, x, with_prior):
    """float64 NumPy restatement of the losses of bgm/base.py:150-153 / :173-180, generator in training mode."""
    z = z.astype(np.float64)
    mu_b = z.mean(axis=0)
    var_b = ((z - mu_b) ** 2).mean(axis=0)
    h = (z - mu_b) / np.sqrt(var_b + 1e-3) * g['bn']['gamma'].astype(np.float64) + g['bn']['beta'].astype(np.float64)
    for W, b in g['hidden']:
        h = h @ W.astype(np.float64) + b.astype(np.float64)
        h = np.where(h > 0, h, 0.2 * h)
    mean = h @ g['mean'][0].astype(np.float64) + g['mean'][1].astype(np.float64)
    s2 = np.logaddexp(0, h @ g['var'][0].astype(np.float64) + g['var'][1].astype(np.float64)) + 1e-6
    loss = (((x - mean) ** 2) / (2 * s2) + 0.5 * np.log(s2)).sum(axis=1).mean()
    if with_prior:
        loss += ((z ** 2).sum(axis=1) / 2).mean()
    return loss


def test_bgm_iterative_oracle_gradients_match_finite_differences():
    from oracle import train_bgm
    rs = np.random.RandomState(5)
    g = onets.init_variational(rs, 3, 6, [8, 8], bias_scale=0.1, bn_random=True)
    z = rs.standard_normal((7, 3)).astype(np.float32)
    x = rs.standard_normal((7, 6)).astype(np.float32)
    # latent gradient, through the batch statistics of the input BatchNormalization
    loss, gz, _ = train_bgm.iter_latent_grad(g, z, x)
    assert abs(loss - _bgm_loss64(g, z, x, True)) < 1e-4 * max(1.0, abs(loss))
    h = 1e-4
    for (r, d) in [(0, 0), (3, 1), (6, 2)]:
        zp, zm = z.astype(np.float64).copy(), z.astype(np.float64).copy()
        zp[r, d] += h
        zm[r, d] -= h
        fd = (_bgm_loss64(g, zp, x, True) - _bgm_loss64(g, zm, x, True)) / (2 * h)
        assert abs(gz[r, d] - fd) < 2e-3 * max(1.0, abs(fd)), (r, d, gz[r, d], fd)
    # parameter gradients of update_g_net: BN gamma and one entry of the variance head
    loss_x, mse, grads, _ = train_bgm.iter_g_grads(g, z, x)
    assert abs(loss_x - _bgm_loss64(g, z, x, False)) < 1e-4 * max(1.0, abs(loss_x))
    import copy
    for which, idx in (("gamma", 1), ("var", (2, 4))):
        gp, gm = copy.deepcopy(g), copy.deepcopy(g)
        if which == "gamma":
            gp['bn']['gamma'] = gp['bn']['gamma'].astype(np.float64); gp['bn']['gamma'][idx] += h
            gm['bn']['gamma'] = gm['bn']['gamma'].astype(np.float64); gm['bn']['gamma'][idx] -= h
            got = grads[0][idx]
        else:
            Wp = gp['var'][0].astype(np.float64); Wp[idx] += h
            Wm = gm['var'][0].astype(np.float64); Wm[idx] -= h
            gp['var'] = (Wp, gp['var'][1]); gm['var'] = (Wm, gm['var'][1])
            got = grads[-2][idx]
        fd = (_bgm_loss64(gp, z, x, False) - _bgm_loss64(gm, z, x, False)) / (2 * h)
        assert abs(got - fd) < 2e-3 * max(1.0, abs(fd)), (which, got, fd)
