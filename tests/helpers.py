"""Shared builders for the parity tests: one synthetic case = params dict + oracle
nets (Keras-layout arrays) + data, and the same weights loaded into the product model."""
import numpy as np

from oracle import nets as onets


def causal_params(v_dim, z_dims, binary=False, g_units=(64,) * 5, f_units=(64, 32, 8),
                  h_units=(64, 32, 8), e_units=(64,) * 5, **extra):
    p = dict(dataset='test', output_dir='/tmp/bgm_b200_test', save_res=False, save_model=False,
             binary_treatment=binary, use_bnn=False, z_dims=list(z_dims), v_dim=v_dim,
             lr_theta=1e-4, lr_z=1e-4, g_units=list(g_units), f_units=list(f_units),
             h_units=list(h_units), e_units=list(e_units), dz_units=[64, 32, 8], lr=2e-4,
             g_d_freq=5, use_z_rec=True, kl_weight=1e-4)
    p.update(extra)
    return p


def causal_nets(params, seed=123, bias_scale=0.1):
    rs = np.random.RandomState(seed)
    zd = sum(params['z_dims'])
    d0, d1, d2, _ = params['z_dims']
    return dict(
        g=onets.init_mlp(rs, [zd] + list(params['g_units']) + [params['v_dim'] + 1], bias_scale),
        e=onets.init_mlp(rs, [params['v_dim']] + list(params['e_units']) + [zd], bias_scale),
        f=onets.init_mlp(rs, [d0 + d1 + 1] + list(params['f_units']) + [2], bias_scale),
        h=onets.init_mlp(rs, [d0 + d2] + list(params['h_units']) + [2], bias_scale))


def causal_data(n, v_dim, binary=False, seed=0):
    rs = np.random.RandomState(seed)
    v = rs.standard_normal((n, v_dim)).astype(np.float32)
    if binary:
        x = (rs.uniform(size=(n, 1)) < 0.5).astype(np.float32)
    else:
        x = rs.exponential(size=(n, 1)).astype(np.float32)
    y = (x + 0.5 * v[:, :1] + rs.standard_normal((n, 1))).astype(np.float32)
    return x, y, v


def product_model(params, nets, engine=None):
    """engine: None / 'auto', 'simt' or 'tensor' (CausalBGM.set_sampler_engine)."""
    from bayesgm_b200 import CausalBGM
    m = CausalBGM(params=params, random_seed=None)
    if engine:
        m.set_sampler_engine(engine)
    flat = lambda layers: [a for W, b in layers for a in (W, b)]
    m.set_weights(g=flat(nets['g']), e=flat(nets['e']), f=flat(nets['f']), h=flat(nets['h']))
    return m


def injected_noise(n, zd, T, seed=7):
    rs = np.random.RandomState(seed)
    return dict(z0=rs.standard_normal((n, zd)).astype(np.float32),
                eps=rs.standard_normal((T, n, zd)).astype(np.float32),
                u=rs.uniform(size=(T, n)))


from oracle.philox import philox4x32_10, philox_normal4, M0, M1, W0, W1  # noqa: E402,F401


def bgm_params(x_dim, z_dim, g_units=(64,) * 5, **extra):
    p = dict(dataset='test', output_dir='/tmp/bgm_b200_test', save_res=False, save_model=False,
             use_bnn=False, x_dim=x_dim, z_dim=z_dim, g_units=list(g_units), e_units=[64] * 5,
             dz_units=[64, 32, 8], dx_units=[64, 32, 8], lr=1e-3, lr_theta=5e-3, lr_z=5e-3, gamma=0.0,
             alpha=0.0, g_d_freq=1, kl_weight=5e-5)
    p.update(extra)
    return p


def bgm_oracle_net(params, seed=21, bn_random=True):
    rs = np.random.RandomState(seed)
    return onets.init_variational(rs, params['z_dim'], params['x_dim'], list(params['g_units']),
                                  bias_scale=0.1, bn_random=bn_random)


def bgm_product_model(params, p):
    from bayesgm_b200 import BGM
    m = BGM(params=params, random_seed=None)
    w = [p['bn']['gamma'], p['bn']['beta'], p['bn']['mean'], p['bn']['var']]
    for W, b in p['hidden'] + [p['mean'], p['var']]:
        w += [W, b]
    m.set_weights(g=w)
    return m


def hmc_noise(n, zd, T, seed=9):
    rs = np.random.RandomState(seed)
    return dict(z0=rs.standard_normal((n, zd)).astype(np.float32),
                momentum=rs.standard_normal((T, n, zd)).astype(np.float32),
                log_u=np.log(rs.uniform(size=(T, n))).astype(np.float32))
