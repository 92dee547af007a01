"""Generates the committed golden vectors under tests/golden/.

Two sources:
  * `reference_datasets.npz` -- outputs of the REFERENCE ITSELF (the part of it that
    imports without TensorFlow: bayesgm.datasets, NumPy legacy RNG), so it needs
    /root/reference; run in the build container only.
  * `causal_*.npz`, `bgm_*.npz` -- outputs of the CPU oracle (oracle/) on seeded inputs.
    The reference's model code cannot run here (TF 2.10 / TFP 0.18 absent) and its tests
    hold no golden vector for this path, so these pin the oracle against drift and give
    the GPU tests a fixture that does not depend on the oracle's code at run time.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import causal, bgm, nets as onets  # noqa: E402
from helpers import causal_params, causal_nets, causal_data, injected_noise  # noqa: E402


def reference_datasets():
    ref = "/root/reference/src"
    if not os.path.isdir(ref):
        print("skip reference_datasets.npz: /root/reference absent")
        return
    sys.path.insert(0, ref)
    from bayesgm.datasets import Sim_Hirano_Imbens_sampler, Gaussian_sampler
    s = Sim_Hirano_Imbens_sampler(N=1000, v_dim=10, seed=0)
    x, y, v = s.load_all()
    b1 = s.next_batch()
    g = Gaussian_sampler(mean=np.zeros(5), sd=1.0)
    gb = g.get_batch(4)
    idx = np.random.choice(1000, 32, replace=False)      # the egm_init index draw (:406) after reseed
    np.savez_compressed(os.path.join(HERE, "reference_datasets.npz"),
                        x=x[:64], y=y[:64], v=v[:64], x_sum=np.float64(x.sum(dtype=np.float64)),
                        v_sum=np.float64(v.sum(dtype=np.float64)), batch_x=b1[0], batch_v=b1[2],
                        gauss_head=g.load_all()[:8], gauss_batch=gb, choice_idx=idx)
    sys.path.remove(ref)
    for k in [k for k in sys.modules if k.startswith("bayesgm")]:
        del sys.modules[k]


def causal_case(name, n, v_dim, z_dims, binary, burn_in, n_keep, q_sd, **extra):
    params = causal_params(v_dim, z_dims, binary, **extra)
    nets = causal_nets(params)
    data = causal_data(n, v_dim, binary)
    zd = sum(z_dims)
    z = np.random.RandomState(11).standard_normal((n, zd)).astype(np.float32)
    lp = causal.log_posterior(params, nets, *data, z)
    nz = injected_noise(n, zd, burn_in + n_keep)
    s, tr = causal.mh_sampler(params, nets, data, q_sd=q_sd, burn_in=burn_in, n_keep=n_keep,
                              noise=causal.InjectedNoise(**nz), return_trace=True)
    xv = None if binary else np.array([0.0, 1.0, 2.5])
    eff = causal.infer_from_latent_posterior(params, nets, s, xv, sample_y=False)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), n=n, v_dim=v_dim, z_dims=np.array(z_dims),
                        binary=binary, burn_in=burn_in, n_keep=n_keep, q_sd=q_sd, z=z, logp=lp, samples=s,
                        accept=np.array(tr['accept']), lp_prop=np.array(tr['lp_prop']),
                        lp_cur=np.array(tr['lp_cur']), effect=eff,
                        x_values=np.zeros(0) if xv is None else xv)


def bgm_case(name, n, x_dim, z_dim, units, miss, burn_in, n_mcmc, L, step):
    rs = np.random.RandomState(21)
    p = onets.init_variational(rs, z_dim, x_dim, units, bias_scale=0.1, bn_random=True)
    x = rs.standard_normal((n, x_dim)).astype(np.float32)
    w = (rs.uniform(size=(n, x_dim)) >= miss).astype(np.float32)
    w[:, 0] = 1.0
    T = burn_in + n_mcmc
    z0 = rs.standard_normal((n, z_dim)).astype(np.float32)
    mom = rs.standard_normal((T, n, z_dim)).astype(np.float32)
    log_u = np.log(rs.uniform(size=(T, n))).astype(np.float32)
    lp, g = bgm.log_posterior_and_grad(p, z0, x, w)
    s, tr = bgm.hmc_sampler(p, x, w, z0=z0, n_mcmc=n_mcmc, burn_in=burn_in, step_size=step,
                            num_leapfrog_steps=L, momentum=mom, log_u=log_u, return_trace=True)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), n=n, x_dim=x_dim, z_dim=z_dim,
                        units=np.array(units), burn_in=burn_in, n_mcmc=n_mcmc, L=L, step=step, x=x, w=w,
                        z0=z0, logp=lp, grad=g, samples=s, accept=np.array(tr['accept']),
                        steps=np.array(tr['step'], np.float32), log_accept=np.array(tr['log_accept']))


if __name__ == "__main__":
    reference_datasets()
    causal_case("causal_cont_zd5", 48, 200, [1, 1, 1, 2], False, 6, 10, 0.3)
    causal_case("causal_binary_zd18", 40, 100, [3, 6, 3, 6], True, 6, 10, 0.3)
    bgm_case("bgm_x10_z3", 40, 10, 3, [64] * 5, 0.0, 8, 6, 10, 0.01)
    bgm_case("bgm_x70_z10_mcar", 36, 70, 10, [64] * 5, 0.3, 8, 6, 5, 0.02)
    print(sorted(os.listdir(HERE)))
