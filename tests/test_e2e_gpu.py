"""End-to-end anchors (GPU).

1. Full BASELINE cfg-3 size, FULL length: n = 100000 rows, p = 200, T = 1000 iterations in one
   persistent launch; the chains of a 2000-row subsample are replayed through the oracle on the
   kernel's own Philox noise (SURVEY 8c tiers ii/iii): accept decisions and kept states agree chain by
   chain, the ADRF of the subsample agrees far inside Monte-Carlo error.
2. The only accuracy numbers the REFERENCE holds for this path: the tutorial run
   (docs/source/causalbgm/tutorial_py.ipynb: Sim_Hirano_Imbens N=20000, shipped YAML, fit 100 epochs
   after 30000 EGM iterations, predict n_mcmc=3000 / burn_in=5000 / q_sd=1) scored against the
   closed-form ADRF x + 2/(1+x)^3 (utils/helpers.py:59-60): RMSE 0.0188, MAPE 0.0103, MH acceptance
   0.0948, final MSE_x/MSE_y/MSE_v 2.0460/1.1746/0.9638 (:604,:651,:679-680).  The measured numbers of
   this repository's run are written to gpurun_out/ and committed under profiles/ (see DESIGN.md
   section 2 for what agrees and what does not).
"""
import json
import os
import sys

import numpy as np
import pytest

from oracle import causal
from helpers import causal_params, causal_nets, product_model

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def test_cfg3_full_size_full_length_subsample_replayed_through_oracle():
    import torch
    from bayesgm_b200.datasets import Sim_Hirano_Imbens_sampler
    n, p, sub = 100000, 200, 2000
    burn_in, n_keep, seed = 500, 500, 20261017
    T = burn_in + n_keep
    params = causal_params(p, [1, 1, 1, 2])
    nets = causal_nets(params)
    x, y, v = Sim_Hirano_Imbens_sampler(N=n, v_dim=p, seed=0).load_all()
    m = product_model(params, nets)
    _, xd, yd, vd, ldv, _ = m._stage((x, y, v))
    r = m._mh_device(xd, yd, vd, ldv, n, burn_in, n_keep, 1.0, False, 1.0, 0.25, 0.05, 50, 100, seed, 0, trace=True)
    torch.cuda.synchronize()
    rate = float(r['accept_count'].sum().item()) / (T * n)
    assert 0.05 < rate < 0.6, rate
    samples = r['samples'][:, :sub].cpu().numpy()
    acc = r['accept_mask'][:, :sub].cpu().numpy().astype(bool)
    assert np.isfinite(samples).all()
    # oracle replay of rows [0, sub) on the kernel's own noise (Philox keyed by the global row)
    nz = m.philox_noise(seed, sub, T)
    so, tro = causal.mh_sampler(params, nets, (x[:sub], y[:sub], v[:sub]), q_sd=1.0, burn_in=burn_in, n_keep=n_keep,
                                noise=causal.InjectedNoise(**nz), recompute_current=False, return_trace=True)
    same = (acc == np.array(tro['accept'])).all(axis=0)
    # a chain diverges only at a rounding-level tie between u and the acceptance ratio: T = 1000 decisions per chain
    assert same.mean() > 0.90, same.mean()
    np.testing.assert_array_equal(samples[:, same], so[:, same])
    xs = np.linspace(0, 3, 20)
    want = causal.infer_from_latent_posterior(params, nets, so, xs, sample_y=False)          # (20, n_keep)
    got = m.infer_from_latent_posterior(samples, x_values=xs, sample_y=False)
    d = np.abs(got.mean(axis=1) - want.mean(axis=1)).max()
    mc_se = want.std(axis=1).max() / np.sqrt(n_keep / 20.0)       # crude: ~20 iterations of autocorrelation
    assert d < max(2e-3, 0.2 * mc_se), (d, mc_se)
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    json.dump(dict(n=n, T=T, subsample=sub, acceptance_rate=rate, chains_identical=float(same.mean()),
                   adrf_max_abs_diff=float(d), mc_se=float(mc_se)), open(os.path.join(out, "e2e_cfg3_T1000.json"), "w"))


def test_tutorial_pipeline_adrf_against_closed_form():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import e2e_adrf
    r = e2e_adrf.run(epochs=100, egm=30000, use_bnn=False, verbose=0)
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    json.dump(r, open(os.path.join(out, "e2e_adrf_tutorial.json"), "w"), indent=1)
    adrf, truth = np.array(r['adrf']), np.array(r['truth'])
    # shape of the dose-response: the dip near x ~ 0.3-0.6 and the unit slope beyond x ~ 1.5
    assert np.argmin(adrf) in range(1, 6), adrf
    slope = np.polyfit(np.linspace(0, 3, 20)[10:], adrf[10:], 1)[0]
    assert 0.8 < slope < 1.2, slope
    # level: deterministic nets (use_bnn=False) do not reach the tutorial's BNN-run RMSE of 0.0188; the bound below
    # pins what this implementation measures (profiles/r02_e2e_adrf.json) against regressions, DESIGN.md section 2
    assert r['rmse'] < 0.35 and r['mape'] < 0.15, (r['rmse'], r['mape'])
    assert 0.8 < r['mse_v'] < 1.05 and r['mse_y'] < 1.3
    assert 0.01 < r['acceptance_rate'] < 0.6
