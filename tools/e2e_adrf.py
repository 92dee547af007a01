#!/usr/bin/env python
"""End-to-end anchor on numbers the REFERENCE holds: the tutorial run of
docs/source/causalbgm/tutorial_py.ipynb (Sim_Hirano_Imbens, N=20000, v_dim=200, shipped YAML,
fit(epochs=100, egm_n_iter=30000) -> predict(n_mcmc=3000, burn_in=5000, x_values=linspace(0,3,20),
q_sd=1.0, bs=20000)) scored against the closed-form ADRF x + 2/(1+x)^3
(src/bayesgm/utils/helpers.py:59-60).  The tutorial prints RMSE 0.0188 / MAPE 0.0103
(:679-680) and an MH acceptance rate of 0.0948 (:651).

    python tools/e2e_adrf.py [--epochs 100] [--egm 30000] [--bnn 0|1] [--out gpurun_out/e2e_adrf.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def shipped_params(use_bnn):
    """src/configs/Sim_Hirano_Imbens.yaml, verbatim except use_bnn / output switches."""
    return dict(dataset='Sim_Hirano_Imbens', output_dir='/tmp/bgm_b200_e2e', save_res=False, save_model=False,
                binary_treatment=False, use_bnn=bool(use_bnn), z_dims=[1, 1, 1, 7], v_dim=200, lr_theta=0.0001,
                lr_z=0.0001, g_units=[64] * 5, f_units=[64, 32, 8], h_units=[64, 32, 8], kl_weight=0.0001,
                lr=0.0002, g_d_freq=5, use_z_rec=True, e_units=[64] * 5, dz_units=[64, 32, 8])


def true_adrf(x):
    return x + 2.0 / (1.0 + x) ** 3                      # utils/helpers.py:59-60


def run(epochs=100, egm=30000, use_bnn=False, n=20000, n_mcmc=3000, burn_in=5000, seed=123, verbose=1,
        epochs_per_eval=10):
    from bayesgm_b200 import CausalBGM
    from bayesgm_b200.datasets import Sim_Hirano_Imbens_sampler
    x, y, v = Sim_Hirano_Imbens_sampler(N=n, v_dim=200).load_all()
    model = CausalBGM(params=shipped_params(use_bnn), random_seed=seed)
    t0 = time.perf_counter()
    model.fit(data=(x, y, v), epochs=epochs, epochs_per_eval=epochs_per_eval, use_egm_init=egm > 0, egm_n_iter=egm,
              egm_batches_per_eval=500, verbose=verbose)
    t_fit = time.perf_counter() - t0
    xs = np.linspace(0, 3, 20)
    t0 = time.perf_counter()
    adrf, interval = model.predict(data=(x, y, v), alpha=0.01, n_mcmc=n_mcmc, burn_in=burn_in, x_values=xs,
                                   q_sd=1.0, bs=n, verbose=verbose)
    t_pred = time.perf_counter() - t0
    truth = true_adrf(xs)
    rmse = float(np.sqrt(np.mean((adrf - truth) ** 2)))
    mape = float(np.mean(np.abs((adrf - truth) / truth)))
    cover = float(np.mean((interval[:, 0] <= truth) & (truth <= interval[:, 1])))
    causal_pre, mse_x, mse_y, mse_v = model.evaluate(data=(x, y, v), data_z=model.data_z)
    return dict(n=n, epochs=epochs, egm_n_iter=egm, use_bnn=bool(use_bnn), n_mcmc=n_mcmc, burn_in=burn_in,
                rmse=rmse, mape=mape, interval_coverage=cover, acceptance_rate=getattr(model, 'last_acceptance_rate', None),
                mse_x=float(mse_x), mse_y=float(mse_y), mse_v=float(mse_v), fit_seconds=t_fit, predict_seconds=t_pred,
                adrf=[float(a) for a in adrf], truth=[float(a) for a in truth],
                interval=[[float(a), float(b)] for a, b in interval],
                best_epoch=getattr(model, 'best_epoch', None),
                tutorial=dict(rmse=0.0188, mape=0.0103, acceptance_rate=0.0948, mse_x=2.0460, mse_y=1.1746, mse_v=0.9638,
                              source="docs/source/causalbgm/tutorial_py.ipynb:604,651,679-680"))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=100)
    ap.add_argument("--egm", type=int, default=30000)
    ap.add_argument("--bnn", type=int, default=0)
    ap.add_argument("--n", type=int, default=20000)
    ap.add_argument("--n_mcmc", type=int, default=3000)
    ap.add_argument("--burn_in", type=int, default=5000)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    r = run(a.epochs, a.egm, a.bnn, a.n, a.n_mcmc, a.burn_in)
    print(json.dumps(r))
    if a.out:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        with open(a.out, "w") as f:
            json.dump(r, f, indent=1)
