"""Tiny run of the sampler (tensor engine, z_dim = 5: causal_mh_tc16_kernel<8, true>) and of predict() for
`compute-sanitizer --tool memcheck python tools/sanitize_l1.py`."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import causal_params, causal_nets, causal_data, product_model
params = causal_params(200, [1, 1, 1, 2])
m = product_model(params, causal_nets(params), 'tensor')
print(m.sampler_info()['kernel'])
data = causal_data(300, 200)
s = m.metropolis_hastings_sampler(data, q_sd=0.5, burn_in=8, n_keep=8, seed=3, verbose=0)
out = m.predict(data, alpha=0.05, n_mcmc=8, burn_in=8, x_values=np.linspace(0, 3, 4), q_sd=1.0, sample_y=True, bs=300, verbose=0)
print("ok", s.shape, np.isfinite(s).all())
