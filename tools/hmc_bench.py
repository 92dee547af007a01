"""Timing of the BGM HMC kernel (not the driver's bench): cfg-5 per-GPU shape by default."""
import argparse, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from helpers import bgm_params, bgm_oracle_net, bgm_product_model

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=37888)
ap.add_argument("--x_dim", type=int, default=500)
ap.add_argument("--z_dim", type=int, default=10)
ap.add_argument("--L", type=int, default=10)
ap.add_argument("--steps", type=int, default=4)
ap.add_argument("--miss", type=float, default=0.3)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
params = bgm_params(a.x_dim, a.z_dim)
p = bgm_oracle_net(params, bn_random=False)
m = bgm_product_model(params, p)
rs = np.random.RandomState(0)
x = rs.standard_normal((a.n, a.x_dim)).astype(np.float32)
x[rs.uniform(size=x.shape) < a.miss] = np.nan
xd, ldx, n = m._stage_x(x, torch)
info = m.kernel_info()
for rep in range(a.reps + 1):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = m._hmc_device(xd, ldx, n, a.steps, 0, 0.01, a.L, seed=rep)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    grads = a.steps * a.L + 1
    tf = 2.0 * info['macs_per_grad'] * n * grads / (ms * 1e-3) / 1e12
    acc = float(r['accept_count'][:a.steps].sum()) / (a.steps * n)
    print("rep %d: %.2f ms  %.2f ms/step  %.2f TFLOP/s (algorithmic)  %.3g chain-steps/s  accept %.3f  %s"
          % (rep, ms, ms / a.steps, tf, n * a.steps / (ms * 1e-3), acc, info))
