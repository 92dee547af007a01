"""Host-side profile (cProfile) of BGM.predict at the cfg-5 per-GPU workload (62500 rows, x_dim 500, 30 % MCAR)."""
import cProfile, pstats, io, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayesgm_b200 import BGM
from bayesgm_b200.datasets import simulate_z_hetero
n = int(sys.argv[1]) if len(sys.argv) > 1 else 62500
X, Y = simulate_z_hetero(n=n, k=10, d=499, seed=42)
data = np.c_[X, Y].astype(np.float32)
data[np.random.RandomState(1).rand(*data.shape) < 0.3] = np.nan
P = dict(dataset='cfg5', output_dir='/tmp/bgm_b200_bench', save_res=False, save_model=False, use_bnn=False, x_dim=500,
         z_dim=10, g_units=[64] * 5, e_units=[64] * 5, dz_units=[64, 32, 8], dx_units=[64, 32, 8], lr=1e-3, lr_theta=5e-3,
         lr_z=5e-3, gamma=0.0, alpha=0.0, g_d_freq=1, kl_weight=5e-5)
m = BGM(params=P, random_seed=123)
kw = dict(alpha=0.05, bs=1000, n_mcmc=100, burn_in=100, step_size=0.01, num_leapfrog_steps=10, verbose=0)
m.predict(data, seed=1, **kw)
torch.cuda.synchronize(); t0 = time.perf_counter()
m.predict(data, seed=2, **kw)
print("predict ms", (time.perf_counter() - t0) * 1e3)
pr = cProfile.Profile(); pr.enable()
m.predict(data, seed=3, **kw)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(18); print(s.getvalue()[:4000])
