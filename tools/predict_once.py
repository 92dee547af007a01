"""Three predict() calls at the bench workload (for an ncu launch list)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from bayesgm_b200 import CausalBGM
x, y, v = bench.make_data(0)
m = CausalBGM(params=bench.params(), random_seed=123)
xh, yh, vh = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in (x, y, v)]
for i in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    m.predict((xh, yh, vh), alpha=0.01, n_mcmc=500, burn_in=500, x_values=bench.X_VALUES, q_sd=1.0, sample_y=True, bs=100000, seed=i, verbose=0)
    print("predict", (time.perf_counter() - t0) * 1e3, "ms")
