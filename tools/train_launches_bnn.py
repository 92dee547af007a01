"""EGM iterations and iterative-phase mini-batches of the shipped Bayesian-net model only (for an ncu launch
list: sum of kernel time per step against the wall-clock step time of tools/train_bench.py)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import torch
from e2e_adrf import shipped_params
from bayesgm_b200 import CausalBGM
from bayesgm_b200.datasets import Sim_Hirano_Imbens_sampler
x, y, v = Sim_Hirano_Imbens_sampler(N=2000, v_dim=200).load_all()
m = CausalBGM(params=shipped_params(True), random_seed=1)
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 5
m.egm_init((x, y, v), egm_n_iter=1, batch_size=32, egm_batches_per_eval=10 ** 9, verbose=0, eval_during=False)
torch.cuda.synchronize()
t0 = time.perf_counter()
m.egm_init((x, y, v), egm_n_iter=iters - 1, batch_size=32, egm_batches_per_eval=10 ** 9, verbose=0, eval_during=False)
torch.cuda.synchronize()
print("EGM iteration wall ms", (time.perf_counter() - t0) * 1e3 / iters)
