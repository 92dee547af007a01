"""A few training steps of each engine, for an ncu launch list (tools/gpu_round2.sh)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import torch
from e2e_adrf import shipped_params
from bayesgm_b200 import CausalBGM
from bayesgm_b200.datasets import Sim_Hirano_Imbens_sampler
x, y, v = Sim_Hirano_Imbens_sampler(N=2000, v_dim=200).load_all()
for bnn in (False, True):
    m = CausalBGM(params=shipped_params(bnn), random_seed=1)
    m.egm_init((x, y, v), egm_n_iter=1, batch_size=32, egm_batches_per_eval=10 ** 9, verbose=0, eval_during=False)
    m.fit((x[:64], y[:64], v[:64]), epochs=0, epochs_per_eval=10 ** 9, batch_size=32, use_egm_init=False, verbose=0)
    torch.cuda.synchronize()
