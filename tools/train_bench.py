"""Training throughput at the tutorial's size (Sim_Hirano_Imbens N=20000, shipped YAML; reference level: ~55
mini-batches/s in the iterative phase, docs/source/causalbgm/tutorial_py.ipynb:372): EGM iterations/s (5 disc + 1
gen steps) and iterative-phase mini-batches/s (update_g/h/f + latent step), for the fused engine (use_bnn=False),
the layered engine on deterministic nets, and the layered engine on the shipped Bayesian nets.  One JSON line each."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import torch
from e2e_adrf import shipped_params
from bayesgm_b200 import CausalBGM
from bayesgm_b200.datasets import Sim_Hirano_Imbens_sampler

x, y, v = Sim_Hirano_Imbens_sampler(N=20000, v_dim=200).load_all()
xd, yd, vd = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (x, y, v)]
for name, bnn, layered, bs in (("fused single-CTA kernels, use_bnn=False", False, False, 32),
                               ("layered engine, use_bnn=False", False, True, 32),
                               ("layered engine, use_bnn=True (shipped)", True, True, 32),
                               ("layered engine, use_bnn=True, batch 256 (iterative phase)", True, True, 256)):
    m = CausalBGM(params=shipped_params(bnn), random_seed=1)
    m._set_layered(layered)
    out = dict(config=name, n=20000, batch_size=bs)
    if bs <= 32:
        m.egm_init((xd, yd, vd), egm_n_iter=99, batch_size=bs, egm_batches_per_eval=10 ** 9, verbose=0, eval_during=False)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        iters = 1000
        m.egm_init((xd, yd, vd), egm_n_iter=iters - 1, batch_size=bs, egm_batches_per_eval=10 ** 9, verbose=0, eval_during=False)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        out.update(egm_iterations_per_s=iters / dt, egm_mini_batches_per_s=6 * iters / dt)
    # iterative phase: one epoch = ceil(n / bs) mini-batches (+ one evaluate)
    t0 = time.perf_counter()
    m.fit((xd, yd, vd), epochs=1, epochs_per_eval=10 ** 9, batch_size=bs, use_egm_init=False, verbose=0)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    nb = 2 * int(np.ceil(20000 / bs))
    out.update(iterative_mini_batches_per_s=nb / dt, iterative_rows_per_s=2 * 20000 / dt,
               reference_level="~55 mini-batches/s of 32 rows (tutorial tqdm, hardware unstated)")
    print(json.dumps(out), flush=True)
