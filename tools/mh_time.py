"""Sampler timing for other shapes than the bench workload (not the driver's bench).
usage: python tools/mh_time.py --z_dims 1 1 1 7 --v_dim 200 --n 100000 [--binary]"""
import argparse, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import causal_params, causal_nets, causal_data, product_model

ap = argparse.ArgumentParser()
ap.add_argument("--z_dims", type=int, nargs=4, default=[1, 1, 1, 7])
ap.add_argument("--v_dim", type=int, default=200)
ap.add_argument("--n", type=int, default=100000)
ap.add_argument("--T", type=int, default=200)
ap.add_argument("--binary", action="store_true")
a = ap.parse_args()
params = causal_params(a.v_dim, a.z_dims, a.binary)
nets = causal_nets(params)
data = causal_data(a.n, a.v_dim, a.binary)
for engine in ("simt", "tensor"):
    m = product_model(params, nets, engine)
    _, x, y, v, ldv, n = m._stage(data)
    aux = m._aux(v, ldv, n)
    for rep in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = m._mh_device(x, y, v, ldv, n, a.T // 2, a.T // 2, 1.0, False, 1.0, 0.25, 0.05, 50, 100, seed=rep, row_offset=0, aux=aux)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    print("%s %s: %.2f ms for T=%d -> %.3g samples/s (accept %.3f)" % (
        engine, m.sampler_info()['kernel'], ms, a.T, n * a.T / (ms * 1e-3), float(r['accept_count'].sum()) / (a.T * n)))
