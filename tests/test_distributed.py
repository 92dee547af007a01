"""world_size-2 gloo tests (CPU) of the host logic of the sharded path: row partition,
global-row-keyed noise, and the single end-of-run all-reduce of the ADRF partial sums.
The device work is stood in for by the oracle (the product has no CPU path)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_rows_partitions_exactly():
    from bayesgm_b200.shard import shard_rows
    for n in (0, 1, 7, 100000, 100003):
        for world in (1, 2, 3, 8):
            blocks = [shard_rows(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bayesgm_b200.shard import shard_rows, merge_adrf, finish_adrf, all_reduce_sum
    from oracle import causal
    from helpers import causal_params, causal_nets, causal_data, philox_normal4
    params = causal_params(12, [1, 1, 1, 2])
    nets = causal_nets(params)
    n, T, burn = 101, 6, 2
    x, y, v = causal_data(n, 12)
    xv = np.array([0.5, 2.0])

    def run(lo, hi):   # noise keyed by GLOBAL row, like the kernel's Philox streams
        rows = np.arange(lo, hi)
        z0 = np.concatenate([philox_normal4(3, rows, 0xFFFFFFFF, 0, g) for g in range(2)], 1)[:, :5]
        eps = np.stack([np.concatenate([philox_normal4(3, rows, t, 0, g) for g in range(2)], 1)[:, :5]
                        for t in range(T)])
        u = np.random.RandomState(0).uniform(size=(T, n))[:, lo:hi]
        zs = causal.mh_sampler(params, nets, (x[lo:hi], y[lo:hi], v[lo:hi]), q_sd=0.5, burn_in=burn,
                               n_keep=T - burn, noise=causal.InjectedNoise(z0, eps, u))
        be = causal.infer_from_latent_posterior(params, nets, zs, xv, sample_y=False)
        return torch.from_numpy(be.astype(np.float64)) * (hi - lo)      # per-(dose,sample) sums
    lo, hi = shard_rows(n, rank, world)
    ce = merge_adrf(run(lo, hi), hi - lo)
    adrf, interval = finish_adrf(ce, 0.1)
    # scalar all-reduce used by the adaptive rules
    cnt = all_reduce_sum(torch.tensor([float(hi - lo)], dtype=torch.float64))
    if rank == 0:
        full = (run(0, n) / n).float().numpy()
        want, want_int = finish_adrf(full, 0.1)
        np.savez(out, adrf=adrf, want=want, interval=interval, want_int=want_int, cnt=cnt.numpy())
    dist.destroy_process_group()


def test_sharded_adrf_equals_single_process(tmp_path):
    out = str(tmp_path / "r.npz")
    port = 29500 + (os.getpid() % 1000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    r = np.load(out)
    np.testing.assert_allclose(r['adrf'], r['want'], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(r['interval'], r['want_int'], rtol=1e-6, atol=1e-6)
    assert r['cnt'][0] == 101
