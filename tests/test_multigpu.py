"""The collectives of SURVEY section 8(e) on real NCCL (2 GPUs): ADRF partial-sum all-reduce in
`CausalBGM.predict(group=)`, gradient all-reduce in the EGM steps, the per-step scalar all-reduce of
the shared HMC step size.  Skipped on a 1-GPU box (run with `gpurun --gpus 2`); the host logic of
the same paths is covered on CPU/gloo by tests/test_distributed.py."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _need_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")


def _init(rank, world, port):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    return torch, dist


def _predict_worker(rank, world, port, out):
    torch, dist = _init(rank, world, port)
    from bayesgm_b200.shard import shard_rows
    from helpers import causal_params, causal_nets, causal_data, product_model
    params = causal_params(200, [1, 1, 1, 2])
    nets = causal_nets(params)
    n = 1001
    x, y, v = causal_data(n, 200)
    xs = np.array([0.3, 1.0, 2.2])
    lo, hi = shard_rows(n, rank, world)
    m = product_model(params, nets)
    kw = dict(alpha=0.1, n_mcmc=40, burn_in=20, x_values=xs, q_sd=0.5, sample_y=True, seed=77, verbose=0)
    adrf, interval = m.predict((x[lo:hi], y[lo:hi], v[lo:hi]), bs=10 ** 6, group=dist.group.WORLD, row_offset=lo, **kw)
    # the kept states of the shard are the shard's rows of the single-GPU run, bit for bit (Philox keyed by global row)
    r_sh = m._mh_device(*m._stage((x[lo:hi], y[lo:hi], v[lo:hi]))[1:], 20, 40, 0.5, False, 1.0, 0.25, 0.05, 50, 100, 77, lo)
    r_all = m._mh_device(*m._stage((x, y, v))[1:], 20, 40, 0.5, False, 1.0, 0.25, 0.05, 50, 100, 77, 0)
    same_states = bool(torch.equal(r_sh['samples'], r_all['samples'][:, lo:hi]))
    gathered = [None] * world
    dist.all_gather_object(gathered, (adrf, interval, same_states))
    if rank == 0:
        want, want_int = m.predict((x, y, v), bs=10 ** 6, **kw)
        np.savez(out, adrf0=gathered[0][0], adrf1=gathered[1][0], int0=gathered[0][1], want=want, want_int=want_int,
                 same=np.array([g[2] for g in gathered]))
    dist.destroy_process_group()


def test_sharded_predict_equals_single_gpu(tmp_path):
    _need_two_gpus()
    import torch.multiprocessing as mp
    out = str(tmp_path / "p.npz")
    mp.spawn(_predict_worker, args=(2, 29650 + os.getpid() % 300, out), nprocs=2, join=True)
    r = np.load(out)
    assert r['same'].all(), "kept states of a shard differ from the single-GPU run"
    np.testing.assert_array_equal(r['adrf0'], r['adrf1'])                     # every rank holds the reduced result
    np.testing.assert_allclose(r['adrf0'], r['want'], rtol=2e-6, atol=2e-6)   # float64 partial sums, other order
    np.testing.assert_allclose(r['int0'], r['want_int'], rtol=2e-6, atol=2e-6)


def _egm_worker(rank, world, port, out):
    torch, dist = _init(rank, world, port)
    import ctypes as C
    from bayesgm_b200 import _lib
    from helpers import causal_params, causal_nets, causal_data, product_model
    params = causal_params(200, [1, 1, 1, 2])
    nets = causal_nets(params)
    rs = np.random.RandomState(5)
    batches = []
    for r in range(world):
        x, y, v = causal_data(32, 200, seed=10 + r)
        batches.append((rs.standard_normal((32, 5)).astype(np.float32), v, x, y))
    m = product_model(params, nets)
    # the discriminator is initialised from NumPy's global generator: give every model the same one
    dz0 = [np.random.RandomState(9).standard_normal(a.shape).astype(np.float32) * 0.3 + (1.0 if i % 4 == 2 else 0.0)
           for i, a in enumerate(m.dz_net.trainable_list())]
    m.set_weights(dz=dz0)
    z, v, x, y = batches[rank]
    m.train_gen_step(z, v, x, y, group=dist.group.WORLD)
    m.train_disc_step(z, v, epsilon=0.3, group=dist.group.WORLD)
    w = m.get_weights()
    flat = np.concatenate([a.ravel() for k in ('g', 'e', 'f', 'h', 'dz') for a in w[k]])
    gathered = [None] * world
    dist.all_gather_object(gathered, flat)
    if rank == 0:
        # single process: average of the two gradients, then the same Adam step
        ref = product_model(params, nets)
        ref.set_weights(dz=dz0)
        tr = ref._device_trainer()
        for group_id, which in ((0, 'gen'), (1, 'disc')):
            tot = None
            for (bz, bv, bx, by) in batches:
                _, g = ref.gradients(which, bz, bv, bx, by, epsilon=0.3)
                tot = g.copy() if tot is None else tot + g
            ref._grad_tensor(group_id).copy_(torch.from_numpy(tot).cuda())
            _lib.call("bgm_train_adam", tr, group_id, 1.0 / world, _lib.stream_ptr())
            ref._trainer_dirty = True
        rw = ref.get_weights()
        want = np.concatenate([a.ravel() for k in ('g', 'e', 'f', 'h', 'dz') for a in rw[k]])
        start = np.concatenate([a.ravel() for k in 'gefh' for W, b in nets[k] for a in (W, b)])
        np.savez(out, r0=gathered[0], r1=gathered[1], want=want, moved=np.abs(gathered[0][:start.size] - start).max())
    dist.destroy_process_group()


def test_egm_gradients_are_averaged_over_ranks(tmp_path):
    _need_two_gpus()
    import torch.multiprocessing as mp
    out = str(tmp_path / "e.npz")
    mp.spawn(_egm_worker, args=(2, 29950 + os.getpid() % 300, out), nprocs=2, join=True)
    r = np.load(out)
    assert r['moved'] > 0
    np.testing.assert_array_equal(r['r0'], r['r1'])          # replicas stay identical
    np.testing.assert_array_equal(r['r0'], r['want'])        # = Adam on the mean gradient of the two batches


def _hmc_worker(rank, world, port, out):
    torch, dist = _init(rank, world, port)
    from bayesgm_b200.shard import shard_rows
    from helpers import bgm_params, bgm_oracle_net, bgm_product_model
    bp = bgm_params(70, 10)
    m = bgm_product_model(bp, bgm_oracle_net(bp))
    rs = np.random.RandomState(0)
    n = 777
    x = rs.standard_normal((n, 70)).astype(np.float32)
    x[rs.uniform(size=x.shape) < 0.3] = np.nan
    lo, hi = shard_rows(n, rank, world)
    xs, ldx, ns = m._stage_x(x[lo:hi], torch)
    r = m._hmc_device(xs, ldx, ns, 6, 20, 0.05, 5, seed=11, row_offset=lo, group=dist.group.WORLD, n_total=n, trace=True)
    xa, ldxa, na = m._stage_x(x, torch)
    ra = m._hmc_device(xa, ldxa, na, 6, 20, 0.05, 5, seed=11, trace=True)
    ok_states = bool(torch.equal(r['samples'], ra['samples'][:, lo:hi]))
    ok_steps = bool(torch.equal(r['step_trace'], ra['step_trace']))
    gathered = [None] * world
    dist.all_gather_object(gathered, (ok_states, ok_steps, float(r['step'][0]), float(ra['step'][0])))
    if rank == 0:
        np.savez(out, ok=np.array([[g[0], g[1]] for g in gathered]), steps=np.array([[g[2], g[3]] for g in gathered]))
    dist.destroy_process_group()


def test_hmc_shared_step_size_is_all_reduced(tmp_path):
    _need_two_gpus()
    import torch.multiprocessing as mp
    out = str(tmp_path / "h.npz")
    mp.spawn(_hmc_worker, args=(2, 30250 + os.getpid() % 300, out), nprocs=2, join=True)
    r = np.load(out)
    assert r['ok'].all(), (r['ok'], r['steps'])
    assert r['steps'][0, 0] != 0.05                          # the step size did adapt
