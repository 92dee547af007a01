"""Host-side sharding helpers for one-process-per-GPU runs (torch.distributed).

The posterior samplers shard by observation (rows): every rank owns a contiguous block
of rows, Philox noise is keyed by the GLOBAL row index so the union of the shards equals
the single-GPU run, and the only collectives are (a) one all-reduce of the ADRF partial
sums at the end of `CausalBGM.predict` (causalbgm/base.py:660-663 combines `bs` slices
the same way: row-count-weighted mean), and (b) one scalar all-reduce per adaptation
event (MH q_sd window counts / HMC step-size statistic).  Works on any backend: NCCL on
the GPUs, gloo in the CPU tests.
"""
import numpy as np


def shard_rows(n, rank, world):
    """Contiguous, balanced row block of `rank`: the first n % world ranks get one extra row."""
    base, extra = divmod(int(n), int(world))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def all_reduce_sum(t, group=None):
    """In-place sum over ranks of a torch tensor (no-op without an initialised group)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and (group is not None or dist.get_world_size() > 1):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def merge_adrf(sums, n_rows, group=None):
    """sums: (n_x, n_keep) float64 tensor of per-(dose, sample) sums over this rank's rows;
    returns the global per-(dose, sample) means as float32 NumPy, like :663."""
    import torch
    cnt = torch.tensor([float(n_rows)], dtype=torch.float64, device=sums.device)
    all_reduce_sum(sums, group)
    all_reduce_sum(cnt, group)
    return (sums / cnt).float().cpu().numpy()


def finish_adrf(ce, alpha):
    """causalbgm/base.py:665-667: point estimate and (1-alpha) interval over the kept samples."""
    adrf = np.mean(ce, axis=1)
    up = np.quantile(ce, 1 - alpha / 2, axis=1)
    lo = np.quantile(ce, alpha / 2, axis=1)
    return adrf, np.stack([lo, up], axis=1)
