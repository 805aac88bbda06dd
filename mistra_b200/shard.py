"""Multi-GPU partitioning of an ensemble: whole columns are dealt
out in contiguous blocks, one block per rank (keeps a column's 148 cells
together for the host scatter); cells are independent within a chemistry step
(kpp.f90:4310-4470 reads layer-k data only), so the data path needs no
collective - only the per-call diagnostics are reduced."""
from __future__ import annotations

import numpy as np


def column_block(total_cols, world, rank):
    """(first column, number of columns) owned by `rank`; blocks differ by at most 1."""
    base, extra = divmod(total_cols, world)
    n = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, n


def diagnostics_vector(ierr, stats):
    """[cells, failed cells, sum Nstp, sum Nacc, sum Nrej, sum Nsng, max Nstp] as float64."""
    ierr = np.asarray(ierr)
    stats = np.asarray(stats).reshape(-1, 8)
    return np.array([ierr.size, int((ierr != 1).sum()), stats[:, 2].sum(), stats[:, 3].sum(),
                     stats[:, 4].sum(), stats[:, 7].sum(), stats[:, 2].max() if ierr.size else 0],
                    dtype=np.float64)


def reduce_diagnostics(dist, vec, device=None):
    """All-reduce the diagnostics vector over the process group (sum, last entry max)."""
    import torch
    t = torch.as_tensor(vec, dtype=torch.float64, device=device).clone()
    s = t.clone()
    dist.all_reduce(s, op=dist.ReduceOp.SUM)
    mx = t[-1:].clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    s[-1] = mx[0]
    return s.cpu().numpy()
