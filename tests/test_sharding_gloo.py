"""N>1 path on CPU: two gloo ranks shard an ensemble by column blocks, each
integrates its own cells (the oracle stands in for the device here) and only the
diagnostics are all-reduced; the union must equal the single-process run."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total_cols, outdir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from mistra_b200 import shard, synthetic
    from oracle import kpp_oracle as ko
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, n = shard.column_block(total_cols, world, rank)
    ens = synthetic.GasEnsemble(n, col0=first)
    out, ierr, stats, hexit, _ = ko.integrate(0, ens.rconst(), ens.fix, ens.var, nthreads=2)
    diag = shard.reduce_diagnostics(dist, shard.diagnostics_vector(ierr, stats))
    np.savez(os.path.join(outdir, "rank%d.npz" % rank), out=out, stats=stats, diag=diag, first=first, n=n)
    dist.barrier()
    dist.destroy_process_group()


def test_column_blocks_partition_the_ensemble():
    from mistra_b200 import shard
    for total in (1, 7, 8, 10000):
        for world in (1, 2, 4, 8):
            blocks = [shard.column_block(total, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and sum(n for _, n in blocks) == total
            for (f0, n0), (f1, _) in zip(blocks, blocks[1:]):
                assert f0 + n0 == f1
            assert max(n for _, n in blocks) - min(n for _, n in blocks) <= 1


def test_two_rank_gloo_run_equals_single_process(tmp_path, oracle):
    from mistra_b200 import shard, synthetic
    total_cols, world = 5, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, total_cols, str(tmp_path)), nprocs=world, join=True)
    ens = synthetic.GasEnsemble(total_cols)
    out, ierr, stats, _, _ = oracle.integrate(0, ens.rconst(), ens.fix, ens.var, nthreads=4)
    parts = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    assert [int(p["n"]) for p in parts] == [3, 2]
    assert np.array_equal(np.concatenate([p["out"] for p in parts]), out)
    assert np.array_equal(np.concatenate([p["stats"] for p in parts]), stats)
    want = shard.diagnostics_vector(ierr, stats)
    for p in parts:
        assert np.array_equal(p["diag"], want)


def _worker_columns(rank, world, port, total_cols, outdir):
    """The column-wise operators (difc on the chemistry arrays, drive copies) shard the same way: whole columns
    per rank, nothing exchanged."""
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from mistra_b200 import shard
    from oracle import difc_oracle as dfo
    from tests.test_difc_oracle import inputs, run
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, n = shard.column_block(total_cols, world, rank)
    c, fields = inputs(total_cols, 17, n=40, sizes=(9, 4, 3 * 4, 2 * 4))      # every rank builds the same ensemble
    mine = {k: (v[first:first + n] if v.ndim == 2 else v) for k, v in c.items()}
    outs = run(dfo.difc, 60.0, mine, [(a[first:first + n], p) for a, p in fields])
    np.savez(os.path.join(outdir, "cols%d.npz" % rank), first=first, n=n, **{"o%d" % i: o for i, o in enumerate(outs)})
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_column_operators_equal_single_process(tmp_path):
    from oracle import difc_oracle as dfo
    from tests.test_difc_oracle import inputs, run
    total_cols, world = 7, 2
    mp.spawn(_worker_columns, args=(world, _free_port(), total_cols, str(tmp_path)), nprocs=world, join=True)
    c, fields = inputs(total_cols, 17, n=40, sizes=(9, 4, 3 * 4, 2 * 4))
    ref = run(dfo.difc, 60.0, c, fields)
    parts = [np.load(os.path.join(str(tmp_path), "cols%d.npz" % r)) for r in range(world)]
    assert [int(p["n"]) for p in parts] == [4, 3]
    for i, r in enumerate(ref):
        assert np.array_equal(np.concatenate([p["o%d" % i] for p in parts]), r)
