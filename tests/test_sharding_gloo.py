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
