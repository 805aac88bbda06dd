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


def _settling_inputs(total_cols):
    from mistra_b200 import kon, sed
    g = kon.kon_grid(nka=8, nkt=9)
    d = sed.synthetic_columns(g, total_cols, n=30, nf=20, seed=23, j2=5, j6=3, populated=0.4)
    return g, d


def _driver_inputs(total_cols):
    from mistra_b200 import driver
    d = driver.synthetic_columns(total_cols, 16, 12, 4, 7, 3, 29)
    cfg = dict(nf=12, halo=True, iod=False, lpBuys13_0D=False, neula=0, box=False, n_bl=2, kinv=9, dt_ch=10.0)
    return cfg, d, np.array([2, -1, 5], dtype=np.int32), np.array([1e-9, 1e-9, -2e-10])


def _driver_call(fn, cfg, d, adv_row, xadv, sl=slice(None)):
    return fn(cfg, *[d[k][sl] for k in ("u0", "t", "p", "rho", "cm3", "am3", "xm1", "conv2", "cm", "cloud", "photol_j")],
              adv_row, xadv, d["s1"][sl], d["s3"][sl])


def _worker_settling(rank, world, port, total_cols, outdir):
    """sedp / sedl and the layer loop of kpp_driver shard by whole columns as well; the per-rank lists of layers per
    mechanism become the global ones by adding the rank's first layer row."""
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from mistra_b200 import shard
    from oracle import driver_oracle as dvo, sed_oracle as sdo
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, n = shard.column_block(total_cols, world, rank)
    sl = slice(first, first + n)
    g, d = _settling_inputs(total_cols)
    ff, dg = sdo.sedp(g, 10.0, 20, d["detw"], d["deta"], d["t"][sl], d["p"][sl], d["vd"][sl], d["ff"][sl], d["diag"][sl])
    s1 = sdo.sedl(10.0, 20, 3, d["detw"], d["deta"], d["t"][sl], d["p"][sl], d["rc"][sl], d["vt"][sl], d["vdm"][sl], d["sl1"][sl])
    cfg, dd, adv_row, xadv = _driver_inputs(total_cols)
    o = _driver_call(dvo.layers, cfg, dd, adv_row, xadv, sl)
    nlev = dd["t"].shape[1]
    np.savez(os.path.join(outdir, "sed%d.npz" % rank), ff=ff, dg=dg, sl1=s1, mech=o["mech"], cb1=o["cb1"], scal=o["scal"],
             **{"lay%d" % m: o["layers"][m] + first * nlev for m in range(3)})
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_settling_and_driver_equal_single_process(tmp_path):
    from oracle import driver_oracle as dvo, sed_oracle as sdo
    total_cols, world = 5, 2
    mp.spawn(_worker_settling, args=(world, _free_port(), total_cols, str(tmp_path)), nprocs=world, join=True)
    g, d = _settling_inputs(total_cols)
    ff, dg = sdo.sedp(g, 10.0, 20, d["detw"], d["deta"], d["t"], d["p"], d["vd"], d["ff"], d["diag"])
    s1 = sdo.sedl(10.0, 20, 3, d["detw"], d["deta"], d["t"], d["p"], d["rc"], d["vt"], d["vdm"], d["sl1"])
    cfg, dd, adv_row, xadv = _driver_inputs(total_cols)
    o = _driver_call(dvo.layers, cfg, dd, adv_row, xadv)
    parts = [np.load(os.path.join(str(tmp_path), "sed%d.npz" % r)) for r in range(world)]
    cat = lambda k: np.concatenate([p[k] for p in parts])
    assert np.array_equal(cat("ff"), ff) and np.array_equal(cat("dg"), dg) and np.array_equal(cat("sl1"), s1)
    assert (ff != d["ff"]).any()
    assert np.array_equal(cat("mech"), o["mech"]) and np.array_equal(cat("cb1"), o["cb1"]) and np.array_equal(cat("scal"), o["scal"])
    for m in range(3):
        assert np.array_equal(cat("lay%d" % m), o["layers"][m])
