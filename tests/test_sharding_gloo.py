"""Multi-process path on the CPU: world_size-2 gloo, cells sharded, spectra all-reduced (the kernel is replaced by the
oracle here -- the GPU version of this test is in test_gpu_parity.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from common import load_golden, problem_from_recipe
from is3d_b200 import distributed, tables

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 8, 9, 1000, 1000003):
        for g in (1, 2, 4, 8):
            spans = [distributed.shard_bounds(n, r, g) for r in range(g)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and a <= b
            assert max(b - a for a, b in spans) == (-(-n // g) if n else 0)


def _oracle_kernel(flags, cells, species, grid, df_tables, laguerre, memory="host", **kw):
    from oracle import cf_oracle as cfo
    dN, skipped, bd = cfo.smooth(flags, cells, species, grid, df_tables, laguerre)
    return dN, dict(cells_skipped_udsigma=skipped, cells_feqmod_breakdown=bd)


def _worker(rank, world, port, name, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fx = tables.load_fixture()
    gold = load_golden(name)
    fl, cells, sp, g, tab, gla = problem_from_recipe(gold["recipe"], fx)      # global tables built before sharding
    dN, st = distributed.smooth_spectra_sharded(fl, cells, sp, g, tab, gla, kernel=_oracle_kernel)
    np.save(os.path.join(out_dir, "dN_%d.npy" % rank), dN)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["s3_df1", "s3_df4"])
def test_two_rank_gloo_matches_single(tmp_path, name):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, name, str(tmp_path)), nprocs=2, join=True)
    gold = load_golden(name)
    d0 = np.load(tmp_path / "dN_0.npy"); d1 = np.load(tmp_path / "dN_1.npy")
    assert np.array_equal(d0, d1)                                   # every rank holds the reduced spectra
    nz = gold["dN"] != 0
    # summation order changes with the number of shards: bounded by a few ulp because all terms are >= 0
    assert np.max(np.abs(d0[nz] - gold["dN"][nz]) / gold["dN"][nz]) < 1e-13
    assert np.all(d0[~nz] == 0)


# ------------------------------------------------------------------------------------------ operation = 0 (spacetime distributions)
def _oracle_spacetime(flags, cells, species, grid, df_tables, laguerre, bins, memory="host", **kw):
    from oracle import cf_oracle as cfo
    out, skipped = cfo.spacetime(flags, cells, species, grid, df_tables, bins, laguerre)
    bd = out.pop("breakdown", 0)
    return out, dict(cells_skipped_udsigma=skipped, cells_feqmod_breakdown=bd)


def _worker_spacetime(rank, world, port, name, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from common import load_spacetime, spacetime_problem
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fx = tables.load_fixture()
    gold = load_spacetime(name)
    fl, cells, sp, g, tab, gla, bins, _ = spacetime_problem(gold["recipe"], fx)
    res, st = distributed.spacetime_distributions_sharded(fl, cells, sp, g, tab, gla, bins, kernel=_oracle_spacetime)
    np.savez(os.path.join(out_dir, "st_%d.npz" % rank), **res)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_spacetime(tmp_path):
    from common import load_spacetime, spacetime_close
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker_spacetime, args=(2, port, "dx3_df4", str(tmp_path)), nprocs=2, join=True)
    r0 = np.load(tmp_path / "st_0.npz"); r1 = np.load(tmp_path / "st_1.npz")
    for k in distributed.SPACETIME_KEYS:
        assert np.array_equal(r0[k], r1[k])
    spacetime_close({k: r0[k] for k in r0.files}, load_spacetime("dx3_df4"))
