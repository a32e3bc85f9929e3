"""N>1 host logic on CPU: world_size-2 gloo processes shard a fiber batch, broadcast the
cores and all-gather the values; the reassembled result must equal the single-process one.
The per-fiber 'backup' here is the CPU oracle (checker role only) so the values are real."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from c3sc_b200 import configs, sharding, synthetic     # noqa: E402


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, F, q):
    from helpers import make_ft, make_port
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = configs.get_config("skidding5d", n=8, rank=3)
    oport = make_port(cfg)
    ranks = cfg.ranks()
    # rank 0 owns the cores; the others start with garbage and must receive them
    cores = synthetic.random_cores(cfg.ngrid, ranks) if rank == 0 else [np.full_like(c, np.nan) for c in synthetic.random_cores(cfg.ngrid, ranks)]
    flat = torch.from_numpy(np.concatenate(cores))
    sharding.broadcast_cores(flat, src=0)
    sizes = [c.size for c in cores]
    cores = [a.copy() for a in np.split(flat.numpy(), np.cumsum(sizes)[:-1])]
    from oracle import pyoracle as po
    ft = po.FT(cfg.ngrid, ranks, cores)
    dv, fi = synthetic.random_fibers(cfg.ngrid, F)
    sdv, sfi, nreal = sharding.shard_fibers(dv, fi, world, rank)
    vals, _ = oport.vi_batch(ft, sdv, sfi, nthreads=1)
    local = torch.from_numpy(np.ascontiguousarray(vals).reshape(-1))
    full = sharding.gather_values(local, F, cfg.n)
    if rank == 0:
        q.put((full.numpy().reshape(F, cfg.n).copy(), nreal))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("F", [16, 17, 1])
def test_two_rank_sharding_matches_single_process(built, F):
    from helpers import make_ft, make_port
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, F, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, nreal = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cfg = configs.get_config("skidding5d", n=8, rank=3)
    oport = make_port(cfg)
    ranks, cores, ft = make_ft(cfg)
    dv, fi = synthetic.random_fibers(cfg.ngrid, F)
    want, _ = oport.vi_batch(ft, dv, fi, nthreads=1)
    assert np.array_equal(got, want)
    assert nreal == sharding.shard_range(F, 2, 0)[1]


def test_shard_ranges_partition_the_batch():
    for F in (0, 1, 7, 8, 9, 6480, 65536):
        for world in (1, 2, 4, 8):
            covered = []
            for r in range(world):
                lo, hi = sharding.shard_range(F, world, r)
                assert 0 <= lo <= hi <= F and hi - lo <= sharding.shard_size(F, world)
                covered += list(range(lo, hi))
            assert covered == list(range(F))
