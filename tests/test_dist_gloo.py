"""N>1 host logic on CPU: world_size-2 gloo run of the sharding + edge-collective plumbing that
bench.py uses (blocks sharded by contiguous ranges, no data-path collective; one all-reduce of
per-rank checksums/timings).  The per-rank "compute" here is the oracle on one block."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, PKG_NAME


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nblocks, q):
    import importlib
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = importlib.import_module(PKG_NAME)
    import oracle_py
    o = oracle_py.Oracle(1024, oracle_py.default_coeff_modulus(4096))
    lo, hi = pkg.shard_range(nblocks, rank, world)
    acc = 0
    for b in range(lo, hi):
        x = o.fill_uniform(b * 64, 64)          # rank-local generation: no scatter
        o.encrypted_dct(x)
        acc ^= o.checksum(x) & 0x7FFFFFFFFFFFFFFF
    t = torch.tensor([acc, hi - lo], dtype=torch.int64)
    cs = t[:1].clone()
    dist.all_reduce(cs, op=dist.ReduceOp.BXOR)
    cnt = t[1:].clone()
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    tm = torch.tensor([float(rank + 1)])
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    if rank == 0:
        q.put((int(cs.item()), int(cnt.item()), float(tm.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_rank(oracle_mod):
    nblocks, world = 5, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nblocks, q)) for r in range(world)]
    for p in procs:
        p.start()
    cs, cnt, tmax = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    o = oracle_mod.Oracle(1024, oracle_mod.default_coeff_modulus(4096))
    ref = 0
    for b in range(nblocks):
        x = o.fill_uniform(b * 64, 64)
        o.encrypted_dct(x)
        ref ^= o.checksum(x) & 0x7FFFFFFFFFFFFFFF
    assert cnt == nblocks and cs == ref and tmax == float(world)
