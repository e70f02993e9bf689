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


def _ingest_worker(rank, world, port, total, chunk, unit, q):
    import importlib
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ingest = importlib.import_module(PKG_NAME + ".ingest")
    src = torch.arange(total * unit, dtype=torch.int64) if rank == 0 else None
    dst = torch.full((total * unit,), -1, dtype=torch.int64) if rank == 0 else None
    seen = []

    def compute(inp, out, n_units):
        # the stand-in for the DCT: an invertible map that also records which rank handled what
        seen.append((int(inp[0]) // unit, n_units))
        out[:n_units * unit] = inp[:n_units * unit] * 3 + (rank + 1)

    done = ingest.scatter_compute_gather(dist, rank, world, total, chunk, unit, src, dst,
                                         lambda n: torch.empty(n, dtype=torch.int64), compute)
    counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([done], dtype=torch.int64))
    if rank == 0:
        q.put((dst.clone(), [int(c.item()) for c in counts], seen))
    dist.barrier()
    dist.destroy_process_group()


def test_one_rank_ingest_schedule_delivers_every_chunk_once_and_in_place():
    """the NCCL scatter / compute / gather pipeline of ingest.py on gloo, world size 2 and 3: ragged
    shares (7 units, chunks of 2), every unit computed exactly once by the rank that owns it, results
    land at the unit's position of the ingest rank's output stream"""
    import importlib
    ingest = importlib.import_module(PKG_NAME + ".ingest")
    for world in (2, 3):
        total, chunk, unit = 7, 2, 5
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_ingest_worker, args=(r, world, port, total, chunk, unit, q)) for r in range(world)]
        for p in procs:
            p.start()
        dst, counts, seen0 = q.get(timeout=240)
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
        assert sum(counts) == total
        want = torch.empty(total * unit, dtype=torch.int64)
        for r in range(world):
            lo, hi = ingest.shard_range(total, r, world)
            assert counts[r] == hi - lo
            want[lo * unit:hi * unit] = torch.arange(lo * unit, hi * unit) * 3 + (r + 1)
        assert (dst == want).all()
        lo0, hi0 = ingest.shard_range(total, 0, world)
        assert seen0 == [(a, b - a) for a, b in ingest.chunk_ranges(lo0, hi0, chunk)]
