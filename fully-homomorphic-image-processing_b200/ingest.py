"""One-rank ingest over NCCL: scatter chunks of input ciphertext blocks from the ingest rank to every
rank, compute, gather the results back -- the edge path `north_star` names ("NCCL over NVLink only to
scatter input ciphertexts and gather results").

The reference ingests one stream in one process (homo/server_jpeg.cpp:113-124 reads the blocks of the
image in order and writes them back in order, :146-153).  On N GPUs the blocks are independent, so the
ingest rank cuts its stream into N contiguous shares (shard_range), each share into chunks of whole
blocks, and runs a three-stage software pipeline in rounds: in round j every other rank receives chunk j
of its share (ncclRecv), computes chunk j - 1 and sends the result of chunk j - 2 back (ncclSend), all
three at once; the ingest rank posts the matching sends / receives of a round as ONE group
(batch_isend_irecv = ncclGroupStart/End) and computes its own share meanwhile.  Two buffers per
direction per rank: a chunk is >= 64 MiB of whole ciphertexts, so the transfers run at NVLink
rate and overlap the kernels.

The schedule is plain torch.distributed point-to-point code: it runs on `gloo` with CPU tensors too,
which is how tests/test_dist_gloo.py checks it (every chunk arrives once, in order, results land in
the right place) without a GPU.  Nothing here touches the arithmetic."""
from __future__ import annotations


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """contiguous share [lo, hi) of `total` units for `rank`: the package's shard_range (r N / G)"""
    return (rank * total) // world, ((rank + 1) * total) // world


def chunk_ranges(lo: int, hi: int, chunk: int) -> list[tuple[int, int]]:
    return [(a, min(a + chunk, hi)) for a in range(lo, hi, chunk)]


def scatter_compute_gather(dist, rank: int, world: int, total_units: int, chunk_units: int, unit_words: int,
                           src, dst, make_buffer, compute, ingest_rank: int = 0):
    """Runs the pipeline over `total_units` units (blocks) of `unit_words` elements each.

    src / dst: on the ingest rank, flat tensors of total_units * unit_words elements (input stream, output
    stream); ignored elsewhere.  make_buffer(n_elements) -> a flat tensor on this rank's device.
    compute(inp, out, n_units): out[:n] <- f(inp[:n]) for n_units units, enqueued on the current stream
    (may be asynchronous; the point-to-point ops of torch.distributed order themselves against it).
    Returns the number of units this rank computed."""
    shares = [shard_range(total_units, r, world) for r in range(world)]
    chunks = [chunk_ranges(lo, hi, chunk_units) for lo, hi in shares]
    rounds = max(len(c) for c in chunks) + 2
    me = chunks[rank]
    done = 0
    if rank == ingest_rank:
        for j in range(rounds):
            ops = []
            for r in range(world):
                if r == ingest_rank:
                    continue
                if j < len(chunks[r]):
                    a, b = chunks[r][j]
                    ops.append(dist.P2POp(dist.isend, src[a * unit_words:b * unit_words], r))
                if 0 <= j - 2 < len(chunks[r]):
                    a, b = chunks[r][j - 2]
                    ops.append(dist.P2POp(dist.irecv, dst[a * unit_words:b * unit_words], r))
            reqs = dist.batch_isend_irecv(ops) if ops else []
            if 0 <= j - 1 < len(me):          # the ingest rank's own share: no transfer at all
                a, b = me[j - 1]
                compute(src[a * unit_words:b * unit_words], dst[a * unit_words:b * unit_words], b - a)
                done += b - a
            for q in reqs:
                q.wait()
        return done
    inbuf = [make_buffer(chunk_units * unit_words) for _ in range(2)]
    outbuf = [make_buffer(chunk_units * unit_words) for _ in range(2)]
    for j in range(rounds):
        ops = []
        if j < len(me):
            a, b = me[j]
            ops.append(dist.P2POp(dist.irecv, inbuf[j % 2][:(b - a) * unit_words], ingest_rank))
        if 0 <= j - 2 < len(me):
            a, b = me[j - 2]
            ops.append(dist.P2POp(dist.isend, outbuf[j % 2][:(b - a) * unit_words], ingest_rank))
        reqs = dist.batch_isend_irecv(ops) if ops else []
        if 0 <= j - 1 < len(me):              # chunk j - 1 arrived in the previous round
            a, b = me[j - 1]
            compute(inbuf[(j - 1) % 2], outbuf[(j - 1) % 2], b - a)
            done += b - a
        for q in reqs:
            q.wait()
    return done
