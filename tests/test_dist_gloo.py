"""CPU tests of the multi-rank host logic with torch.distributed (gloo, world_size 2): the rendezvous
helpers, the unique-id broadcast, max-over-ranks timing, and -- through the oracle -- that sharding the
units by global index range with a counter-based stream gives the same moments as one rank."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import torch.distributed as td
    import oracle
    import parcompfin_b200 as pcf
    from parcompfin_b200 import dist
    job = dist.setup("gloo")
    assert (job.rank, job.world) == (rank, world)
    # 1. rank 0's 128-byte id reaches everyone
    payload = bytes(range(128)) if rank == 0 else None
    got = dist.share_bytes(job, payload)
    # 2. max over ranks
    mx = dist.reduce_scalars(job, [float(rank + 1), 10.0 - rank], "max")
    # 3. sharded Asian pricing through the oracle: rank r prices paths [b, e) of the Philox stream
    N, M, seed = 4001, 12, 77
    b, e = pcf.shard_of(N, rank, world)
    z = oracle.normal_stream(seed, pcf.STREAM_ASIA, b, e - b, M, (1 / M) ** 0.5)
    _, s, s2 = oracle.mc_asia(100, 100, .05, .2, 1, e - b, M, "call", z, moments=True) if e > b else (0, 0.0, 0.0)
    tot = dist.reduce_scalars(job, [s, s2, float(e - b)], "sum")
    # 4. antithetic pairs shard by pair index (mc_amer): the union of shards is every pair exactly once
    pb, pe = pcf.shard_of(N // 2, rank, world)
    cnt = dist.reduce_scalars(job, [float(pe - pb)], "sum")
    dist.barrier(job)
    q.put((rank, got == bytes(range(128)), mx, tot, cnt))
    dist.teardown(job)


def test_two_rank_gloo_sharding_matches_single_rank():
    import torch.multiprocessing as mp
    import oracle
    import parcompfin_b200 as pcf
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    N, M, seed = 4001, 12, 77
    z = oracle.normal_stream(seed, pcf.STREAM_ASIA, 0, N, M, (1 / M) ** 0.5)
    _, s, s2 = oracle.mc_asia(100, 100, .05, .2, 1, N, M, "call", z, moments=True)
    for rank, ok, mx, tot, cnt in res:
        assert ok
        assert mx == [2.0, 10.0]
        assert abs(tot[0] - s) < 1e-12 * abs(s) and abs(tot[1] - s2) < 1e-12 * abs(s2) and tot[2] == N
        assert cnt == [float(N // 2)]


def test_single_process_job_is_a_noop():
    from parcompfin_b200 import dist
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        os.environ.pop(k, None)
    job = dist.setup()
    assert (job.rank, job.world, job.backend) == (0, 1, None)
    assert dist.share_bytes(job, b"x") == b"x"
    assert dist.reduce_scalars(job, [1.0, 2.0]) == [1.0, 2.0]
    dist.barrier(job)
    dist.teardown(job)
