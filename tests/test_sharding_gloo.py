"""The N>1 path on CPU: two gloo ranks shard the counter-RNG workload, score their shards (with the ORACLE -- this is a
test of the host-side sharding logic, the CUDA path has no CPU mode), reduce their scalars and gather their results."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as tmp

import oracle_lib as ol
from mini_parallel_b200 import sharding, synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = sharding.shard_range(rank, world, n)
        qb, qo, rb, ro = synth.make_pairs(lo, hi - lo, 150, 500, 0)         # this rank regenerates only its shard
        res = ol.batch(qb, qo, rb, ro, threads=2, simd=True)
        local = torch.from_numpy(np.stack([res["score"], res["end_i"], res["end_j"]], axis=1).astype(np.int32))
        ms = 10.0 + rank                                                      # stand-in for the CUDA-event time of the rank
        mx = sharding.reduce_scalars([ms], "max")[0]
        tot = sharding.reduce_scalars([float(local[:, 0].sum()), float(hi - lo)], "sum")
        allr = sharding.gather_results(local, n, rank, world)
        q.put((rank, lo, hi, mx, tot, allr.numpy()))
    finally:
        dist.destroy_process_group()


def test_shard_ranges_partition_the_index_space():
    for n in (0, 1, 7, 1000, 100_000_000):
        for world in (1, 2, 3, 4, 8):
            r = [sharding.shard_range(k, world, n) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_two_gloo_ranks_shard_score_reduce_gather():
    n, world = 601, 2
    ctx = tmp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    qb, qo, rb, ro = synth.make_pairs(0, n, 150, 500, 0)
    whole = ol.batch(qb, qo, rb, ro, threads=4, simd=True)
    exp = np.stack([whole["score"], whole["end_i"], whole["end_j"]], axis=1).astype(np.int32)
    assert (got[0][1], got[0][2], got[1][1], got[1][2]) == (0, 300, 300, 601)
    for rank, lo, hi, mx, tot, allr in got:
        assert mx == 11.0                                                     # max over ranks of the per-rank time
        assert tot == [float(exp[:, 0].sum()), float(n)]                      # checksum of checksums
        assert np.array_equal(allr, exp)                                      # sharded == whole, on every rank
