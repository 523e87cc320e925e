"""The N>1 bookkeeping of the read-sharded runs (aligngraph2_b200/shard.py) under torch.distributed with the gloo
backend, world_size 2, on the CPU: contiguous balanced read ranges that tile the batch in rank order, max-over-ranks
timing, summed work counters."""
import os
import socket

import torch.multiprocessing as mp

from aligngraph2_b200.shard import read_range


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_reads, q):
    import torch.distributed as dist
    from aligngraph2_b200.shard import read_range, reduce_measurement
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = read_range(n_reads, rank, world)
    ms, c = reduce_measurement(10.0 + 5 * rank, {"aligned": (hi - lo) * 100, "cells": hi - lo})
    dist.barrier()
    q.put((rank, lo, hi, ms, c))
    dist.destroy_process_group()


def test_read_ranges_tile_the_batch():
    for n, w in ((10, 3), (7, 8), (500000, 8), (1, 2), (0, 4)):
        r = [read_range(n, k, w) for k in range(w)]
        assert r[0][0] == 0 and r[-1][1] == n
        assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
        sizes = [b - a for a, b in r]
        assert max(sizes) - min(sizes) <= 1


def test_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 101, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (_, lo0, hi0, ms0, c0), (_, lo1, hi1, ms1, c1) = out
    assert (lo0, hi0, lo1, hi1) == (0, 51, 51, 101)
    assert ms0 == ms1 == 15.0                      # max over ranks
    assert c0 == c1 == {"aligned": 10100.0, "cells": 101.0}
