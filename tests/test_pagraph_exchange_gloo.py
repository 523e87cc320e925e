"""Host logic of the multi-GPU graph build on CPU (gloo, world_size 2): the owner-partitioned all-to-all of
aligngraph2_b200.pagraph.exchange_streams delivers every owner its vertex range from all ranks in rank order, so that a
stable sort by vertex restores the global (read-major) order inside every vertex -- the property the epsilon join's
exactness rests on."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_vertices, q):
    sys.path.insert(0, ROOT)
    from aligngraph2_b200 import pagraph
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(100 + rank)
    n = 5000 + 777 * rank
    vertex = rng.integers(0, n_vertices, size=n).astype(np.int32)
    order = (np.arange(n) + rank * 1_000_000).astype(np.int32)            # global stream position
    per = -(-n_vertices // world)
    owner = vertex // per
    perm = np.argsort(owner, kind="stable")                                # what ag2_pg_partition does on the device
    counts = np.bincount(owner, minlength=world)
    got_v, got_o = pagraph.exchange_streams([torch.from_numpy(vertex[perm]), torch.from_numpy(order[perm])], counts)
    got_v, got_o = got_v.numpy(), got_o.numpy()
    ok = bool(((got_v // per) == rank).all())
    # inside every vertex the received order must be ascending in the global stream position
    s = np.argsort(got_v, kind="stable")
    ok = ok and all((np.diff(got_o[s][got_v[s] == v]) > 0).all() for v in np.unique(got_v))
    total = torch.tensor([len(got_v)])
    dist.all_reduce(total)
    q.put((rank, ok, int(total.item()), len(got_v)))
    dist.destroy_process_group()


def test_exchange_preserves_global_order_per_vertex():
    world, n_vertices = 2, 97
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, 29611, n_vertices, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] for r in res), res
    assert res[0][2] == 5000 + 5000 + 777 and sum(r[3] for r in res) == res[0][2]


def test_exchange_plan_is_the_transpose():
    from aligngraph2_b200 import pagraph
    c = np.array([[1, 2, 3], [4, 5, 6], [7, 8, 9]])
    assert pagraph.exchange_plan(c, 1) == ([4, 5, 6], [2, 5, 8])
