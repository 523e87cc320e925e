"""ctypes mirror of include/ag2_pagraph.h: PAGraph's A-Bruijn graph build on the GPU and the
traversal of the built graph (row B9) on the host.

``Job`` follows run2() of PAGraph/src/main/pagraph.cpp:69-243 over the file-level entry points: open the input set, then
per config block load_block() + build() (+ dump()).  There is no CPU path: without the CUDA library or a GPU the calls
raise.  ``build_distributed`` is the multi-GPU form (reads sharded over ranks, one all-to-all of vertex tuples by owner
rank) on torch.distributed.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import lib as _lib


class Params(C.Structure):
    _fields_ = [("outer_sample", C.c_int32), ("read_to_ctg_topk", C.c_int32), ("read_to_ref_topk", C.c_int32),
                ("read_to_ctg_ratio", C.c_double), ("read_to_ref_ratio", C.c_double), ("epsilon", C.c_int64),
                ("cov_filter", C.c_int64)]


class Stats(C.Structure):
    _fields_ = [("n_vertices", C.c_int64), ("lanes", C.c_int64 * 2), ("columns", C.c_int64 * 2), ("samples", C.c_int64 * 2),
                ("tuples", C.c_int64 * 2), ("edges_raw", C.c_int64 * 2), ("positions", C.c_int64), ("edges", C.c_int64),
                ("launches", C.c_int64), ("extract_ms", C.c_double), ("join_ms", C.c_double),
                ("join_sort_ms", C.c_double), ("join_cluster_ms", C.c_double), ("join_edges_ms", C.c_double)]

    def as_dict(self):
        return {n: (list(getattr(self, n)) if hasattr(getattr(self, n), "__len__") else getattr(self, n)) for n, _ in self._fields_}


class GraphView(C.Structure):
    _fields_ = [("k", C.c_int32), ("n_vertices", C.c_int64), ("codes", C.c_void_p), ("pos_off", C.c_void_p), ("ctg", C.c_void_p),
                ("ref", C.c_void_p), ("count", C.c_void_p), ("edge_off", C.c_void_p), ("edge_to", C.c_void_p), ("edge_step", C.c_void_p)]


class SeqsView(C.Structure):
    _fields_ = [("n", C.c_int64), ("names", C.POINTER(C.c_char_p)), ("bases", C.c_char_p), ("offs", C.c_void_p)]


class TravelParams(C.Structure):
    _fields_ = [("deviation", C.c_int64), ("error_rate", C.c_double), ("start_split", C.c_double), ("min_len", C.c_int64),
                ("threads", C.c_int32)]


ALN_DTYPE = np.dtype([("query", "<i4"), ("target", "<i4"), ("score", "<u8"), ("qb", "<i8"), ("qe", "<i8"), ("tb", "<i8"),
                      ("te", "<i8"), ("forward", "<i4"), ("ncols", "<i4"), ("q_off", "<i8"), ("t_off", "<i8")])
assert ALN_DTYPE.itemsize == 72

EXPORTS = [
    "ag2_pg_create", "ag2_pg_destroy", "ag2_pg_last_error", "ag2_pg_params_default", "ag2_pg_set_kmers", "ag2_pg_fetch_codes",
    "ag2_pg_set_targets", "ag2_pg_set_reads", "ag2_pg_set_alignments", "ag2_pg_set_filters", "ag2_pg_build", "ag2_pg_extract",
    "ag2_pg_partition", "ag2_pg_stream_dev", "ag2_pg_import_dev", "ag2_pg_join", "ag2_pg_group_exchange", "ag2_pg_group_gather", "ag2_pg_get_stats", "ag2_pg_graph_fetch",
    "ag2_pg_stream", "ag2_pg_job_open", "ag2_pg_job_close", "ag2_pg_job_error", "ag2_pg_job_blocks", "ag2_pg_job_block_ref",
    "ag2_pg_job_handle", "ag2_pg_job_load_block", "ag2_pg_job_dump",
    "ag2_pg_travel_params_default", "ag2_pg_travel", "ag2_pg_job_travel", "ag2_pg_job_write_contig_list",
]

_bound = False


def _L():
    global _bound
    L = _lib.load()
    if _bound:
        return L
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int
    L.ag2_pg_create.argtypes = [i32, C.POINTER(vp)]
    L.ag2_pg_destroy.argtypes = [vp]
    L.ag2_pg_destroy.restype = None
    L.ag2_pg_last_error.argtypes = [vp]
    L.ag2_pg_last_error.restype = C.c_char_p
    L.ag2_pg_params_default.argtypes = [C.POINTER(Params)]
    L.ag2_pg_params_default.restype = None
    L.ag2_pg_set_kmers.argtypes = [vp, vp, i64, C.POINTER(i64)]
    L.ag2_pg_fetch_codes.argtypes = [vp, vp, i64]
    L.ag2_pg_set_targets.argtypes = [vp, vp, i64, vp, i64]
    L.ag2_pg_set_reads.argtypes = [vp, vp, vp, i64, i64]
    L.ag2_pg_set_alignments.argtypes = [vp, i32, vp, i64, vp, i64]
    L.ag2_pg_set_filters.argtypes = [vp, vp, vp, vp]
    for n in ("ag2_pg_build", "ag2_pg_extract", "ag2_pg_join"):
        getattr(L, n).argtypes = [vp, C.POINTER(Params)]
    L.ag2_pg_partition.argtypes = [vp, i32, vp]
    L.ag2_pg_stream_dev.argtypes = [vp, C.POINTER(i64), vp, C.POINTER(i64), vp]
    L.ag2_pg_import_dev.argtypes = [vp, i64, vp, i64, vp]
    L.ag2_pg_group_exchange.argtypes = [vp, i32]
    L.ag2_pg_group_gather.argtypes = [vp, i32]
    L.ag2_pg_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.ag2_pg_graph_fetch.argtypes = [vp, vp, vp, vp, vp, i64, vp, vp, vp, i64]
    L.ag2_pg_stream.argtypes = [vp]
    L.ag2_pg_stream.restype = vp
    L.ag2_pg_job_open.argtypes = [i32, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(vp)]
    L.ag2_pg_job_close.argtypes = [vp]
    L.ag2_pg_job_close.restype = None
    L.ag2_pg_job_error.argtypes = [vp]
    L.ag2_pg_job_error.restype = C.c_char_p
    L.ag2_pg_job_blocks.argtypes = [vp]
    L.ag2_pg_job_block_ref.argtypes = [vp, i32]
    L.ag2_pg_job_block_ref.restype = C.c_char_p
    L.ag2_pg_job_handle.argtypes = [vp]
    L.ag2_pg_job_handle.restype = vp
    L.ag2_pg_job_load_block.argtypes = [vp, i32, i32, i32]
    L.ag2_pg_job_dump.argtypes = [vp, i32, C.c_char_p, i32]
    L.ag2_pg_travel_params_default.argtypes = [C.POINTER(TravelParams)]
    L.ag2_pg_travel_params_default.restype = None
    L.ag2_pg_travel.argtypes = [C.POINTER(GraphView), C.POINTER(SeqsView), C.POINTER(SeqsView), vp, vp, i64, C.POINTER(TravelParams),
                                C.c_char_p, C.c_char_p, vp, C.POINTER(i64)]
    L.ag2_pg_job_travel.argtypes = [vp, i32, C.POINTER(TravelParams), C.c_char_p]
    L.ag2_pg_job_write_contig_list.argtypes = [vp, C.c_char_p]
    _bound = True
    return L


def default_params(epsilon: int = 10, cov: int = 1) -> Params:
    p = Params()
    _L().ag2_pg_params_default(C.byref(p))
    p.epsilon, p.cov_filter = epsilon, cov
    return p


def travel_params(epsilon: int = 10, min_len: int = 50, threads: int = 1) -> TravelParams:
    """PGM/pagraph.cpp:123-126,247-256: deviation = 2 * epsilon, errorRate 0.15, startSplit 0.90."""
    p = TravelParams()
    _L().ag2_pg_travel_params_default(C.byref(p))
    p.deviation, p.min_len, p.threads = 2 * epsilon, min_len, threads
    return p


def _seqs_view(names, seqs):
    arr = (C.c_char_p * max(len(names), 1))(*[n.encode() for n in names])
    bases = b"".join(seqs)
    offs = np.concatenate(([0], np.cumsum([len(x) for x in seqs]))).astype(np.int64)
    v = SeqsView(len(names), arr, bases, offs.ctypes.data)
    return v, (arr, bases, offs)


def travel(graph: "Graph", codes: np.ndarray, k: int, ctgs, refs, use, params: TravelParams, out_dir: str, prefix: str):
    """B9 on a host graph (ag2_pg_travel): ``ctgs`` / ``refs`` = (names, sequences as bytes), ``use`` = [(contig index,
    forward flag)] of the config block.  Writes the reference's files under out_dir; returns the success contig indices."""
    L = _L()
    keep = [np.ascontiguousarray(a) for a in (codes.astype(np.uint64), graph.pos_off.astype(np.int64), graph.ctg.astype(np.uint32),
                                               graph.ref.astype(np.uint32), graph.count.astype(np.uint16), graph.edge_off.astype(np.int64),
                                               graph.edge_to.astype(np.uint32), graph.edge_step.astype(np.int32))]
    gv = GraphView(k, len(codes), *[a.ctypes.data for a in keep])
    cv, ck = _seqs_view(*ctgs)
    rv, rk = _seqs_view(*refs)
    uc = np.array([u[0] for u in use], np.int32)
    uf = np.array([1 if u[1] else 0 for u in use], np.uint8)
    ok = np.zeros(2 * len(use) + 1, np.int32)
    n_ok = C.c_int64()
    rc = L.ag2_pg_travel(C.byref(gv), C.byref(cv), C.byref(rv), uc.ctypes.data, uf.ctypes.data, len(use), C.byref(params),
                         out_dir.encode(), prefix.encode(), ok.ctypes.data, C.byref(n_ok))
    if rc != 0:
        raise _lib.Ag2Error(f"ag2_pg_travel -> {_lib.ERRORS.get(rc, rc)}")
    del ck, rk
    return ok[:n_ok.value].tolist()


class Graph:
    """CSR copy of the device graph (ag2_pg_graph_fetch)."""

    def __init__(self, pos_off, ctg, ref, count, edge_off, edge_to, edge_step):
        self.pos_off, self.ctg, self.ref, self.count = pos_off, ctg, ref, count
        self.edge_off, self.edge_to, self.edge_step = edge_off, edge_to, edge_step


class Job:
    """One pagraph input set (the -k -c -R -p -a arguments of `pagraph`, AlignGraph2.py:414-427)."""

    def __init__(self, kmer: str, ctg: str, ref: str, pre_dir: str, aln: str, device: int = 0):
        self.L = _L()
        self.h = C.c_void_p()
        rc = self.L.ag2_pg_job_open(device, kmer.encode(), ctg.encode(), ref.encode(), pre_dir.encode(), aln.encode(), C.byref(self.h))
        if rc != 0:
            msg = self.L.ag2_pg_job_error(self.h).decode() if self.h else ""
            if self.h:
                self.L.ag2_pg_job_close(self.h)
                self.h = None
            raise _lib.Ag2Error(f"ag2_pg_job_open -> {_lib.ERRORS.get(rc, rc)}: {msg}")
        self.pg = self.L.ag2_pg_job_handle(self.h)

    def _check(self, rc, what):
        if rc != 0:
            raise _lib.Ag2Error(f"{what} -> {_lib.ERRORS.get(rc, rc)}: {self.L.ag2_pg_job_error(self.h).decode()} "
                                f"{self.L.ag2_pg_last_error(self.pg).decode()}")

    @property
    def n_blocks(self) -> int:
        return self.L.ag2_pg_job_blocks(self.h)

    def load_block(self, block: int, rank: int = 0, world: int = 1) -> None:
        self._check(self.L.ag2_pg_job_load_block(self.h, block, rank, world), "ag2_pg_job_load_block")

    def build(self, params: Params) -> Stats:
        self._check(self.L.ag2_pg_build(self.pg, C.byref(params)), "ag2_pg_build")
        return self.stats()

    def extract(self, params: Params) -> Stats:
        self._check(self.L.ag2_pg_extract(self.pg, C.byref(params)), "ag2_pg_extract")
        return self.stats()

    def join(self, params: Params) -> Stats:
        self._check(self.L.ag2_pg_join(self.pg, C.byref(params)), "ag2_pg_join")
        return self.stats()

    def stats(self) -> Stats:
        s = Stats()
        self.L.ag2_pg_get_stats(self.pg, C.byref(s))
        return s

    def dump(self, block: int, path: str, append: bool = False) -> None:
        self._check(self.L.ag2_pg_job_dump(self.h, block, path.encode(), 1 if append else 0), "ag2_pg_job_dump")

    def codes(self) -> np.ndarray:
        out = np.zeros(max(self.stats().n_vertices, 1), np.uint64)
        self._check(self.L.ag2_pg_fetch_codes(self.pg, out.ctypes.data, len(out)), "ag2_pg_fetch_codes")
        return out[:self.stats().n_vertices]

    def travel(self, block: int, params: TravelParams, out_dir: str) -> None:
        """B9 on the graph the handle holds (PAssembly::testTravel5): writes <out_dir>/<block>_<ctg>_<0|1>.txt etc."""
        self._check(self.L.ag2_pg_job_travel(self.h, block, C.byref(params), out_dir.encode()), "ag2_pg_job_travel")

    def write_contig_list(self, out_dir: str) -> None:
        self._check(self.L.ag2_pg_job_write_contig_list(self.h, out_dir.encode()), "ag2_pg_job_write_contig_list")

    def block_ref(self, block: int) -> str:
        return self.L.ag2_pg_job_block_ref(self.h, block).decode()

    def graph(self) -> Graph:
        s = self.stats()
        nv = s.n_vertices
        po, eo = np.zeros(nv + 1, np.int64), np.zeros(nv + 1, np.int64)
        ctg, ref = np.zeros(max(s.positions, 1), np.uint32), np.zeros(max(s.positions, 1), np.uint32)
        cnt = np.zeros(max(s.positions, 1), np.uint16)
        to, step = np.zeros(max(s.edges, 1), np.uint32), np.zeros(max(s.edges, 1), np.int32)
        self._check(self.L.ag2_pg_graph_fetch(self.pg, po.ctypes.data, ctg.ctypes.data, ref.ctypes.data, cnt.ctypes.data, len(ctg),
                                              eo.ctypes.data, to.ctypes.data, step.ctypes.data, len(to)), "ag2_pg_graph_fetch")
        return Graph(po, ctg[:s.positions], ref[:s.positions], cnt[:s.positions], eo, to[:s.edges], step[:s.edges])

    # ---- multi-GPU: streams as torch tensors over the handle's device memory -------------------------------------
    def partition(self, n_owners: int) -> np.ndarray:
        counts = np.zeros(2 * n_owners, np.int64)
        self._check(self.L.ag2_pg_partition(self.pg, n_owners, counts.ctypes.data), "ag2_pg_partition")
        return counts.reshape(2, n_owners)

    def stream_pointers(self):
        nt, ne = C.c_int64(), C.c_int64()
        tp, ep = (C.c_void_p * 3)(), (C.c_void_p * 3)()
        self._check(self.L.ag2_pg_stream_dev(self.pg, C.byref(nt), tp, C.byref(ne), ep), "ag2_pg_stream_dev")
        return nt.value, [tp[i] or 0 for i in range(3)], ne.value, [ep[i] or 0 for i in range(3)]

    def import_streams(self, n_tuples: int, tuple_ptrs, n_edges: int, edge_ptrs) -> None:
        tp, ep = (C.c_void_p * 3)(*tuple_ptrs), (C.c_void_p * 3)(*edge_ptrs)
        self._check(self.L.ag2_pg_import_dev(self.pg, n_tuples, tp, n_edges, ep), "ag2_pg_import_dev")

    def close(self) -> None:
        if self.h:
            self.L.ag2_pg_job_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def exchange_plan(counts_all: np.ndarray, rank: int):
    """Split sizes of the all-to-all from the gathered per-owner counts.

    counts_all[r][o] = items rank r holds for owner o.  Returns (send splits of this rank, receive splits): the receive
    buffer is the rank-major concatenation, i.e. global read order inside every vertex once it is stably sorted by vertex.
    """
    counts_all = np.asarray(counts_all, dtype=np.int64)
    return counts_all[rank].tolist(), counts_all[:, rank].tolist()


def _device_tensor(ptr: int, n: int, device):
    """uint32/int32 view (as int32) of n 4-byte items of library-owned device memory."""
    import torch

    if n == 0:
        return torch.empty(0, dtype=torch.int32, device=device)

    class _Arr:
        __cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (ptr, False), "version": 3}

    return torch.as_tensor(_Arr(), device=device)


def exchange_streams(arrays, counts, group=None):
    """The one exchange step of the graph build: every rank holds its streams already partitioned by owner rank
    (`counts[o]` items for owner o, same split for every array of the stream); one all-to-all per array delivers to
    each owner the items of its vertex range from all ranks, concatenated in RANK order -- which is the global read order,
    because reads are sharded contiguously.  Works on any backend (NCCL on the GPUs, gloo in the CPU tests)."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = arrays[0].device
    mine = torch.as_tensor(np.asarray(counts, dtype=np.int64), device=dev)
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine, group=group)
    allc = torch.stack(gathered).cpu().numpy()
    send, recv = exchange_plan(allc, rank)
    out = []
    for src in arrays:
        dst = torch.empty(int(sum(recv)), dtype=src.dtype, device=dev)
        dist.all_to_all_single(dst, src, output_split_sizes=recv, input_split_sizes=send, group=group)
        out.append(dst)
    return out


def build_distributed(job: Job, block: int, params: Params, group=None) -> Stats:
    """Reads sharded over the ranks of the (NCCL) process group: extract per rank, stable partition by owner rank of the
    vertex, ONE all-to-all per stream array, join per vertex range.  Every rank ends with the vertices it owns; the others
    are empty in its CSR.  Results equal the one-GPU build because the rank-major receive order is the global read order."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    job.load_block(block, rank, world)
    return extract_exchange_join(job, params, group)


def extract_exchange_join(job: Job, params: Params, group=None) -> Stats:
    """build_distributed without the file loading: the block's reads of this rank are loaded already."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    job.extract(params)
    if world == 1:
        return job.join(params)
    counts = job.partition(world)                                  # [2][world]
    dev = torch.device("cuda", torch.cuda.current_device())
    nt, tptr, ne, eptr = job.stream_pointers()
    tup = exchange_streams([_device_tensor(p, nt, dev) for p in tptr], counts[0], group)
    edg = exchange_streams([_device_tensor(p, ne, dev) for p in eptr], counts[1], group)
    torch.cuda.synchronize()
    job.import_streams(tup[0].numel(), [t.data_ptr() for t in tup], edg[0].numel(), [t.data_ptr() for t in edg])
    return job.join(params)


def build_group(jobs, block: int, params: Params):
    """The graph build over the GPUs of ONE process: jobs[r] (one per device; a device may appear twice) takes the r-th
    contiguous range of the block's reads, extracts, and ag2_pg_group_exchange moves every (rank, owner) segment straight
    into the owner's buffer with peer copies over NVLink; then every job joins its vertex range.  Returns the jobs' Stats."""
    n = len(jobs)
    for r, j in enumerate(jobs):
        j.load_block(block, r, n)
        j.extract(params)
    if n > 1:
        L = _L()
        arr = (C.c_void_p * n)(*[j.pg for j in jobs])
        rc = L.ag2_pg_group_exchange(arr, n)
        if rc != 0:
            raise _lib.Ag2Error(f"ag2_pg_group_exchange: {_lib.ERRORS.get(rc, rc)}: {L.ag2_pg_last_error(jobs[0].pg).decode(errors='replace')}")
    return [j.join(params) for j in jobs]


def gather_group(jobs) -> None:
    """ag2_pg_group_gather: the per-GPU vertex tables merged into jobs[0] on the device (peer copies + offset sums)."""
    n = len(jobs)
    L = _L()
    arr = (C.c_void_p * n)(*[j.pg for j in jobs])
    rc = L.ag2_pg_group_gather(arr, n)
    if rc != 0:
        raise _lib.Ag2Error(f"ag2_pg_group_gather: {_lib.ERRORS.get(rc, rc)}: {L.ag2_pg_last_error(jobs[0].pg).decode(errors='replace')}")


def merge_graphs(graphs) -> Graph:
    """Graphs of the jobs of build_group / build_distributed (each holds the vertices its rank owns, the others empty):
    owner ranges ascend with the rank, so the merged payload is the rank-order concatenation and the merged per-vertex
    sizes are the element-wise sums."""
    pos_n = sum(np.diff(g.pos_off) for g in graphs)
    edge_n = sum(np.diff(g.edge_off) for g in graphs)
    cat = lambda name: np.concatenate([getattr(g, name) for g in graphs])
    return Graph(np.concatenate(([0], np.cumsum(pos_n))).astype(np.int64), cat("ctg"), cat("ref"), cat("count"),
                 np.concatenate(([0], np.cumsum(edge_n))).astype(np.int64), cat("edge_to"), cat("edge_step"))


def gather_graph(job: Job, group=None) -> Graph:
    """The all-gather that merges the per-GPU vertex tables before the traversal: every rank receives the
    CSR pieces of all ranks.  Owner ranges ascend with the rank, so the merged payload is the rank-order concatenation and
    the merged per-vertex sizes are the element-wise sums."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    g = job.graph()
    if world == 1:
        return g
    dev = torch.device("cuda", torch.cuda.current_device())

    def gather(a: np.ndarray) -> list:
        n = torch.tensor([a.size], dtype=torch.int64, device=dev)
        sizes = [torch.empty_like(n) for _ in range(world)]
        dist.all_gather(sizes, n, group=group)
        sizes = [int(x.item()) for x in sizes]
        pad = torch.zeros(max(max(sizes), 1), dtype=torch.int64, device=dev)
        pad[:a.size] = torch.from_numpy(a.astype(np.int64)).to(dev)
        outs = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(outs, pad, group=group)
        return [o[:k].cpu().numpy() for o, k in zip(outs, sizes)]

    pos_n = sum(gather(np.diff(g.pos_off)))
    edge_n = sum(gather(np.diff(g.edge_off)))
    cat = lambda a, dt: np.concatenate(gather(a)).astype(dt)
    pos_off = np.concatenate(([0], np.cumsum(pos_n))).astype(np.int64)
    edge_off = np.concatenate(([0], np.cumsum(edge_n))).astype(np.int64)
    return Graph(pos_off, cat(g.ctg, np.uint32), cat(g.ref, np.uint32), cat(g.count, np.uint16),
                 edge_off, cat(g.edge_to, np.uint32), cat(g.edge_step, np.int32))


def graph_dump_text(graph: Graph, codes: np.ndarray, block: int, ref_name: str) -> bytes:
    """The dump format of ag2_pg_job_dump from a Graph held on the host (used with gather_graph)."""
    out = [f"#config {block} {ref_name}\n"]
    po, eo = graph.pos_off, graph.edge_off
    for v in np.nonzero((np.diff(po) > 0) | (np.diff(eo) > 0))[0]:
        pos = " ".join(f"{graph.ctg[i]},{graph.ref[i]},{graph.count[i]}" for i in range(po[v], po[v + 1]))
        edg = " ".join(f"{graph.edge_to[i]},{graph.edge_step[i]}" for i in range(eo[v], eo[v + 1]))
        line = f"V {v} {codes[v]} P {po[v + 1] - po[v]}" + (" " + pos if pos else "") + f" E {eo[v + 1] - eo[v]}" + (" " + edg if edg else "")
        out.append(line + "\n")
    return "".join(out).encode()
