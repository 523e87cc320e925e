"""ctypes bindings for the parity checkers.  TEST INFRASTRUCTURE ONLY.

``Oracle``  -> oracle/libag2_oracle.so (this repo's C restatement; built by oracle/Makefile,
               travels to the GPU box)
``RefLib``  -> oracle/_ref/libref_mecat.so (the unmodified reference sources behind
               oracle/ref_shim.cpp; only exists where /root/reference was available at build time)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package aligngraph2_b200 never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.environ.get("AG2_ORACLE_SO") or os.path.join(HERE, "libag2_oracle.so")   # the override is for the sanitizer build (oracle/asan_check.sh)
REF_SO = os.path.join(HERE, "_ref", "libref_mecat.so")
REF_BIN = os.path.join(HERE, "_ref", "mecat2ref")
REF_KMER_COUNTER = os.path.join(HERE, "_ref", "kmer_counter")

MAX_ALN = 500000  # MC/defs.h:202 MAX_SEQ_SIZE


def build(ref: bool = True) -> None:
    """Build the oracle (.so) and, when /root/reference is present, the reference checkers."""
    subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)
    if ref and os.path.isdir("/root/reference/mecat_plus"):
        subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)


class _Aln(C.Structure):
    _fields_ = [
        ("ok", C.c_int), ("qoff", C.c_int), ("qend", C.c_int), ("toff", C.c_int), ("tend", C.c_int),
        ("aln_size", C.c_int), ("qaln", C.c_void_p), ("taln", C.c_void_p),
        ("cells", C.c_long), ("rows", C.c_long), ("calls", C.c_long),
    ]


def _u8(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint8)


class Oracle:
    """The C restatement (oracle/ag2_oracle.c)."""

    def __init__(self) -> None:
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        self.lib = C.CDLL(ORACLE_SO)
        L = self.lib
        L.orc_xdrop_new.restype = C.c_void_p
        L.orc_xdrop_free.argtypes = [C.c_void_p]
        L.orc_xdrop_block.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                      C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p,
                                      C.POINTER(C.c_int), C.POINTER(C.c_long)]
        L.orc_xdrop_go.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                   C.c_int, C.POINTER(_Aln)]
        L.orc_extend_candidate.argtypes = [C.c_void_p, C.c_char_p, C.c_long, C.c_char_p, C.c_int,
                                           C.c_long, C.c_long, C.POINTER(C.c_long), C.POINTER(_Aln)]
        L.orc_xdrop_counters.argtypes = [C.c_void_p, C.POINTER(C.c_long), C.POINTER(C.c_long),
                                         C.POINTER(C.c_long)]
        self.x = C.c_void_p(L.orc_xdrop_new())

    def __del__(self):
        try:
            self.lib.orc_xdrop_free(self.x)
        except Exception:
            pass

    def block(self, A, M, B, N, forward=True):
        """One block DP.  A, B: uint8 code arrays in *memory* order; for forward=False the block
        starts at the LAST element and runs towards index 0.  Returns (score, ae, be, ops, cells)."""
        A = _u8(A)
        B = _u8(B)
        ops = np.zeros(M + N + 8, dtype=np.uint8)
        ae, be, nops, cells = C.c_int(), C.c_int(), C.c_int(), C.c_long()
        pa = A.ctypes.data + (0 if forward else len(A) - 1)
        pb = B.ctypes.data + (0 if forward else len(B) - 1)
        s = self.lib.orc_xdrop_block(self.x, pa, M, pb, N, int(forward), C.byref(ae), C.byref(be),
                                     ops.ctypes.data, C.byref(nops), C.byref(cells))
        return s, ae.value, be.value, ops[:nops.value].copy(), cells.value

    def go(self, q, qstart, t, tstart, min_aln=1000):
        q = _u8(q)
        t = _u8(t)
        a = _Aln()
        self.lib.orc_xdrop_go(self.x, q.ctypes.data, qstart, len(q), t.ctypes.data, tstart, len(t), min_aln,
                              C.byref(a))
        return _aln_dict(a)

    def extend(self, ref: bytes, read: bytes, loc1: int, loc2: int):
        """extend_candidate on raw ASCII; ``read`` already oriented.  Returns dict with rec."""
        a = _Aln()
        rec = (C.c_long * 4)()
        self.lib.orc_extend_candidate(self.x, ref, len(ref), read, len(read), loc1, loc2, rec, C.byref(a))
        d = _aln_dict(a)
        d.update(qb=rec[0], qe=rec[1], sb=rec[2], se=rec[3])
        return d

    def counters(self):
        c, r, k = C.c_long(), C.c_long(), C.c_long()
        self.lib.orc_xdrop_counters(self.x, C.byref(c), C.byref(r), C.byref(k))
        return c.value, r.value, k.value


def _aln_dict(a: _Aln) -> dict:
    return dict(ok=a.ok, qoff=a.qoff, qend=a.qend, toff=a.toff, tend=a.tend, aln_size=a.aln_size,
                qaln=C.string_at(a.qaln, a.aln_size), taln=C.string_at(a.taln, a.aln_size),
                cells=a.cells, rows=a.rows, calls=a.calls)


def have_ref() -> bool:
    return os.path.exists(REF_SO)


class RefLib:
    """The unmodified reference GapAligner / extend_candidate (oracle/ref_shim.cpp)."""

    def __init__(self) -> None:
        self.lib = C.CDLL(REF_SO)
        L = self.lib
        L.ref_xdrop_new.restype = C.c_void_p
        L.ref_xdrop_free.argtypes = [C.c_void_p]
        L.ref_xdrop_go.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                   C.c_int, C.POINTER(C.c_int), C.c_void_p, C.c_void_p]
        L.ref_xdrop_block.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                      C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p, C.POINTER(C.c_int)]
        L.ref_extend_candidate.argtypes = [C.c_void_p, C.c_char_p, C.c_long, C.c_char_p, C.c_char_p, C.c_int,
                                           C.c_long, C.c_long, C.c_int, C.c_int, C.POINTER(C.c_int),
                                           C.POINTER(C.c_long), C.c_void_p, C.c_void_p]
        self.x = C.c_void_p(L.ref_xdrop_new())
        self._qa = C.create_string_buffer(MAX_ALN)
        self._ta = C.create_string_buffer(MAX_ALN)

    def __del__(self):
        try:
            self.lib.ref_xdrop_free(self.x)
        except Exception:
            pass

    def block(self, A, M, B, N, forward=True):
        # one guard element on each side: the reference reads B[N] (dead value) at the band end
        A = np.concatenate(([0], _u8(A), [0])).astype(np.uint8)
        B = np.concatenate(([0], _u8(B), [0])).astype(np.uint8)
        ops = np.zeros(M + N + 8, dtype=np.uint8)
        ae, be, nops = C.c_int(), C.c_int(), C.c_int()
        pa = A.ctypes.data + (1 if forward else len(A) - 2)
        pb = B.ctypes.data + (1 if forward else len(B) - 2)
        s = self.lib.ref_xdrop_block(self.x, pa, M, pb, N, int(forward), C.byref(ae), C.byref(be),
                                     ops.ctypes.data, C.byref(nops))
        return s, ae.value, be.value, ops[:nops.value].copy()

    def go(self, q, qstart, t, tstart, min_aln=1000):
        q = np.concatenate(([0], _u8(q), [0])).astype(np.uint8)
        t = np.concatenate(([0], _u8(t), [0])).astype(np.uint8)
        out = (C.c_int * 5)()
        ok = self.lib.ref_xdrop_go(self.x, q.ctypes.data + 1, qstart, len(q) - 2, t.ctypes.data + 1, tstart,
                                   len(t) - 2, min_aln, out, self._qa, self._ta)
        n = out[4]
        return dict(ok=ok, qoff=out[0], qend=out[1], toff=out[2], tend=out[3], aln_size=n,
                    qaln=self._qa.raw[:n], taln=self._ta.raw[:n])

    def extend(self, ref: bytes, read: bytes, loc1: int, loc2: int):
        oi = (C.c_int * 4)()
        ol = (C.c_long * 2)()
        ok = self.lib.ref_extend_candidate(self.x, ref, len(ref), read, read, len(read), loc1, loc2, ord("F"), 0,
                                           oi, ol, self._qa, self._ta)
        if not ok:
            return dict(ok=0)
        qa = self._qa.value
        return dict(ok=1, qb=oi[0], qe=oi[1], sb=ol[0], se=ol[1], qaln=qa, taln=self._ta.value, aln_size=len(qa))


class MapperOracle:
    """oracle/ag2_mapper.c: index, votes, seeding, candidates, extension, rescue, second pass -> `.r` records."""

    def __init__(self) -> None:
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        self.lib = C.CDLL(ORACLE_SO)
        self.lib.orc_map_batch.restype = C.c_long
        self.lib.orc_map_batch.argtypes = [C.c_char_p, C.c_long, C.c_char_p, C.c_void_p, C.c_void_p, C.c_long, C.c_int,
                                           C.c_double, C.c_double, C.c_int, C.c_int, C.c_char_p, C.c_void_p]

    @staticmethod
    def upper_ref(genome: bytes) -> bytes:
        """creat_ref_index keeps the FASTA characters and upper-cases those above 'Z' (impl_large.cpp:432-437)."""
        a = np.frombuffer(genome, dtype=np.uint8).copy()
        a[a > 90] -= 32
        return a.tobytes()

    def map_batch(self, genome: bytes, bases: bytes, offsets, ids, r_path: str, cbl=200, alpha=0.5, beta=2.0,
                  maxc=10, num_output=1):
        ref = self.upper_ref(genome)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        st = np.zeros(8, dtype=np.int64)
        n = self.lib.orc_map_batch(ref, len(ref), bases, offsets.ctypes.data, ids.ctypes.data, len(ids), cbl, alpha, beta,
                                   maxc, num_output, r_path.encode(), st.ctypes.data)
        keys = ("cells", "calls", "aligned", "pass2_reads", "insert_loc", "rescue_extensions", "multi_candidate_reads",
                "votes_ne_1")
        return n, dict(zip(keys, (int(v) for v in st)))


class _Index(C.Structure):
    _fields_ = [("R", C.c_long), ("ref", C.c_void_p), ("cbl", C.c_int), ("nblk", C.c_long), ("rcnt", C.c_void_p),
                ("cnt", C.c_void_p), ("off", C.c_void_p), ("pos", C.c_void_p), ("kcount", C.c_void_p), ("vote", C.c_void_p),
                ("ave", C.c_float)]


class _Cand(C.Structure):
    _fields_ = [("loc1", C.c_long), ("loc2", C.c_long), ("left1", C.c_long), ("left2", C.c_long), ("right1", C.c_long),
                ("right2", C.c_long), ("score", C.c_int), ("num1", C.c_int), ("num2", C.c_int), ("chain", C.c_char)]


NCODES = 1 << 26


class IndexOracle:
    """orc_index_build (creat_ref_index + get_vote) and orc_seed_candidates, with numpy views of the arrays."""

    def __init__(self, genome: bytes, bases: bytes, offsets, cbl=200, alpha=0.5, beta=2.0, maxc=10):
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        L = self.lib = C.CDLL(ORACLE_SO)
        L.orc_read_hist13.argtypes = [C.c_char_p, C.c_long, C.c_void_p]
        L.orc_read_index_prefix.restype = C.c_long
        L.orc_read_index_prefix.argtypes = [C.c_void_p, C.c_long]
        L.orc_index_build.restype = C.POINTER(_Index)
        L.orc_index_build.argtypes = [C.c_char_p, C.c_long, C.c_void_p, C.c_int, C.c_double, C.c_double]
        L.orc_index_free.argtypes = [C.POINTER(_Index)]
        L.orc_mapper_new.restype = C.c_void_p
        L.orc_mapper_new.argtypes = [C.POINTER(_Index), C.c_int, C.c_int]
        L.orc_mapper_free.argtypes = [C.c_void_p]
        L.orc_seed_candidates.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.POINTER(_Cand)]
        self.ref = MapperOracle.upper_ref(genome)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        pre = L.orc_read_index_prefix(offsets.ctypes.data, len(offsets) - 1)
        self.rcnt = np.zeros(NCODES, dtype=np.int32)
        L.orc_read_hist13(bases, int(offsets[pre]), self.rcnt.ctypes.data)
        self.ix = L.orc_index_build(self.ref, len(self.ref), self.rcnt.ctypes.data, cbl, alpha, beta)
        ix = self.ix.contents
        self.nblk = ix.nblk
        self.cnt = np.ctypeslib.as_array(C.cast(ix.cnt, C.POINTER(C.c_int32)), (NCODES,))
        self.off = np.ctypeslib.as_array(C.cast(ix.off, C.POINTER(C.c_uint32)), (NCODES + 1,))
        self.pos = np.ctypeslib.as_array(C.cast(ix.pos, C.POINTER(C.c_uint32)), (int(self.off[-1]) + 1,))[:-1]
        self.kcount = np.ctypeslib.as_array(C.cast(ix.kcount, C.POINTER(C.c_int32)), (ix.nblk + 10,))
        self.vote = np.ctypeslib.as_array(C.cast(ix.vote, C.POINTER(C.c_float)), (ix.nblk + 10,))
        self.maxc = maxc
        self.m = C.c_void_p(L.orc_mapper_new(self.ix, maxc, 1))

    def candidates(self, read: bytes, pass_: int = 0):
        out = (_Cand * (self.maxc + 1))()
        n = self.lib.orc_seed_candidates(self.m, read, len(read), pass_, out)
        return [(c.loc1, c.loc2, c.left1, c.left2, c.right1, c.right2, c.score, c.num1, c.num2, ord(c.chain)) for c in out[:n]]

    def __del__(self):
        try:
            self.lib.orc_mapper_free(self.m)
            self.lib.orc_index_free(self.ix)
        except Exception:
            pass


def solid_kmers(bases: bytes, offsets, k: int, threshold: float = 0.2):
    """oracle/ag2_kmer.c: (sorted solid k-mer codes as uint64, min abundance)."""
    if not os.path.exists(ORACLE_SO):
        build(ref=False)
    L = C.CDLL(ORACLE_SO)
    L.orc_solid_kmers.restype = C.c_long
    L.orc_solid_kmers.argtypes = [C.c_char_p, C.c_void_p, C.c_long, C.c_int, C.c_double, C.c_void_p, C.c_long, C.POINTER(C.c_long)]
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    cut = C.c_long()
    n = L.orc_solid_kmers(bases, offsets.ctypes.data, len(offsets) - 1, k, threshold, None, 0, C.byref(cut))
    out = np.empty(n, dtype=np.uint64)
    L.orc_solid_kmers(bases, offsets.ctypes.data, len(offsets) - 1, k, threshold, out.ctypes.data, n, C.byref(cut))
    return out, cut.value


REF_PAGRAPH_DUMP = os.path.join(HERE, "_ref", "pagraph_dump")


def pagraph_dump(d: str, out: str, eps: int = 10, cov: int = 2, solid: str = "solid.bin", ctg: str = "ctg.fasta",
                 ref: str = "ref.fasta", aln: str = "c2r.ref") -> bytes:
    """The restated A-Bruijn build (oracle/ag2_pagraph.cpp) on a pagraph input directory; returns the dump text."""
    if not os.path.exists(ORACLE_SO):
        build(ref=False)
    L = C.CDLL(ORACLE_SO)
    L.ag2o_pagraph_dump.argtypes = [C.c_char_p] * 5 + [C.c_long] * 2 + [C.c_char_p]
    j = lambda n: os.path.join(d, n).encode()
    rc = L.ag2o_pagraph_dump(j(solid), j(ctg), j(ref), d.encode(), j(aln), eps, cov, j(out))
    if rc != 0:
        raise RuntimeError(f"ag2o_pagraph_dump -> {rc}")
    return open(os.path.join(d, out), "rb").read()


def parse_graph_dump(blob: bytes):
    """dump text -> list (one per config) of {vertex idx: (code, [(ctg, ref, count)...], [(to, step)...])}."""
    cfgs = []
    for line in blob.decode().split("\n"):
        if line.startswith("#config"):
            cfgs.append({})
        elif line.startswith("V "):
            t = line.split(" ")
            npos = int(t[4])
            pos = [tuple(int(x) for x in s.split(",")) for s in t[5:5 + npos]]
            nedge = int(t[6 + npos])
            edges = [tuple(int(x) for x in s.split(",")) for s in t[7 + npos:7 + npos + nedge]]
            cfgs[-1][int(t[1])] = (int(t[2]), pos, edges)
    return cfgs


# ---- vanilla MECAT2 DiffAligner ("next"): oracle/ag2_diff.c against oracle/_ref/libref_mecat_vanilla.so ----
REF_DIFF_SO = os.path.join(HERE, "_ref", "libref_mecat_vanilla.so")


def have_diff_ref() -> bool:
    return os.path.exists(REF_DIFF_SO)


def _diff_block_args(Q, T, right_extend):
    """The block as the aligner sees it: a pointer at its first base, running towards lower addresses when not right_extend.
    Returns (buffers kept alive, q pointer, t pointer)."""
    q, t = _u8(Q), _u8(T)
    if right_extend:
        return (q, t), q.ctypes.data, t.ctypes.data
    qr, tr = np.ascontiguousarray(q[::-1]), np.ascontiguousarray(t[::-1])
    return (qr, tr), qr.ctypes.data + max(len(qr) - 1, 0), tr.ctypes.data + max(len(tr) - 1, 0)


class _DiffBase:
    def _block(self, fn, handle, Q, T, right_extend):
        keep, qp, tp = _diff_block_args(Q, T, right_extend)
        out6 = (C.c_int * 6)()
        cap = len(Q) + len(T) + 64
        qs, ts = np.zeros(cap, np.uint8), np.zeros(cap, np.uint8)
        args = [C.c_void_p(qp), len(Q), C.c_void_p(tp), len(T), int(bool(right_extend)), out6, qs.ctypes.data_as(C.c_void_p),
                ts.ctypes.data_as(C.c_void_p)]
        rc = fn(*([handle] + args if handle is not None else args))
        del keep
        n = out6[5]
        return dict(rc=rc, q_s=out6[0], q_e=out6[1], t_s=out6[2], t_e=out6[3], dist=out6[4], n=n, qstr=qs[:n].tobytes(), tstr=ts[:n].tobytes())

    def _go(self, fn, handle, q, qstart, t, tstart, min_aln, extra):
        qq, tt = _u8(q), _u8(t)
        out5 = (C.c_int * 5)()
        cap = len(qq) + len(tt) + 64
        qa, ta = C.create_string_buffer(cap), C.create_string_buffer(cap)
        args = [qq.ctypes.data_as(C.c_void_p), int(qstart), len(qq), tt.ctypes.data_as(C.c_void_p), int(tstart), len(tt), int(min_aln)] + extra + [out5, qa, ta]
        ok = fn(*([handle] + args if handle is not None else args))
        n = out5[4]
        return dict(ok=int(ok), qoff=out5[0], qend=out5[1], toff=out5[2], tend=out5[3], aln_size=n, qaln=qa.raw[:n], taln=ta.raw[:n])


class DiffOracle(_DiffBase):
    """oracle/ag2_diff.c (codes 0..3 in)."""

    def __init__(self) -> None:
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        self.lib = C.CDLL(ORACLE_SO)
        self.lib.orc_diff_block.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_void_p, C.c_void_p]
        self.lib.orc_diff_go.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int),
                                         C.c_char_p, C.c_char_p]

    def block(self, Q, T, right_extend=True):
        return self._block(self.lib.orc_diff_block, None, Q, T, right_extend)

    def go(self, q, qstart, t, tstart, min_aln=0, large_block=0):
        return self._go(self.lib.orc_diff_go, None, q, qstart, t, tstart, min_aln, [int(large_block)])


class DiffRef(_DiffBase):
    """The unmodified vanilla-MECAT2 DiffAligner (oracle/ref_diff_shim.cpp)."""

    def __init__(self, large_block: int = 0) -> None:
        self.lib = C.CDLL(REF_DIFF_SO)
        L = self.lib
        L.ref_diff_new.restype = C.c_void_p
        L.ref_diff_new.argtypes = [C.c_int]
        L.ref_diff_free.argtypes = [C.c_void_p]
        L.ref_diff_block.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_void_p, C.c_void_p]
        L.ref_diff_go.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int),
                                  C.c_char_p, C.c_char_p]
        self.x = C.c_void_p(L.ref_diff_new(int(large_block)))

    def close(self) -> None:
        if self.x:
            self.lib.ref_diff_free(self.x)
            self.x = None

    def block(self, Q, T, right_extend=True):
        return self._block(self.lib.ref_diff_block, self.x, Q, T, right_extend)

    def go(self, q, qstart, t, tstart, min_aln=0):
        return self._go(self.lib.ref_diff_go, self.x, q, qstart, t, tstart, min_aln, [])
