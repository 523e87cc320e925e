"""Host-side mirror of the reference's mecat2ref+ extension interface, over the C ABI.

Reference seams (paths under mecat_plus/MECAT-master_1/src/):
  * ``GapAligner::go`` / ``XdropAligner``      common/gapalign.h:4-25, common/xdrop_gapalign.cpp:359-439
  * ``extend_candidate``                        mecat2ref/mecat2ref_aux.cpp:210-270
  * ``candidate_save`` / ``TempResult``         mecat2ref/mecat2ref_defs.h:90-95, mecat2ref/output.h

``Mecat2RefDevice`` owns one GPU context: load the (concatenated, upper-cased) reference once, load a
batch of reads, then extend any number of candidates.  Results come back as ``TempResult``-like
records; alignment strings are ASCII ``ACGT-`` exactly as the reference writes them.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np

from . import lib as _lib
from .lib import CANDIDATE_DTYPE, NCODES, RECORD_DTYPE, SEED_CAND_DTYPE, Ag2Error, ExtendStats, MapStats


def _as_bytes_array(x) -> np.ndarray:
    if isinstance(x, (bytes, bytearray)):
        return np.frombuffer(x, dtype=np.uint8)
    return np.ascontiguousarray(x, dtype=np.uint8)


class Mecat2RefDevice:
    """One GPU's resident state for the mecat2ref+ hot path (reference + read batch)."""

    def __init__(self, device: int = 0):
        self._L = _lib.load()
        self._ctx = C.c_void_p()
        rc = self._L.ag2_ctx_create(device, C.byref(self._ctx))
        if rc != 0:
            raise Ag2Error(f"ag2_ctx_create(device={device}) failed: {_lib.ERRORS.get(rc, rc)} "
                           "(a CUDA device is required; there is no CPU fallback)")
        self.device = device
        self.n_reads = 0
        self.ref_len = 0
        self._n_cand = 0

    def close(self) -> None:
        if self._ctx:
            self._L.ag2_ctx_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str) -> None:
        if rc != 0:
            msg = self._L.ag2_last_error(self._ctx).decode(errors="replace")
            raise Ag2Error(f"{what}: {_lib.ERRORS.get(rc, rc)}: {msg}")

    # -- residency ---------------------------------------------------------------------------
    def load_reference(self, ref) -> None:
        """``ref``: ASCII bases of the concatenated reference (creat_ref_index, impl_large.cpp:432-437)."""
        a = _as_bytes_array(ref)
        self._check(self._L.ag2_ref_load(self._ctx, a.ctypes.data, a.size), "ag2_ref_load")
        self.ref_len = int(a.size)

    def load_reads(self, reads: Sequence | None = None, *, bases=None, offsets=None) -> None:
        """Either a sequence of ASCII reads, or a concatenated ``bases`` buffer with ``offsets`` (n+1)."""
        if reads is not None:
            arrs = [_as_bytes_array(r) for r in reads]
            offsets = np.zeros(len(arrs) + 1, dtype=np.int64)
            np.cumsum([a.size for a in arrs], out=offsets[1:])
            bases = np.concatenate(arrs) if arrs else np.zeros(0, dtype=np.uint8)
        bases = _as_bytes_array(bases)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = offsets.size - 1
        self._check(self._L.ag2_reads_load(self._ctx, bases.ctypes.data, offsets.ctypes.data, n), "ag2_reads_load")
        self.n_reads = n

    def load_reads_async(self, bases, offsets) -> None:
        """ag2_reads_load_async: returns once the copies are queued; ``bases`` (ideally pinned) must stay alive and unchanged
        until a later call on this device has returned.  extend_batch_into() then starts on the first reads while the
        rest is still being copied."""
        bases = _as_bytes_array(bases)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = offsets.size - 1
        self._pending_bases = bases
        self._check(self._L.ag2_reads_load_async(self._ctx, bases.ctypes.data, offsets.ctypes.data, n), "ag2_reads_load_async")
        self.n_reads = n

    def wait_reads(self) -> None:
        self._check(self._L.ag2_reads_wait(self._ctx), "ag2_reads_wait")
        self._pending_bases = None

    # -- index + seeding -------------------------------------------------------------------------
    def build_index(self, cbl: int = 200, alpha: float = 0.5, beta: float = 2.0) -> None:
        """build_read_index + creat_ref_index + get_vote (impl_large.cpp:258-608) for the loaded reference and
        read batch; cbl / alpha / beta are mecat2ref+'s -z / -l / -u."""
        self._check(self._L.ag2_index_build(self._ctx, cbl, alpha, beta), "ag2_index_build")

    def fetch_index(self) -> dict:
        """Host copies of the device index (tests / inspection)."""
        n_pos, nblk = C.c_int64(), C.c_int64()
        self._check(self._L.ag2_index_fetch(self._ctx, None, None, None, None, 0, C.byref(n_pos), None, None, C.byref(nblk)),
                    "ag2_index_fetch")
        d = dict(rcnt=np.empty(NCODES, np.int32), cnt=np.empty(NCODES, np.int32), off=np.empty(NCODES + 1, np.uint32),
                 pos=np.empty(max(n_pos.value, 1), np.uint32), kcount=np.empty(nblk.value + 10, np.int32),
                 vote=np.empty(nblk.value + 10, np.float32))
        self._check(self._L.ag2_index_fetch(self._ctx, d["rcnt"].ctypes.data, d["cnt"].ctypes.data, d["off"].ctypes.data,
                                            d["pos"].ctypes.data, d["pos"].size, C.byref(n_pos), d["kcount"].ctypes.data,
                                            d["vote"].ctypes.data, C.byref(nblk)), "ag2_index_fetch")
        d["pos"] = d["pos"][:n_pos.value]
        d["nblk"] = nblk.value
        return d

    def seed_candidates(self, pass_: int = 0, maxc: int = 10):
        """Seeding + candidate scoring of every loaded read (reference_mapping :776-991; pass_=1 is the
        reference's second pass).  Returns (cands[n_reads, maxc] of SEED_CAND_DTYPE, ncand[n_reads])."""
        cands = np.zeros((self.n_reads, maxc), dtype=SEED_CAND_DTYPE)
        ncand = np.zeros(self.n_reads, dtype=np.int32)
        self._check(self._L.ag2_seed_candidates(self._ctx, pass_, maxc, cands.ctypes.data, ncand.ctypes.data),
                    "ag2_seed_candidates")
        return cands, ncand

    def extend_seed_candidates(self, maxc: int = 10):
        """extend_candidate over every candidate of the last seed_candidates(maxc=...) call, without a host round
        trip for the candidates.  Returns (records, qaln, saln) like extend()."""
        n = C.c_int64()
        self._check(self._L.ag2_extend_upload_from_seeds(self._ctx, maxc, C.byref(n)), "ag2_extend_upload_from_seeds")
        rec = np.zeros(n.value, dtype=RECORD_DTYPE)
        if n.value == 0:
            return rec, np.zeros(0, np.uint8), np.zeros(0, np.uint8)
        self._n_cand = n.value
        self.run()
        used = C.c_int64()
        self._check(self._L.ag2_extend_fetch(self._ctx, rec.ctypes.data, None, None, 0, C.byref(used)), "ag2_extend_fetch")
        qa = np.empty(max(used.value, 1), dtype=np.uint8)
        sa = np.empty(max(used.value, 1), dtype=np.uint8)
        self._check(self._L.ag2_extend_fetch(self._ctx, rec.ctypes.data, qa.ctypes.data, sa.ctypes.data, qa.size, C.byref(used)),
                    "ag2_extend_fetch")
        return rec, qa[:used.value], sa[:used.value]

    def map_reads(self, maxc: int = 10, num_output: int = 1):
        """reference_mapping()'s loop body for the whole batch (seeding, candidates, extension, rescue, second pass,
        output choice).  Returns (records, qaln, saln): the records in thread-file order."""
        n = C.c_int64()
        self._check(self._L.ag2_map_reads(self._ctx, maxc, num_output, C.byref(n)), "ag2_map_reads")
        rec = np.zeros(n.value, dtype=RECORD_DTYPE)
        used = C.c_int64()
        self._check(self._L.ag2_map_fetch(self._ctx, rec.ctypes.data, None, None, 0, C.byref(used)), "ag2_map_fetch")
        qa = np.empty(max(used.value, 1), dtype=np.uint8)
        sa = np.empty(max(used.value, 1), dtype=np.uint8)
        self._check(self._L.ag2_map_fetch(self._ctx, rec.ctypes.data, qa.ctypes.data, sa.ctypes.data, qa.size, C.byref(used)),
                    "ag2_map_fetch")
        return rec, qa[:used.value], sa[:used.value]

    def map_reads_only(self, maxc: int = 10, num_output: int = 1) -> int:
        """ag2_map_reads without the fetch: results stay on the device; returns the number of records."""
        n = C.c_int64()
        self._check(self._L.ag2_map_reads(self._ctx, maxc, num_output, C.byref(n)), "ag2_map_reads")
        return n.value

    def map_fetch_into(self, rec: np.ndarray, qaln_out=None, saln_out=None) -> int:
        """ag2_map_fetch into caller-owned (ideally pinned) buffers; returns the bytes used per string."""
        used = C.c_int64()
        q = qaln_out.ctypes.data if qaln_out is not None else None
        t = saln_out.ctypes.data if saln_out is not None else None
        cap = qaln_out.size if qaln_out is not None else 0
        self._check(self._L.ag2_map_fetch(self._ctx, rec.ctypes.data, q, t, cap, C.byref(used)), "ag2_map_fetch")
        return used.value

    def map_fetch_packed_into(self, rec: np.ndarray, ops_out: np.ndarray) -> int:
        """ag2_map_fetch_packed into caller-owned buffers (ops_out: uint32 words, 16 columns each); returns the columns used."""
        used = C.c_int64()
        self._check(self._L.ag2_map_fetch_packed(self._ctx, rec.ctypes.data, ops_out.ctypes.data, ops_out.size * 16, C.byref(used)),
                    "ag2_map_fetch_packed")
        return used.value

    def map_stats(self) -> dict:
        s = MapStats()
        self._check(self._L.ag2_map_get_stats(self._ctx, C.byref(s)), "ag2_map_get_stats")
        return {k: getattr(s, k) for k, _ in MapStats._fields_}

    @staticmethod
    def write_thread_file(path: str, rec, qaln, saln, read_ids) -> None:
        """output_temp_result (output.cpp:237-251) for every record: the `<wrk>/N.r` text."""
        with open(path, "wb") as f:
            for r in rec:
                o, n = int(r["aln_off"]), int(r["aln_len"])
                f.write(b"%d\t%c\t%d\t%d\t%d\t%d\t%d\t%d\n" % (read_ids[r["read"]], b"FR"[r["strand"]], r["vscore"], r["qb"], r["qe"],
                                                              r["qs"], r["sb"], r["se"]))
                f.write(qaln[o:o + n].tobytes() + b"\n" + saln[o:o + n].tobytes() + b"\n")

    # -- PAGraph kmer_counter ------------------------------------------------------------------------
    def solid_kmers(self, k: int = 14, threshold: float = 0.2):
        """kmerCounter (PAGraph/src/main/kmer_counter.cpp:19-96) over the loaded read batch: returns
        (ascending solid k-mer codes as uint64, min abundance)."""
        self._check(self._L.ag2_kmer_begin(self._ctx, k), "ag2_kmer_begin")
        self._check(self._L.ag2_kmer_add_reads(self._ctx), "ag2_kmer_add_reads")
        cut, n = C.c_int64(), C.c_int64()
        self._check(self._L.ag2_kmer_solid(self._ctx, threshold, C.byref(cut), C.byref(n)), "ag2_kmer_solid")
        out = np.empty(max(n.value, 1), dtype=np.uint64)
        self._check(self._L.ag2_kmer_fetch(self._ctx, out.ctypes.data, out.size), "ag2_kmer_fetch")
        return out[:n.value], cut.value

    # -- extension ---------------------------------------------------------------------------
    @staticmethod
    def make_candidates(read, strand, loc1, loc2, score=None) -> np.ndarray:
        n = len(read)
        c = np.zeros(n, dtype=CANDIDATE_DTYPE)
        c["read"], c["strand"], c["loc1"], c["loc2"] = read, strand, loc1, loc2
        c["score"] = 0 if score is None else score
        return c

    def extend(self, cand: np.ndarray, qaln_out: np.ndarray | None = None, saln_out: np.ndarray | None = None):
        """extend_candidate over a batch.  Returns (records, qaln, saln): records is a RECORD_DTYPE
        array (one per candidate, ``ok`` as GapAligner::go returns), strings are uint8 ASCII buffers
        indexed by ``aln_off``/``aln_len``."""
        cand = np.ascontiguousarray(cand, dtype=CANDIDATE_DTYPE)
        n = cand.size
        rec = np.zeros(n, dtype=RECORD_DTYPE)
        self.upload_candidates(cand)
        self.run()
        used = C.c_int64()
        self._check(self._L.ag2_extend_fetch(self._ctx, rec.ctypes.data, None, None, 0, C.byref(used)), "ag2_extend_fetch")
        if qaln_out is None or qaln_out.size < used.value:
            qaln_out = np.empty(max(used.value, 1), dtype=np.uint8)
        if saln_out is None or saln_out.size < used.value:
            saln_out = np.empty(max(used.value, 1), dtype=np.uint8)
        self._check(self._L.ag2_extend_fetch(self._ctx, rec.ctypes.data, qaln_out.ctypes.data, saln_out.ctypes.data,
                                             qaln_out.size, C.byref(used)), "ag2_extend_fetch")
        return rec, qaln_out[:used.value], saln_out[:used.value]

    def extend_batch_into(self, cand: np.ndarray, rec: np.ndarray, qaln_out: np.ndarray, saln_out: np.ndarray) -> int:
        """The single C-ABI call (ag2_xdrop_extend_batch) with caller-owned host buffers."""
        used = C.c_int64()
        self._check(self._L.ag2_xdrop_extend_batch(self._ctx, cand.ctypes.data, cand.size, rec.ctypes.data,
                                                   qaln_out.ctypes.data, saln_out.ctypes.data, qaln_out.size,
                                                   C.byref(used)), "ag2_xdrop_extend_batch")
        return used.value

    def extend_batch_packed_into(self, cand: np.ndarray, rec: np.ndarray, ops_out: np.ndarray) -> int:
        """ag2_xdrop_extend_batch_packed: records + 2-bit alignment ops (uint32 words, 16 columns each) into caller-owned host
        buffers; returns the columns used."""
        used = C.c_int64()
        self._check(self._L.ag2_xdrop_extend_batch_packed(self._ctx, cand.ctypes.data, cand.size, rec.ctypes.data, ops_out.ctypes.data,
                                                          ops_out.size * 16, C.byref(used)), "ag2_xdrop_extend_batch_packed")
        return used.value

    def fetch_packed(self, n: int):
        """ag2_extend_fetch_packed after run(): (records, ops words, columns)."""
        rec = np.zeros(n, dtype=RECORD_DTYPE)
        used = C.c_int64()
        self._check(self._L.ag2_extend_fetch_packed(self._ctx, rec.ctypes.data, None, 0, C.byref(used)), "ag2_extend_fetch_packed")
        ops = np.zeros((used.value + 15) // 16 + 1, dtype=np.uint32)
        self._check(self._L.ag2_extend_fetch_packed(self._ctx, rec.ctypes.data, ops.ctypes.data, ops.size * 16, C.byref(used)),
                    "ag2_extend_fetch_packed")
        return rec, ops, used.value

    def expand_alignments(self, rec: np.ndarray, ops: np.ndarray, bases: np.ndarray, offsets: np.ndarray, ref: np.ndarray, columns: int,
                          threads: int = 4):
        """ag2_expand_alignments (host only): the two ASCII strings of every ok record from the ops."""
        qa = np.zeros(max(columns, 1), dtype=np.uint8)
        sa = np.zeros(max(columns, 1), dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        rc = self._L.ag2_expand_alignments(rec.ctypes.data, rec.size, ops.ctypes.data, bases.ctypes.data, offsets.ctypes.data,
                                           ref.ctypes.data, qa.ctypes.data, sa.ctypes.data, threads)
        if rc != 0:
            raise Ag2Error(f"ag2_expand_alignments: {_lib.ERRORS.get(rc, rc)}")
        return qa, sa

    def upload_candidates(self, cand: np.ndarray) -> None:
        cand = np.ascontiguousarray(cand, dtype=CANDIDATE_DTYPE)
        self._check(self._L.ag2_extend_upload(self._ctx, cand.ctypes.data, cand.size), "ag2_extend_upload")
        self._n_cand = cand.size

    def run(self) -> None:
        """Launch the extension kernels on resident data (no host<->device traffic but scalars)."""
        self._check(self._L.ag2_extend_run(self._ctx), "ag2_extend_run")

    def stats(self) -> dict:
        s = ExtendStats()
        self._check(self._L.ag2_extend_get_stats(self._ctx, C.byref(s)), "ag2_extend_get_stats")
        return {k: getattr(s, k) for k, _ in ExtendStats._fields_}

    @property
    def stream(self) -> int:
        return int(self._L.ag2_ctx_stream(self._ctx) or 0)


def record_strings(rec, qaln: np.ndarray, saln: np.ndarray):
    """(qmap, smap) bytes of one record."""
    o, n = int(rec["aln_off"]), int(rec["aln_len"])
    return qaln[o:o + n].tobytes(), saln[o:o + n].tobytes()
