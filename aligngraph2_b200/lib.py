"""ctypes binding of libag2_b200.so -- the C ABI declared in include/ag2_b200.h.

There is no CPU path: if the CUDA library is missing or no GPU is present the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
SO = os.environ.get("AG2_B200_LIB") or os.path.join(PKG, "libag2_b200.so")   # the override is for kernel-variant experiments

AG2_OK = 0
ERRORS = {-1: "AG2_ENODEV", -2: "AG2_ECUDA", -3: "AG2_EINVAL", -4: "AG2_ENOMEM", -5: "AG2_ESTATE", -6: "AG2_ECAP"}

# numpy views of the ABI structs (ag2_candidate / ag2_record / ag2_extend_stats)
CANDIDATE_DTYPE = np.dtype([("read", "<i4"), ("strand", "<i4"), ("loc1", "<i8"), ("loc2", "<i4"), ("score", "<i4")])
RECORD_DTYPE = np.dtype([("ok", "<i4"), ("read", "<i4"), ("strand", "<i4"), ("vscore", "<i4"), ("qb", "<i4"),
                         ("qe", "<i4"), ("qs", "<i4"), ("aln_len", "<i4"), ("sb", "<i8"), ("se", "<i8"),
                         ("aln_off", "<i8")])
SEED_CAND_DTYPE = np.dtype([("loc1", "<i8"), ("loc2", "<i8"), ("left1", "<i8"), ("left2", "<i8"), ("right1", "<i8"),
                            ("right2", "<i8"), ("score", "<i4"), ("num1", "<i4"), ("num2", "<i4"), ("chain", "<i4")])
NCODES = 1 << 26
assert CANDIDATE_DTYPE.itemsize == 24 and RECORD_DTYPE.itemsize == 56 and SEED_CAND_DTYPE.itemsize == 64


class ExtendStats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("cells", "rows", "blocks", "aligned", "columns", "wide_chains",
                                        "interior", "launches")] + [("kernel_ms", C.c_double), ("lane_chains", C.c_int64), ("slots", C.c_int64)]


class MapStats(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("total_ms", "seed_ms", "extend_ms", "plan_ms", "rescue_ms", "pass2_ms", "pair_kernel_ms")] + [
        (n, C.c_int64) for n in ("n_reads", "n_records", "n_candidates", "n_rescue", "n_pass2_reads", "seed_overflow1", "seed_overflow2",
                                 "cells", "launches")]


EXPORTS = [
    "ag2_device_count", "ag2_host_alloc", "ag2_host_free", "ag2_ctx_create", "ag2_ctx_destroy", "ag2_last_error", "ag2_version", "ag2_ref_load", "ag2_reads_load", "ag2_reads_load_async", "ag2_reads_wait",
    "ag2_xdrop_extend_batch", "ag2_xdrop_extend_batch_packed", "ag2_extend_fetch_packed", "ag2_expand_alignments", "ag2_extend_upload", "ag2_extend_run", "ag2_extend_fetch", "ag2_extend_get_stats",
    "ag2_ctx_stream", "ag2_index_build", "ag2_index_fetch", "ag2_seed_candidates",
    "ag2_extend_upload_from_seeds", "ag2_map_reads", "ag2_map_fetch", "ag2_map_fetch_packed", "ag2_map_get_stats",
    "ag2_kmer_begin", "ag2_kmer_add_reads", "ag2_kmer_merge", "ag2_kmer_solid", "ag2_kmer_fetch",
]


class Ag2Error(RuntimeError):
    pass


_lib = None


def load() -> C.CDLL:
    """dlopen the in-tree CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO):
        raise Ag2Error(f"{SO} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(nvcc, sm_100a).  aligngraph2_b200 has no CPU fallback.")
    L = C.CDLL(SO)
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int
    L.ag2_device_count.argtypes = [C.POINTER(i32)]
    L.ag2_host_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
    L.ag2_host_free.argtypes = [vp]
    L.ag2_host_free.restype = None
    L.ag2_ctx_create.argtypes = [i32, C.POINTER(vp)]
    L.ag2_ctx_destroy.argtypes = [vp]
    L.ag2_ctx_destroy.restype = None
    L.ag2_last_error.argtypes = [vp]
    L.ag2_last_error.restype = C.c_char_p
    L.ag2_version.restype = C.c_char_p
    L.ag2_ref_load.argtypes = [vp, vp, i64]
    L.ag2_reads_load.argtypes = [vp, vp, vp, i64]
    L.ag2_reads_load_async.argtypes = [vp, vp, vp, i64]
    L.ag2_reads_wait.argtypes = [vp]
    L.ag2_xdrop_extend_batch.argtypes = [vp, vp, i64, vp, vp, vp, i64, C.POINTER(i64)]
    L.ag2_xdrop_extend_batch_packed.argtypes = [vp, vp, i64, vp, vp, i64, C.POINTER(i64)]
    L.ag2_extend_fetch_packed.argtypes = [vp, vp, vp, i64, C.POINTER(i64)]
    L.ag2_expand_alignments.argtypes = [vp, i64, vp, vp, vp, vp, vp, vp, i32]
    L.ag2_extend_upload.argtypes = [vp, vp, i64]
    L.ag2_extend_run.argtypes = [vp]
    L.ag2_extend_fetch.argtypes = [vp, vp, vp, vp, i64, C.POINTER(i64)]
    L.ag2_extend_get_stats.argtypes = [vp, C.POINTER(ExtendStats)]
    L.ag2_index_build.argtypes = [vp, i32, C.c_double, C.c_double]
    L.ag2_index_fetch.argtypes = [vp, vp, vp, vp, vp, i64, C.POINTER(i64), vp, vp, C.POINTER(i64)]
    L.ag2_seed_candidates.argtypes = [vp, i32, i32, vp, vp]
    L.ag2_extend_upload_from_seeds.argtypes = [vp, i32, C.POINTER(i64)]
    L.ag2_map_reads.argtypes = [vp, i32, i32, C.POINTER(i64)]
    L.ag2_map_fetch.argtypes = [vp, vp, vp, vp, i64, C.POINTER(i64)]
    L.ag2_map_fetch_packed.argtypes = [vp, vp, vp, i64, C.POINTER(i64)]
    L.ag2_map_get_stats.argtypes = [vp, C.POINTER(MapStats)]
    L.ag2_kmer_begin.argtypes = [vp, i32]
    L.ag2_kmer_add_reads.argtypes = [vp]
    L.ag2_kmer_merge.argtypes = [vp, vp]
    L.ag2_kmer_solid.argtypes = [vp, C.c_double, C.POINTER(i64), C.POINTER(i64)]
    L.ag2_kmer_fetch.argtypes = [vp, vp, i64]
    L.ag2_ctx_stream.argtypes = [vp]
    L.ag2_ctx_stream.restype = vp
    _lib = L
    return L
