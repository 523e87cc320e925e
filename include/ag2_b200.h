/*
 * ag2_b200.h -- C ABI of the B200-native mecat2ref+ / PAGraph hot path (libag2_b200.so).
 *
 * The reference (Godotcoffee/AlignGraph2) has no FFI layer: its hot path sits behind process
 * boundaries (AlignGraph2.py:265-277 spawns mecat2ref+).  In-process, the narrowest seams are
 *   - GapAligner::go            mecat_plus/MECAT-master_1/src/common/gapalign.h:4-25
 *   - extend_candidate          mecat_plus/MECAT-master_1/src/mecat2ref/mecat2ref_aux.cpp:210-270
 *   - meap_ref_impl_large       mecat_plus/MECAT-master_1/src/mecat2ref/mecat2ref.cpp:602,995
 * and the entry points below cut exactly there.  INTEGRATION.md shows the
 * patch a maintainer would apply to reference_mapping() to call them.
 *
 * Conventions: extern "C"; plain pointers and sizes; caller-owned HOST buffers unless a name says
 * `_dev`; every function returns 0 on success or a negative AG2_E* code (ag2_last_error() gives
 * the text); no exceptions cross the boundary; one ag2_ctx per GPU, not thread-safe.
 * There is no CPU fallback: without a CUDA device ag2_ctx_create fails with AG2_ENODEV.
 */
#ifndef AG2_B200_H
#define AG2_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AG2_OK 0
#define AG2_ENODEV (-1)   /* no usable CUDA device */
#define AG2_ECUDA (-2)    /* CUDA runtime error */
#define AG2_EINVAL (-3)   /* bad argument */
#define AG2_ENOMEM (-4)   /* device or host allocation failed */
#define AG2_ESTATE (-5)   /* call order (e.g. extend before reference/reads are loaded) */
#define AG2_ECAP (-6)     /* caller buffer too small */

typedef struct ag2_ctx ag2_ctx;

/* One candidate = one call of extend_candidate (mecat2ref_aux.cpp:210): `candidate_save`
 * (mecat2ref_defs.h:90-95) reduced to the fields the extension reads. */
typedef struct ag2_candidate {
    int32_t read;    /* index into the loaded read batch */
    int32_t strand;  /* 0 = 'F' (read as given), 1 = 'R' (reverse complement, impl_large.cpp:799-833) */
    int64_t loc1;    /* 1-based reference position of the seed (candidate_save::loc1) */
    int32_t loc2;    /* 0-based position in the oriented read   (candidate_save::loc2) */
    int32_t score;   /* candidate_save::score, copied to the record (TempResult::vscore) */
} ag2_candidate;

/* One extension result = TempResult (output.h) without the strings.  ok mirrors the return of
 * GapAligner::go (qe - qb >= 1000, mecat2ref_aux.cpp:240).  Fields other than ok are only
 * meaningful when ok != 0; aln_off/aln_len locate the two alignment strings. */
typedef struct ag2_record {
    int32_t ok;
    int32_t read;
    int32_t strand;
    int32_t vscore;
    int32_t qb, qe, qs;   /* strand-local read coordinates, read length */
    int32_t aln_len;      /* alignment columns (strings are NOT NUL terminated) */
    int64_t sb, se;       /* global (concatenated) reference coordinates */
    int64_t aln_off;      /* offset of this record's columns in qaln_out / saln_out */
} ag2_record;

/* Device-side work counters of the last extend call (for the roofline: SURVEY.md 8d). */
typedef struct ag2_extend_stats {
    int64_t cells;        /* DP cell evaluations == inner-loop iterations of xdrop_align */
    int64_t rows;         /* DP rows */
    int64_t blocks;       /* xdrop_align calls */
    int64_t aligned;      /* sum of (qe - qb) over ok records */
    int64_t columns;      /* alignment columns emitted over ok records */
    int64_t wide_chains;  /* extension directions rerun on the wide (row-parallel) kernel */
    int64_t interior;     /* wide kernel: rows that needed the interior-pruned-cell fix-up */
    int64_t launches;     /* kernel launches made by the call */
    double kernel_ms;     /* CUDA-event time of the dominant kernel (xdrop_pair_kernel) */
    int64_t lane_chains;  /* extension directions the pair kernel handed on (to the wide kernel when few, else to the lane kernel) */
    int64_t slots;        /* window slots the pair kernel evaluated (8 per executed group for each of a warp's 64 directions, running or not); cells / slots = how full they were */
} ag2_extend_stats;

/* Number of CUDA devices this process sees.  The hosts shard reads over them, one ag2_ctx and one host thread per
 * device (the counterpart of the reference's -t worker threads, mecat2ref_impl_large.cpp:2072-2087, which share one
 * index; here every device holds its own copy). */
int ag2_device_count(int *count);
int ag2_ctx_create(int device, ag2_ctx **ctx);
void ag2_ctx_destroy(ag2_ctx *ctx);
const char *ag2_last_error(const ag2_ctx *ctx);
const char *ag2_version(void);

/* Page-locked host memory for the buffers of the calls below (records, alignment strings, read bases): device copies from
 * and to such memory run at full PCIe speed and do not stall on page faults of fresh allocations.  Plain malloc'ed
 * buffers work everywhere too. */
int ag2_host_alloc(size_t bytes, void **ptr);
void ag2_host_free(void *ptr);

/* Reference: ASCII bases, already concatenated over chromosomes and upper-cased the way
 * creat_ref_index does (impl_large.cpp:432-437).  Packed 2 bits/base on the device. */
int ag2_ref_load(ag2_ctx *ctx, const char *ref, int64_t ref_len);

/* Read batch: ASCII bases, reads concatenated; offs has n+1 entries (offs[0] = 0).  Packed 2
 * bits/base on the device with a side bitmask for bases the reverse strand must not complement
 * (anything but upper-case ACGT). */
int ag2_reads_load(ag2_ctx *ctx, const char *bases, const int64_t *offs, int64_t n_reads);

/* The same batch, but the call returns as soon as the copies are queued: the bases go up in pieces on a copy stream.
 * `bases` must stay valid and unchanged until ag2_reads_wait() or any later call on this context has returned (`offs` is
 * consumed before the call returns).  ag2_xdrop_extend_batch, given candidates in read order, starts extending the first
 * reads while the later ones are still in flight (the reference has no counterpart: its load_fastq batch,
 * mecat2ref_impl_large.cpp:1965-1991, is in host memory already).  Every other entry point waits for the whole batch. */
int ag2_reads_load_async(ag2_ctx *ctx, const char *bases, const int64_t *offs, int64_t n_reads);
int ag2_reads_wait(ag2_ctx *ctx);

/* extend_candidate x n.  rec_out[n]; qaln_out/saln_out receive ASCII ACGT- columns of the ok
 * records back to back (record i at [aln_off, aln_off + aln_len)); aln_cap is the capacity of
 * each string buffer, *aln_used the bytes written.  AG2_ECAP if too small (rec_out is still
 * filled, so the caller can size the buffers and call ag2_extend_fetch).
 * Results travel home while later candidates are still being extended: one kernel launch covers the whole batch and
 * raises a flag per finished output chunk ("streamed" form; after ag2_reads_load_async the kernel also starts before
 * the reads are all up).  Environment AG2_E2E_PATH=chunked (read per call) selects one launch per output chunk instead,
 * which needs device workspace for a chunk only; the library falls back to it by itself when the batch's workspace
 * does not fit.  The bytes written are the same. */
int ag2_xdrop_extend_batch(ag2_ctx *ctx, const ag2_candidate *cand, int64_t n, ag2_record *rec_out,
                           char *qaln_out, char *saln_out, int64_t aln_cap, int64_t *aln_used);

/* The same call with the alignments as 2-BIT OPS instead of two ASCII strings: column j of the pool is bits 2 * (j % 16) of
 * ops_out[j / 16]; 0 = a base of both sequences, 1 = '-' in the read string, 2 = '-' in the reference string.  Together with
 * the read and the reference (which the caller has) this is the whole content of TempResult::qmap / smap at an eighth of
 * the bytes that cross PCIe; ag2_expand_alignments rebuilds the strings on the host when they are needed (bit-identical to
 * what ag2_xdrop_extend_batch returns).  cap_columns = capacity of ops_out in columns (16 per word); aln_off of the
 * records indexes columns; output chunks start on word boundaries, so *columns_used includes a little padding. */
int ag2_xdrop_extend_batch_packed(ag2_ctx *ctx, const ag2_candidate *cand, int64_t n, ag2_record *rec_out, uint32_t *ops_out,
                                  int64_t cap_columns, int64_t *columns_used);
int ag2_extend_fetch_packed(ag2_ctx *ctx, ag2_record *rec_out, uint32_t *ops_out, int64_t cap_columns, int64_t *columns_used);
/* Host-only (no GPU work): qaln_out / saln_out [aln_off, aln_off + aln_len) of every ok record from the ops, the ASCII reads
 * of the batch (as given to ag2_reads_load) and the reference (as given to ag2_ref_load); `threads` host threads. */
int ag2_expand_alignments(const ag2_record *rec, int64_t n, const uint32_t *ops, const char *read_bases, const int64_t *read_offs,
                          const char *ref, char *qaln_out, char *saln_out, int threads);

/* Split form of the same call for callers that keep inputs resident in HBM:
 * _upload copies candidates, _run launches the kernels on resident data (no host traffic),
 * _fetch copies records and strings back. */
int ag2_extend_upload(ag2_ctx *ctx, const ag2_candidate *cand, int64_t n);
int ag2_extend_run(ag2_ctx *ctx);
int ag2_extend_fetch(ag2_ctx *ctx, ag2_record *rec_out, char *qaln_out, char *saln_out, int64_t aln_cap,
                     int64_t *aln_used);
int ag2_extend_get_stats(ag2_ctx *ctx, ag2_extend_stats *out);

/* ---- index build: build_read_index + creat_ref_index + get_vote (mecat2ref_impl_large.cpp:258-608) ----
 * Uses the loaded reference and the loaded read batch (its first <= 100 000 reads / 1e9 characters, as the
 * reference does).  cbl / alpha / beta are mecat2ref+'s -z / -l / -u.  The index stays on the device. */
int ag2_index_build(ag2_ctx *ctx, int cbl, double alpha, double beta);

/* Copies of the device index for inspection; any pointer may be NULL.  rcnt/cnt: 4^13 masked counts
 * (countin1 / countin), off: 4^13+1 CSR offsets, pos: *n_pos 1-based k-mer starts (allloc), kcount/vote:
 * *nblk + 10 entries (sim::k_count / sim::vote). */
int ag2_index_fetch(ag2_ctx *ctx, int32_t *rcnt, int32_t *cnt, uint32_t *off, uint32_t *pos, int64_t pos_cap,
                    int64_t *n_pos, int32_t *kcount, float *vote, int64_t *nblk);

/* candidate_save (mecat2ref_defs.h:90-95) */
typedef struct ag2_seed_candidate {
    int64_t loc1, loc2, left1, left2, right1, right2;
    int32_t score, num1, num2;
    int32_t chain;   /* 'F' or 'R' */
} ag2_seed_candidate;

/* Seeding + candidate scoring of every loaded read (reference_mapping, mecat2ref_impl_large.cpp:776-991;
 * pass = 1 is the reference's second pass :1049-1260).  out[r * maxc + i], i < ncand[r], in canidate_loc[]
 * order; maxc is mecat2ref+'s -n (<= 16). */
int ag2_seed_candidates(ag2_ctx *ctx, int pass, int maxc, ag2_seed_candidate *out, int32_t *ncand);

/* Makes every seed candidate of the last ag2_seed_candidates call (same maxc) an extension candidate, on the
 * device (reads in order, canidate_loc[] order inside a read); then ag2_extend_run / ag2_extend_fetch. */
int ag2_extend_upload_from_seeds(ag2_ctx *ctx, int maxc, int64_t *n_candidates);

/* ---- the whole per-read path: the loop body of reference_mapping() (mecat2ref_impl_large.cpp:776-1316) ----
 * for every loaded read: seeding, candidate scoring, extension of every candidate, rescue_clipped_align,
 * output_results, and the second pass for reads without any alignment.  maxc / num_output are mecat2ref+'s
 * -n / -b.  The records come back in thread-file order (read order, output_results order inside a read):
 * exactly the sequence of output_temp_result() calls of a `-t 1` run.  ag2_record.read is the index of the read
 * in the loaded batch.  Needs ag2_index_build. */
int ag2_map_reads(ag2_ctx *ctx, int maxc, int num_output, int64_t *n_records);

/* Where the last ag2_map_reads call spent its time (host wall clock between the stages' synchronisation points) and
 * what it processed; the counterpart of the reference's "The Mapping Time" line (mecat2ref_impl_large.cpp:2133-2138),
 * per stage. */
typedef struct ag2_map_stats {
    double total_ms;
    double seed_ms;        /* seeding + candidate scoring (A5-A7), first pass */
    double extend_ms;      /* extend_candidate over every candidate (A8-A10) */
    double plan_ms;        /* rescue_clipped_align's searches (A11) */
    double rescue_ms;      /* extension of the rescue candidates, linking, output choice (A11-A12) */
    double pass2_ms;       /* everything of the second pass (:1049-1315) */
    double pair_kernel_ms; /* CUDA-event time of xdrop_pair_kernel, all launches of the call */
    int64_t n_reads, n_records, n_candidates, n_rescue, n_pass2_reads;
    int64_t seed_overflow1, seed_overflow2; /* reads handed from the first seeding launch to the second / to the thread path */
    int64_t cells;         /* DP cells of all extensions */
    int64_t launches;      /* kernel launches of the extension batches */
} ag2_map_stats;
int ag2_map_get_stats(ag2_ctx *ctx, ag2_map_stats *out);
int ag2_map_fetch(ag2_ctx *ctx, ag2_record *rec_out, char *qaln_out, char *saln_out, int64_t aln_cap, int64_t *aln_used);
/* the same with the alignments as 2-bit ops (see ag2_xdrop_extend_batch_packed, ag2_expand_alignments) */
int ag2_map_fetch_packed(ag2_ctx *ctx, ag2_record *rec_out, uint32_t *ops_out, int64_t cap_columns, int64_t *columns_used);

/* ---- PAGraph kmer_counter (PAGraph/src/main/kmer_counter.cpp:19-96) ----
 * ag2_kmer_begin zeroes the 4^k abundance table (k <= 15); ag2_kmer_add_reads adds every k-mer of the loaded read
 * batch (call ag2_reads_load + ag2_kmer_add_reads per batch); ag2_kmer_solid applies the cut -- the smallest abundance
 * a with 1 - bins(abundance <= a) / 4^k <= threshold -- and selects the k-mers with abundance >= cut;
 * ag2_kmer_fetch copies their codes (2 bits per base, first base most significant, A0 C1 G2 T3) in ascending order,
 * the order `kmer_counter -t 1` writes. */
int ag2_kmer_begin(ag2_ctx *ctx, int k);
int ag2_kmer_add_reads(ag2_ctx *ctx);
/* dst's table += src's table (both after ag2_kmer_begin with the same k): the merge of the per-GPU histograms when the read
 * batches were spread over several GPUs; reads src's table over NVLink peer memory. */
int ag2_kmer_merge(ag2_ctx *dst, ag2_ctx *src);
int ag2_kmer_solid(ag2_ctx *ctx, double threshold, int64_t *min_abundance, int64_t *n_solid);
int ag2_kmer_fetch(ag2_ctx *ctx, uint64_t *codes_out, int64_t cap);

/* The CUDA stream the context launches on (cudaStream_t as void*), for callers that time with
 * their own events. */
void *ag2_ctx_stream(ag2_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif
