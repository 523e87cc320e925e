/*
 * ag2_pagraph.h -- C ABI of the B200-native PAGraph A-Bruijn graph build (libag2_b200.so), SURVEY.md 8a rows B2-B8.
 *
 * The reference (Godotcoffee/AlignGraph2, PAGraph/) has no FFI layer; `pagraph` is spawned as a process
 * (AlignGraph2.py:414-427).  In-process the build is the object pair
 *   PABruijnGraph        PAGraph/src/tools/graph/PABruijnGraph.hpp:24-132   (vertex set, positions, edges, merges)
 *   PositionProcessor    PAGraph/src/tools/position/PositionProcessor.hpp:21-116 (preProcess / process)
 * driven by run2() (PAGraph/src/main/pagraph.cpp:69-243).  The entry points below replace exactly what run2 does
 * between "Building original pa Graph" (:155) and PAssembly::testTravel5 (:225): one ag2_pg = one PABruijnGraph plus
 * the PositionProcessor state of the current config block.  INTEGRATION.md shows the patch to run2.
 *
 * Conventions as in ag2_b200.h: extern "C", plain pointers and sizes, caller-owned HOST buffers unless a name says
 * `_dev`, 0 or a negative AG2_E* code, no exceptions across the boundary, one handle per GPU, not thread-safe.
 * No CPU fallback: without a CUDA device ag2_pg_create fails with AG2_ENODEV.
 */
#ifndef AG2_PAGRAPH_H
#define AG2_PAGRAPH_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ag2_pg ag2_pg;

/* One record of a `.ref` alignment file = AlignInf (PAGraph/src/tools/align/AlignInf.hpp:18-45) with names resolved
 * to database indices (-1 = name not in the database, or a header line that did not parse: such records still take
 * part in the score sort, AlignmentHelper.cpp:30-37).  The two alignment lines stay text -- normally the file itself,
 * handed over as one buffer: record i's query line is text[q_off .. q_off + ncols), its target line
 * text[t_off .. t_off + ncols); ParseAlignTools::parseDiff (ParseAlignTools.cpp:8-26) runs on the device. */
typedef struct ag2_pg_aln {
    int32_t query;      /* read index (read->contig, read->reference) or contig index (contig->reference) */
    int32_t target;     /* contig index or reference index */
    uint64_t score;     /* AlignInf::_score: atoll(vscore) for Mecat files, query span for the contig->reference file */
    int64_t qb, qe;     /* header columns 5,6 (forward-strand query coordinates) */
    int64_t tb, te;     /* header columns 8,9 */
    int32_t forward;    /* header column 3 == "F" */
    int32_t ncols;      /* alignment columns (length of the query line) */
    int64_t q_off;      /* offset of the query line in the text buffer */
    int64_t t_off;      /* offset of the target line */
} ag2_pg_aln;

#define AG2_PG_READ_TO_CTG 0
#define AG2_PG_READ_TO_REF 1
#define AG2_PG_CTG_TO_REF 2

/* pagraph.cpp:110-125; ag2_pg_params_default() fills the values run2 hard-codes */
typedef struct ag2_pg_params {
    int32_t outer_sample;        /* 3 */
    int32_t read_to_ctg_topk;    /* -1 */
    int32_t read_to_ref_topk;    /* -1 */
    double read_to_ctg_ratio;    /* 0.35 */
    double read_to_ref_ratio;    /* 0.10 */
    int64_t epsilon;             /* --epsilon (posError), 10 */
    int64_t cov_filter;          /* -v, pagraph default 1, the pipeline passes 2 */
} ag2_pg_params;

typedef struct ag2_pg_stats {
    int64_t n_vertices;          /* dense table size (PABruijnGraph::availableKmerNumber) */
    int64_t lanes[2];            /* (read, strand) pairs processed, per phase (0 read->contig, 1 read->reference) */
    int64_t columns[2];          /* alignment columns walked */
    int64_t samples[2];          /* sampled k-mer positions == addPosition calls */
    int64_t tuples[2];           /* DualPos appended to vertices */
    int64_t edges_raw[2];        /* addEdge calls */
    int64_t positions;           /* positions left after both merges ("merge pos" reduces tuples to this) */
    int64_t edges;               /* edges left after mergeEdge */
    int64_t launches;            /* kernel launches of the last build */
    double extract_ms, join_ms;  /* CUDA-event times of the two stages of the last build */
    double join_sort_ms, join_cluster_ms, join_edges_ms;  /* join_ms split: tuple sort by vertex; first-fit clustering + (ctg, ref) order + compaction; edge sort / unique */
} ag2_pg_stats;

int ag2_pg_create(int device, ag2_pg **pg);
void ag2_pg_destroy(ag2_pg *pg);
const char *ag2_pg_last_error(const ag2_pg *pg);
void ag2_pg_params_default(ag2_pg_params *p);

/* B2.  PABruijnGraph::PABruijnGraph (PABruijnGraph.cpp:10-45) over FileKmerIterator (FileKmerIterator.cpp:11-44):
 * `words` is the whole solid_kmer_set.bin file as uint64 -- the iterator re-reads the file from byte 0, so the leading
 * size_t k is one of the "k-mers" and gets a vertex.  Sorted and made unique on the device; dense vertex index = rank. */
int ag2_pg_set_kmers(ag2_pg *pg, const uint64_t *words, int64_t n_words, int64_t *n_vertices);
int ag2_pg_fetch_codes(ag2_pg *pg, uint64_t *codes_out, int64_t cap);

/* B6.  PositionMapper (PositionMapper.cpp:8-47) for the contig and the reference database. */
int ag2_pg_set_targets(ag2_pg *pg, const int64_t *ctg_len, int64_t n_ctg, const int64_t *ref_len, int64_t n_ref);

/* The read database of the current config block (AutoSeqDatabase + CompressedSeq: anything but CcGgTt is A).
 * ASCII, concatenated, offs[n+1].  With several GPUs each rank passes its own contiguous read range, first_read being
 * the database index of its first read (0 on one GPU).  Call before ag2_pg_set_alignments. */
int ag2_pg_set_reads(ag2_pg *pg, const char *bases, const int64_t *offs, int64_t n_reads, int64_t first_read);

/* B3.  One alignment database in FILE order; the library does the reference's two std::sorts (database by score,
 * MecatAlignDatabase.cpp:19 / MummerAlignDatabaseV2.cpp:48; per query by score, Aligner.cpp:53-55).  The order of equal
 * scores comes out of the sort of the WHOLE database and the coverage filter counts every record, so with several GPUs
 * every rank passes every record (`query` = database index of the read); only the records of the rank's own reads need
 * their text (ncols = 0 otherwise). */
int ag2_pg_set_alignments(ag2_pg *pg, int which, const ag2_pg_aln *alns, int64_t n, const char *text, int64_t text_len);

/* Aligner::clearRefFilter/setRefFilter/clearCtgFilter/setCtgFilter (pagraph.cpp:205-217): one byte per sequence. */
int ag2_pg_set_filters(ag2_pg *pg, const uint8_t *ref_flag, const uint8_t *ctg_flag, const uint8_t *ctg_forward);

/* B4-B8 in one call: resetAllNodes, PositionProcessor::preProcess and ::process (PositionProcessor.cpp:57-148):
 * contig->reference position table, read->contig tuples, read->reference tuples, mergeEdge, mergeKmerPosition,
 * sortKmerPosition.  The graph stays on the device. */
int ag2_pg_build(ag2_pg *pg, const ag2_pg_params *params);

/* The same build in stages, for the multi-GPU path (reads sharded over ranks):
 *   ag2_pg_extract       tuples and edges of this rank's reads, both phases, in (phase, read, strand, position) order
 *   ag2_pg_partition     stable partition of both streams by owner rank of the vertex (owner = vertex / ceil(nV / n));
 *                        counts[0..n) tuples and counts[n..2n) edges per owner
 *   ag2_pg_stream_dev    device pointers of the partitioned streams (3 x uint32 tuple arrays: vertex, ctg, ref;
 *                        3 x 32-bit edge arrays: from, to, step) for the caller's all-to-all
 *   ag2_pg_import_dev    replaces the streams with the received ones (rank-major concatenation == global read order)
 *   ag2_pg_join          B8 on whatever streams the handle holds
 * ag2_pg_build == extract + join. */
int ag2_pg_extract(ag2_pg *pg, const ag2_pg_params *params);
int ag2_pg_partition(ag2_pg *pg, int n_owners, int64_t *counts);
int ag2_pg_stream_dev(ag2_pg *pg, int64_t *n_tuples, void **tuple_dev3, int64_t *n_edges, void **edge_dev3);
int ag2_pg_import_dev(ag2_pg *pg, int64_t n_tuples, void *const *tuple_dev3, int64_t n_edges, void *const *edge_dev3);
int ag2_pg_join(ag2_pg *pg, const ag2_pg_params *params);
/* partition + exchange + import for the n handles of ONE process (one per GPU, reads sharded contiguously in handle
 * order): every (rank, owner) segment goes straight into the owner's buffer by a peer copy over NVLink, no host staging.
 * Then ag2_pg_join on every handle; handle o ends up with the vertices it owns, the others empty in its CSR.  (Across
 * processes the same exchange is one NCCL all-to-all per array: aligngraph2_b200/pagraph.py::build_distributed.) */
int ag2_pg_group_exchange(ag2_pg *const *pgs, int n);
/* after ag2_pg_join on every handle: the merge of the per-GPU vertex tables into handle 0 (the all-gather before the
 * traversal of SURVEY 8e, to the one rank that traverses): payload concatenated in handle order by peer copies, offsets
 * summed.  Handle 0 then holds the whole graph; the others keep their part. */
int ag2_pg_group_gather(ag2_pg *const *pgs, int n);

int ag2_pg_get_stats(ag2_pg *pg, ag2_pg_stats *out);

/* The graph as CSR over dense vertex indices: vertex v holds positions [pos_off[v], pos_off[v+1]) sorted by
 * (ctg, ref) (KMerAdjNode::sortPosition) and edges [edge_off[v], edge_off[v+1]) sorted by (to, step).
 * Any pointer may be NULL; pos_off / edge_off have n_vertices + 1 entries. */
int ag2_pg_graph_fetch(ag2_pg *pg, int64_t *pos_off, uint32_t *ctg, uint32_t *ref, uint16_t *count, int64_t pos_cap,
                       int64_t *edge_off, uint32_t *edge_to, int32_t *edge_step, int64_t edge_cap);

void *ag2_pg_stream(ag2_pg *pg);

/* ---- file level: what run2 does around the objects above (pagraph.cpp:127-218) ------------------------------------
 * ag2_pg_job_open      loads solid_kmer_set.bin (-k), the contig (-c) and reference (-R) databases, the contig->reference
 *                      alignments (-a) and <pre dir>/config.txt (-p); creates the ag2_pg.  *job is set even when the call
 *                      fails, so that ag2_pg_job_error() can say why; close it either way.
 * ag2_pg_job_load_block  reads + read->contig + read->reference files and the filters of config block `block`
 *                      (the loop body at pagraph.cpp:167-218); rank / world give this process a contiguous range of the reads
 * then ag2_pg_build (or the staged calls) on ag2_pg_job_handle(), and
 * ag2_pg_job_dump      the graph as text: "#config <block> <ref>" then, per vertex that holds anything,
 *                      "V <idx> <code> P <n> {ctg,ref,count}.. E <m> {to,step}.."; the parity tests compare it with the same dump of the reference classes. */
typedef struct ag2_pg_job ag2_pg_job;
int ag2_pg_job_open(int device, const char *kmer_path, const char *ctg_path, const char *ref_path, const char *pre_dir,
                    const char *ctg_to_ref_path, ag2_pg_job **job);
void ag2_pg_job_close(ag2_pg_job *job);
const char *ag2_pg_job_error(const ag2_pg_job *job);
int ag2_pg_job_blocks(const ag2_pg_job *job);
const char *ag2_pg_job_block_ref(const ag2_pg_job *job, int block);
ag2_pg *ag2_pg_job_handle(ag2_pg_job *job);
int ag2_pg_job_load_block(ag2_pg_job *job, int block, int rank, int world);
int ag2_pg_job_dump(ag2_pg_job *job, int block, const char *path, int append);

/* ---- B9: traversal and assembly of the walks -----------------------------------------------------------------------
 * PAssembly::testTravel5 (PAGraph/src/tools/graph/PAssembly.cpp:11-336) over PAlgorithm::travelSequence
 * (graph/PAlgorithm.cpp:145-426), graphTravel / walkStraight / classifySuccessors (graph/PAlgorithm.tcc:37-298) and
 * PAlgorithm::seqToString (:428-489).  A host stage by design (SURVEY 8e "replicas only": 2 x contigs independent walks,
 * each a sequential pointer chase) over the graph the device built; it needs no device and has no device variant. */
typedef struct ag2_pg_graph_view {       /* what ag2_pg_graph_fetch + ag2_pg_fetch_codes return */
    int32_t k;                           /* k-mer size (first word of solid_kmer_set.bin) */
    int64_t n_vertices;
    const uint64_t *codes;               /* ascending; dense vertex index = rank */
    const int64_t *pos_off;              /* n_vertices + 1 */
    const uint32_t *ctg, *ref;           /* packed positions (PositionMapper::dualToSingle), sorted by (ctg, ref) per vertex */
    const uint16_t *count;               /* abundance of each position */
    const int64_t *edge_off;             /* n_vertices + 1 */
    const uint32_t *edge_to;             /* sorted by (to, step) per vertex */
    const int32_t *edge_step;
} ag2_pg_graph_view;

typedef struct ag2_pg_seqs {             /* a sequence database (AutoSeqDatabase): names = first token after '>' */
    int64_t n;
    const char *const *names;
    const char *bases;                   /* ASCII as in the file, concatenated; anything but CcGgTt reads as A (CompressedSeq.cpp:16-26) */
    const int64_t *offs;                 /* n + 1 */
} ag2_pg_seqs;

typedef struct ag2_pg_travel_params {    /* PGM/pagraph.cpp:123-126,247-256 */
    int64_t deviation;                   /* 2 * --epsilon */
    double error_rate;                   /* 0.15 */
    double start_split;                  /* 0.90 */
    int64_t min_len;                     /* -l, 50 */
    int32_t threads;                     /* -t: start vertices per round = min(t, 8) (PAlgorithm.cpp:146), and the host threads used */
} ag2_pg_travel_params;

void ag2_pg_travel_params_default(ag2_pg_travel_params *p);

/* One config block: walks the contigs (use_ctg[i], use_forward[i]) -- usedCtg of PGM/pagraph.cpp:209-214 -- and writes
 * <out_dir>/<prefix><ctg>_<0|1>.txt for each, then <..>.fasta / .help / .con for every chain that connects contigs or
 * extends one (PAssembly.cpp:229-333).  ok_ctg[0..*n_ok) = contig indices of the returned success set in its std::set order
 * (a contig walked in both orientations appears twice); room for 2 * n_use entries. */
int ag2_pg_travel(const ag2_pg_graph_view *g, const ag2_pg_seqs *ctgs, const ag2_pg_seqs *refs, const int32_t *use_ctg,
                  const uint8_t *use_forward, int64_t n_use, const ag2_pg_travel_params *params, const char *out_dir,
                  const char *prefix, int32_t *ok_ctg, int64_t *n_ok);

/* The same on the job's current block: fetches the graph of the handle, prefix "<block>_", and remembers the success set.
 * ag2_pg_job_write_contig_list writes <out_dir>/contig.txt (PGM/pagraph.cpp:265-269) from all blocks travelled so far. */
int ag2_pg_job_travel(ag2_pg_job *job, int block, const ag2_pg_travel_params *params, const char *out_dir);
int ag2_pg_job_write_contig_list(ag2_pg_job *job, const char *out_dir);

#ifdef __cplusplus
}
#endif
#endif
