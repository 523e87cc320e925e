/*
 * ag2_oracle.h -- CPU restatement of the mecat2ref+ hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is the parity checker for aligngraph2_b200: a plain-C restatement of the reference
 * algorithm (reference file:line cited at every function in ag2_oracle.c).  It is pinned against
 * the UNMODIFIED reference compiled into oracle/_ref/ (oracle/Makefile, tests/test_oracle_pinned.py)
 * and against the golden vectors in tests/golden/ that were produced by that reference build
 * (tests/golden/gen_golden.py).  The reference ships no tests or golden vectors of its own for
 * this path, so those two are the pin.
 *
 * Nothing in the product (aligngraph2_b200/, include/) may include, link or call this file; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 */
#ifndef AG2_ORACLE_H
#define AG2_ORACLE_H

#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- X-drop extension ----------------------------------------------- */
typedef struct orc_aln {
    int ok;              /* go() return value: qend - qoff >= min_aln */
    int qoff, qend, toff, tend, aln_size;
    char *qaln, *taln;   /* ASCII ACGT-, NUL terminated, owned by the orc_xdrop */
    long cells;          /* inner-loop cell evaluations of this go() (the roofline's C) */
    long rows;           /* DP rows evaluated */
    long calls;          /* block DP calls */
} orc_aln;

typedef struct orc_xdrop orc_xdrop;
orc_xdrop *orc_xdrop_new(void);
void orc_xdrop_free(orc_xdrop *x);

/* One block DP.  A/B are code arrays (0..3); forward=0 reads A[-i], B[-i].  ops[] gets one op per
 * traceback step in walk order (3 sub, 0 gap-in-A, 6 gap-in-B).  Returns the best score. */
int orc_xdrop_block(orc_xdrop *x, const uint8_t *A, int M, const uint8_t *B, int N, int forward,
                    int *ae, int *be, uint8_t *ops, int *nops, long *cells);

/* GapAligner::go on codes 0..3. */
int orc_xdrop_go(orc_xdrop *x, const uint8_t *query, int qstart, int qsize,
                 const uint8_t *target, int tstart, int tsize, int min_aln, orc_aln *out);

/* extend_candidate: window extraction + go + record coordinates.  ref/read are raw ASCII.
 * loc1 is the 1-based reference position of the seed, loc2 the 0-based read position.
 * Returns ok; on ok fills rec = {qb, qe, sb, se} (sb/se global reference coordinates). */
int orc_extend_candidate(orc_xdrop *x, const char *ref, long ref_size, const char *read, int read_len,
                         long loc1, long loc2, long rec[4], orc_aln *out);

/* cumulative counters over the lifetime of the orc_xdrop */
void orc_xdrop_counters(const orc_xdrop *x, long *cells, long *rows, long *calls);

/* ---- index (rows A2-A4), seeding + candidates (A5-A7), rescue + output (A11, A12): ag2_mapper.c ---- */
typedef struct orc_index {
    long R;              /* concatenated reference length */
    char *ref;           /* reference characters as creat_ref_index keeps them */
    int cbl;             /* similarity block size (-z) */
    long nblk;           /* R / cbl + 1 */
    int *rcnt;           /* masked read 13-mer counts [4^13]      (countin1) */
    int *cnt;            /* masked reference 13-mer counts [4^13] (countin) */
    uint32_t *off;       /* CSR bucket offsets [4^13 + 1] */
    uint32_t *pos;       /* 1-based k-mer start positions, ascending inside a bucket (allloc) */
    int *kcount;         /* per similarity block (sim::k_count), 10 zero entries of slack */
    float *vote;         /* per similarity block (sim::vote), 10 zero entries of slack */
    float ave;
} orc_index;

typedef struct orc_cand { /* candidate_save, mecat2ref_defs.h:90-95 */
    long loc1, loc2, left1, left2, right1, right2;
    int score, num1, num2;
    char chain;
} orc_cand;

typedef struct orc_block { /* Back_List, mecat2ref_defs.h:84-88 */
    short score, score2, loczhi[20], seedno[20], seednum;
    int index;
} orc_block;

typedef struct orc_mapper orc_mapper;

void orc_read_hist13(const char *seq, long n, int *counts);
long orc_read_index_prefix(const long *offs, long nreads);
orc_index *orc_index_build(const char *ref, long R, const int *rcnt, int cbl, double alpha, double beta);
void orc_index_free(orc_index *ix);
orc_mapper *orc_mapper_new(const orc_index *ix, int maxc, int num_output);
void orc_mapper_free(orc_mapper *m);
/* reference_mapping's loop body for one read; appends `.r` records to out (may be NULL) */
int orc_map_read(orc_mapper *m, int read_id, const char *read, int len, FILE *out);
/* seeding + candidate scan of one read without extension (pass 0, or 1 = the reference's second pass) */
int orc_seed_candidates(orc_mapper *m, const char *read, int len, int pass, orc_cand *out);
/* pass-1 candidates of the read orc_map_read saw last, in canidate_loc[] order */
int orc_mapper_last_candidates(const orc_mapper *m, orc_cand *out, int *pass2);
long orc_mapper_cells(const orc_mapper *m);
long orc_mapper_calls(const orc_mapper *m);
long orc_mapper_aligned(const orc_mapper *m);
long orc_map_batch(const char *ref, long R, const char *reads, const long *offs, const int *ids, long n, int cbl, double alpha,
                   double beta, int maxc, int num_output, const char *r_path, long *stats);

/* ---- PAGraph kmer_counter (row B1): ag2_kmer.c ---- */
long orc_solid_kmers(const char *reads, const long *offs, long n_reads, int k, double threshold, uint64_t *codes_out, long cap,
                     long *min_abundance);

/* ---- vanilla MECAT2 DiffAligner (row N2, "next"): ag2_diff.c ---- */
/* one block the way dw_in_one_direction aligns it; out6 = q_s q_e t_s t_e dist n; strings are codes 0..4 (4 = gap) */
int orc_diff_block(const uint8_t *Q, int q_len, const uint8_t *T, int t_len, int right_extend, int *out6, uint8_t *qstr, uint8_t *tstr);
/* DiffAligner::go; out5 = qoff qend toff tend aln_size; ASCII strings */
int orc_diff_go(const uint8_t *query, int qstart, int qsize, const uint8_t *target, int tstart, int tsize, int min_aln_size,
                int large_block, int *out5, char *qaln, char *taln);

#ifdef __cplusplus
}
#endif
#endif
