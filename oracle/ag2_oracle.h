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

#ifdef __cplusplus
}
#endif
#endif
