/*
 * ag2_diff.c -- plain-C restatement of vanilla MECAT2's DiffAligner.  TEST INFRASTRUCTURE ONLY.
 *
 * SURVEY row N2 ("next"): AlignGraph2.py:232-239,478-485 runs thirdparty/mecat's mecat2ref, whose gapped aligner is
 * DiffAligner -- the O(ND) furthest-reaching-point difference algorithm with a band, block by block
 * (thirdparty/mecat/src/common/diff_gapalign.cpp; byte-identical in algorithm to mecat_plus/.../common/diff_gapalign.cpp,
 * which mecat2ref+ carries but never reaches).  This file is the first step of that row and only that: the checker a CUDA
 * path will be held against.  No product code calls it; nothing under aligngraph2_b200/ implements N2 yet.
 *
 * Pinned (tests/test_oracle_diff.py) against oracle/_ref/libref_mecat_vanilla.so = the unmodified reference sources
 * behind oracle/ref_diff_shim.cpp, on fresh random inputs and on tests/golden/diff_*.npz generated from it.
 *
 * Reference lines are those of thirdparty/mecat/src/common/diff_gapalign.cpp ("D:") and gapalign.cpp ("G:").
 */
#include "ag2_oracle.h"

#include <stdlib.h>
#include <string.h>

#define DIFF_GAP 4            /* GAP_CODE, defs.h:197 */
#define DIFF_TAIL_MATCH 4     /* kTailMatchBP, D:225 */

typedef struct {
    int d, k;       /* the node: d differences, diagonal k = x - y */
    int from_k;     /* diagonal of its predecessor at d - 1 */
    int x1, y1;     /* where the step from the predecessor lands */
    int x2, y2;     /* end of the run of matches from there */
} diff_node;

typedef struct {
    int q_s, q_e, t_s, t_e, dist, n;
} diff_aln;

static inline char seq_at(const char *s, int i, int fwd) { return fwd ? s[i] : s[-i]; } /* extract_char, gapalign.h:27-35 */

/* first node not before (d, k) in the order the search appends them: d ascending, k ascending (D:12-37) */
static const diff_node *find_node(const diff_node *nodes, long n, int d, int k)
{
    long lo = 0, hi = n;
    while (lo < hi) {
        const long mid = lo + (hi - lo) / 2;
        const diff_node *m = &nodes[mid];
        const int before = m->d == d ? m->k < k : m->d < d;
        if (before) lo = mid + 1;
        else hi = mid;
    }
    return &nodes[lo];
}

/* GetAlignString (D:39-104): back from node (d, k) to the origin, then forward emitting columns (codes, 4 = gap). */
static void diff_traceback(const char *Q, int qn, const char *T, int tn, const diff_node *nodes, long n_nodes, int (*pts)[2], int d,
                           int k, int fwd, diff_aln *a, char *qstr, char *tstr)
{
    int np = 0;
    for (int cd = d, ck = k; cd >= 0 && np < qn + tn + 1; --cd) {
        const diff_node *nd = find_node(nodes, n_nodes, cd, ck);
        pts[np][0] = nd->x2;
        pts[np][1] = nd->y2;
        ++np;
        pts[np][0] = nd->x1;
        pts[np][1] = nd->y1;
        ++np;
        ck = nd->from_k;
    }
    --np;
    int cx = pts[np][0], cy = pts[np][1], pos = 0;
    a->q_s = cx;
    a->t_s = cy;
    while (np > 0) {
        --np;
        const int nx = pts[np][0], ny = pts[np][1];
        if (nx == cx && ny == cy) continue;
        if (nx == cx) {                 /* target bases against gaps */
            for (int i = 0; i < ny - cy; ++i) {
                qstr[pos + i] = DIFF_GAP;
                tstr[pos + i] = seq_at(T, cy + i, fwd);
            }
            pos += ny - cy;
        } else if (ny == cy) {          /* query bases against gaps */
            for (int i = 0; i < nx - cx; ++i) {
                qstr[pos + i] = seq_at(Q, cx + i, fwd);
                tstr[pos + i] = DIFF_GAP;
            }
            pos += nx - cx;
        } else {                        /* a run of matches */
            for (int i = 0; i < nx - cx; ++i) qstr[pos + i] = seq_at(Q, cx + i, fwd);
            for (int i = 0; i < ny - cy; ++i) tstr[pos + i] = seq_at(T, cy + i, fwd);
            pos += ny - cy;
        }
        cx = nx;
        cy = ny;
    }
    a->n = pos;
}

/* Align (D:107-219) as dw_in_one_direction calls it: tol = band tolerance, strings wanted.  Returns the reference's return
 * value (an end of either block reached). */
static int diff_block(const char *Q, int qn, const char *T, int tn, int tol, int fwd, diff_aln *a, char *qstr, char *tstr)
{
    const int max_d = (int)(.3 * (qn + tn));
    const int off = max_d, band = tol * 2;
    /* furthest x per diagonal and x + y per diagonal (the reference's V and U, 4096 zeroed ints each: D:229-230) */
    int *far_x = (int *)calloc((size_t)2 * max_d + 8, sizeof(int)), *far_sum = (int *)calloc((size_t)2 * max_d + 8, sizeof(int));
    const long node_cap = (long)(max_d + 1) * (tol + 3) + 16;
    diff_node *nodes = (diff_node *)malloc((size_t)node_cap * sizeof(diff_node));
    int(*pts)[2] = (int(*)[2])malloc(((size_t)2 * max_d + 8) * sizeof(int[2]));
    long n_nodes = 0, end_nodes = 0, best_nodes = -1;
    int lo = 0, hi = 0, best_sum = -1, best_x = -1, best_y = -1, best_d = qn + tn + 100, best_k = 0;
    int x = -1, y = -1, k = 0, d, done = 0;
    memset(a, 0, sizeof *a);

    for (d = 0; d < max_d; ++d) {
        if (hi - lo > band) break;
        for (k = lo; k <= hi; k += 2) {
            int from;
            if (k == lo || (k != hi && far_x[k - 1 + off] < far_x[k + 1 + off])) {   /* from the diagonal above: a target base */
                from = k + 1;
                x = far_x[k + 1 + off];
            } else {                                                                 /* from the diagonal below: a query base */
                from = k - 1;
                x = far_x[k - 1 + off] + 1;
            }
            y = x - k;
            diff_node *nd = &nodes[n_nodes];
            nd->d = d;
            nd->k = k;
            nd->x1 = x;
            nd->y1 = y;
            while (x < qn && y < tn && seq_at(Q, x, fwd) == seq_at(T, y, fwd)) {
                ++x;
                ++y;
            }
            nd->x2 = x;
            nd->y2 = y;
            nd->from_k = from;
            ++n_nodes;
            far_x[k + off] = x;
            far_sum[k + off] = x + y;
            if (x + y > best_sum) {
                best_sum = x + y;
                best_x = x;
                best_y = y;
                best_d = d;
                best_k = k;
                best_nodes = n_nodes;
            }
            if (x >= qn || y >= tn) {
                done = 1;
                end_nodes = n_nodes;
                break;
            }
        }
        /* the band for d + 1: diagonals whose x + y is within tol of the best (D:168-175) */
        int nlo = hi, nhi = lo;
        for (int k2 = lo; k2 <= hi; k2 += 2)
            if (far_sum[k2 + off] >= best_sum - tol) {
                if (k2 < nlo) nlo = k2;
                if (k2 > nhi) nhi = k2;
            }
        hi = nhi + 1;
        lo = nlo - 1;
        if (done) {
            a->q_e = x;
            a->t_e = y;
            a->dist = d;
            a->n = (x + y + d) / 2;
            diff_traceback(Q, qn, T, tn, nodes, end_nodes, pts, d, k, fwd, a, qstr, tstr);
            break;
        }
    }
    if (!done && best_x > 0) {   /* neither end reached: the node with the largest x + y (D:193-203) */
        a->q_e = best_x;
        a->t_e = best_y;
        a->dist = best_d;
        a->n = (best_x + best_y + best_d) / 2;
        diff_traceback(Q, qn, T, tn, nodes, best_nodes, pts, best_d, best_k, fwd, a, qstr, tstr);
    }
    free(far_x);
    free(far_sum);
    free(nodes);
    free(pts);
    return a->q_e == qn || a->t_e == tn;
}

/* trim_mismatch_end (G:47-68) on code strings */
static int diff_trim_tail(const char *qa, const char *ta, int n, int want, int *qcnt, int *tcnt, int *acnt)
{
    int m = 0, k, q = 0, t = 0, ac = 0;
    for (k = n - 1; k >= 0 && m < want; --k) {
        ++ac;
        if (qa[k] != DIFF_GAP) ++q;
        if (ta[k] != DIFF_GAP) ++t;
        m = qa[k] == ta[k] ? m + 1 : 0;
    }
    *qcnt = q;
    *tcnt = t;
    *acnt = ac;
    return m == want && k > 0;
}

typedef struct {
    char *q, *t;
    int n;
} diff_store;

/* dw_in_one_direction (D:221-292); q / t point at the direction's first base (towards lower addresses when !fwd) */
static void diff_direction(const char *q, int qsize, const char *t, int tsize, int seg, int fwd, diff_store *st, char *qstr, char *tstr)
{
    int qidx = 0, tidx = 0;
    for (;;) {
        /* retrieve_next_aln_block (G:9-45) */
        const int qleft = qsize - qidx, tleft = tsize - tidx;
        int qblk, tblk, last;
        if (qleft < seg + 100 || tleft < seg + 100) {
            const int tq = (int)(tleft + tleft * 0.2), qt = (int)(qleft + qleft * 0.2);
            qblk = qleft < tq ? qleft : tq;
            tblk = tleft < qt ? tleft : qt;
            last = 1;
        } else {
            qblk = tblk = seg;
            last = 0;
        }
        const char *Q = fwd ? q + qidx : q - qidx, *T = fwd ? t + tidx : t - tidx;
        diff_aln a;
        diff_block(Q, qblk, T, tblk, (int)(0.3 * (qblk > tblk ? qblk : tblk)), fwd, &a, qstr, tstr);
        int qcnt, tcnt, acnt;
        if (!diff_trim_tail(qstr, tstr, a.n, DIFF_TAIL_MATCH, &qcnt, &tcnt, &acnt)) break;
        const int full_map = qblk - a.q_e <= 20 || tblk - a.t_e <= 20;
        const int stop = last || !full_map;
        if (stop) {          /* the last block of the direction keeps its tail matches */
            qcnt -= DIFF_TAIL_MATCH;
            tcnt -= DIFF_TAIL_MATCH;
            acnt -= DIFF_TAIL_MATCH;
        }
        const int keep = a.n - acnt;
        memcpy(st->q + st->n, qstr, (size_t)keep);
        memcpy(st->t + st->n, tstr, (size_t)keep);
        st->n += keep;
        if (stop) break;
        qidx += a.q_e - qcnt;
        tidx += a.t_e - tcnt;
    }
}

int orc_diff_block(const uint8_t *Q, int q_len, const uint8_t *T, int t_len, int right_extend, int *out6, uint8_t *qstr, uint8_t *tstr)
{
    diff_aln a;
    const int tol = (int)(0.3 * (q_len > t_len ? q_len : t_len));
    const int rc = diff_block((const char *)Q, q_len, (const char *)T, t_len, tol, right_extend, &a, (char *)qstr, (char *)tstr);
    out6[0] = a.q_s;
    out6[1] = a.q_e;
    out6[2] = a.t_s;
    out6[3] = a.t_e;
    out6[4] = a.dist;
    out6[5] = a.n;
    return rc;
}

/* DiffAligner::go (D:294-349): left of (qstart, tstart) backwards, right of it forwards; ASCII out. */
int orc_diff_go(const uint8_t *query, int qstart, int qsize, const uint8_t *target, int tstart, int tsize, int min_aln_size,
                int large_block, int *out5, char *qaln, char *taln)
{
    const int seg = large_block ? 1000 : 500;       /* DiffAlignParameters::init, diff_gapalign.h:25-47 */
    const size_t cap = (size_t)qsize + (size_t)tsize + 4096;
    diff_store left = {(char *)malloc(cap), (char *)malloc(cap), 0}, right = {(char *)malloc(cap), (char *)malloc(cap), 0};
    char *qstr = (char *)malloc(8192), *tstr = (char *)malloc(8192);   /* a block's strings: <= 2 x 1200 columns */
    const char *q = (const char *)query, *t = (const char *)target;
    diff_direction(q + qstart - 1, qstart, t + tstart - 1, tstart, seg, 0, &left, qstr, tstr);
    diff_direction(q + qstart, qsize - qstart, t + tstart, tsize - tstart, seg, 1, &right, qstr, tstr);
    static const char letters[] = "ACGT-";
    int n = 0, qi = 0, ti = 0;
    for (int k = left.n - 1; k >= 0; --k, ++n) {
        qaln[n] = letters[(unsigned char)left.q[k]];
        taln[n] = letters[(unsigned char)left.t[k]];
        qi += qaln[n] != '-';
        ti += taln[n] != '-';
    }
    out5[0] = qstart - qi;
    out5[2] = tstart - ti;
    qi = ti = 0;
    for (int k = 0; k < right.n; ++k, ++n) {
        qaln[n] = letters[(unsigned char)right.q[k]];
        taln[n] = letters[(unsigned char)right.t[k]];
        qi += qaln[n] != '-';
        ti += taln[n] != '-';
    }
    out5[1] = qstart + qi;
    out5[3] = tstart + ti;
    out5[4] = n;
    qaln[n] = taln[n] = '\0';
    free(left.q);
    free(left.t);
    free(right.q);
    free(right.t);
    free(qstr);
    free(tstr);
    return n >= min_aln_size;
}
