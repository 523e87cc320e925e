/*
 * ag2_oracle.c -- CPU restatement of the mecat2ref+ hot path.  TEST INFRASTRUCTURE ONLY
 * (see ag2_oracle.h for the rules on who may call it).
 *
 * Paths below are relative to /root/reference/mecat_plus/MECAT-master_1/src/ :
 *   MC/  = common/      M2R/ = mecat2ref/
 *
 * Pinned: tests/test_oracle_pinned.py runs every function here against the unmodified reference
 * build in oracle/_ref/ (when present) and against tests/golden/ vectors produced by it.
 */
#include "ag2_oracle.h"

#include <stdlib.h>
#include <string.h>

/* MC/xdrop_gapalign.cpp:8 ; MC/xdrop_gapalign.h:98-114 (reward 1, penalty -1, gap_open 0,
 * gap_extend 1, x_dropoff 30, block 500) ; MC/xdrop_gapalign.h:26-34 (traceback byte values) */
#define NEG_INF (-100000000)
#define REWARD 1
#define PENALTY (-1)
#define GAP_OPEN 0
#define GAP_EXT 1
#define XDROP 30
#define BLK 500
#define OP_SUB 3
#define OP_GAP_A 0 /* gap in the query: consumes a target base */
#define OP_GAP_B 6 /* gap in the target: consumes a query base */
#define OP_MASK 7
#define FLAG_EXT_A 0x10
#define FLAG_EXT_B 0x40
#define GAP_CODE 4 /* MC/defs.h:197 */

typedef struct {
    int h; /* best score ending in a match/mismatch here  (BlastGapDP::best) */
    int e; /* best score ending in a gap here             (BlastGapDP::best_gap) */
} colstate;

typedef struct {
    char *p;
    long n, cap;
} strbuf;

struct orc_xdrop {
    colstate *col;
    long col_cap;
    uint8_t *tb;
    long tb_cap;
    uint8_t *ops;
    long ops_cap;
    strbuf lq, lt, rq, rt, tq, tt; /* left/right/tmp aligned code strings */
    char *qaln, *taln;
    long aln_cap;
    uint8_t *qcodes, *tcodes;
    long qc_cap, tc_cap;
    long cells, rows, calls;
};

static void sb_reserve(strbuf *s, long need)
{
    if (need <= s->cap) return;
    long c = s->cap ? s->cap : 1024;
    while (c < need) c *= 2;
    s->p = (char *)realloc(s->p, (size_t)c);
    s->cap = c;
}
static void sb_push(strbuf *s, char c)
{
    sb_reserve(s, s->n + 1);
    s->p[s->n++] = c;
}
static void sb_append(strbuf *s, const char *p, long n)
{
    sb_reserve(s, s->n + n);
    memcpy(s->p + s->n, p, (size_t)n);
    s->n += n;
}

orc_xdrop *orc_xdrop_new(void)
{
    orc_xdrop *x = (orc_xdrop *)calloc(1, sizeof(*x));
    return x;
}

void orc_xdrop_free(orc_xdrop *x)
{
    if (!x) return;
    free(x->col); free(x->tb); free(x->ops);
    free(x->lq.p); free(x->lt.p); free(x->rq.p); free(x->rt.p); free(x->tq.p); free(x->tt.p);
    free(x->qaln); free(x->taln); free(x->qcodes); free(x->tcodes);
    free(x);
}

void orc_xdrop_counters(const orc_xdrop *x, long *cells, long *rows, long *calls)
{
    if (cells) *cells = x->cells;
    if (rows) *rows = x->rows;
    if (calls) *calls = x->calls;
}

static inline int seq_at(const uint8_t *s, int i, int forward) /* MC/gapalign.h:27-35 extract_char */
{
    return forward ? s[i] : s[-i];
}

/*
 * xdrop_align -- MC/xdrop_gapalign.cpp:11-213.
 *
 * Semi-global affine X-drop DP from (0,0).  Row a = a query bases consumed, column b = b target
 * bases consumed.  One traceback byte per evaluated cell.  Things that are easy to get wrong and
 * that this restatement keeps on purpose:
 *   - best is a ROW-MAJOR running maximum: cells later in the same row are pruned against it (:109);
 *   - a pruned cell that is not the band's first cell only has its h set to NEG_INF (:111): its e
 *     keeps the stale value of an earlier row, and the row's running horizontal gap score is NOT
 *     decremented across it (:109-112 skip :120-133);
 *   - the traceback byte of a pruned cell is still written (:139), without extension flags;
 *   - the band never reaches column N unless the row-0 initialisation did (:147 `b_size < N`);
 *   - the last loop iteration reads B[b_size-1] even when that is index N (:85) -- the value is
 *     dead; here the read is skipped instead of performed out of bounds.
 */
/* Band-geometry statistics for kernel design (test infrastructure only): [0] histogram of row widths,
 * [1] per-block maximum row width, [2] per-block maximum of (band end - window base) for a window whose base
 * follows `first` in steps of 8 columns, checked every 4 rows. */
long orc_band_hist[3][256];

int orc_xdrop_block(orc_xdrop *x, const uint8_t *A, int M, const uint8_t *B, int N, int forward,
                    int *ae_out, int *be_out, uint8_t *ops, int *nops, long *cells_out)
{
    *ae_out = 0;
    *be_out = 0;
    *nops = 0;
    if (cells_out) *cells_out = 0;
    if (M <= 0 || N <= 0) return 0;

    const int goe = GAP_OPEN + GAP_EXT;
    const int xd = XDROP < goe ? goe : XDROP; /* :48 */
    const long stride = (long)N + 2;
    if (x->col_cap < N + 2) {
        x->col_cap = N + 2;
        x->col = (colstate *)realloc(x->col, sizeof(colstate) * (size_t)x->col_cap);
    }
    if (x->tb_cap < (long)(M + 1) * stride) {
        x->tb_cap = (long)(M + 1) * stride;
        x->tb = (uint8_t *)realloc(x->tb, (size_t)x->tb_cap);
    }
    colstate *col = x->col;
    uint8_t *tb = x->tb;
    long cells = 0, rows = 0;

    /* row 0  (:53-67) */
    int s = -goe, i;
    col[0].h = 0;
    col[0].e = -goe;
    for (i = 1; i <= N; ++i) {
        if (s < -xd) break;
        col[i].h = s;
        col[i].e = s - goe;
        s -= GAP_EXT;
        tb[i] = OP_GAP_A;
    }
    int bsize = i, first = 0, best = 0, ae = 0, be = 0;
    int st_base = 0, st_maxw = 0, st_maxneed = bsize;

    for (int a = 1; a <= M; ++a) { /* :69-165 */
        if ((a & 3) == 0 && first - st_base >= 8) st_base += 8;
        if (bsize - first > st_maxw) st_maxw = bsize - first;
        orc_band_hist[0][bsize - first > 255 ? 255 : bsize - first]++;
        const int ac = seq_at(A, a - 1, forward);
        uint8_t *row = tb + (long)a * stride;
        int diag = NEG_INF, hgap = NEG_INF, last = first, b;
        ++rows;
        for (b = first; b < bsize; ++b) { /* :84-140 */
            const int vgap_in = col[b].e;
            int next_diag = NEG_INF;
            if (b < N) { /* dead read at b == N in the reference (:85) */
                const int bc = seq_at(B, b, forward);
                next_diag = col[b].h + (ac == bc ? REWARD : PENALTY);
            }
            int sc = diag, vgap = vgap_in;
            uint8_t op = OP_SUB;
            ++cells;
            if (sc < vgap) { op = OP_GAP_B; sc = vgap; }
            if (sc < hgap) { op = OP_GAP_A; sc = hgap; }
            if (best - sc > xd) {
                if (first == b) ++first;
                else col[b].h = NEG_INF;
            } else {
                last = b;
                if (sc > best) { best = sc; ae = a; be = b; }
                vgap -= GAP_EXT;
                if (vgap < sc - goe) col[b].e = sc - goe;
                else { col[b].e = vgap; op += FLAG_EXT_A; }
                hgap -= GAP_EXT;
                if (hgap < sc - goe) hgap = sc - goe;
                else op += FLAG_EXT_B;
                col[b].h = sc;
            }
            diag = next_diag;
            row[b] = op;
        }
        if (first == bsize) break; /* :142 */
        if (last < bsize - 1) {
            bsize = last + 1; /* :144-145 */
        } else {
            while (hgap >= best - xd && bsize < N) { /* :147-153 */
                col[bsize].h = hgap;
                col[bsize].e = hgap - goe;
                hgap -= GAP_EXT;
                row[bsize] = OP_GAP_A;
                ++bsize;
            }
        }
        if (bsize < N) { /* :160-164 */
            col[bsize].h = NEG_INF;
            col[bsize].e = NEG_INF;
            ++bsize;
        }
        if (bsize - st_base > st_maxneed) st_maxneed = bsize - st_base;
    }
    orc_band_hist[1][st_maxw > 255 ? 255 : st_maxw]++;
    orc_band_hist[2][st_maxneed > 255 ? 255 : st_maxneed]++;

    /* traceback (:170-210), expanded to one op per step */
    int a = ae, b = be, n = 0;
    int cur = OP_SUB;
    while (a > 0 || b > 0) {
        const uint8_t t = tb[(long)a * stride + b];
        if (cur == OP_GAP_A) {
            cur = (t & FLAG_EXT_A) ? OP_GAP_A : (t & OP_MASK);
        } else if (cur == OP_GAP_B) {
            cur = (t & FLAG_EXT_B) ? OP_GAP_B : (t & OP_MASK);
        } else {
            cur = t & OP_MASK;
        }
        if (cur == OP_GAP_A) --b;
        else if (cur == OP_GAP_B) --a;
        else { --a; --b; }
        ops[n++] = (uint8_t)cur;
    }
    *nops = n;
    *ae_out = ae;
    *be_out = be;
    if (cells_out) *cells_out = cells;
    x->cells += cells;
    x->rows += rows;
    x->calls += 1;
    return best;
}

/* trim_mismatch_end -- MC/gapalign.cpp:47-68 */
static int trim_tail(const char *q, const char *t, long n, int want, int *qcnt, int *tcnt, int *acnt)
{
    int m = 0;
    long k;
    *qcnt = *tcnt = *acnt = 0;
    for (k = n - 1; k >= 0 && m < want; --k) {
        ++*acnt;
        if (q[k] != GAP_CODE) ++*qcnt;
        if (t[k] != GAP_CODE) ++*tcnt;
        if (q[k] == t[k]) ++m;
        else m = 0;
    }
    return m == want && k > 0;
}

/*
 * align_ex -- MC/xdrop_gapalign.cpp:263-357, with retrieve_next_aln_block (MC/gapalign.cpp:9-45)
 * and script_to_aligned_string (MC/xdrop_gapalign.cpp:215-261) folded in.
 * q/t point at the first base of the extension; backward extensions read q[-i], t[-i].
 * Output strings hold codes 0..3 and GAP_CODE, in extension order.
 */
static void extend_one_direction(orc_xdrop *x, const uint8_t *q, int qsize, const uint8_t *t, int tsize,
                                 int forward, strbuf *qout, strbuf *tout)
{
    int qidx = 0, tidx = 0;
    const int inc = forward ? 1 : -1;
    qout->n = tout->n = 0;
    if (x->ops_cap < 4096) {
        x->ops_cap = 4096;
        x->ops = (uint8_t *)realloc(x->ops, (size_t)x->ops_cap);
    }
    for (;;) {
        const int qleft = qsize - qidx, tleft = tsize - tidx;
        int qblk, tblk, last_block;
        if (qleft < BLK + 100 || tleft < BLK + 100) { /* MC/gapalign.cpp:26-29 */
            const int qcap = (int)(tleft + tleft * 0.2);
            const int tcap = (int)(qleft + qleft * 0.2);
            qblk = qleft < qcap ? qleft : qcap;
            tblk = tleft < tcap ? tleft : tcap;
            last_block = 1;
        } else {
            qblk = BLK;
            tblk = BLK;
            last_block = 0;
        }
        const uint8_t *Q = forward ? q + qidx : q - qidx;
        const uint8_t *T = forward ? t + tidx : t - tidx;
        if (x->ops_cap < qblk + tblk + 8) {
            x->ops_cap = qblk + tblk + 8;
            x->ops = (uint8_t *)realloc(x->ops, (size_t)x->ops_cap);
        }
        int ae, be, nops;
        orc_xdrop_block(x, Q, qblk, T, tblk, forward, &ae, &be, x->ops, &nops, NULL);

        /* ops are in walk order (end -> origin); replay from the origin (:230) */
        x->tq.n = x->tt.n = 0;
        const uint8_t *qp = Q, *tp = T;
        for (int i = nops - 1; i >= 0; --i) {
            switch (x->ops[i]) {
            case OP_SUB:
                sb_push(&x->tq, (char)*qp); sb_push(&x->tt, (char)*tp);
                qp += inc; tp += inc;
                break;
            case OP_GAP_A:
                sb_push(&x->tq, GAP_CODE); sb_push(&x->tt, (char)*tp);
                tp += inc;
                break;
            default: /* OP_GAP_B */
                sb_push(&x->tq, (char)*qp); sb_push(&x->tt, GAP_CODE);
                qp += inc;
                break;
            }
        }
        const int full_map = (qblk - ae <= 20) || (tblk - be <= 20); /* :334-335 */
        if (!full_map || last_block) {
            sb_append(qout, x->tq.p, x->tq.n);
            sb_append(tout, x->tt.p, x->tt.n);
            break;
        }
        int qcnt, tcnt, acnt;
        if (!trim_tail(x->tq.p, x->tt.p, x->tq.n, 4, &qcnt, &tcnt, &acnt)) break; /* :349: block dropped */
        sb_append(qout, x->tq.p, x->tq.n - acnt);
        sb_append(tout, x->tt.p, x->tt.n - acnt);
        qidx += ae - qcnt;
        tidx += be - tcnt;
    }
}

/* XdropAligner::go -- MC/xdrop_gapalign.cpp:359-439 */
int orc_xdrop_go(orc_xdrop *x, const uint8_t *query, int qstart, int qsize,
                 const uint8_t *target, int tstart, int tsize, int min_aln, orc_aln *out)
{
    static const char dec[] = "ACGT-";
    const long c0 = x->cells, r0 = x->rows, k0 = x->calls;
    extend_one_direction(x, query + qstart - 1, qstart, target + tstart - 1, tstart, 0, &x->lq, &x->lt);
    extend_one_direction(x, query + qstart, qsize - qstart, target + tstart, tsize - tstart, 1, &x->rq, &x->rt);

    const long need = x->lq.n + x->rq.n + 2;
    if (x->aln_cap < need) {
        x->aln_cap = need * 2;
        x->qaln = (char *)realloc(x->qaln, (size_t)x->aln_cap);
        x->taln = (char *)realloc(x->taln, (size_t)x->aln_cap);
    }
    long idx = 0;
    int i = 0, j = 0;
    /* the farthest left column is skipped (:401-402: n = size-1, k starts at n-1) */
    for (long k = x->lq.n - 2; k >= 0; --k, ++idx) {
        int c = x->lq.p[k];
        if (c != GAP_CODE) ++i;
        x->qaln[idx] = dec[c];
        c = x->lt.p[k];
        if (c != GAP_CODE) ++j;
        x->taln[idx] = dec[c];
    }
    out->qoff = qstart - i;
    out->toff = tstart - j;
    i = j = 0;
    for (long k = 0; k < x->rq.n; ++k, ++idx) {
        int c = x->rq.p[k];
        if (c != GAP_CODE) ++i;
        x->qaln[idx] = dec[c];
        c = x->rt.p[k];
        if (c != GAP_CODE) ++j;
        x->taln[idx] = dec[c];
    }
    x->qaln[idx] = 0;
    x->taln[idx] = 0;
    out->aln_size = (int)idx;
    out->qend = qstart + i;
    out->tend = tstart + j;
    out->qaln = x->qaln;
    out->taln = x->taln;
    out->cells = x->cells - c0;
    out->rows = x->rows - r0;
    out->calls = x->calls - k0;
    out->ok = (out->qend - out->qoff >= min_aln);
    return out->ok;
}

/* get_dna_encode_table + the ">3 -> 0" clamp: MC/defs.cpp:3-36, M2R/mecat2ref_aux.cpp:195-197 */
static inline uint8_t encode_base(unsigned char c)
{
    switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return 0;
    }
}

/* extend_candidate + extract_sequences -- M2R/mecat2ref_aux.cpp:171-270 */
int orc_extend_candidate(orc_xdrop *x, const char *ref, long ref_size, const char *read, int read_len,
                         long loc1, long loc2, long rec[4], orc_aln *out)
{
    const int read_start = (int)loc2;
    const long ref_start = loc1 - 1;
    const long L1 = read_start, R1 = read_len - read_start;
    const long L2 = ref_start, R2 = ref_size - ref_start;
    const long L = L1 < L2 ? L1 : L2, R = R1 < R2 ? R1 : R2;
    long lcap = (long)(L * 1.2), rcap = (long)(R * 1.2);
    const long left = L2 < lcap ? L2 : lcap;
    const long right = R2 < rcap ? R2 : rcap;
    const long tsize = left + right;
    if (x->tc_cap < tsize + 1) {
        x->tc_cap = tsize + 1;
        x->tcodes = (uint8_t *)realloc(x->tcodes, (size_t)x->tc_cap);
    }
    if (x->qc_cap < read_len + 1) {
        x->qc_cap = read_len + 1;
        x->qcodes = (uint8_t *)realloc(x->qcodes, (size_t)x->qc_cap);
    }
    const char *rs = ref + ref_start - left;
    for (long k = 0; k < tsize; ++k) x->tcodes[k] = encode_base((unsigned char)rs[k]);
    for (int k = 0; k < read_len; ++k) x->qcodes[k] = encode_base((unsigned char)read[k]);
    int ok = orc_xdrop_go(x, x->qcodes, read_start, read_len, x->tcodes, (int)left, (int)tsize, 1000, out);
    rec[0] = out->qoff;
    rec[1] = out->qend;
    rec[2] = ref_start - left + out->toff;
    rec[3] = ref_start - left + out->tend;
    return ok;
}
