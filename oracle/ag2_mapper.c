/*
 * ag2_mapper.c -- CPU restatement of mecat2ref+'s index build, seeding, candidate scoring, rescue and
 * per-read output.  TEST INFRASTRUCTURE ONLY (see ag2_oracle.h).
 *
 * Follows /root/reference/mecat_plus/MECAT-master_1/src/mecat2ref/ (M2R/):
 *   mecat2ref_impl_large.cpp  build_read_index :258-399, creat_ref_index :402-566, get_vote :568-608,
 *                             transnum_buchang :95-121, insert_loc :123-170, insert_loc3 :211-256,
 *                             find_location3 :609-693, reference_mapping :696-1327
 *   mecat2ref_aux.cpp         find_location2 :92-170, fill_clipped_candidate :272-299,
 *                             find_left/right_clipped_candidate :301-361, rescue_clipped_align :396-538,
 *                             output_results :541-559
 *   output.cpp                output_temp_result :237-251
 *
 * Pinned: tests/test_oracle_mapper.py compares the thread file this writes with the `<wrk>/1.r` the
 * unmodified reference binary (oracle/_ref/mecat2ref -t 1) writes for the same inputs, byte for byte,
 * and with the committed golden copy under tests/golden/.
 *
 * The float/double mix of the reference's consistency tests is kept expression by expression
 * (x86-64 SSE2, no FMA): see ddf_ok_f (float subtraction, find_location*) vs ddf_ok_d (double, insert_loc*).
 */
#include "ag2_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define SEED_LEN 13
#define NCODES (1 << (2 * SEED_LEN))
#define ZV 1000
#define ZVS 2000
#define SM 20
#define SI 21
#define SVM 100000
#define MAXSTR 1000000000L
#define CLIPPED 2000
#define DDFS_CUTOFF 0.25 /* ddfs_cutoff_pacbio, never switched (impl_large.cpp:18-20) */

static inline int atct(unsigned char c) /* atcttrans (:75-82): A0 T1 C2 G3, else 4 */
{
    switch (c) {
    case 'A': return 0;
    case 'T': return 1;
    case 'C': return 2;
    case 'G': return 3;
    default: return 4;
    }
}

/* rolling 13-mer scan shared by build_read_index (:340-372) and creat_ref_index (:465-493, :514-550) */
typedef void (*kmer_cb)(void *ctx, unsigned code, long i);
static void scan_kmers(const char *seq, long n, kmer_cb cb, void *ctx)
{
    unsigned eit = 0;
    long start = 0;
    const int leftnum = 34 - 2 * SEED_LEN;
    for (long i = 0; i < n; ++i) {
        const int t = atct((unsigned char)seq[i]);
        if (t == 4) {
            eit = 0;
            start = 0;
            continue;
        }
        eit = (eit << 2) + (unsigned)t;
        ++start;
        if (start >= SEED_LEN) {
            cb(ctx, eit, i);
            eit = (eit << leftnum) >> leftnum;
        }
    }
}

static void count_cb(void *ctx, unsigned code, long i)
{
    (void)i;
    ++((int *)ctx)[code];
}

static void mask_counts(int *c) /* sumvalue_x (:84-93): counts above 128 become 0 */
{
    for (long i = 0; i < NCODES; ++i)
        if (c[i] > 128) c[i] = 0;
}

/* build_read_index: masked 13-mer counts of the concatenated read prefix.  counts[4^13] is overwritten. */
void orc_read_hist13(const char *seq, long n, int *counts)
{
    memset(counts, 0, sizeof(int) * (size_t)NCODES);
    scan_kmers(seq, n, count_cb, counts);
    mask_counts(counts);
}

/* which reads enter the read index: the first <= 100 000 reads while the running length (+1 per read)
 * stays below 1e9 (:277) */
long orc_read_index_prefix(const long *offs, long nreads)
{
    long lenl = 0, k = 0;
    while (k < nreads && k < SVM && lenl < MAXSTR) {
        lenl += (offs[k + 1] - offs[k]) + 1;
        ++k;
    }
    return k;
}

typedef struct {
    orc_index *ix;
    int *fill;
} fill_ctx;

static void fill_cb(void *vctx, unsigned code, long i)
{
    fill_ctx *c = (fill_ctx *)vctx;
    orc_index *ix = c->ix;
    if (ix->cnt[code] > 0) ix->pos[ix->off[code] + (uint32_t)c->fill[code]++] = (uint32_t)(i + 2 - SEED_LEN);
    const long nn = (i + 2 - SEED_LEN) / ix->cbl;
    if (ix->rcnt[code] > 0) ix->kcount[nn] += ix->rcnt[code]; /* (:543-546) */
}

/* creat_ref_index + get_vote.  ref: concatenated, already upper-cased where > 'Z' (:432-437). */
orc_index *orc_index_build(const char *ref, long R, const int *rcnt, int cbl, double alpha, double beta)
{
    orc_index *ix = (orc_index *)calloc(1, sizeof(*ix));
    ix->R = R;
    ix->cbl = cbl;
    ix->ref = (char *)malloc((size_t)R + 1);
    memcpy(ix->ref, ref, (size_t)R);
    ix->ref[R] = 0;
    ix->rcnt = (int *)malloc(sizeof(int) * (size_t)NCODES);
    memcpy(ix->rcnt, rcnt, sizeof(int) * (size_t)NCODES);
    ix->cnt = (int *)calloc((size_t)NCODES, sizeof(int));
    scan_kmers(ix->ref, R, count_cb, ix->cnt);
    mask_counts(ix->cnt);
    ix->off = (uint32_t *)malloc(sizeof(uint32_t) * ((size_t)NCODES + 1));
    uint32_t sum = 0;
    for (long i = 0; i < NCODES; ++i) {
        ix->off[i] = sum;
        sum += (uint32_t)ix->cnt[i];
    }
    ix->off[NCODES] = sum;
    ix->pos = (uint32_t *)malloc(sizeof(uint32_t) * ((size_t)sum + 1));
    ix->nblk = R / cbl + 1;
    ix->kcount = (int *)calloc((size_t)ix->nblk + 10, sizeof(int));
    ix->vote = (float *)calloc((size_t)ix->nblk + 10, sizeof(float));
    fill_ctx fc = {ix, (int *)calloc((size_t)NCODES, sizeof(int))};
    scan_kmers(ix->ref, R, fill_cb, &fc);
    free(fc.fill);
    /* get_vote (:568-608) */
    long total = 0;
    for (long j = 0; j < ix->nblk; ++j) total += ix->kcount[j];
    const float ave = (float)(total / ix->nblk); /* integer division first */
    ix->ave = ave;
    for (long j = 0; j < ix->nblk; ++j) {
        if (ave == 0 || ix->kcount[j] == 0) {
            ix->vote[j] = 1;
        } else {
            float dev = ix->kcount[j] / ave;
            if (dev < alpha * 2) {
            } else if (dev > beta) {
            } else {
                dev = 1;
            }
            ix->vote[j] = dev;
        }
    }
    return ix;
}

void orc_index_free(orc_index *ix)
{
    if (!ix) return;
    free(ix->ref); free(ix->rcnt); free(ix->cnt); free(ix->off); free(ix->pos); free(ix->kcount); free(ix->vote);
    free(ix);
}

/* ---- per-read mapping ------------------------------------------------------------------------ */
typedef struct {
    int qid, qoff, qend;
    int parent_id, id, prev_id, next_id;
    char valid, qdir;
    long soff, send;
} aln_info; /* AlignInfo, mecat2ref_aux.h:9-20 */

typedef struct {
    int read_id, vscore, qb, qe, qs;
    char dir;
    long sb, se;
    char *qmap, *smap;
} temp_result;

struct orc_mapper {
    const orc_index *ix;
    int maxc, num_output;
    long nblk;
    orc_block *db[2];
    int *index_list[2];
    short *index_score[2];
    int nb[2];
    char *rev;
    long rev_cap;
    orc_xdrop *x;
    temp_result *results;
    aln_info *alns;
    int nresults, naln;
    orc_cand *cands;
    int ncand;
    orc_cand *last_cands; /* pass-1 candidates of the last read (for seeding parity checks) */
    int last_ncand, last_pass2;
    long cells, calls, aligned;
    long n_insert, n_rescue_ext, n_multi_cand, n_records2; /* branch coverage counters for the tests */
};

orc_mapper *orc_mapper_new(const orc_index *ix, int maxc, int num_output)
{
    orc_mapper *m = (orc_mapper *)calloc(1, sizeof(*m));
    m->ix = ix;
    m->maxc = maxc;
    m->num_output = num_output;
    m->nblk = ix->R / ZV + 5;
    for (int s = 0; s < 2; ++s) {
        m->db[s] = (orc_block *)calloc((size_t)m->nblk, sizeof(orc_block));
        m->index_list[s] = (int *)malloc(sizeof(int) * (size_t)m->nblk);
        m->index_score[s] = (short *)malloc(sizeof(short) * (size_t)m->nblk);
        for (long i = 0; i < m->nblk; ++i) m->db[s][i].index = -1;
    }
    m->x = orc_xdrop_new();
    m->results = (temp_result *)calloc((size_t)maxc + 6, sizeof(temp_result));
    m->alns = (aln_info *)calloc((size_t)maxc + 6, sizeof(aln_info));
    m->cands = (orc_cand *)calloc((size_t)maxc + 1, sizeof(orc_cand));
    m->last_cands = (orc_cand *)calloc((size_t)maxc + 1, sizeof(orc_cand));
    return m;
}

void orc_mapper_free(orc_mapper *m)
{
    if (!m) return;
    for (int s = 0; s < 2; ++s) {
        free(m->db[s]); free(m->index_list[s]); free(m->index_score[s]);
    }
    for (int i = 0; i < m->maxc + 6; ++i) {
        free(m->results[i].qmap); free(m->results[i].smap);
    }
    free(m->results); free(m->alns); free(m->cands); free(m->last_cands); free(m->rev);
    orc_xdrop_free(m->x);
    free(m);
}

long orc_mapper_cells(const orc_mapper *m) { long c; orc_xdrop_counters(m->x, &c, NULL, NULL); return c; }
long orc_mapper_calls(const orc_mapper *m) { long c; orc_xdrop_counters(m->x, NULL, NULL, &c); return c; }
long orc_mapper_aligned(const orc_mapper *m) { return m->aligned; }

int orc_mapper_last_candidates(const orc_mapper *m, orc_cand *out, int *pass2)
{
    memcpy(out, m->last_cands, sizeof(orc_cand) * (size_t)m->last_ncand);
    if (pass2) *pass2 = m->last_pass2;
    return m->last_ncand;
}

/* the two flavours of the distance-difference-factor test */
static inline int ddf_ok_f(int dloc, int dseed, float len) /* find_location2/3: float arithmetic throughout */
{
    return fabsf(dloc / (dseed * len) - 1) < DDFS_CUTOFF;
}
static inline int ddf_ok_d(int dloc, int dseed, float len) /* insert_loc*: "- 1.0" promotes to double */
{
    return fabs(dloc / (dseed * len) - 1.0) < DDFS_CUTOFF;
}

/* insert_loc (:123-170) and insert_loc3 (:211-256): same code, block length 1000 / 2000 */
static void insert_loc(const orc_index *ix, orc_block *spr, int loc, int seedn, float len, long templong, long zvl)
{
    int list_loc[SI], list_score[SI], list_seed[SI], i, j, minval, mini;
    float score_sim[SI];
    for (i = 0; i < SM; i++) {
        list_loc[i] = spr->loczhi[i];
        list_seed[i] = spr->seedno[i];
        list_score[i] = 0;
    }
    list_loc[SM] = loc;
    list_seed[SM] = seedn;
    list_score[SM] = 0;
    mini = -1;
    minval = 10000;
    for (i = 0; i < SM; i++)
        for (j = i + 1; j < SI; j++)
            if (list_seed[j] - list_seed[i] > 0 && list_loc[j] - list_loc[i] > 0 &&
                ddf_ok_d(list_loc[j] - list_loc[i], list_seed[j] - list_seed[i], len)) {
                list_score[i]++;
                list_score[j]++;
            }
    for (i = 0; i < SI; i++) {
        const int _loc = (int)(templong * zvl + list_loc[i]);
        const int nn = _loc / ix->cbl;
        score_sim[i] = list_score[i] / ix->vote[nn];
    }
    for (i = 0; i < SI; i++)
        if (minval > score_sim[i]) {
            minval = (int)score_sim[i]; /* int minval: truncation is part of the behaviour */
            mini = i;
        }
    if (mini == SM) {
        spr->loczhi[SM - 1] = (short)loc;
        spr->seedno[SM - 1] = (short)seedn;
    } else if (mini < SM) {
        for (i = mini; i < SM; i++) {
            spr->loczhi[i] = (short)list_loc[i + 1];
            spr->seedno[i] = (short)list_seed[i + 1];
        }
        spr->score--;
    }
}

/* find_location3 (:609-693) when sc != NULL, find_location2 (aux.cpp:92-170) when ix == NULL */
static int find_location(const orc_index *ix, const int *t_loc, const int *t_seedn, int *t_score, long *loc, int k,
                         int *rep_loc, float len, int read_len1, long start_loc)
{
    int i, j, maxval = 0, maxi = 0, rep = 0, lasti = 0;
    for (i = 0; i < k; i++) t_score[i] = 0;
    for (i = 0; i < k - 1; i++)
        for (j = i + 1; j < k; j++)
            if (t_seedn[j] - t_seedn[i] > 0 && t_loc[j] - t_loc[i] > 0 && t_loc[j] - t_loc[i] < read_len1 &&
                ddf_ok_f(t_loc[j] - t_loc[i], t_seedn[j] - t_seedn[i], len)) {
                t_score[i]++;
                t_score[j]++;
            }
    if (ix)
        for (i = 0; i < k; i++) {
            const long nn = (start_loc + t_loc[i]) / ix->cbl;
            t_score[i] = (int)(t_score[i] / ix->vote[nn]);
        }
    for (i = 0; i < k; i++) {
        if (maxval < t_score[i]) {
            maxval = t_score[i];
            maxi = i;
            rep = 0;
        } else if (maxval == t_score[i]) {
            rep++;
            lasti = i;
        }
    }
    for (i = 0; i < 4; i++) loc[i] = 0;
    if (maxval >= 5 && rep == maxval) {
        loc[0] = t_loc[maxi], loc[1] = t_seedn[maxi];
        *rep_loc = maxi;
        loc[2] = t_loc[lasti], loc[3] = t_seedn[lasti];
        return 1;
    } else if (maxval >= 5 && rep != maxval) {
        for (j = 0; j < maxi; j++)
            if (t_seedn[maxi] - t_seedn[j] > 0 && t_loc[maxi] - t_loc[j] > 0 && t_loc[maxi] - t_loc[j] < read_len1 &&
                ddf_ok_f(t_loc[maxi] - t_loc[j], t_seedn[maxi] - t_seedn[j], len)) {
                if (loc[0] == 0) {
                    loc[0] = t_loc[j];
                    loc[1] = t_seedn[j];
                    *rep_loc = j;
                } else {
                    loc[2] = t_loc[j];
                    loc[3] = t_seedn[j];
                }
            }
        j = maxi;
        if (loc[0] == 0) {
            loc[0] = t_loc[j];
            loc[1] = t_seedn[j];
            *rep_loc = j;
        } else {
            loc[2] = t_loc[j];
            loc[3] = t_seedn[j];
        }
        for (j = maxi + 1; j < k; j++)
            if (t_seedn[j] - t_seedn[maxi] > 0 && t_loc[j] - t_loc[maxi] > 0 && t_loc[j] - t_loc[maxi] <= read_len1 &&
                ddf_ok_f(t_loc[j] - t_loc[maxi], t_seedn[j] - t_seedn[maxi], len)) {
                if (loc[0] == 0) {
                    loc[0] = t_loc[j];
                    loc[1] = t_seedn[j];
                    *rep_loc = j;
                } else {
                    loc[2] = t_loc[j];
                    loc[3] = t_seedn[j];
                }
            }
        return 1;
    }
    return 0;
}

/* seeding (:842-878 / :1104-1147) + candidate scan (:882-991 / :1151-1260) for one strand */
static void seed_and_scan(orc_mapper *m, const char *seq, int read_len, int BC, int strand, long zv, int thresh)
{
    const orc_index *ix = m->ix;
    orc_block *database = m->db[strand];
    int *index_list = m->index_list[strand];
    short *index_score = m->index_score[strand];
    int j = 0;
    /* transnum_buchang (:95-121) */
    const int cleave_num = (read_len - SEED_LEN) / BC + 1;
    for (int k = 0; k < cleave_num; k++) {
        int eit = 0;
        const int start = k * BC;
        for (int q = 0; q < SEED_LEN; q++) {
            const int t = (start + q < read_len) ? atct((unsigned char)seq[start + q]) : 4; /* the NUL ends a short read */
            if (t == 4) {
                eit = -1;
                break;
            }
            eit = (eit << 2) + t;
        }
        if (eit < 0) continue;
        const int count1 = ix->cnt[eit];
        const uint32_t *lead = ix->pos + ix->off[eit];
        for (int i = 0; i < count1; i++, lead++) {
            const long templong = (long)(*lead) / zv;
            const long u_k = (long)(*lead) % zv;
            orc_block *spr = database + templong;
            if (spr->score == 0 || spr->seednum < k + 1) {
                const long loc = ++(spr->score);
                if (loc <= SM) {
                    spr->loczhi[loc - 1] = (short)u_k;
                    spr->seedno[loc - 1] = (short)(k + 1);
                } else {
                    insert_loc(ix, spr, (int)u_k, k + 1, (float)BC, templong, zv);
                    ++m->n_insert;
                }
                long s_k;
                if (templong > 0) s_k = spr->score + (spr - 1)->score;
                else s_k = spr->score;
                if (spr->index == -1) {
                    index_list[j] = (int)templong;
                    index_score[j] = (short)s_k;
                    spr->index = j;
                    j++;
                } else {
                    index_score[spr->index] = (short)s_k;
                }
                spr->score2 = spr->score;
            }
            spr->seednum = (short)(k + 1);
        }
    }
    m->nb[strand] = j;
    const int cc1 = j;
    int temp_list[200], temp_seedn[200], temp_score[200];
    for (int i = 0; i < cc1; i++) {
        if (!(index_score[i] > thresh)) continue;
        const int bid = index_list[i];
        orc_block *spr = database + bid;
        if (spr->score == 0) continue;
        long s_k = spr->score, loc = 0, start_loc = (long)bid * zv;
        if (bid > 0) {
            loc = (spr - 1)->score;
            if (loc > 0) start_loc = (long)(bid - 1) * zv;
        }
        long u_k = 0;
        if (loc == 0) {
            for (int q = 0; q < s_k && q < SM; q++) {
                temp_list[u_k] = spr->loczhi[q];
                temp_seedn[u_k] = spr->seedno[q];
                u_k++;
            }
        } else {
            const orc_block *spr1 = spr - 1;
            for (int q = 0; q < loc && q < SM; q++) {
                temp_list[u_k] = spr1->loczhi[q];
                temp_seedn[u_k] = spr1->seedno[q];
                u_k++;
            }
            for (int q = 0; q < s_k && q < SM; q++) {
                temp_list[u_k] = spr->loczhi[q] + (int)zv;
                temp_seedn[u_k] = spr->seedno[q];
                u_k++;
            }
        }
        long location_loc[4];
        int repeat_loc = 0;
        if (!find_location(ix, temp_list, temp_seedn, temp_score, location_loc, (int)u_k, &repeat_loc, (float)BC, read_len, start_loc))
            continue;
        if (temp_score[repeat_loc] < 6) continue;
        orc_cand ct;
        memset(&ct, 0, sizeof ct);
        ct.score = temp_score[repeat_loc];
        const int loc_seed = temp_seedn[repeat_loc];
        location_loc[0] = start_loc + location_loc[0];
        location_loc[1] = (location_loc[1] - 1) * BC;
        const long loc_list = location_loc[0];
        const long left_length1 = location_loc[0] + SEED_LEN - 1, right_length1 = ix->R - location_loc[0];
        const long left_length2 = location_loc[1] + SEED_LEN - 1, right_length2 = read_len - location_loc[1];
        const int num1 = (int)(left_length1 >= left_length2 ? left_length2 : left_length1);
        const int num2 = (int)(right_length1 >= right_length2 ? right_length2 : right_length1);
        int seedcount = 0;
        ct.loc1 = location_loc[0];
        ct.num1 = num1;
        ct.loc2 = location_loc[1];
        ct.num2 = num2;
        ct.left1 = left_length1;
        ct.left2 = left_length2;
        ct.right1 = right_length1;
        ct.right2 = right_length2;
        /* consistent seeds in the blocks to the left (:950-961) and right (:963-973) */
        {
            long ub = bid - 2;
            int k = num1 / (int)zv;
            for (orc_block *s1 = spr - 2; ub >= 0 && k >= 0; s1--, k--, ub--)
                if (s1->score > 0) {
                    const long sl = ub * zv;
                    const int scnt = s1->score < SM ? s1->score : SM;
                    int sk = 0;
                    for (int q = 0; q < scnt; q++)
                        if (fabs((loc_list - sl - s1->loczhi[q]) / ((loc_seed - s1->seedno[q]) * BC * 1.0) - 1.0) < DDFS_CUTOFF) {
                            seedcount++;
                            sk++;
                        }
                    if (sk * 1.0 / scnt > 0.4) s1->score = 0;
                }
        }
        {
            long ub = bid + 1;
            int k = num2 / (int)zv;
            for (orc_block *s1 = spr + 1; k > 0; s1++, k--, ub++)
                if (s1->score > 0) {
                    const long sl = ub * zv;
                    const int scnt = s1->score < SM ? s1->score : SM;
                    int sk = 0;
                    for (int q = 0; q < scnt; q++)
                        if (fabs((sl + s1->loczhi[q] - loc_list) / ((s1->seedno[q] - loc_seed) * BC * 1.0) - 1.0) < DDFS_CUTOFF) {
                            seedcount++;
                            sk++;
                        }
                    if (sk * 1.0 / scnt > 0.4) s1->score = 0;
                }
        }
        ct.score += seedcount;
        ct.chain = strand == 0 ? 'F' : 'R';
        /* keep the MAXC best, ties after equals (:978-990) */
        int low = 0, high = m->ncand - 1;
        while (low <= high) {
            const int mid = (low + high) / 2;
            if (mid >= m->ncand || m->cands[mid].score < ct.score) high = mid - 1;
            else low = mid + 1;
        }
        if (m->ncand < m->maxc) {
            for (int q = m->ncand - 1; q > high; q--) m->cands[q + 1] = m->cands[q];
        } else {
            for (int q = m->ncand - 2; q > high; q--) m->cands[q + 1] = m->cands[q];
        }
        if (high + 1 < m->maxc) m->cands[high + 1] = ct;
        if (m->ncand < m->maxc) m->ncand++;
    }
}

static void reset_blocks(orc_mapper *m)
{
    for (int s = 0; s < 2; ++s)
        for (int t = 0; t < m->nb[s]; ++t) {
            orc_block *b = m->db[s] + m->index_list[s][t];
            b->score = 0;
            b->score2 = 0;
            b->index = -1;
        }
}

/* extend_candidate (aux.cpp:210-270) on top of orc_extend_candidate */
static int extend_cand(orc_mapper *m, const orc_cand *can, const char *fwd, const char *rev, int read_name, int read_len,
                       int with_alns)
{
    long rec[4];
    orc_aln a;
    const char *rd = can->chain == 'F' ? fwd : rev;
    if (!orc_extend_candidate(m->x, m->ix->ref, m->ix->R, rd, read_len, can->loc1, can->loc2, rec, &a)) return 0;
    temp_result *r = &m->results[m->nresults++];
    r->read_id = read_name;
    r->dir = can->chain;
    r->vscore = can->score;
    r->qb = (int)rec[0];
    r->qe = (int)rec[1];
    r->qs = read_len;
    r->sb = rec[2];
    r->se = rec[3];
    r->qmap = (char *)realloc(r->qmap, (size_t)a.aln_size + 1);
    r->smap = (char *)realloc(r->smap, (size_t)a.aln_size + 1);
    memcpy(r->qmap, a.qaln, (size_t)a.aln_size + 1);
    memcpy(r->smap, a.taln, (size_t)a.aln_size + 1);
    if (with_alns) {
        aln_info *ai = &m->alns[m->naln++];
        ai->qoff = r->qb;
        ai->qend = r->qe;
        ai->qdir = r->dir;
        ai->soff = r->sb;
        ai->send = r->se;
        ai->valid = 1;
        ai->id = m->nresults - 1;
        ai->prev_id = ai->next_id = ai->parent_id = -1;
    }
    return 1;
}

static void sort_alns(aln_info *a, int n) /* std::sort on <= 16 elements = insertion sort; key: span descending */
{
    for (int i = 1; i < n; ++i) {
        aln_info v = a[i];
        int j = i;
        if ((v.qend - v.qoff) > (a[0].qend - a[0].qoff)) {
            memmove(a + 1, a, sizeof(aln_info) * (size_t)i);
            a[0] = v;
            continue;
        }
        while ((v.qend - v.qoff) > (a[j - 1].qend - a[j - 1].qoff)) {
            a[j] = a[j - 1];
            --j;
        }
        a[j] = v;
    }
}

static int contained(const aln_info *a, const aln_info *b) /* AlignInfoContained, aux.h:28-42 */
{
    const int extra = 100;
    return a->qdir == b->qdir && b->qoff + extra >= a->qoff && b->qend <= a->qend + extra && b->soff + extra >= a->soff &&
           b->send <= a->send + extra;
}
static int full_align(const aln_info *a, int qsize) { return a->qend - a->qoff >= qsize * 0.9; }

static int left_clipped(const aln_info *a, const aln_info *b) /* is_left_clipped_align, aux.cpp:363-377 */
{
    if (a->qdir != b->qdir) return 0;
    if (abs(b->qend - a->qoff) <= 200 && a->soff - b->send > -200 && a->soff - b->send < 10000) return 1;
    if (labs(b->send - a->soff) <= 200 && a->qoff - b->qend > -200 && a->qoff - b->qend < 10000) return 1;
    return 0;
}
static int right_clipped(const aln_info *a, const aln_info *b) /* is_right_clipped_align, aux.cpp:379-393 */
{
    if (a->qdir != b->qdir) return 0;
    if (abs(a->qend - b->qoff) <= 200 && b->soff - a->send > -200 && b->soff - a->send < 10000) return 1;
    if (labs(a->send - b->soff) <= 200 && b->qoff - a->qend > -200 && b->qoff - a->qend < 10000) return 1;
    return 0;
}

static int fill_clipped(const orc_block *block, long bid, orc_cand *can, char chain, int read_size, int BC, int block_size)
{
    int seedn[SM], boff[SM], score[SM], rep_loc = 0;
    long locations[4];
    const int n = block->score2 < SM ? block->score2 : SM;
    for (int i = 0; i < n; ++i) {
        seedn[i] = block->seedno[i];
        boff[i] = block->loczhi[i];
        score[i] = 0;
    }
    if (find_location(NULL, boff, seedn, score, locations, n, &rep_loc, (float)BC, read_size, 0)) {
        can->score = score[rep_loc];
        can->chain = chain;
        can->loc1 = bid * block_size + locations[0];
        can->loc2 = (locations[1] - 1) * BC;
        return 1;
    }
    return 0;
}

static int find_left_clipped(const aln_info *aln, orc_cand *can, const orc_block *database, int block_size, int read_size, int BC)
{
    if (aln->qoff <= CLIPPED || aln->soff <= CLIPPED) return 0;
    const int n1 = aln->qoff / block_size;
    int n2 = (int)(aln->soff / block_size);
    int n = n1 < n2 ? n1 : n2;
    int max_score = 0;
    const orc_block *block = NULL;
    long bid = -1;
    for (--n2; n >= 0 && n2 >= 0; --n, --n2)
        if (database[n2].score2 > max_score) {
            max_score = database[n2].score2;
            block = database + n2;
            bid = n2;
        }
    if (block && block->score2 > 4) return fill_clipped(block, bid, can, aln->qdir, read_size, BC, block_size);
    return 0;
}

static int find_right_clipped(const aln_info *aln, orc_cand *can, const orc_block *database, int block_size, int read_size,
                              long ref_size, int BC)
{
    if (read_size - aln->qend <= CLIPPED || ref_size - aln->send <= CLIPPED) return 0;
    const int n1 = (read_size - aln->qend) / block_size;
    const int n2 = (int)((ref_size - aln->send) / block_size);
    int n = n1 < n2 ? n1 : n2;
    int max_score = 0;
    long bid = -1;
    const orc_block *block = NULL;
    long k = aln->send / block_size + 1;
    for (; n >= 0; --n, ++k)
        if (database[k].score2 > max_score) {
            max_score = database[k].score2;
            block = database + k;
            bid = k;
        }
    if (block && block->score2 > 4) return fill_clipped(block, bid, can, aln->qdir, read_size, BC, block_size);
    return 0;
}

/* rescue_clipped_align, aux.cpp:396-538 */
static void rescue(orc_mapper *m, const char *fwd, const char *rev, int read_name, int read_len, int block_size, int BC)
{
    aln_info *alnv = m->alns;
    int naln = m->naln;
    sort_alns(alnv, naln);
    for (int i = 0; i < naln - 1; ++i) {
        if (!alnv[i].valid) continue;
        for (int j = i + 1; j < naln; ++j) {
            if (!alnv[j].valid) continue;
            if (contained(&alnv[i], &alnv[j])) alnv[j].valid = 0;
        }
    }
    int k = 0;
    for (int i = 0; i < naln; ++i)
        if (alnv[i].valid) alnv[k++] = alnv[i];
    naln = k;
    m->naln = naln;
    if (naln == 0) return; /* the reference reads alnv[0] uninitialised here (:428); nothing follows either way */
    if (full_align(&alnv[0], read_len)) return;
    for (int i = 0; i < naln - 1; ++i) {
        if (alnv[i].parent_id != -1) continue;
        for (int j = i + 1; j < naln; ++j) {
            if (alnv[j].parent_id != -1) continue;
            if (alnv[i].prev_id != -1 && left_clipped(&alnv[i], &alnv[j])) {
                alnv[i].prev_id = alnv[j].id;
                alnv[j].parent_id = alnv[i].id;
            }
            if (alnv[i].next_id != -1 && right_clipped(&alnv[i], &alnv[j])) {
                alnv[i].next_id = alnv[j].id;
                alnv[j].parent_id = alnv[i].id;
            }
        }
    }
    const int n = naln < 3 ? naln : 3;
    k = 0;
    orc_cand can;
    memset(&can, 0, sizeof can);
    for (int i = 0; i < n; ++i) {
        if (alnv[i].parent_id != -1) continue;
        const orc_block *database = alnv[i].qdir == 'F' ? m->db[0] : m->db[1];
        for (int side = 0; side < 2; ++side) {
            int found;
            if (side == 0) found = alnv[i].prev_id == -1 && find_left_clipped(&alnv[i], &can, database, block_size, read_len, BC);
            else found = alnv[i].next_id == -1 && find_right_clipped(&alnv[i], &can, database, block_size, read_len, m->ix->R, BC);
            if (!found) continue;
            ++m->n_rescue_ext;
            if (!extend_cand(m, &can, fwd, rev, read_name, read_len, 0)) continue;
            const temp_result *rs = &m->results[m->nresults - 1];
            aln_info *ai = &alnv[naln + k];
            ai->qoff = rs->qb;
            ai->qend = rs->qe;
            ai->qdir = rs->dir;
            ai->soff = rs->sb;
            ai->send = rs->se;
            ai->valid = 1;
            ai->id = m->nresults - 1;
            ai->prev_id = ai->next_id = ai->parent_id = -1;
            if (side == 0 ? left_clipped(&alnv[i], ai) : right_clipped(&alnv[i], ai)) {
                ai->parent_id = alnv[i].id;
                if (side == 0) alnv[i].prev_id = ai->id;
                else alnv[i].next_id = ai->id;
                ++k;
            }
        }
    }
    if (!k) return;
    naln += k;
    sort_alns(alnv, naln);
    k = 0;
    for (int i = 0; i < naln; ++i)
        if (full_align(&alnv[i], read_len)) {
            alnv[i].parent_id = alnv[i].prev_id = alnv[i].next_id = -1;
            ++k;
        }
    if (k) naln = k;
    m->naln = naln;
}

static void write_result(orc_mapper *m, const temp_result *r, FILE *out) /* output_temp_result, output.cpp:237-251 */
{
    if (out)
        fprintf(out, "%d\t%c\t%d\t%d\t%d\t%d\t%ld\t%ld\n%s\n%s\n", r->read_id, r->dir, r->vscore, r->qb, r->qe, r->qs, r->sb, r->se,
                r->qmap, r->smap);
    m->aligned += r->qe - r->qb;
}

static int output_results(orc_mapper *m, FILE *out) /* aux.cpp:541-559 */
{
    int n = 0, written = 0;
    for (int i = 0; i < m->naln && n < m->num_output; ++i) {
        if (m->alns[i].parent_id != -1) continue;
        write_result(m, &m->results[m->alns[i].id], out);
        ++written;
        if (m->alns[i].prev_id != -1) { write_result(m, &m->results[m->alns[i].prev_id], out); ++written; }
        if (m->alns[i].next_id != -1) { write_result(m, &m->results[m->alns[i].next_id], out); ++written; }
        ++n;
    }
    return written;
}

/* seeding + candidate scan only (no extension), pass 0 or the second pass; returns #candidates in out[maxc] */
int orc_seed_candidates(orc_mapper *m, const char *read, int len, int pass, orc_cand *out)
{
    if (m->rev_cap < len + 1) {
        m->rev_cap = len + 1;
        m->rev = (char *)realloc(m->rev, (size_t)m->rev_cap);
    }
    for (int i = 0; i < len; ++i) {
        char c = read[len - 1 - i];
        switch (c) {
        case 'A': c = 'T'; break;
        case 'T': c = 'A'; break;
        case 'C': c = 'G'; break;
        case 'G': c = 'C'; break;
        default: break;
        }
        m->rev[i] = c;
    }
    m->rev[len] = 0;
    int BC = pass == 0 ? 5 + len / 1000 : 5;
    if (BC > 20) BC = 20;
    const long zv = pass == 0 ? ZV : ZVS;
    m->ncand = 0;
    seed_and_scan(m, read, len, BC, 0, zv, pass == 0 ? 6 : 4);
    seed_and_scan(m, m->rev, len, BC, 1, zv, pass == 0 ? 6 : 4);
    reset_blocks(m);
    memcpy(out, m->cands, sizeof(orc_cand) * (size_t)m->ncand);
    return m->ncand;
}

/* one read through reference_mapping's loop body (:776-1316).  Returns the number of records written. */
int orc_map_read(orc_mapper *m, int read_id, const char *read, int len, FILE *out)
{
    if (m->rev_cap < len + 1) {
        m->rev_cap = len + 1;
        m->rev = (char *)realloc(m->rev, (size_t)m->rev_cap);
    }
    for (int i = 0; i < len; ++i) { /* reverse, complement upper-case ACGT only (:799-833) */
        char c = read[len - 1 - i];
        switch (c) {
        case 'A': c = 'T'; break;
        case 'T': c = 'A'; break;
        case 'C': c = 'G'; break;
        case 'G': c = 'C'; break;
        default: break;
        }
        m->rev[i] = c;
    }
    m->rev[len] = 0;
    int written = 0;
    m->last_pass2 = 0;
    for (int pass = 0; pass < 2; ++pass) {
        int BC = pass == 0 ? 5 + len / 1000 : 5;
        if (BC > 20) BC = 20;
        const long zv = pass == 0 ? ZV : ZVS;
        m->ncand = 0;
        seed_and_scan(m, read, len, BC, 0, zv, pass == 0 ? 6 : 4);
        seed_and_scan(m, m->rev, len, BC, 1, zv, pass == 0 ? 6 : 4);
        if (m->ncand > 1) ++m->n_multi_cand;
        if (pass == 0) {
            m->last_ncand = m->ncand;
            memcpy(m->last_cands, m->cands, sizeof(orc_cand) * (size_t)m->ncand);
        }
        m->naln = 0;
        m->nresults = 0;
        for (int i = 0; i < m->ncand; ++i) extend_cand(m, &m->cands[i], read, m->rev, read_id, len, 1);
        const int naln_ext = m->naln;
        rescue(m, read, m->rev, read_id, len, (int)zv, BC);
        written += output_results(m, out);
        reset_blocks(m);
        if (naln_ext != 0) break; /* the second pass only runs when no candidate extended (:1049) */
        m->last_pass2 = 1;
    }
    return written;
}

/* The mapping part of meap_ref_impl_large (:1994-2149) for in-memory inputs, one "thread file":
 * read index over the prefix of the batch, reference index + votes, then every read in order.
 * ids[i] is the read id chang_fastqfile assigned (mecat2ref.cpp:298,319).  Returns records written. */
long orc_map_batch(const char *ref, long R, const char *reads, const long *offs, const int *ids, long n, int cbl, double alpha,
                   double beta, int maxc, int num_output, const char *r_path, long *stats /* [8]: cells calls aligned pass2_reads insert_loc rescue_extensions multi_candidate_reads votes!=1 */)
{
    int *rc = (int *)malloc(sizeof(int) * (size_t)NCODES);
    const long pre = orc_read_index_prefix(offs, n);
    orc_read_hist13(reads, offs[pre], rc);
    orc_index *ix = orc_index_build(ref, R, rc, cbl, alpha, beta);
    free(rc);
    orc_mapper *m = orc_mapper_new(ix, maxc, num_output);
    FILE *out = r_path ? fopen(r_path, "w") : NULL;
    long written = 0, pass2 = 0;
    char *buf = NULL;
    long cap = 0;
    for (long i = 0; i < n; ++i) {
        const long len = offs[i + 1] - offs[i];
        if (cap < len + 1) {
            cap = len + 1;
            buf = (char *)realloc(buf, (size_t)cap);
        }
        memcpy(buf, reads + offs[i], (size_t)len);
        buf[len] = 0;
        written += orc_map_read(m, ids[i], buf, (int)len, out);
        pass2 += m->last_pass2;
    }
    if (out) fclose(out);
    if (stats) {
        stats[0] = orc_mapper_cells(m);
        stats[1] = orc_mapper_calls(m);
        stats[2] = orc_mapper_aligned(m);
        stats[3] = pass2;
        stats[4] = m->n_insert;
        stats[5] = m->n_rescue_ext;
        stats[6] = m->n_multi_cand;
        long masked = 0, votes = 0;
        for (long i = 0; i < NCODES; ++i) masked += (ix->cnt[i] == 0 && ix->off[i + 1] == ix->off[i]) ? 0 : 0;
        for (long j = 0; j < ix->nblk; ++j) votes += ix->vote[j] != 1.0f;
        stats[7] = votes;
        (void)masked;
    }
    free(buf);
    orc_mapper_free(m);
    orc_index_free(ix);
    return written;
}
