// seed_device.cuh -- device code of mecat2ref+'s seeding and candidate scoring.
//
// One THREAD maps one read: it runs the reference's seeding loop (M2R/mecat2ref_impl_large.cpp:842-878),
// candidate scan (:882-991), find_location3 (:609-693) and insert_loc (:123-170) statement by statement,
// forward strand first, then the reverse strand, into one top-MAXC candidate list -- the order the
// reference's ties depend on.  The only structural change is the block table: the reference gives every
// worker a DENSE `Back_List[ref_len/1000]` per strand (46 MB at 250 Mb); here every read gets a small
// open-addressing hash table keyed by block id, sized from the read's own hit count, in a scratch
// arena.  A block that was never touched reads as score = 0, exactly what the dense table holds after the
// per-read reset (:1036-1047); the first-touch order (`index_list`) is kept in an array as in the reference.
//
// The reference's float/double mix is kept expression by expression with round-to-nearest intrinsics
// (no FMA contraction): ddf_ok_f is find_location3's float test, ddf_ok_d insert_loc's double test.
//
// Compiled by nvcc into libag2_b200.so and, unchanged with -DAG2_EMU, by g++ for the CPU-side tests.
#pragma once

#ifdef AG2_EMU
#include "warp_emu.h"
#include <cmath>
#else
#include <cuda_runtime.h>
#endif
#include <stdint.h>

namespace ag2 {

constexpr int kSeedLen = 13;
constexpr int kSM = 20, kSI = 21;
constexpr int kMaxCand = 16; // capacity of the per-read candidate list; -n (MAXC) must be <= this

#ifdef AG2_EMU
inline float fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
inline float fmul_rn(float a, float b) { volatile float r = a * b; return r; }
inline float fsub_rn(float a, float b) { volatile float r = a - b; return r; }
inline double ddiv_rn(double a, double b) { volatile double r = a / b; return r; }
inline double dmul_rn(double a, double b) { volatile double r = a * b; return r; }
inline double dsub_rn(double a, double b) { volatile double r = a - b; return r; }
#else
__device__ __forceinline__ float fdiv_rn(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float fmul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fsub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double ddiv_rn(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ double dmul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dsub_rn(double a, double b) { return __dsub_rn(a, b); }
#endif

struct RefIndex {           // device-resident creat_ref_index + get_vote result
    int64_t ref_len;
    const int32_t *cnt;     // masked 13-mer counts [4^13]            (countin)
    const uint32_t *off;    // CSR offsets [4^13 + 1]
    const uint32_t *pos;    // 1-based k-mer starts, ascending per bucket (allloc)
    const float *vote;      // per similarity block                    (sim::vote)
    int32_t cbl;            // similarity block size (-z)
};

struct BackList {           // Back_List, M2R/mecat2ref_defs.h:84-88
    int16_t score, score2, loczhi[kSM], seedno[kSM], seednum;
    int32_t index;
};

struct alignas(8) BlockSlot {   // open addressing over 8-byte slots; the 92-byte Back_Lists live in a pool, one per block touched
    int32_t key;            // block id, -1 = empty
    int32_t val;            // index into BlockTable::pool
};

struct SeedCand {           // candidate_save, M2R/mecat2ref_defs.h:90-95
    int64_t loc1, loc2, left1, left2, right1, right2;
    int32_t score, num1, num2;
    int32_t chain;          // 'F' / 'R'
};

struct BlockTable {
    BlockSlot *slots;
    BackList *pool;         // [hits + 1]: a hit touches at most one new block
    mutable int32_t n_pool;
    uint32_t mask;          // capacity - 1 (power of two)
    int32_t *index_list;    // first-touch order
    int16_t *index_score;
    int32_t n_index;
};

__device__ __forceinline__ uint32_t block_hash(int32_t k) { return (uint32_t)k * 2654435761u; }

__device__ __forceinline__ BackList *table_find(const BlockTable &t, int32_t key)
{
    uint32_t h = (block_hash(key) >> 7) & t.mask;
    for (;;) {
        const BlockSlot s = t.slots[h];
        if (s.key == key) return &t.pool[s.val];
        if (s.key == -1) return nullptr;
        h = (h + 1) & t.mask;
    }
}

__device__ __forceinline__ BackList *table_get(const BlockTable &t, int32_t key) // find or create
{
    uint32_t h = (block_hash(key) >> 7) & t.mask;
    for (;;) {
        const BlockSlot s = t.slots[h];
        if (s.key == key) return &t.pool[s.val];
        if (s.key == -1) {
            const BlockSlot fresh = {key, t.n_pool};
            t.slots[h] = fresh;
            BackList *b = &t.pool[t.n_pool++];
            b->score = 0;
            b->score2 = 0;
            b->seednum = 0;
            b->index = -1;
            return b;
        }
        h = (h + 1) & t.mask;
    }
}

__device__ __forceinline__ int table_score(const BlockTable &t, int32_t key)
{
    const BackList *b = key >= 0 ? table_find(t, key) : nullptr;
    return b ? b->score : 0;
}

__device__ __forceinline__ bool ddf_ok_f(int dloc, int dseed, float len)
{
    return fabsf(fsub_rn(fdiv_rn((float)dloc, fmul_rn((float)dseed, len)), 1.0f)) < 0.25;
}
__device__ __forceinline__ bool ddf_ok_d(int dloc, int dseed, float len)
{
    return fabs(dsub_rn((double)fdiv_rn((float)dloc, fmul_rn((float)dseed, len)), 1.0)) < 0.25;
}

// atcttrans (:75-82) on a base stored as A0 C1 G2 T3: A0 T1 C2 G3
__device__ __forceinline__ int atct_of_code(int c) { return (0x1320 >> (4 * c)) & 3; }

// insert_loc (:123-170) / insert_loc3 (:211-256)
__device__ void insert_loc(const RefIndex &ix, BackList *spr, int loc, int seedn, float len, int64_t templong, int64_t zvl)
{
    int list_loc[kSI], list_score[kSI], list_seed[kSI], i, j, minval, mini;
    float score_sim[kSI];
    for (i = 0; i < kSM; i++) {
        list_loc[i] = spr->loczhi[i];
        list_seed[i] = spr->seedno[i];
        list_score[i] = 0;
    }
    list_loc[kSM] = loc;
    list_seed[kSM] = seedn;
    list_score[kSM] = 0;
    mini = -1;
    minval = 10000;
    for (i = 0; i < kSM; i++)
        for (j = i + 1; j < kSI; j++)
            if (list_seed[j] - list_seed[i] > 0 && list_loc[j] - list_loc[i] > 0 &&
                ddf_ok_d(list_loc[j] - list_loc[i], list_seed[j] - list_seed[i], len)) {
                list_score[i]++;
                list_score[j]++;
            }
    for (i = 0; i < kSI; i++) {
        const int _loc = (int)(templong * zvl + list_loc[i]);
        const int nn = _loc / ix.cbl;
        score_sim[i] = fdiv_rn((float)list_score[i], ix.vote[nn]);
    }
    for (i = 0; i < kSI; i++)
        if ((float)minval > score_sim[i]) {
            minval = (int)score_sim[i];
            mini = i;
        }
    if (mini == kSM) {
        spr->loczhi[kSM - 1] = (int16_t)loc;
        spr->seedno[kSM - 1] = (int16_t)seedn;
    } else if (mini < kSM) {
        for (i = mini; i < kSM; i++) {
            spr->loczhi[i] = (int16_t)list_loc[i + 1];
            spr->seedno[i] = (int16_t)list_seed[i + 1];
        }
        spr->score--;
    }
}

// Second half of find_location3 (:628-693) / find_location2: t_score holds the (vote-scaled) consistency counts; picks the
// best seed, its first and last consistent partners.  Shared by the one-thread-per-read path and the CTA path.
__device__ int find_location_choose(const int *t_loc, const int *t_seedn, const int *t_score, int64_t *loc, int k, int *rep_loc, float len,
                                    int read_len1)
{
    int i, j, maxval = 0, maxi = 0, rep = 0, lasti = 0;
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

// find_location3 (:609-693)
__device__ int find_location3(const RefIndex &ix, const int *t_loc, const int *t_seedn, int *t_score, int64_t *loc, int k,
                              int *rep_loc, float len, int read_len1, int64_t start_loc)
{
    int i, j;
    for (i = 0; i < k; i++) t_score[i] = 0;
    for (i = 0; i < k - 1; i++)
        for (j = i + 1; j < k; j++)
            if (t_seedn[j] - t_seedn[i] > 0 && t_loc[j] - t_loc[i] > 0 && t_loc[j] - t_loc[i] < read_len1 &&
                ddf_ok_f(t_loc[j] - t_loc[i], t_seedn[j] - t_seedn[i], len)) {
                t_score[i]++;
                t_score[j]++;
            }
    for (i = 0; i < k; i++) {
        const int64_t nn = (start_loc + t_loc[i]) / ix.cbl;
        t_score[i] = (int)fdiv_rn((float)t_score[i], ix.vote[nn]);
    }
    return find_location_choose(t_loc, t_seedn, t_score, loc, k, rep_loc, len, read_len1);
}

// One seed of a read strand: the 13-mer at oriented position `start` in atct code, or -1 if it holds a base
// that is not upper-case ACGT (transnum_buchang :95-121).  strand 1 = reversed, ACGT complemented.
__device__ __forceinline__ int seed_code(const uint32_t *reads2, const uint32_t *irr, int64_t roff, int rlen, int strand, int start)
{
    int eit = 0;
    for (int q = 0; q < kSeedLen; ++q) {
        const int p = start + q;
        if (p >= rlen) return -1;
        const int64_t fp = roff + (strand ? rlen - 1 - p : p);
        if ((irr[fp >> 5] >> (fp & 31)) & 1u) return -1;
        int c = (int)((reads2[fp >> 4] >> (2 * (fp & 15))) & 3u);
        if (strand) c ^= 3;
        eit = (eit << 2) + atct_of_code(c);
    }
    return eit;
}

// number of index hits of a read strand = upper bound of the blocks it can touch (sizes the hash table)
__device__ int64_t count_hits(const RefIndex &ix, const uint32_t *reads2, const uint32_t *irr, int64_t roff, int rlen, int strand, int BC)
{
    const int cleave_num = (rlen - kSeedLen) / BC + 1;
    int64_t hits = 0;
    for (int k = 0; k < cleave_num; ++k) {
        const int code = seed_code(reads2, irr, roff, rlen, strand, k * BC);
        if (code >= 0) hits += ix.cnt[code];
    }
    return hits;
}

// seeding (:842-878) + candidate scan (:882-991) of one strand; cands[0..ncand) is the read's shared list
__device__ void seed_and_scan(const RefIndex &ix, const uint32_t *reads2, const uint32_t *irr, int64_t roff, int rlen, int strand,
                              int BC, int64_t zv, int thresh, int maxc, BlockTable &tb, SeedCand *cands, int &ncand)
{
    int j = 0;
    const int cleave_num = (rlen - kSeedLen) / BC + 1;
    for (int k = 0; k < cleave_num; k++) {
        const int eit = seed_code(reads2, irr, roff, rlen, strand, k * BC);
        if (eit < 0) continue;
        const int count1 = ix.cnt[eit];
        const uint32_t *lead = ix.pos + ix.off[eit];
        for (int i = 0; i < count1; i++, lead++) {
            const int64_t templong = (int64_t)(*lead) / zv;
            const int64_t u_k = (int64_t)(*lead) % zv;
            BackList *spr = table_get(tb, (int32_t)templong);
            if (spr->score == 0 || spr->seednum < k + 1) {
                const int loc = ++(spr->score);
                if (loc <= kSM) {
                    spr->loczhi[loc - 1] = (int16_t)u_k;
                    spr->seedno[loc - 1] = (int16_t)(k + 1);
                } else {
                    insert_loc(ix, spr, (int)u_k, k + 1, (float)BC, templong, zv);
                }
                int s_k = spr->score;
                if (templong > 0) s_k += table_score(tb, (int32_t)templong - 1);
                if (spr->index == -1) {
                    tb.index_list[j] = (int32_t)templong;
                    tb.index_score[j] = (int16_t)s_k;
                    spr->index = j;
                    j++;
                } else {
                    tb.index_score[spr->index] = (int16_t)s_k;
                }
                spr->score2 = spr->score;
            }
            spr->seednum = (int16_t)(k + 1);
        }
    }
    tb.n_index = j;
    int temp_list[2 * kSM], temp_seedn[2 * kSM], temp_score[2 * kSM];
    for (int i = 0; i < j; i++) {
        if (!(tb.index_score[i] > thresh)) continue;
        const int bid = tb.index_list[i];
        BackList *spr = table_find(tb, bid);
        if (spr->score == 0) continue;
        const int s_k = spr->score;
        int loc = 0;
        int64_t start_loc = (int64_t)bid * zv;
        const BackList *spr1 = nullptr;
        if (bid > 0) {
            spr1 = table_find(tb, bid - 1);
            loc = spr1 ? spr1->score : 0;
            if (loc > 0) start_loc = (int64_t)(bid - 1) * zv;
        }
        int u_k = 0;
        if (loc == 0) {
            for (int q = 0; q < s_k && q < kSM; q++) {
                temp_list[u_k] = spr->loczhi[q];
                temp_seedn[u_k] = spr->seedno[q];
                u_k++;
            }
        } else {
            for (int q = 0; q < loc && q < kSM; q++) {
                temp_list[u_k] = spr1->loczhi[q];
                temp_seedn[u_k] = spr1->seedno[q];
                u_k++;
            }
            for (int q = 0; q < s_k && q < kSM; q++) {
                temp_list[u_k] = spr->loczhi[q] + (int)zv;
                temp_seedn[u_k] = spr->seedno[q];
                u_k++;
            }
        }
        int64_t location_loc[4];
        int repeat_loc = 0;
        if (!find_location3(ix, temp_list, temp_seedn, temp_score, location_loc, u_k, &repeat_loc, (float)BC, rlen, start_loc)) continue;
        if (temp_score[repeat_loc] < 6) continue;
        SeedCand ct;
        ct.score = temp_score[repeat_loc];
        const int loc_seed = temp_seedn[repeat_loc];
        location_loc[0] = start_loc + location_loc[0];
        location_loc[1] = (location_loc[1] - 1) * BC;
        const int64_t loc_list = location_loc[0];
        ct.left1 = location_loc[0] + kSeedLen - 1;
        ct.right1 = ix.ref_len - location_loc[0];
        ct.left2 = location_loc[1] + kSeedLen - 1;
        ct.right2 = rlen - location_loc[1];
        ct.num1 = (int)(ct.left1 >= ct.left2 ? ct.left2 : ct.left1);
        ct.num2 = (int)(ct.right1 >= ct.right2 ? ct.right2 : ct.right1);
        ct.loc1 = location_loc[0];
        ct.loc2 = location_loc[1];
        int seedcount = 0;
        { // consistent seeds in the blocks to the left (:950-961)
            int64_t ub = bid - 2;
            for (int k = ct.num1 / (int)zv; ub >= 0 && k >= 0; k--, ub--) {
                BackList *s1 = table_find(tb, (int32_t)ub);
                if (!s1 || s1->score <= 0) continue;
                const int64_t sl = ub * zv;
                const int scnt = s1->score < kSM ? s1->score : kSM;
                int sk = 0;
                for (int q = 0; q < scnt; q++)
                    if (fabs(dsub_rn(ddiv_rn((double)(loc_list - sl - s1->loczhi[q]), dmul_rn((double)((loc_seed - s1->seedno[q]) * BC), 1.0)), 1.0)) < 0.25) {
                        seedcount++;
                        sk++;
                    }
                if (ddiv_rn(dmul_rn((double)sk, 1.0), (double)scnt) > 0.4) s1->score = 0;
            }
        }
        { // and to the right (:963-973)
            int64_t ub = bid + 1;
            for (int k = ct.num2 / (int)zv; k > 0; k--, ub++) {
                BackList *s1 = table_find(tb, (int32_t)ub);
                if (!s1 || s1->score <= 0) continue;
                const int64_t sl = ub * zv;
                const int scnt = s1->score < kSM ? s1->score : kSM;
                int sk = 0;
                for (int q = 0; q < scnt; q++)
                    if (fabs(dsub_rn(ddiv_rn((double)(sl + s1->loczhi[q] - loc_list), dmul_rn((double)((s1->seedno[q] - loc_seed) * BC), 1.0)), 1.0)) < 0.25) {
                        seedcount++;
                        sk++;
                    }
                if (ddiv_rn(dmul_rn((double)sk, 1.0), (double)scnt) > 0.4) s1->score = 0;
            }
        }
        ct.score += seedcount;
        ct.chain = strand == 0 ? 'F' : 'R';
        // keep the MAXC best, ties after equals (:978-990)
        int low = 0, high = ncand - 1;
        while (low <= high) {
            const int mid = (low + high) / 2;
            if (mid >= ncand || cands[mid].score < ct.score) high = mid - 1;
            else low = mid + 1;
        }
        if (ncand < maxc) {
            for (int q = ncand - 1; q > high; q--) cands[q + 1] = cands[q];
        } else {
            for (int q = ncand - 2; q > high; q--) cands[q + 1] = cands[q];
        }
        if (high + 1 < maxc) cands[high + 1] = ct;
        if (ncand < maxc) ncand++;
    }
}

} // namespace ag2

namespace ag2 {

// scratch bytes a read needs for one strand's block table (hits = that strand's index hit count)
__device__ __forceinline__ uint32_t table_capacity(int64_t hits)
{
    uint32_t cap = 16;
    while ((int64_t)cap < 2 * hits + 2) cap <<= 1;
    return cap;
}
__device__ __forceinline__ int64_t table_bytes(int64_t hits)
{
    const int64_t cap = table_capacity(hits);
    int64_t b = cap * (int64_t)sizeof(BlockSlot) + (hits + 1) * (int64_t)(sizeof(BackList) + sizeof(int32_t) + sizeof(int16_t));
    return (b + 15) & ~(int64_t)15;
}

// lays the table of a strand with `hits` index hits out in scratch (table_bytes(hits) bytes, 16-byte aligned) and empties it
__device__ __forceinline__ void table_init(BlockTable &tb, uint8_t *scratch, int64_t hits, bool with_index)
{
    const uint32_t cap = table_capacity(hits);
    tb.slots = reinterpret_cast<BlockSlot *>(scratch);
    tb.mask = cap - 1;
    tb.pool = reinterpret_cast<BackList *>(scratch + (size_t)cap * sizeof(BlockSlot));
    tb.n_pool = 0;
    int32_t *il = reinterpret_cast<int32_t *>(tb.pool + hits + 1);
    tb.index_list = with_index ? il : nullptr;
    tb.index_score = with_index ? reinterpret_cast<int16_t *>(il + hits + 1) : nullptr;
    tb.n_index = 0;
    const BlockSlot empty = {-1, 0};
    for (uint32_t i = 0; i < cap; ++i) tb.slots[i] = empty;
}

__device__ __forceinline__ int seed_stride(int rlen, int pass)
{
    int BC = pass == 0 ? 5 + rlen / 1000 : 5; // (:786-787, :1055)
    return BC > 20 ? 20 : BC;
}

// reference_mapping's seeding + candidate scan for one read (both strands), pass 0 (1000-bp blocks,
// threshold > 6) or pass 1 (the reference's second pass: stride 5, 2000-bp blocks, threshold > 4).
// scratch must hold max(table_bytes(hits_F), table_bytes(hits_R)) bytes, 16-byte aligned.
__device__ int map_read_candidates(const RefIndex &ix, const uint32_t *reads2, const uint32_t *irr, int64_t roff, int rlen,
                                   int pass, int maxc, uint8_t *scratch, SeedCand *cands)
{
    const int BC = seed_stride(rlen, pass);
    const int64_t zv = pass == 0 ? 1000 : 2000;
    const int thresh = pass == 0 ? 6 : 4;
    int ncand = 0;
    for (int strand = 0; strand < 2; ++strand) {
        const int64_t hits = count_hits(ix, reads2, irr, roff, rlen, strand, BC);
        BlockTable tb;
        table_init(tb, scratch, hits, true);
        seed_and_scan(ix, reads2, irr, roff, rlen, strand, BC, zv, thresh, maxc, tb, cands, ncand);
    }
    return ncand;
}

} // namespace ag2
