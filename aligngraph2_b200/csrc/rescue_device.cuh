// rescue_device.cuh -- device code of the per-read bookkeeping around the extensions (SURVEY.md 8a rows
// A11, A12): rescue_clipped_align (M2R/mecat2ref_aux.cpp:396-538) with find_left/right_clipped_candidate
// (:301-361), fill_clipped_candidate (:272-299), find_location2 (:92-170), and output_results (:541-559).
//
// The reference interleaves "look for a clipped candidate" and "extend it" per alignment; the searches only
// read the alignment itself and the seeding tables, never the outcome of an earlier rescue extension, so one
// thread per read can PLAN all (at most 6) rescue candidates first (plan_read), the extensions run as one more
// batch of the extension kernel, and finish_read replays the reference's linking / final sort / output choice
// in the original order.  The seeding tables of the strand are rebuilt on demand (seeding only, no candidate
// scan: the scan changes `score` but the rescue reads `score2`, loczhi, seedno) -- it is needed for the few
// reads whose best alignment covers < 90 % of the read.
#pragma once

#include "seed_device.cuh"
#include "xdrop_device.cuh"

namespace ag2 {

constexpr int kClipped = 2000;              // CLIPPED, mecat2ref_aux.h:91
constexpr int kMaxAlns = kMaxCand + 6;      // alns[MAXC + 6], results[MAXC + 6] (impl_large.cpp:731-734)
constexpr int kMaxRescue = 6;

struct AlnInfo {                            // AlignInfo, mecat2ref_aux.h:9-20
    int32_t qoff, qend;
    int32_t parent_id, id, prev_id, next_id;
    int32_t valid, qdir;                    // qdir 'F' / 'R'
    int64_t soff, send;
};

struct RescueCand {
    int32_t on;                             // a candidate was found for this slot
    int32_t chain, score;
    int32_t loc2;
    int64_t loc1;
};

struct ReadPlan {
    int32_t naln;                           // alignments after the containment filter
    int32_t naln_ext;                       // successful extensions of the seed candidates (0 -> second pass)
    int32_t nres;                           // results so far (ids 0..nres-1)
    int32_t n_rescue;                       // rescue candidates planned (slots with on != 0)
    AlnInfo alns[kMaxAlns];
    int64_t res_ref[kMaxAlns];              // result id -> index into the record pool
    RescueCand rescue[kMaxRescue];          // slot 2*i + side, i < 3, side 0 = left / 1 = right
};

__device__ __forceinline__ void sort_alns(AlnInfo *a, int n) // std::sort on <= 16 elements: insertion sort, span descending
{
    for (int i = 1; i < n; ++i) {
        const AlnInfo v = a[i];
        const int key = v.qend - v.qoff;
        int j = i;
        while (j > 0 && key > (a[j - 1].qend - a[j - 1].qoff)) {
            a[j] = a[j - 1];
            --j;
        }
        a[j] = v;
    }
}

__device__ __forceinline__ bool aln_contained(const AlnInfo &a, const AlnInfo &b) // AlignInfoContained, aux.h:28-42
{
    const int extra = 100;
    return a.qdir == b.qdir && b.qoff + extra >= a.qoff && b.qend <= a.qend + extra && b.soff + extra >= a.soff &&
           b.send <= a.send + extra;
}
__device__ __forceinline__ bool aln_full(const AlnInfo &a, int qsize) { return (double)(a.qend - a.qoff) >= dmul_rn((double)qsize, 0.9); }

__device__ __forceinline__ int64_t labs64(int64_t v) { return v < 0 ? -v : v; }
__device__ __forceinline__ bool left_clipped(const AlnInfo &a, const AlnInfo &b) // aux.cpp:363-377
{
    if (a.qdir != b.qdir) return false;
    if (abs(b.qend - a.qoff) <= 200 && a.soff - b.send > -200 && a.soff - b.send < 10000) return true;
    if (labs64(b.send - a.soff) <= 200 && a.qoff - b.qend > -200 && a.qoff - b.qend < 10000) return true;
    return false;
}
__device__ __forceinline__ bool right_clipped(const AlnInfo &a, const AlnInfo &b) // aux.cpp:379-393
{
    if (a.qdir != b.qdir) return false;
    if (abs(a.qend - b.qoff) <= 200 && b.soff - a.send > -200 && b.soff - a.send < 10000) return true;
    if (labs64(a.send - b.soff) <= 200 && b.qoff - a.qend > -200 && b.qoff - a.qend < 10000) return true;
    return false;
}

// find_location2 (aux.cpp:92-170): find_location3 without the similarity vote
__device__ int find_location2(const int *t_loc, const int *t_seedn, int *t_score, int64_t *loc, int k, int *rep_loc, float len,
                              int read_len1)
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
    return find_location_choose(t_loc, t_seedn, t_score, loc, k, rep_loc, len, read_len1);
}

// seeding only (impl_large.cpp:842-878): the strand's block table as rescue_clipped_align sees it
__device__ void seed_only(const RefIndex &ix, const uint32_t *reads2, const uint32_t *irr, int64_t roff, int rlen, int strand, int BC,
                          int64_t zv, BlockTable &tb)
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
                if (spr->index == -1) spr->index = j++;
                spr->score2 = spr->score;
            }
            spr->seednum = (int16_t)(k + 1);
        }
    }
    tb.n_index = j;
}

__device__ __forceinline__ int table_score2(const BlockTable &t, int64_t key)
{
    const BackList *b = key >= 0 ? table_find(t, (int32_t)key) : nullptr;
    return b ? b->score2 : 0;
}

__device__ bool fill_clipped(const BackList *block, int64_t bid, RescueCand &can, int chain, int read_size, int BC, int block_size)
{
    int seedn[kSM], boff[kSM], score[kSM], rep_loc = 0;
    int64_t locations[4];
    const int n = block->score2 < kSM ? block->score2 : kSM;
    for (int i = 0; i < n; ++i) {
        seedn[i] = block->seedno[i];
        boff[i] = block->loczhi[i];
        score[i] = 0;
    }
    if (find_location2(boff, seedn, score, locations, n, &rep_loc, (float)BC, read_size)) {
        can.on = 1;
        can.score = score[rep_loc];
        can.chain = chain;
        can.loc1 = bid * block_size + locations[0];
        can.loc2 = (int32_t)((locations[1] - 1) * BC);
        return true;
    }
    return false;
}

__device__ bool find_left_clipped(const AlnInfo &aln, RescueCand &can, const BlockTable &tb, int block_size, int read_size, int BC)
{
    if (aln.qoff <= kClipped || aln.soff <= kClipped) return false;
    const int n1 = aln.qoff / block_size;
    int64_t n2 = aln.soff / block_size;
    int64_t n = n1 < n2 ? n1 : n2;
    int max_score = 0;
    int64_t bid = -1;
    for (--n2; n >= 0 && n2 >= 0; --n, --n2) {
        const int s2 = table_score2(tb, n2);
        if (s2 > max_score) {
            max_score = s2;
            bid = n2;
        }
    }
    if (bid >= 0) {
        const BackList *block = table_find(tb, (int32_t)bid);
        if (block->score2 > 4) return fill_clipped(block, bid, can, aln.qdir, read_size, BC, block_size);
    }
    return false;
}

__device__ bool find_right_clipped(const AlnInfo &aln, RescueCand &can, const BlockTable &tb, int block_size, int read_size,
                                   int64_t ref_size, int BC)
{
    if (read_size - aln.qend <= kClipped || ref_size - aln.send <= kClipped) return false;
    const int n1 = (read_size - aln.qend) / block_size;
    const int64_t n2 = (ref_size - aln.send) / block_size;
    int64_t n = n1 < n2 ? n1 : n2;
    int max_score = 0;
    int64_t bid = -1;
    int64_t k = aln.send / block_size + 1;
    for (; n >= 0; --n, ++k) {
        const int s2 = table_score2(tb, k);
        if (s2 > max_score) {
            max_score = s2;
            bid = k;
        }
    }
    if (bid >= 0) {
        const BackList *block = table_find(tb, (int32_t)bid);
        if (block->score2 > 4) return fill_clipped(block, bid, can, aln.qdir, read_size, BC, block_size);
    }
    return false;
}

// First half of rescue_clipped_align (:416-445): the read's alignments sorted by span, contained ones dropped.  recs: the
// records of this read's seed candidates in canidate_loc[] order, rec_base their index in the record pool.  Returns true
// if the searches for clipped candidates (:446-520) are due: they need the strand's seeding table.
__device__ bool plan_alns(const Record *recs, int ncand, int64_t rec_base, int rlen, ReadPlan &P)
{
    P.naln = P.naln_ext = P.nres = P.n_rescue = 0;
    for (int s = 0; s < kMaxRescue; ++s) P.rescue[s].on = 0;
    for (int c = 0; c < ncand; ++c) {
        if (!recs[c].ok) continue;
        AlnInfo &ai = P.alns[P.naln++];
        ai.qoff = recs[c].qb;
        ai.qend = recs[c].qe;
        ai.qdir = recs[c].strand ? 'R' : 'F';
        ai.soff = recs[c].sb;
        ai.send = recs[c].se;
        ai.valid = 1;
        ai.id = P.nres;
        ai.prev_id = ai.next_id = ai.parent_id = -1;
        P.res_ref[P.nres++] = rec_base + c;
    }
    P.naln_ext = P.naln;
    if (P.naln == 0) return false;
    AlnInfo *alnv = P.alns;
    int naln = P.naln;
    sort_alns(alnv, naln);
    for (int i = 0; i < naln - 1; ++i) {
        if (!alnv[i].valid) continue;
        for (int j = i + 1; j < naln; ++j) {
            if (!alnv[j].valid) continue;
            if (aln_contained(alnv[i], alnv[j])) alnv[j].valid = 0;
        }
    }
    int k = 0;
    for (int i = 0; i < naln; ++i)
        if (alnv[i].valid) alnv[k++] = alnv[i];
    naln = k;
    P.naln = naln;
    return !aln_full(alnv[0], rlen);
    // (:430-444) links clipped pairs only when prev_id / next_id != -1, which is never the case here: no effect
}

// The searches of rescue_clipped_align (:446-520) on the one-thread-per-read path: the strand's table is rebuilt in
// `scratch` (global memory).  P as plan_alns left it.
__device__ void plan_search(const RefIndex &ix, const uint32_t *reads2, const uint32_t *irr, int64_t roff, int rlen, int pass,
                            uint8_t *scratch, ReadPlan &P)
{
    AlnInfo *alnv = P.alns;
    const int n = P.naln < 3 ? P.naln : 3;
    const int BC = seed_stride(rlen, pass);
    const int64_t zv = pass == 0 ? 1000 : 2000;
    int have_strand = -1;
    BlockTable tb;
    P.n_rescue = 0;
    for (int s = 0; s < kMaxRescue; ++s) P.rescue[s].on = 0;
    for (int i = 0; i < n; ++i) {
        const int strand = alnv[i].qdir == 'F' ? 0 : 1;
        if (have_strand != strand) {
            const int64_t hits = count_hits(ix, reads2, irr, roff, rlen, strand, BC);
            table_init(tb, scratch, hits, false);
            seed_only(ix, reads2, irr, roff, rlen, strand, BC, zv, tb);
            have_strand = strand;
        }
        if (find_left_clipped(alnv[i], P.rescue[2 * i], tb, (int)zv, rlen, BC)) ++P.n_rescue;
        if (find_right_clipped(alnv[i], P.rescue[2 * i + 1], tb, (int)zv, rlen, ix.ref_len, BC)) ++P.n_rescue;
    }
}

__device__ void plan_read(const RefIndex &ix, const uint32_t *reads2, const uint32_t *irr, int64_t roff, int rlen, int pass,
                          const Record *recs, int ncand, int64_t rec_base, uint8_t *scratch, ReadPlan &P)
{
    if (plan_alns(recs, ncand, rec_base, rlen, P)) plan_search(ix, reads2, irr, roff, rlen, pass, scratch, P);
}

// Second half of rescue_clipped_align (:446-537) + output_results (:541-559).  rrec[s]: the record of rescue slot
// s (only read where P.rescue[s].on), rrec_ref[s] its index in the record pool.  out_refs receives the record
// pool indices to write, in order; returns how many.
__device__ int finish_read(ReadPlan &P, const Record *rrec, const int64_t *rrec_ref, int rlen, int num_output, int64_t *out_refs)
{
    AlnInfo *alnv = P.alns;
    int naln = P.naln;
    if (naln > 0 && P.n_rescue > 0) {
        int k = 0;
        const int n = naln < 3 ? naln : 3;
        for (int i = 0; i < n; ++i)
            for (int side = 0; side < 2; ++side) {
                const int s = 2 * i + side;
                if (!P.rescue[s].on || !rrec[s].ok) continue;
                const int id = P.nres++;
                P.res_ref[id] = rrec_ref[s];
                AlnInfo &ai = alnv[naln + k];
                ai.qoff = rrec[s].qb;
                ai.qend = rrec[s].qe;
                ai.qdir = rrec[s].strand ? 'R' : 'F';
                ai.soff = rrec[s].sb;
                ai.send = rrec[s].se;
                ai.valid = 1;
                ai.id = id;
                ai.prev_id = ai.next_id = ai.parent_id = -1;
                if (side == 0 ? left_clipped(alnv[i], ai) : right_clipped(alnv[i], ai)) {
                    ai.parent_id = alnv[i].id;
                    if (side == 0) alnv[i].prev_id = ai.id;
                    else alnv[i].next_id = ai.id;
                    ++k;
                }
            }
        if (k) {
            naln += k;
            sort_alns(alnv, naln);
            k = 0;
            for (int i = 0; i < naln; ++i)
                if (aln_full(alnv[i], rlen)) {
                    alnv[i].parent_id = alnv[i].prev_id = alnv[i].next_id = -1;
                    ++k;
                }
            if (k) naln = k;
        }
        P.naln = naln;
    }
    int n = 0, nout = 0;
    for (int i = 0; i < naln && n < num_output; ++i) {
        if (alnv[i].parent_id != -1) continue;
        out_refs[nout++] = P.res_ref[alnv[i].id];
        if (alnv[i].prev_id != -1) out_refs[nout++] = P.res_ref[alnv[i].prev_id];
        if (alnv[i].next_id != -1) out_refs[nout++] = P.res_ref[alnv[i].next_id];
        ++n;
    }
    return nout;
}

} // namespace ag2
