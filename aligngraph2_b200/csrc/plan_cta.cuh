// plan_cta.cuh -- the searches of rescue_clipped_align (M2R/mecat2ref_aux.cpp:446-520: find_left/right_clipped_candidate
// :301-361, fill_clipped_candidate :272-299) for the few reads whose best alignment covers < 90 % of the read, one CTA
// per read: the strand's seeding table is rebuilt as sorted runs in shared memory (seed_cta_build: no candidate scan, so
// a run's score is the reference's `score2`), and thread 0 looks the blocks beyond the clip up in it.
#pragma once

#include "rescue_device.cuh"
#include "seed_cta.cuh"

namespace ag2 {

struct PlanCtaArgs {
    RefIndex ix;
    const uint32_t *reads2, *irr;
    const int64_t *read_off;
    const int32_t *read_len;
    const int32_t *reads;          // pass-1 list of reads (item k -> read), or null
    const int32_t *work;           // items to do
    const unsigned *n_work_dev;
    int pass, cap;
    unsigned *next;
    ReadPlan *plans;               // [item]
    int64_t *n_rescue;             // [item]
    int32_t *ovf;
    unsigned *ovf_count;
    uint32_t *heavy_pool;
};

// run that holds block b, or -1
__device__ __forceinline__ int run_find(const RunView &v, int nruns, int64_t b)
{
    if (b < 0) return -1;
    const int r = run_lower_bound(v, nruns, b);
    return r < nruns && (int64_t)ev_block(v.ev[v.run_start[r]]) == b ? r : -1;
}

__device__ bool cta_fill_clipped(const RunView &rv, const int16_t *run_score, int r, int64_t bid, RescueCand &can, int chain, int read_size,
                                 int BC, int block_size)
{
    int seedn[kSM], boff[kSM], score[kSM], rep_loc = 0;
    int64_t locations[4];
    const int n = run_score[r] < kSM ? run_score[r] : kSM;
    for (int i = 0; i < n; ++i) {
        run_entry(rv, r, i, boff[i], seedn[i]);
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

__device__ bool cta_find_left_clipped(const AlnInfo &aln, RescueCand &can, const RunView &rv, const int16_t *run_score, int nruns,
                                      int block_size, int read_size, int BC)
{
    if (aln.qoff <= kClipped || aln.soff <= kClipped) return false;
    const int n1 = aln.qoff / block_size;
    int64_t n2 = aln.soff / block_size;
    int64_t n = n1 < n2 ? n1 : n2;
    int max_score = 0, best = -1;
    int64_t bid = -1;
    for (--n2; n >= 0 && n2 >= 0; --n, --n2) {
        const int r = run_find(rv, nruns, n2);
        const int s2 = r >= 0 ? run_score[r] : 0;
        if (s2 > max_score) {
            max_score = s2;
            bid = n2;
            best = r;
        }
    }
    if (bid >= 0 && run_score[best] > 4) return cta_fill_clipped(rv, run_score, best, bid, can, aln.qdir, read_size, BC, block_size);
    return false;
}

__device__ bool cta_find_right_clipped(const AlnInfo &aln, RescueCand &can, const RunView &rv, const int16_t *run_score, int nruns,
                                       int block_size, int read_size, int64_t ref_size, int BC)
{
    if (read_size - aln.qend <= kClipped || ref_size - aln.send <= kClipped) return false;
    const int n1 = (read_size - aln.qend) / block_size;
    const int64_t n2 = (ref_size - aln.send) / block_size;
    int64_t n = n1 < n2 ? n1 : n2;
    int max_score = 0, best = -1;
    int64_t bid = -1;
    int64_t k = aln.send / block_size + 1;
    for (; n >= 0; --n, ++k) {
        const int r = run_find(rv, nruns, k);
        const int s2 = r >= 0 ? run_score[r] : 0;
        if (s2 > max_score) {
            max_score = s2;
            bid = k;
            best = r;
        }
    }
    if (bid >= 0 && run_score[best] > 4) return cta_fill_clipped(rv, run_score, best, bid, can, aln.qdir, read_size, BC, block_size);
    return false;
}

// Body of plan_cta_kernel.  plans[k] is as plan_alns left it; fills its rescue slots.
__device__ void plan_cta_body(const PlanCtaArgs &a, uint8_t *smem)
{
    const int tid = threadIdx.x;
    const unsigned n_work = *a.n_work_dev;
    const int64_t zv = a.pass == 0 ? 1000 : 2000;
    const int block_bits = seed_block_bits(a.ix.ref_len, zv);
    uint32_t *pool = a.heavy_pool + (size_t)blockIdx.x * (a.cap / (kSM + 1) + 1) * kHeavyWords;
    {
        SeedCtaSmem sm = seed_cta_carve(smem, a.cap);
        if (tid == 0) {
            stage_barrier_init(sm.bar);
            sm.misc[5] = 0;           // phase of the staging barrier
        }
        __syncthreads();
    }
    for (;;) {
        SeedCtaSmem sm0 = seed_cta_carve(smem, a.cap);
        if (tid == 0) sm0.misc[3] = (int)atomicAdd(a.next, 1u);
        __syncthreads();
        const unsigned w = (unsigned)sm0.misc[3];
        if (w >= n_work) break;
        const int64_t k = a.work[w];
        const int64_t r = a.reads ? a.reads[k] : k;
        const int rlen = a.read_len[r];
        const int64_t roff = a.read_off[r];
        const int BC = seed_stride(rlen, a.pass);
        ReadPlan &P = a.plans[k];
        if (tid == 0) {
            P.n_rescue = 0;
            for (int s = 0; s < kMaxRescue; ++s) P.rescue[s].on = 0;
        }
        __syncthreads();
        const int n = P.naln < 3 ? P.naln : 3;
        int have_strand = -1, nruns = 0, found = 0;
        bool ok = true;
        SeedCtaSmem sm = sm0;
        for (int i = 0; i < n && ok; ++i) {
            const int strand = P.alns[i].qdir == 'F' ? 0 : 1;
            if (have_strand != strand) {
                __syncthreads();
                sm = seed_cta_carve(smem, a.cap);
                nruns = seed_cta_build(a.ix, a.reads2, a.irr, roff, rlen, strand, BC, zv, a.cap, block_bits, sm, pool);
                if (nruns < 0) {
                    ok = false;
                    break;
                }
                have_strand = strand;
            }
            if (tid == 0) {
                const RunView rv = {sm.ev, sm.run_start, pool};
                const AlnInfo aln = P.alns[i];
                RescueCand c0 = P.rescue[2 * i], c1 = P.rescue[2 * i + 1];
                if (cta_find_left_clipped(aln, c0, rv, sm.run_score, nruns, (int)zv, rlen, BC)) ++found;
                if (cta_find_right_clipped(aln, c1, rv, sm.run_score, nruns, (int)zv, rlen, a.ix.ref_len, BC)) ++found;
                P.rescue[2 * i] = c0;
                P.rescue[2 * i + 1] = c1;
            }
        }
        if (tid == 0) {
            if (ok) {
                P.n_rescue = found;
                a.n_rescue[k] = found;
            } else {
                P.n_rescue = 0;
                for (int s = 0; s < kMaxRescue; ++s) P.rescue[s].on = 0;
                a.n_rescue[k] = 0;
                a.ovf[atomicAdd(a.ovf_count, 1u)] = (int32_t)k;
            }
        }
        __syncthreads();
    }
}

} // namespace ag2
