// map_kernels.cuh -- launch wrappers of the per-read bookkeeping (rescue_device.cuh): one thread per read.
// Used by ag2_map_reads, the replacement of reference_mapping()'s loop body (impl_large.cpp:776-1316) for a
// whole read batch:  seed -> extend -> plan rescue -> extend -> link + choose output -> (second pass for the
// reads without any alignment) -> records in thread-file order.
#pragma once

#include "plan_cta.cuh"

namespace ag2 {

constexpr int kOutCap = 3 * kMaxAlns; // output_results writes at most 3 records per kept alignment

// reads[k] (or k itself when reads == nullptr) is the read index of work item k
__device__ __forceinline__ int64_t item_read(const int32_t *reads, int64_t k) { return reads ? reads[k] : k; }

// ---- seeding + candidate scoring: one CTA per read (seed_cta.cuh) ----
__global__ void AG2_SEED_BOUNDS seed_cta_kernel(SeedCtaArgs a)
{
    extern __shared__ __align__(16) uint8_t seed_smem[];
    seed_cta_body(a, seed_smem);
}

// ---- the same on the one-thread-per-read path (seed_device.cuh), for the items the CTA path reported as overflow ----
// work[w] is the item of work entry w; need / need_prefix are indexed by w, the results by item
__global__ void seed_need_sub_kernel(RefIndex ix, const uint32_t *reads2, const uint32_t *irr, const int64_t *read_off,
                                     const int32_t *read_len, const int32_t *reads, const int32_t *work, int64_t n, int pass, int64_t *need)
{
    for (int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; w < n; w += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = item_read(reads, item_read(work, w));
        const int rlen = read_len[r];
        const int BC = seed_stride(rlen, pass);
        const int64_t a = table_bytes(count_hits(ix, reads2, irr, read_off[r], rlen, 0, BC));
        const int64_t b = table_bytes(count_hits(ix, reads2, irr, read_off[r], rlen, 1, BC));
        need[w] = a > b ? a : b;
    }
}

__global__ void seed_map_sub_kernel(RefIndex ix, const uint32_t *reads2, const uint32_t *irr, const int64_t *read_off,
                                    const int32_t *read_len, const int32_t *reads, const int32_t *work, int64_t first, int64_t n, int pass,
                                    int maxc, const int64_t *need_prefix, uint8_t *scratch, SeedCand *cands, int32_t *ncand)
{
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
        const int64_t w = first + q;
        const int64_t k = item_read(work, w);
        const int64_t r = item_read(reads, k);
        SeedCand local[kMaxCand + 1];
        const int nc = map_read_candidates(ix, reads2, irr, read_off[r], read_len[r], pass, maxc,
                                           scratch + (need_prefix[w] - need_prefix[first]), local);
        ncand[k] = nc;
        for (int i = 0; i < nc; ++i) cands[k * maxc + i] = local[i];
    }
}

__global__ void seeds_to_candidates_sub_kernel(const SeedCand *cands, const int32_t *ncand, const int64_t *prefix, const int32_t *reads,
                                               int64_t n, int maxc, Candidate *out)
{
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        const int64_t base = prefix[k];
        for (int i = 0; i < ncand[k]; ++i) {
            const SeedCand &c = cands[k * maxc + i];
            Candidate o;
            o.read = (int32_t)item_read(reads, k);
            o.strand = c.chain == 'F' ? 0 : 1;
            o.loc1 = c.loc1;
            o.loc2 = (int32_t)c.loc2;
            o.score = c.score;
            out[base + i] = o;
        }
    }
}

// rescue planning, first half (plan_alns): one thread per item; the items that need the seeding table go to `list`
__global__ void plan_alns_kernel(const int32_t *read_len, const int32_t *reads, int64_t n, const int32_t *ncand, const int64_t *cand_prefix,
                                 const Record *pool, int64_t pool_base, ReadPlan *plans, int64_t *n_rescue, int32_t *list, unsigned *list_count)
{
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = item_read(reads, k);
        const int64_t rec_base = pool_base + cand_prefix[k];
        n_rescue[k] = 0;
        if (plan_alns(pool + rec_base, ncand[k], rec_base, read_len[r], plans[k])) list[atomicAdd(list_count, 1u)] = (int32_t)k;
    }
}

// second half, one CTA per listed item (plan_cta.cuh)
__global__ void __launch_bounds__(kSeedCtaThreads) plan_cta_kernel(PlanCtaArgs a)
{
    extern __shared__ __align__(16) uint8_t seed_smem[];
    plan_cta_body(a, seed_smem);
}

// second half on the one-thread-per-read path, for the items the CTA path reported as overflow
__global__ void plan_search_sub_kernel(RefIndex ix, const uint32_t *reads2, const uint32_t *irr, const int64_t *read_off, const int32_t *read_len,
                                       const int32_t *reads, const int32_t *work, int64_t first, int64_t n, int pass,
                                       const int64_t *need_prefix, uint8_t *scratch, ReadPlan *plans, int64_t *n_rescue)
{
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
        const int64_t w = first + q;
        const int64_t k = work[w];
        const int64_t r = item_read(reads, k);
        plan_search(ix, reads2, irr, read_off[r], read_len[r], pass, scratch + (need_prefix[w] - need_prefix[first]), plans[k]);
        n_rescue[k] = plans[k].n_rescue;
    }
}

__global__ void rescue_gather_kernel(const ReadPlan *plans, const int64_t *rescue_prefix, const int32_t *reads, int64_t n, Candidate *out)
{
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        int64_t o = rescue_prefix[k];
        for (int s = 0; s < kMaxRescue; ++s) {
            const RescueCand &c = plans[k].rescue[s];
            if (!c.on) continue;
            Candidate cd;
            cd.read = (int32_t)item_read(reads, k);
            cd.strand = c.chain == 'F' ? 0 : 1;
            cd.loc1 = c.loc1;
            cd.loc2 = c.loc2;
            cd.score = c.score;
            out[o++] = cd;
        }
    }
}

__global__ void finish_kernel(ReadPlan *plans, const int64_t *rescue_prefix, const int32_t *reads, int64_t n, const int32_t *read_len,
                              const Record *pool, int64_t rescue_pool_base, int num_output, int64_t *out_refs, int32_t *nout,
                              int32_t *need_pass2)
{
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = item_read(reads, k);
        Record rrec[kMaxRescue];
        int64_t rref[kMaxRescue];
        int64_t o = rescue_pool_base + rescue_prefix[k];
        for (int s = 0; s < kMaxRescue; ++s) {
            rrec[s].ok = 0;
            rref[s] = -1;
            if (!plans[k].rescue[s].on) continue;
            rrec[s] = pool[o];
            rref[s] = o;
            ++o;
        }
        int64_t refs[kOutCap];
        const int no = finish_read(plans[k], rrec, rref, read_len[r], num_output, refs);
        nout[r] = no;
        for (int i = 0; i < no; ++i) out_refs[r * kOutCap + i] = refs[i];
        if (need_pass2) need_pass2[r] = plans[k].naln_ext == 0 ? 1 : 0;
    }
}

// stream compaction helpers (flags -> list)
__global__ void flags_to_i64_kernel(const int32_t *flags, int64_t n, int64_t *out)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = flags[i];
}
__global__ void compact_reads_kernel(const int32_t *flags, const int64_t *prefix, int64_t n, int32_t *list)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if (flags[i]) list[prefix[i]] = (int32_t)i;
}

// records in thread-file order: read order, output_results order inside a read
__global__ void gather_output_kernel(const int64_t *out_refs, const int32_t *nout, const int64_t *out_prefix, int64_t n_reads,
                                     const Record *pool, Record *out)
{
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n_reads; r += (int64_t)gridDim.x * blockDim.x)
        for (int i = 0; i < nout[r]; ++i) out[out_prefix[r] + i] = pool[out_refs[r * kOutCap + i]];
}

} // namespace ag2
