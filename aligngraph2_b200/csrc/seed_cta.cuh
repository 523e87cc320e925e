// seed_cta.cuh -- seeding and candidate scoring of mecat2ref+, one CTA per read.
//
// What is computed is the reference's seeding loop (M2R/mecat2ref_impl_large.cpp:842-878), candidate scan (:882-991),
// find_location3 (:609-693) and insert_loc (:123-170); the form is not the reference's.  The reference walks the index
// hits of a read strand one by one and updates a dense per-worker table of 1000-bp blocks (`Back_List`, 46 MB per strand
// at 250 Mb).  Here a strand is a STREAM OF EVENTS that is sorted:
//
//   1. every thread takes seeds, looks up their buckets and the hits are expanded 128 at a time; a hit is an event
//      (block, seed number, offset in block) if it is the first hit of its seed in its block -- which is all the
//      reference's `score == 0 || seednum < k + 1` test says, because a bucket is ascending (:853);
//   2. the events (64-bit words in shared memory) are sorted by block with a stable LSD radix sort (warp `match.any`
//      ranking, one histogram per warp); a run of equal blocks IS that block's Back_List: its first 20 events are
//      loczhi[] / seedno[], its length is `score` -- in the order the reference appended them;
//   3. a block with more than 20 events ("heavy": low-error reads, the stride-5 second pass) is replayed sequentially by
//      one thread through insert_loc, its surviving 20 entries go to a small pool in global memory and the score after
//      every event is kept (the neighbour's `index_score` needs the score it had AT THAT TIME);
//   4. `index_score` (:863-872) of block b = its final score + the score of block b - 1 when b was last updated = the
//      number of events of run(b - 1) whose seed is not later than the last seed of run(b): a count inside the
//      adjacent run, no table;
//   5. blocks above the threshold are visited in first-touch order (`index_list`): their first event's serial number
//      orders them; warp 0 runs the candidate scan over them -- the O(k^2) consistency votes of find_location3 spread
//      over the lanes, the neighbour blocks harvested one per lane -- and keeps the top-MAXC list exactly as :978-990.
//
// A read strand that does not fit the CTA's shared memory (more events than `cap`) is reported in an overflow list; the
// host reruns those reads with a larger `cap` and, beyond that, on the one-thread-per-read path (seed_device.cuh), which
// holds its table in global memory.
//
// Compiled by nvcc into libag2_b200.so and, unchanged with -DAG2_EMU (one warp per CTA), by g++ for the CPU-side tests.
#pragma once

#include "seed_device.cuh"

namespace ag2 {

#ifdef AG2_EMU
constexpr int kSeedCtaThreads = 32;
#elif defined(AG2_SEED_THREADS)
constexpr int kSeedCtaThreads = AG2_SEED_THREADS;       // experiments/runs: occupancy sweeps
#else
constexpr int kSeedCtaThreads = 256;
#endif
#ifdef AG2_SEED_MIN_CTAS
#define AG2_SEED_BOUNDS __launch_bounds__(kSeedCtaThreads, AG2_SEED_MIN_CTAS)
#else
#define AG2_SEED_BOUNDS __launch_bounds__(kSeedCtaThreads)   // a stated minimum of 1 CTA let ptxas take 118 registers: 2 CTAs per SM
#endif
constexpr int kSeedCtaWarps = kSeedCtaThreads / 32;
constexpr int kSeedDigitBits = 9;                       // radix of the sort: 512 bins per warp -- two passes up to 262 144 blocks (250 Mb)
constexpr int kSeedBins = 1 << kSeedDigitBits;
constexpr int kNullBlock = 0x7fffff;                    // block field of a hit that is not an event (not the first of its seed in its block)
constexpr int kSeedMaxCap = 32768;                      // the serial number of an event has 15 bits
constexpr int kHeavyWords = kSM;                        // pool entry of a heavy block: its 20 surviving (seed, offset) pairs

// event word: [63:41] block  [40:26] seed number k + 1  [25:15] offset in block  [14:0] serial number e
// (for a heavy run the last field is reused: event 1 holds the run's pool slot, events >= 20 the score after them)
__device__ __forceinline__ uint64_t ev_pack(uint32_t block, int seedno, int u, int e)
{
    return ((uint64_t)block << 41) | ((uint64_t)seedno << 26) | ((uint64_t)u << 15) | (uint64_t)e;
}
__device__ __forceinline__ int ev_block(uint64_t x) { return (int)(x >> 41); }
__device__ __forceinline__ int ev_seed(uint64_t x) { return (int)((x >> 26) & 0x7fffu); }
__device__ __forceinline__ int ev_u(uint64_t x) { return (int)((x >> 15) & 0x7ffu); }
__device__ __forceinline__ int ev_e(uint64_t x) { return (int)(x & 0x7fffu); }
__device__ __forceinline__ uint64_t ev_with_e(uint64_t x, int e) { return (x & ~(uint64_t)0x7fffu) | (uint64_t)e; }

struct SeedCtaArgs {
    RefIndex ix;
    const uint32_t *reads2, *irr;
    const int64_t *read_off;
    const int32_t *read_len;
    const int32_t *reads;          // pass-1 list of reads (item k -> read), or null: item k is read k
    const int32_t *work;           // list of items to do, or null: all of [0, n_work)
    const unsigned *n_work_dev;    // its length on the device (null: n_work)
    unsigned n_work;
    int pass, maxc, cap;
    unsigned *next;                // work counter
    SeedCand *cands;               // [item][maxc]
    int32_t *ncand;                // [item]
    int32_t *ovf;                  // items whose strand did not fit `cap`
    unsigned *ovf_count;
    uint32_t *heavy_pool;          // [CTA][cap / 21 + 1][kHeavyWords]
};

// shared memory of a CTA, carved from one dynamic allocation
struct SeedCtaSmem {
    uint64_t *ev;                  // events; sorted by block after sort_events
    uint64_t *aux;                 // the sort's second buffer; afterwards: run_start, run_score, qin, qout
    uint16_t *run_start;           // [nruns + 1]
    int16_t *run_score;            // [nruns]  the block's `score` (the scan zeroes it)
    uint32_t *qin, *qout;          // blocks above the threshold: (first serial << 15) | run
    uint16_t *hist;                // [512][warps]
    uint32_t *wsum;                // block_scan scratch
    int *misc;
    int *tl_loc, *tl_seed, *tl_score;
    SeedCand *cands;
    uint64_t *bar;                 // mbarrier of the read staging (TMA bulk copy)
};

__host__ __device__ __forceinline__ size_t seed_cta_smem_bytes(int cap)
{
    return (size_t)cap * 16 + kSeedBins * kSeedCtaWarps * 2 + 40 * 4 + 32 * 4 + 3 * 2 * kSM * 4 + (kMaxCand + 1) * sizeof(SeedCand) + 16 + 64;
}

__device__ __forceinline__ SeedCtaSmem seed_cta_carve(uint8_t *p, int cap)
{
    SeedCtaSmem s;
    s.ev = reinterpret_cast<uint64_t *>(p);
    s.aux = s.ev + cap;
    uint8_t *q = reinterpret_cast<uint8_t *>(s.aux + cap);
    s.cands = reinterpret_cast<SeedCand *>(q);
    q += (kMaxCand + 1) * sizeof(SeedCand);
    s.bar = reinterpret_cast<uint64_t *>(q);
    q += 16;
    s.hist = reinterpret_cast<uint16_t *>(q);
    q += kSeedBins * kSeedCtaWarps * 2;
    s.wsum = reinterpret_cast<uint32_t *>(q);
    q += 40 * 4;
    s.misc = reinterpret_cast<int *>(q);
    q += 32 * 4;
    s.tl_loc = reinterpret_cast<int *>(q);
    s.tl_seed = s.tl_loc + 2 * kSM;
    s.tl_score = s.tl_seed + 2 * kSM;
    s.run_start = nullptr;
    s.run_score = nullptr;
    s.qin = s.qout = nullptr;
    return s;
}

// the aliases inside the buffer the sort left free
__device__ __forceinline__ void seed_cta_alias(SeedCtaSmem &s, uint64_t *free_buf, int cap)
{
    uint8_t *q = reinterpret_cast<uint8_t *>(free_buf);
    s.run_start = reinterpret_cast<uint16_t *>(q);            // cap + 1 entries fit 2 * cap + 2 bytes; qin starts at 2 * cap + 4
    s.run_score = reinterpret_cast<int16_t *>(q + 2 * (size_t)cap + 4);
    s.qin = reinterpret_cast<uint32_t *>(q + 4 * (size_t)cap + 8);
    s.qout = s.qin + (cap / 2 - 4);
}
__device__ __forceinline__ int seed_cta_qcap(int cap) { return cap / 2 - 4; }

// exclusive prefix of v over the threads of the CTA (thread order) and the total; two barriers
__device__ __forceinline__ int block_scan_excl(int v, int &total, uint32_t *wsum)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) wsum[warp] = (uint32_t)inc;
    __syncthreads();
    int base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kSeedCtaWarps; ++w) {
        const int s = (int)wsum[w];
        if (w < warp) base += s;
        tot += s;
    }
    __syncthreads();
    total = tot;
    return base + inc - v;
}

#ifdef AG2_EMU
inline unsigned __match_any_sync(unsigned, unsigned v)
{
    const uint64_t *buf = warp_emu::exchange(v);
    unsigned r = 0;
    for (int l = 0; l < 32; ++l) r |= (unsigned)(buf[l] == (uint64_t)v) << l;
    return r;
}
inline unsigned __brev(unsigned v)
{
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) r |= ((v >> i) & 1u) << (31 - i);
    return r;
}
#endif

// Stable LSD radix sort of src[0..n) by bits [lo, hi) of the 64-bit words; returns the buffer that holds the result.
// Every warp owns a contiguous part of the array and a histogram of its own; inside a batch of 32 the rank of an element
// among equals is its lane order (match.any), so equal keys keep their order.  Two batches are in flight per step (their
// loads and match.any do not depend on each other; only the histogram updates are ordered).
__device__ uint64_t *sort_events(uint64_t *src, uint64_t *dst, int n, int lo, int hi, uint16_t *hist, uint32_t *wsum)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = ((n + kSeedCtaWarps - 1) / kSeedCtaWarps + 63) & ~63;
    const int seg_lo = min(n, warp * per), seg_hi = min(n, seg_lo + per);
    const unsigned lt = (1u << lane) - 1u;
    const int passes = (hi - lo + kSeedDigitBits - 1) / kSeedDigitBits;
    const int width = (hi - lo + passes - 1) / passes;           // digits of equal width: fewer bins to scan
    for (int shift = lo; shift < hi; shift += width) {
        const int bits = min(width, hi - shift);
        const unsigned dmask = (1u << bits) - 1u;
        const int nbins = 1 << bits;
        for (int i = tid; i < nbins * kSeedCtaWarps; i += kSeedCtaThreads) hist[i] = 0;
        __syncthreads();
        for (int b = seg_lo; b < seg_hi; b += 64) {
            const int i0 = b + lane, i1 = b + 32 + lane;
            const bool v0 = i0 < seg_hi, v1 = i1 < seg_hi;
            const unsigned d0 = v0 ? (unsigned)(src[i0] >> shift) & dmask : (unsigned)kSeedBins;
            const unsigned d1 = v1 ? (unsigned)(src[i1] >> shift) & dmask : (unsigned)kSeedBins;
            const unsigned p0 = __match_any_sync(0xffffffffu, d0), p1 = __match_any_sync(0xffffffffu, d1);
            if (v0 && (p0 & lt) == 0) hist[d0 * kSeedCtaWarps + warp] += (uint16_t)__popc(p0);
            __syncwarp();
            if (v1 && (p1 & lt) == 0) hist[d1 * kSeedCtaWarps + warp] += (uint16_t)__popc(p1);
            __syncwarp();
        }
        __syncthreads();
        {   // exclusive scan in (digit, warp) order
            const int per_t = (nbins * kSeedCtaWarps + kSeedCtaThreads - 1) / kSeedCtaThreads;   // <= 16
            const int e_lo = min(nbins * kSeedCtaWarps, tid * per_t), e_hi = min(nbins * kSeedCtaWarps, e_lo + per_t);
            unsigned sum = 0;
            for (int q = e_lo; q < e_hi; ++q) sum += hist[q];
            int total;
            unsigned run = (unsigned)block_scan_excl((int)sum, total, wsum);
            for (int q = e_lo; q < e_hi; ++q) {
                const unsigned c = hist[q];
                hist[q] = (uint16_t)run;
                run += c;
            }
        }
        __syncthreads();
        for (int b = seg_lo; b < seg_hi; b += 64) {
            const int i0 = b + lane, i1 = b + 32 + lane;
            const bool v0 = i0 < seg_hi, v1 = i1 < seg_hi;
            const uint64_t x0 = v0 ? src[i0] : 0, x1 = v1 ? src[i1] : 0;
            const unsigned d0 = v0 ? (unsigned)(x0 >> shift) & dmask : (unsigned)kSeedBins;
            const unsigned d1 = v1 ? (unsigned)(x1 >> shift) & dmask : (unsigned)kSeedBins;
            const unsigned p0 = __match_any_sync(0xffffffffu, d0), p1 = __match_any_sync(0xffffffffu, d1);
            unsigned off = 0;
            if (v0) off = hist[d0 * kSeedCtaWarps + warp];
            __syncwarp();
            if (v0) {
                dst[off + (unsigned)__popc(p0 & lt)] = x0;
                if ((p0 & lt) == 0) hist[d0 * kSeedCtaWarps + warp] = (uint16_t)(off + (unsigned)__popc(p0));
            }
            __syncwarp();
            if (v1) off = hist[d1 * kSeedCtaWarps + warp];
            __syncwarp();
            if (v1) {
                dst[off + (unsigned)__popc(p1 & lt)] = x1;
                if ((p1 & lt) == 0) hist[d1 * kSeedCtaWarps + warp] = (uint16_t)(off + (unsigned)__popc(p1));
            }
            __syncwarp();
        }
        __syncthreads();
        uint64_t *t = src;
        src = dst;
        dst = t;
    }
    return src;
}

// ---- the read's 2-bit codes and its "irregular base" mask, staged in shared memory by the TMA (1-D bulk copies that
// complete on an mbarrier: one thread issues them, nobody's registers are involved) ----
__device__ __forceinline__ void stage_barrier_init(uint64_t *bar)
{
#ifndef AG2_EMU
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#else
    *bar = 0;
#endif
}
// thread 0: both copies, `bytes2` + `bytesi` (multiples of 16) expected on the barrier
__device__ __forceinline__ void stage_issue(uint64_t *bar, void *dst2, const void *src2, unsigned bytes2, void *dsti, const void *srci, unsigned bytesi)
{
#ifndef AG2_EMU
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the buffer was last written by ordinary stores (the previous sort)
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes2 + bytesi) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"((unsigned)__cvta_generic_to_shared(dst2)),
                 "l"(src2), "r"(bytes2), "r"(b)
                 : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"((unsigned)__cvta_generic_to_shared(dsti)),
                 "l"(srci), "r"(bytesi), "r"(b)
                 : "memory");
#else
    memcpy(dst2, src2, bytes2);
    memcpy(dsti, srci, bytesi);
    (void)bar;
#endif
}
__device__ __forceinline__ void stage_wait(uint64_t *bar, unsigned parity)
{
#ifndef AG2_EMU
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("{\n\t.reg .pred p;\n\tSTAGE_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra STAGE_WAIT;\n\t}" ::"r"(b), "r"(parity)
                 : "memory");
#else
    (void)bar;
    (void)parity;
#endif
}

// 13 bases at oriented positions start .. start + 12 of a read strand as a seed code (atcttrans: A0 T1 C2 G3, first base
// most significant), or -1 if one of them is not upper-case ACGT or lies behind the read (transnum_buchang :95-121).
// roff2 / roffi: position of the read's first base in the `reads2` / `irr` arrays handed in (global: the same packed
// offset; staged: the few bases the 16-byte aligned copies start early).
__device__ __forceinline__ int seed_code_fast(const uint32_t *reads2, const uint32_t *irr, int64_t roff2, int64_t roffi, int rlen, int strand, int start)
{
    if (start + kSeedLen > rlen) return -1;
    const int lo = strand ? rlen - start - kSeedLen : start;   // lowest file position of the window
    {
        const int64_t f0 = roffi + lo;
        const int64_t w = f0 >> 5;
        const int sh = (int)(f0 & 31);
        const uint64_t bits = ((uint64_t)irr[w + 1] << 32 | irr[w]) >> sh;
        if (bits & 0x1fffu) return -1;
    }
    const int64_t f0 = roff2 + lo;
    const int64_t w = f0 >> 4;
    const int sh = 2 * (int)(f0 & 15);
    uint32_t x = (uint32_t)((((uint64_t)reads2[w + 1] << 32) | reads2[w]) >> sh) & 0x3ffffffu;   // base f0 in bits 1:0
    if (strand) {
        x ^= 0x3ffffffu;                                     // complement; the highest file position is the first base already
    } else {
        x = __brev(x) >> 6;                                  // first base most significant ...
        x = ((x & 0x1555555u) << 1) | ((x >> 1) & 0x1555555u);   // ... with the bits of every pair back in order
    }
    // A0 C1 G2 T3 -> A0 T1 C2 G3: (hi, lo) -> (hi ^ lo, hi)
    const uint32_t hi = (x >> 1) & 0x1555555u, lo2 = x & 0x1555555u;
    return (int)(((hi ^ lo2) << 1) | hi);
}

// find_location3's votes (:614-627) for entry x of the pooled list, by the thread that owns x
__device__ __forceinline__ int fl3_vote(const RefIndex &ix, const int *t_loc, const int *t_seedn, int k, int x, float len, int read_len1,
                                        int64_t start_loc)
{
    int cnt = 0;
    for (int y = 0; y < k; ++y) {
        if (y == x) continue;
        const int i = y < x ? y : x, j = y < x ? x : y;
        if (t_seedn[j] - t_seedn[i] > 0 && t_loc[j] - t_loc[i] > 0 && t_loc[j] - t_loc[i] < read_len1 &&
            ddf_ok_f(t_loc[j] - t_loc[i], t_seedn[j] - t_seedn[i], len))
            ++cnt;
    }
    const int64_t nn = (start_loc + t_loc[x]) / ix.cbl;
    return (int)fdiv_rn((float)cnt, ix.vote[nn]);
}

// entry q of run r's Back_List (q < min(score, 20)): light runs read their events, heavy runs their pool entry
struct RunView {
    const uint64_t *ev;
    const uint16_t *run_start;
    const uint32_t *pool;
};
__device__ __forceinline__ void run_entry(const RunView &v, int r, int q, int &loc, int &seedn)
{
    const int s = v.run_start[r], m = v.run_start[r + 1] - s;
    if (m <= kSM) {
        const uint64_t x = v.ev[s + q];
        loc = ev_u(x);
        seedn = ev_seed(x);
    } else {
        const uint32_t w = v.pool[(size_t)ev_e(v.ev[s + 1]) * kHeavyWords + q];
        loc = (int)(w & 0x7ffu);
        seedn = (int)(w >> 11);
    }
}

// first run with block >= b
__device__ __forceinline__ int run_lower_bound(const RunView &v, int nruns, int64_t b)
{
    int lo = 0, hi = nruns;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((int64_t)ev_block(v.ev[v.run_start[mid]]) < b) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// Steps 1-3 for one strand: events, sort, runs, heavy blocks.  Returns the number of runs, or -1 if the strand does not
// fit `cap`.  On return sm.ev is sorted, sm.run_start / sm.run_score are set; `pool` receives the heavy blocks.
__device__ int seed_cta_build(const RefIndex &ix, const uint32_t *reads2, const uint32_t *irr, int64_t roff, int rlen, int strand, int BC,
                              int64_t zv, int cap, int block_bits, SeedCtaSmem &sm, uint32_t *pool)
{
    const int tid = threadIdx.x;
    const int cleave_num = (rlen - kSeedLen) / BC + 1;
    if (cleave_num > 0x7fff) return -1;   // the seed number has 15 bits here (reads beyond the reference's RM = 100 000 characters): thread path
    int n_ev = 0;
    bool overflow = false;
    uint64_t *ev = sm.ev;
    // 1. events: every thread expands the bucket of its seed in place; a hit that is not the first of its seed in its
    //    block stays in the stream as a null event (it sorts behind everything and the serial numbers keep their order)
    if (tid == 0) sm.misc[4] = 0;     // null events
    // the read's codes and mask go to shared memory first (into the sort's second buffer, idle until the events are
    // complete) when they fit: one TMA bulk copy each, 16-byte aligned start and length
    const uint32_t *codes = reads2, *mask = irr;
    int64_t roff2 = roff, roffi = roff;
    {
        const int64_t a2 = roff & ~(int64_t)63, ai = roff & ~(int64_t)127;      // 64 bases = 16 bytes of codes, 128 bases = 16 bytes of mask
        const unsigned bytes2 = (unsigned)(((roff + rlen - a2 + 15) / 16 * 4 + 8 + 15) & ~15), bytesi = (unsigned)(((roff + rlen - ai + 31) / 32 * 4 + 8 + 15) & ~15);
        if ((size_t)bytes2 + bytesi <= (size_t)cap * 8) {
            uint8_t *dst2 = reinterpret_cast<uint8_t *>(sm.aux), *dsti = dst2 + bytes2;
            const unsigned parity = (unsigned)sm.misc[5] & 1u;
            __syncthreads();          // everybody is done with what the buffer held
            if (tid == 0) {
                stage_issue(sm.bar, dst2, reads2 + (a2 >> 4), bytes2, dsti, irr + (ai >> 5), bytesi);
                sm.misc[5] = (int)(parity ^ 1u);
            }
            stage_wait(sm.bar, parity);
            codes = reinterpret_cast<const uint32_t *>(dst2);
            mask = reinterpret_cast<const uint32_t *>(dsti);
            roff2 = roff - a2;
            roffi = roff - ai;
        }
    }
    __syncthreads();
    const uint32_t zv32 = (uint32_t)zv;
    int nulls = 0;
    for (int k0 = 0; k0 < cleave_num && !overflow; k0 += kSeedCtaThreads) {
        const int k = k0 + tid;
        uint32_t o0 = 0, cnt = 0;
        if (k < cleave_num) {
            const int code = seed_code_fast(codes, mask, roff2, roffi, rlen, strand, k * BC);
            if (code >= 0) {
                o0 = ix.off[code];
                cnt = ix.off[code + 1] - o0;      // the CSR is built from the masked counts: 0 for a bucket above 128
            }
        }
        int total;
        const int hb = n_ev + block_scan_excl((int)cnt, total, sm.wsum);
        if (n_ev + total > cap) {
            overflow = true;
            break;
        }
        const uint32_t *p = ix.pos + o0;
        uint32_t prev_blk = 0xffffffffu;
        for (uint32_t i = 0; i < cnt; ++i) {
            const uint32_t pos = p[i];
            const uint32_t blk = zv32 == 1000u ? pos / 1000u : pos / 2000u;
            const uint32_t u = pos - blk * zv32;
            const bool acc = blk != prev_blk;         // a bucket is ascending: the first hit of this seed in this block (:853)
            prev_blk = blk;
            ev[hb + (int)i] = acc ? ev_pack(blk, k + 1, (int)u, hb + (int)i) : ev_pack((uint32_t)kNullBlock, 0, 0, hb + (int)i);
            nulls += acc ? 0 : 1;
        }
        n_ev += total;
    }
    if (nulls) atomicAdd(reinterpret_cast<unsigned *>(&sm.misc[4]), (unsigned)nulls);
    __syncthreads();
    if (overflow) return -1;
    const int n_all = n_ev;
    n_ev -= sm.misc[4];
    // 2. sort by block; the runs
    ev = sort_events(sm.ev, sm.aux, n_all, 41, 41 + block_bits, sm.hist, sm.wsum);   // the null events end up behind n_ev
    uint64_t *free_buf = ev == sm.ev ? sm.aux : sm.ev;
    sm.ev = ev;
    sm.aux = free_buf;
    seed_cta_alias(sm, free_buf, cap);
    const int per = (n_ev + kSeedCtaThreads - 1) / kSeedCtaThreads;
    const int c_lo = min(n_ev, tid * per), c_hi = min(n_ev, c_lo + per);
    int heads = 0;
    for (int i = c_lo; i < c_hi; ++i) heads += (i == 0 || ev_block(ev[i]) != ev_block(ev[i - 1])) ? 1 : 0;
    int nruns;
    int ri = block_scan_excl(heads, nruns, sm.wsum);
    for (int i = c_lo; i < c_hi; ++i)
        if (i == 0 || ev_block(ev[i]) != ev_block(ev[i - 1])) sm.run_start[ri++] = (uint16_t)i;
    if (tid == 0) {
        sm.run_start[nruns] = (uint16_t)n_ev;
        sm.misc[0] = 0;     // heavy pool slots used
    }
    __syncthreads();
    // 3. scores; heavy blocks replayed through insert_loc
    for (int r = tid; r < nruns; r += kSeedCtaThreads) {
        const int s = sm.run_start[r], m = sm.run_start[r + 1] - s;
        if (m <= kSM) {
            sm.run_score[r] = (int16_t)m;
            continue;
        }
        BackList bl;
        bl.score = bl.score2 = 0;
        bl.seednum = 0;
        bl.index = -1;
        const int64_t templong = ev_block(ev[s]);
        for (int i = 0; i < m; ++i) {
            const uint64_t x = ev[s + i];
            const int loc = ++bl.score;
            if (loc <= kSM) {
                bl.loczhi[loc - 1] = (int16_t)ev_u(x);
                bl.seedno[loc - 1] = (int16_t)ev_seed(x);
            } else {
                insert_loc(ix, &bl, ev_u(x), ev_seed(x), (float)BC, templong, zv);
            }
            if (i >= kSM) ev[s + i] = ev_with_e(x, bl.score);   // the score after this event
        }
        const int slot = atomicAdd(reinterpret_cast<unsigned *>(&sm.misc[0]), 1u);
        ev[s + 1] = ev_with_e(ev[s + 1], slot);
        for (int q = 0; q < kSM; ++q) pool[(size_t)slot * kHeavyWords + q] = ((uint32_t)(uint16_t)bl.seedno[q] << 11) | (uint32_t)(uint16_t)bl.loczhi[q];
        sm.run_score[r] = bl.score;
    }
    __syncthreads();
    return nruns;
}

// the score run p had once every event with a seed number <= seedno had been applied
__device__ __forceinline__ int run_score_at(const uint64_t *ev, const uint16_t *run_start, int p, int seedno)
{
    const int s = run_start[p], m = run_start[p + 1] - s;
    int c = 0;
    while (c < m && ev_seed(ev[s + c]) <= seedno) ++c;
    if (c == 0) return 0;
    if (m <= kSM || c <= kSM) return c;
    return ev_e(ev[s + c - 1]);
}

// Seeding + candidate scan of one strand (steps 1-5); cands[0 .. *ncand) is the read's shared list in shared memory.
// Returns false if the strand does not fit.
__device__ bool seed_cta_strand(const RefIndex &ix, const uint32_t *reads2, const uint32_t *irr, int64_t roff, int rlen, int strand, int BC,
                                int64_t zv, int thresh, int maxc, int cap, int block_bits, SeedCtaSmem &sm, uint32_t *pool)
{
    const int tid = threadIdx.x, lane = tid & 31;
    const int nruns = seed_cta_build(ix, reads2, irr, roff, rlen, strand, BC, zv, cap, block_bits, sm, pool);
    if (nruns < 0) return false;
    const uint64_t *ev = sm.ev;
    // 4. index_score (:863-872): the blocks above the threshold
    if (tid == 0) sm.misc[1] = 0;
    __syncthreads();
    const int qcap = seed_cta_qcap(cap);
    for (int r = tid; r < nruns; r += kSeedCtaThreads) {
        const int s = sm.run_start[r], e1 = sm.run_start[r + 1];
        const int b = ev_block(ev[s]);
        int sk = sm.run_score[r];
        if (r > 0 && b > 0 && ev_block(ev[sm.run_start[r - 1]]) == b - 1) sk += run_score_at(ev, sm.run_start, r - 1, ev_seed(ev[e1 - 1]));
        if ((int16_t)sk > thresh) {
            const unsigned q = atomicAdd(reinterpret_cast<unsigned *>(&sm.misc[1]), 1u);
            if ((int)q < qcap) sm.qin[q] = ((uint32_t)ev_e(ev[s]) << 15) | (uint32_t)r;
        }
    }
    __syncthreads();
    const int nq = sm.misc[1];
    if (nq > qcap) return false;
    // first-touch order: rank by the first event's serial number (all different)
    for (int i = tid; i < nq; i += kSeedCtaThreads) {
        const uint32_t me = sm.qin[i];
        int rank = 0;
        for (int j = 0; j < nq; ++j) rank += sm.qin[j] < me ? 1 : 0;
        sm.qout[rank] = me;
    }
    __syncthreads();
    // 5. candidate scan (:882-991) by warp 0
    if (tid < 32) {
        RunView rv = {ev, sm.run_start, pool};
        int ncand = sm.misc[2];
        for (int qi = 0; qi < nq; ++qi) {
            const int r = (int)(sm.qout[qi] & 0x7fffu);
            const int s_k = sm.run_score[r];
            if (s_k == 0) continue;
            const int bid = ev_block(ev[sm.run_start[r]]);
            int loc = 0;
            if (bid > 0 && r > 0 && ev_block(ev[sm.run_start[r - 1]]) == bid - 1) loc = sm.run_score[r - 1];
            int64_t start_loc = (int64_t)(loc > 0 ? bid - 1 : bid) * zv;
            const int n1 = loc > 0 ? min(loc, kSM) : 0, n2 = min(s_k, kSM), u_k = n1 + n2;
            __syncwarp();   // lane 0 may still be reading the previous block's lists (find_location_choose); a shuffle orders no memory
            for (int x = lane; x < u_k; x += 32) {
                int l, sd;
                if (x < n1) run_entry(rv, r - 1, x, l, sd);
                else {
                    run_entry(rv, r, x - n1, l, sd);
                    if (n1 > 0) l += (int)zv;
                }
                sm.tl_loc[x] = l;
                sm.tl_seed[x] = sd;
            }
            __syncwarp();
            for (int x = lane; x < u_k; x += 32) sm.tl_score[x] = fl3_vote(ix, sm.tl_loc, sm.tl_seed, u_k, x, (float)BC, rlen, start_loc);
            __syncwarp();
            int64_t location_loc[4] = {0, 0, 0, 0};
            int repeat_loc = 0, flag = 0;
            if (lane == 0) flag = find_location_choose(sm.tl_loc, sm.tl_seed, sm.tl_score, location_loc, u_k, &repeat_loc, (float)BC, rlen);
            flag = __shfl_sync(0xffffffffu, flag, 0);
            if (!flag) continue;
            repeat_loc = __shfl_sync(0xffffffffu, repeat_loc, 0);
            location_loc[0] = __shfl_sync(0xffffffffu, location_loc[0], 0);
            location_loc[1] = __shfl_sync(0xffffffffu, location_loc[1], 0);
            if (sm.tl_score[repeat_loc] < 6) continue;
            SeedCand ct;
            ct.score = sm.tl_score[repeat_loc];
            const int loc_seed = sm.tl_seed[repeat_loc];
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
            {   // consistent seeds in the blocks bid - 2 - num1 / zv .. bid - 2 (:950-961), one block per lane
                const int64_t b_hi = (int64_t)bid - 2, b_lo = b_hi - ct.num1 / (int)zv;
                if (b_hi >= 0) {
                    const int r_lo = run_lower_bound(rv, nruns, b_lo < 0 ? 0 : b_lo), r_hi = run_lower_bound(rv, nruns, b_hi + 1);
                    for (int p = r_lo + lane; p < r_hi; p += 32) {
                        const int sc = sm.run_score[p];
                        if (sc <= 0) continue;
                        const int64_t sl = (int64_t)ev_block(ev[sm.run_start[p]]) * zv;
                        const int scnt = sc < kSM ? sc : kSM;
                        int sk = 0;
                        for (int q = 0; q < scnt; q++) {
                            int l, sd;
                            run_entry(rv, p, q, l, sd);
                            if (fabs(dsub_rn(ddiv_rn((double)(loc_list - sl - l), dmul_rn((double)((loc_seed - sd) * BC), 1.0)), 1.0)) < 0.25) sk++;
                        }
                        seedcount += sk;
                        if (ddiv_rn(dmul_rn((double)sk, 1.0), (double)scnt) > 0.4) sm.run_score[p] = 0;
                    }
                }
            }
            {   // and in the blocks bid + 1 .. bid + num2 / zv (:963-973)
                const int64_t b_lo = (int64_t)bid + 1, b_hi = (int64_t)bid + ct.num2 / (int)zv;
                const int r_lo = run_lower_bound(rv, nruns, b_lo), r_hi = run_lower_bound(rv, nruns, b_hi + 1);
                for (int p = r_lo + lane; p < r_hi; p += 32) {
                    const int sc = sm.run_score[p];
                    if (sc <= 0) continue;
                    const int64_t sl = (int64_t)ev_block(ev[sm.run_start[p]]) * zv;
                    const int scnt = sc < kSM ? sc : kSM;
                    int sk = 0;
                    for (int q = 0; q < scnt; q++) {
                        int l, sd;
                        run_entry(rv, p, q, l, sd);
                        if (fabs(dsub_rn(ddiv_rn((double)(sl + l - loc_list), dmul_rn((double)((sd - loc_seed) * BC), 1.0)), 1.0)) < 0.25) sk++;
                    }
                    seedcount += sk;
                    if (ddiv_rn(dmul_rn((double)sk, 1.0), (double)scnt) > 0.4) sm.run_score[p] = 0;
                }
            }
            seedcount = __reduce_add_sync(0xffffffffu, seedcount);
            __syncwarp();
            ct.score += seedcount;
            ct.chain = strand == 0 ? 'F' : 'R';
            if (lane == 0) {   // keep the MAXC best, ties after equals (:978-990)
                SeedCand *cands = sm.cands;
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
            }
            if (ncand < maxc) ncand++;
            __syncwarp();
        }
        __syncwarp();       // every lane has read misc[2] (no block above the threshold: nothing in between)
        if (lane == 0) sm.misc[2] = ncand;
    }
    __syncthreads();
    return true;
}

__device__ __forceinline__ int seed_block_bits(int64_t ref_len, int64_t zv)
{
    const int64_t nb = ref_len / zv + 2;
    int bits = 1;
    while (((int64_t)1 << bits) < nb) ++bits;
    return bits;
}

// Body of seed_cta_kernel: a persistent CTA takes items from the work counter.
__device__ void seed_cta_body(const SeedCtaArgs &a, uint8_t *smem)
{
    const int tid = threadIdx.x;
    const unsigned n_work = a.n_work_dev ? *a.n_work_dev : a.n_work;
    const int64_t zv = a.pass == 0 ? 1000 : 2000;
    const int thresh = a.pass == 0 ? 6 : 4;
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
        SeedCtaSmem sm = seed_cta_carve(smem, a.cap);
        if (tid == 0) sm.misc[3] = (int)atomicAdd(a.next, 1u);
        __syncthreads();
        const unsigned w = (unsigned)sm.misc[3];
        if (w >= n_work) break;
        const int64_t k = a.work ? a.work[w] : (int64_t)w;
        const int64_t r = a.reads ? a.reads[k] : k;
        const int rlen = a.read_len[r];
        const int64_t roff = a.read_off[r];
        const int BC = seed_stride(rlen, a.pass);
        if (tid == 0) sm.misc[2] = 0;    // candidates so far
        __syncthreads();
        bool ok = true;
        for (int strand = 0; strand < 2 && ok; ++strand) {
            SeedCtaSmem s2 = seed_cta_carve(smem, a.cap);
            ok = seed_cta_strand(a.ix, a.reads2, a.irr, roff, rlen, strand, BC, zv, thresh, a.maxc, a.cap, block_bits, s2, pool);
        }
        if (ok) {
            const int nc = sm.misc[2];
            if (tid == 0) a.ncand[k] = nc;
            for (int i = tid; i < nc * (int)(sizeof(SeedCand) / 8); i += kSeedCtaThreads)
                reinterpret_cast<uint64_t *>(a.cands + k * a.maxc)[i] = reinterpret_cast<const uint64_t *>(sm.cands)[i];
        } else if (tid == 0) {
            a.ncand[k] = 0;
            a.ovf[atomicAdd(a.ovf_count, 1u)] = (int32_t)k;
        }
        __syncthreads();
    }
}

} // namespace ag2
