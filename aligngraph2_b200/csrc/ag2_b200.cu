// ag2_b200.cu -- kernels and the C ABI (include/ag2_b200.h) of the B200-native mecat2ref+ hot path.
//
// Built for sm_100a only (see __graft_entry__.build()).  No CPU fallback: every entry point needs a
// live CUDA context and returns AG2_ENODEV / AG2_ECUDA otherwise.
#include "../../include/ag2_b200.h"
#include "xdrop_device.cuh"
#include "xdrop_lane.cuh"
#include "xdrop_pair.cuh"
#include "index_kernels.cuh"
#include "map_kernels.cuh"
#include "kmer_kernels.cuh"

#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <thread>
#include <chrono>

using namespace ag2;

static_assert(sizeof(ag2_candidate) == sizeof(Candidate), "ag2_candidate layout");
static_assert(sizeof(ag2_record) == sizeof(Record), "ag2_record layout");

namespace {

constexpr int kWideK = 23;        // 736 columns: covers every possible block (N <= 718)
constexpr int kWideWarps = 2;
constexpr int kMaxStreamChunks = 256;

// ---------------------------------------------------------------------------------------------
// packing kernels (A8: get_dna_encode_table + ">3 -> 0", MC/defs.cpp:3-36, mecat2ref_aux.cpp:195-197)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned encode_ascii(unsigned c, unsigned &irregular)
{
    unsigned code = 0;
    irregular = 1;
    switch (c) {
    case 'A': code = 0; irregular = 0; break;
    case 'C': code = 1; irregular = 0; break;
    case 'G': code = 2; irregular = 0; break;
    case 'T': code = 3; irregular = 0; break;
    case 'a': code = 0; break;
    case 'c': code = 1; break;
    case 'g': code = 2; break;
    case 't': code = 3; break;
    default: break;
    }
    return code;
}

// one thread = 32 bases of one read = two 2-bit words + one "irregular" word.
// poff[r] = packed base offset of read r (multiple of 32); offs[r] = ASCII offset.
// piece_done / piece_flag (may be null): the CTA that finishes last raises the flag a running streamed extension waits for
// (wait_for_read, xdrop_device.cuh) -- by the kernel itself, not by a memset behind it, which the driver may run as a kernel
// of its own with another shared-memory carve-out.
__global__ void pack_reads_kernel(const char *__restrict__ ascii, const int64_t *__restrict__ offs,
                                  const int64_t *__restrict__ poff, int64_t n_reads, int64_t g_first, int64_t n_groups,
                                  uint32_t *__restrict__ out2, uint32_t *__restrict__ irr_out, unsigned int *piece_done, int *piece_flag)
{
    for (int64_t g = g_first + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < g_first + n_groups; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t pb = g << 5;
        int64_t lo = 0, hi = n_reads - 1; // last read with poff[r] <= pb
        while (lo < hi) {
            const int64_t mid = (lo + hi + 1) >> 1;
            if (poff[mid] <= pb) lo = mid;
            else hi = mid - 1;
        }
        const int64_t len = offs[lo + 1] - offs[lo];
        const int64_t local = pb - poff[lo];
        const char *src = ascii + offs[lo] + local;
        uint32_t w0 = 0, w1 = 0, ir = 0;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            if (local + i < len) {
                unsigned irregular;
                const unsigned code = encode_ascii((unsigned char)src[i], irregular);
                if (i < 16) w0 |= code << (2 * i);
                else w1 |= code << (2 * (i - 16));
                ir |= irregular << i;
            }
        }
        out2[2 * g] = w0;
        out2[2 * g + 1] = w1;
        irr_out[g] = ir;
    }
    if (piece_flag) {
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0 && atomicAdd(piece_done, 1u) + 1 == gridDim.x) {
            __threadfence();
            *reinterpret_cast<volatile int *>(piece_flag) = 1;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// extension kernels
// ---------------------------------------------------------------------------------------------
__global__ void extend_setup_kernel(const Candidate *cand, int64_t n, PackedSeqs sq, int64_t n_reads, ExtGeom *geom,
                                    int64_t *caps, int64_t *nmeta)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        ExtGeom g;
        int64_t m;
        caps[i] = setup_one(cand[i], sq, n_reads, g, m);
        nmeta[i] = m;
        geom[i] = g;
    }
}

// single-CTA exclusive scan of int64 (n up to a few million; plumbing, not a hot kernel).
// out[i] = sum in[0..i), out[n] = total.
__global__ void exclusive_scan_i64(const int64_t *in, int64_t n, int64_t *out)
{
    __shared__ int64_t part[1024];
    const int t = threadIdx.x;
    const int64_t per = (n + blockDim.x - 1) / blockDim.x;
    const int64_t lo = min(n, t * per), hi = min(n, lo + per);
    int64_t s = 0;
    for (int64_t i = lo; i < hi; ++i) s += in[i];
    part[t] = s;
    __syncthreads();
    if (t == 0) {
        int64_t acc = 0;
        for (unsigned i = 0; i < blockDim.x; ++i) {
            const int64_t v = part[i];
            part[i] = acc;
            acc += v;
        }
        out[n] = acc;
    }
    __syncthreads();
    s = part[t];
    for (int64_t i = lo; i < hi; ++i) {
        const int64_t v = in[i];
        out[i] = s;
        s += v;
    }
}

// The same for one output chunk of a streamed run: 128 threads, no static shared array -- it has to fit beside the resident
// pair kernel.  out[i] = sum in[0..i); *total = sum in[0..n).
__global__ void __launch_bounds__(128) scan_chunk_kernel(const int64_t *in, int64_t n, int64_t *out, int64_t *total)
{
    __shared__ int64_t part[128];
    const int t = threadIdx.x;
    const int64_t per = (n + 127) / 128;
    const int64_t lo = min(n, t * per), hi = min(n, lo + per);
    int64_t s = 0;
    for (int64_t i = lo; i < hi; ++i) s += in[i];
    part[t] = s;
    __syncthreads();
    if (t == 0) {
        int64_t acc = 0;
        for (int i = 0; i < 128; ++i) {
            const int64_t v = part[i];
            part[i] = acc;
            acc += v;
        }
        *total = acc;
    }
    __syncthreads();
    s = part[t];
    for (int64_t i = lo; i < hi; ++i) {
        const int64_t v = in[i];
        out[i] = s;
        s += v;
    }
}

__global__ void set_slots_kernel(ExtGeom *geom, const int64_t *prefix, const int64_t *meta_prefix, int64_t first, int64_t n)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        geom[first + i].slot = prefix[first + i] - prefix[first];
        geom[first + i].meta = meta_prefix[first + i] - meta_prefix[first];
    }
}

// Longest-first order of a chunk's directions: key = columns reserved for the direction (proportional to the bases it
// can extend over), so the directions started last, when the queue runs dry, are the shortest ones.
__global__ void chain_keys_kernel(const ExtGeom *geom, int64_t n_chains, uint32_t *keys, int32_t *ids)
{
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n_chains; t += (int64_t)gridDim.x * blockDim.x) {
        const ExtGeom g = geom[t >> 1];
        const int32_t cap = g.valid ? ((t & 1) ? g.capR : g.capL) : 0;
        keys[t] = (uint32_t)max(cap, 0);
        ids[t] = (int32_t)t;
    }
}

// Streamed run: the directions of ALL output chunks share one queue.  A direction of chunk c that needs `cap` units of a
// slot's time has to start by (c + 1) * delta - cap for its chunk to be complete on schedule (delta = a chunk's share of a
// slot's time): ascending order of that latest start.  With one chunk this is longest-first.
__global__ void stream_keys_kernel(const ExtGeom *geom, int64_t n_chains, int32_t chunk_cn, int64_t delta, uint32_t *keys, int32_t *ids)
{
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n_chains; t += (int64_t)gridDim.x * blockDim.x) {
        const ExtGeom g = geom[t >> 1];
        const int64_t cap = g.valid ? max((t & 1) ? g.capR : g.capL, 0) : 0;
        const int64_t c = (t >> 1) / chunk_cn;
        const int64_t key = (c + 1) * delta - cap + ((int64_t)1 << 30);
        keys[t] = (uint32_t)(key < 0 ? 0 : key > (int64_t)0xffffffffu ? (int64_t)0xffffffffu : key);
        ids[t] = (int32_t)t;
    }
}

// Lane path (xdrop_lane.cuh): one thread per extension direction, 32 directions in lock step per
// warp, refilled from a global queue at block boundaries.
__global__ void __launch_bounds__(kLaneThreads) xdrop_lane_kernel(LaneArgs g)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    LaneSmem &sm = *reinterpret_cast<LaneSmem *>(smem_raw);
    const int tid = threadIdx.x;
    uint8_t *scratch = g.scratch + ((size_t)blockIdx.x * kLaneThreads + tid) * kLaneScratch;
    lane_kernel_body(g, sm, tid, scratch);
}

// Pair path (xdrop_pair.cuh): two directions per thread in packed 16-bit arithmetic, 64 directions in lock step
// per warp.  The dominant kernel; the lane kernel above restarts the directions it hands over.
__device__ __forceinline__ void pair_kernel_entry(const LaneArgs &g);
__global__ void __launch_bounds__(kPairThreads, 8) xdrop_pair_kernel(LaneArgs g) { pair_kernel_entry(g); }
// The same built for nine CTAs per SM: ptxas then takes 96 registers (16 bytes spilled).  An experiment of the streamed form
// (AG2_STREAM_LEAN_KERNEL=1): eight CTAs that leave a quarter of the register file to the small kernels beside them.
__global__ void __launch_bounds__(kPairThreads, 9) xdrop_pair_kernel_lean(LaneArgs g) { pair_kernel_entry(g); }
__device__ __forceinline__ void pair_kernel_entry(const LaneArgs &g)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    PairSmem &sm = *reinterpret_cast<PairSmem *>(smem_raw);
    const int tid = threadIdx.x;
    uint8_t *scratch = g.scratch + (size_t)blockIdx.x * kPairCtaScratch; // the CTA's: traceback of its 128 directions interleaved
    if (g.done_ctas && tid == 0) reinterpret_cast<volatile unsigned int *>(g.done_ctas)[1] = 1u; // a producer runs: the consumer may take tickets
    pair_kernel_body(g, sm, tid, scratch);
    if (g.done_ctas) { // tells the concurrent consumer that this CTA will publish nothing more
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            atomicAdd(g.done_ctas, 1u);
        }
    }
}

// One warp per record: the two directions' workspace areas -> dense strings; fills aln_off.
__global__ void assemble_kernel(Record *rec, const ExtGeom *geom, const ChainResult *res, const uint32_t *meta,
                                const int64_t *dense_off, int64_t dense_base, int64_t first, int64_t n, const char *ws_q,
                                const char *ws_t, char *out_q, char *out_t, unsigned long long *aligned,
                                unsigned long long *columns)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t k = warp; k < n; k += nwarps) {
        const int64_t i = first + k;
        const int64_t dst = dense_base + dense_off[i];
        if (lane == 0) rec[i].aln_off = dst;
        if (!rec[i].ok) continue;
        assemble_record(geom[i], res[2 * i], res[2 * i + 1], meta, ws_q, ws_t, out_q + dst, out_t + dst, lane);
        if (lane == 0) {
            atomicAdd(aligned, (unsigned long long)(rec[i].qe - rec[i].qb));
            atomicAdd(columns, (unsigned long long)rec[i].aln_len);
        }
    }
}

// Alignment columns as 2-bit ops, 16 per word (column j of the dense pool at bits 2 * (j % 16) of word j / 16): 0 = a base of
// both, 1 = '-' in the read string, 2 = '-' in the reference string.  With the read and the reference the caller has, this
// is the whole content of the two ASCII strings at an eighth of the bytes (ag2_expand_alignments rebuilds them).
__global__ void pack_ops_kernel(const char *__restrict__ q, const char *__restrict__ t, int64_t w_lo, int64_t w_hi, int64_t col_end,
                                uint32_t *__restrict__ ops)
{
    for (int64_t w = w_lo + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; w < w_hi; w += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c0 = w << 4;
        uint32_t v = 0;
        if (c0 + 16 <= col_end) {
            const uint4 a = *reinterpret_cast<const uint4 *>(q + c0), b = *reinterpret_cast<const uint4 *>(t + c0);
            const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const uint32_t qc = (aw[k >> 2] >> (8 * (k & 3))) & 0xffu, tc = (bw[k >> 2] >> (8 * (k & 3))) & 0xffu;
                v |= (qc == '-' ? 1u : tc == '-' ? 2u : 0u) << (2 * k);
            }
        } else {
            for (int k = 0; k < 16 && c0 + k < col_end; ++k) v |= (q[c0 + k] == '-' ? 1u : t[c0 + k] == '-' ? 2u : 0u) << (2 * k);
        }
        ops[w] = v;
    }
}

// Wide path (any band a block can have): the int32 row kernel with K columns per lane.
template <int K, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) xdrop_chains_kernel(ChainArgs g)
{
    __shared__ WarpSmem smem[WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpSmem &sm = smem[warp];
    uint8_t *tb = g.tb + ((size_t)blockIdx.x * WARPS + warp) * g.tb_stride;
    ChainCounters ctr = {0, 0, 0, 0, 0};
    for (;;) {
        unsigned long long t = 0;
        if (lane == 0) t = atomicAdd(g.next, 1ull);
        t = __shfl_sync(kFull, t, 0);
        if ((int64_t)t >= g.n_chains) break;
        const int64_t chain = g.queue ? (int64_t)g.queue[t] : (int64_t)t;
        // most bands that left the pair kernel's 96-column window still fit 128 columns: rows ~5 x shorter than with K = 23
        bool done = continue_handed_over<K>(g, chain, t, sm, tb, lane, ctr);
        if (!done && K > kNarrowK && g.try_narrow) done = run_chain<kNarrowK>(g, chain, sm, tb, lane, ctr);
        if (!done) done = run_chain<K>(g, chain, sm, tb, lane, ctr);
        if (!done) {
            ctr.wide += 1;
            if (lane == 0 && g.wide_count) g.wide_queue[atomicAdd(g.wide_count, 1u)] = (int32_t)chain;
        }
    }
    if (lane == 0) {
        atomicAdd(&g.counters->cells, ctr.cells);
        atomicAdd(&g.counters->rows, ctr.rows);
        atomicAdd(&g.counters->blocks, ctr.blocks);
        atomicAdd(&g.counters->interior, ctr.interior);
        atomicAdd(&g.counters->wide, ctr.wide);
    }
}

// The wide path as a CONSUMER that runs beside the pair kernel: a few CTAs, launched first, poll the hand-over queue and
// rerun each published direction at once (one warp per direction), so that no hand-over is left for after the pair kernel
// when the GPU would sit idle behind a handful of sequential directions.  Tickets are taken only while producers are
// active; [min(*next, *count), *count) is what the host still has to run afterwards.
constexpr unsigned long long kConsumerStartNs = 1500ull * 1000 * 1000;
template <int K, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) xdrop_stream_kernel(ChainArgs g, const unsigned int *done_ctas, unsigned int n_producers)
{
    __shared__ WarpSmem smem[WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpSmem &sm = smem[warp];
    uint8_t *tb = g.tb + ((size_t)blockIdx.x * WARPS + warp) * g.tb_stride;
    ChainCounters ctr = {0, 0, 0, 0, 0};
    const volatile int32_t *queue = g.queue;
    const volatile unsigned int *done = done_ctas;
    // No ticket before a producer is seen running beside this kernel (done_ctas[1]).  Where kernels do not overlap -- a
    // profiler that serialises launches -- the producers start only after this kernel has ended: it leaves after
    // kConsumerStartNs without a ticket and the host's post-pass runs every hand-over.
    int go = 1;
    if (lane == 0 && done[1] == 0) {
        unsigned long long t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while (done[1] == 0) {
            __nanosleep(2000);
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > kConsumerStartNs) {
                go = 0;
                break;
            }
        }
    }
    go = __shfl_sync(kFull, go, 0);
    while (go) {
        int chain = -1;
        unsigned long long t = 0;
        if (lane == 0 && *done < n_producers) {
            t = atomicAdd(g.next, 1ull);
            for (;;) {
                chain = queue[t];
                if (chain >= 0) break;
                if (*done >= n_producers) { // everything that will ever be published is visible now
                    __threadfence();
                    chain = queue[t];
                    break;
                }
                __nanosleep(4000);
            }
        }
        chain = __shfl_sync(kFull, chain, 0);
        t = __shfl_sync(kFull, t, 0);
        if (chain < 0) break;
        __threadfence();
        bool ok = continue_handed_over<K>(g, chain, t, sm, tb, lane, ctr);
        if (!ok && K > kNarrowK && g.try_narrow) ok = run_chain<kNarrowK>(g, chain, sm, tb, lane, ctr);
        if (!ok) ok = run_chain<K>(g, chain, sm, tb, lane, ctr);
        if (!ok) ctr.wide += 1;
    }
    if (lane == 0) {
        atomicAdd(&g.counters->cells, ctr.cells);
        atomicAdd(&g.counters->rows, ctr.rows);
        atomicAdd(&g.counters->blocks, ctr.blocks);
        atomicAdd(&g.counters->interior, ctr.interior);
        atomicAdd(&g.counters->wide, ctr.wide);
    }
}

__global__ void extend_finalize_kernel(const Candidate *cand, const ExtGeom *geom, const ChainResult *res,
                                       const int32_t *read_len, int64_t first, int64_t n, Record *rec,
                                       int64_t *str_begin, int64_t *ok_len)
{
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = first + k;
        const Candidate c = cand[i];
        const ExtGeom g = geom[i];
        Record o;
        int64_t sb;
        const int rlen = g.valid ? read_len[c.read] : 0;
        finalize_one(c, g, res[2 * i], res[2 * i + 1], rlen, o, sb);
        rec[i] = o;
        str_begin[i] = sb;
        ok_len[i] = o.ok ? o.aln_len : 0;
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

} // namespace

struct ag2_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    std::string err;

    DevBuf ascii;                       // staging for ASCII uploads
    DevBuf ref2, ref_irr, ref_ascii, ref_offs;
    int64_t ref_len = 0;
    // index (A2-A4) and seeding (A5-A7)
    DevBuf ix_rcnt, ix_cnt, ix_off, ix_pos, ix_fill, ix_tiles, ix_kcount, ix_vote;
    int64_t ix_npos = 0, ix_nblk = 0;
    int ix_cbl = 0;
    bool ref_indexed = false, votes_ready = false;
    int64_t read_prefix_len = 0;       // bytes of the concatenated reads that enter the read index (A2)
    DevBuf seed_need, seed_prefix, seed_scratch, seed_cands, seed_ncand;
    DevBuf seed_ctl, seed_ovf1, seed_ovf2, seed_pool, plan_list;   // CTA-per-read seeding / rescue planning: control words, overflow lists, heavy-block pool
    ag2_map_stats map_stats = {};
    unsigned seed_overflow[2] = {0, 0};   // items the last seeding stage passed on to the second CTA launch / to the thread path
    // ag2_kmer_* (PAGraph kmer_counter)
    DevBuf km_table, km_hist, km_flags, km_offs, km_out;
    int km_k = 0;
    int64_t km_n_solid = 0;
    // ag2_map_reads
    DevBuf rec_pool, map_cand, map_cand_prefix, map_plans, map_rescue_n, map_rescue_prefix, map_out_refs, map_nout, map_flags,
        map_list, map_out_prefix, map_out_rec;
    int64_t map_n_out = 0;
    bool mapped = false;
    size_t seed_scratch_limit = 0;        // 0 = from the free memory (seed_limit); else a fixed limit in bytes
    DevBuf reads2, reads_irr, read_off, read_len, ascii_offs;
    int64_t n_reads = 0, read_bases = 0;
    int32_t read_len_p90 = 0;          // 90th percentile of the read lengths (of a sample): what the seeding tables are sized for

    DevBuf cand, geom, caps, prefix, nmeta, meta_prefix, meta, res, rec, str_begin, ok_len, dense_off;
    int64_t n_cand = 0;
    std::vector<int64_t> h_prefix, h_meta_prefix;
    DevBuf ws_q, ws_t;                  // workspace strings (one chunk of candidates)
    DevBuf out_q, out_t;                // dense output strings
    DevBuf out_ops;                     // the same as 2-bit ops (packed output)
    int64_t out_total = 0;
    DevBuf tb, tb_wide, tb_pair;
    DevBuf wide_queue, lane_queue, lane_resume, defer_queue, defer_resume;
    DevBuf tb_stream;                   // traceback scratch of the consumer kernel's warps
    cudaEvent_t side_done = nullptr;
    DevBuf order_keys, order_keys2, order_ids, order_queue, order_tmp;   // longest-first queue of the pair kernel
    DevBuf scalars;                     // ChainCounters + work counters + totals
    size_t ws_limit = (size_t)24 << 30; // bytes per workspace string per chunk (resident runs: one chunk, no drain tails)
    // host-buffer runs (ag2_xdrop_extend_batch), workspace bytes per OUTPUT chunk.  Chunked form: one launch per chunk, and a
    // launch lasts at least as long as its longest direction, so a chunk must hold several times the 132 k directions the GPU
    // runs at once (6 GiB = 2 chunks at configs[1]: 606 ms per step; 1 GiB = 13 chunks: 1187 ms).  Streamed form: one launch,
    // the chunk is only the unit that is finalised, assembled and copied home while the kernel goes on.
    size_t ws_limit_chunked = (size_t)6 << 30;
    size_t ws_limit_streamed = (size_t)1 << 30;
    size_t ws_call = 0;                   // the limit of the call in progress (AG2_WS_STREAMED overrides either default)
    int64_t piece_bytes = (int64_t)256 << 20;   // ASCII bytes per upload piece of ag2_reads_load
    cudaStream_t copy_stream = nullptr;
    cudaStream_t side_stream = nullptr;   // high priority: the consumer of the pair kernel's hand-overs
    cudaStream_t in_stream = nullptr;     // read uploads + packing
    struct ReadPiece {
        int64_t end_read;                 // reads [previous end, end_read) are packed once `ready` has happened
        cudaEvent_t ready;
    };
    std::vector<ReadPiece> pieces;
    DevBuf piece_flag, piece_end;         // device copies for the streamed run: flag k != 0 once piece k is packed
    std::vector<int64_t> h_piece_end;
    DevBuf chunk_count;                   // streamed run: finished directions per output chunk
    int *chunk_flag = nullptr;            // host-mapped, kMaxStreamChunks ints
    cudaStream_t post_stream = nullptr;   // streamed run: finalize / assemble of finished chunks beside the extension kernel
    cudaEvent_t post_done = nullptr;
    cudaEvent_t chunk_done_ev = nullptr;  // streamed run: pair kernel + consumer finished
    std::vector<cudaEvent_t> piece_events;
    bool reads_pending = false;           // an asynchronous load may still be in flight
    std::vector<int64_t> h_poff;
    std::vector<int32_t> h_lens;
    cudaEvent_t chunk_done = nullptr;
    bool ran = false;
    int64_t stats_lane_chains = 0;      // directions the pair kernel handed to the lane kernel (last run)
    int64_t stats_direct_wide = 0;      // of those, how many went straight to the wide kernel

    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> chain_events;
    ag2_extend_stats stats = {};
};

namespace {

struct Scalars {
    ChainCounters ctr;
    unsigned long long next_fast, next_wide, next_pair, next_post;
    unsigned int wide_count, lane_count, stream_error, pad_;
    unsigned int pair_done, pair_started;   // adjacent: the pair kernel gets &pair_done and raises pair_done[1] when its first CTA runs
    unsigned int defer_ctl[4];              // deferred long last blocks of the pair kernel: slots reserved, slots claimed, warps in their main phase
    unsigned long long aligned, columns;
};

int fail(ag2_ctx *c, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf;
    return code;
}

#define CK(call)                                                                                    \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            return fail(ctx, e_ == cudaErrorMemoryAllocation ? AG2_ENOMEM : AG2_ECUDA, "%s: %s",    \
                        #call, cudaGetErrorString(e_));                                             \
    } while (0)

int reserve(ag2_ctx *ctx, DevBuf &b, size_t bytes)
{
    if (bytes <= b.cap) return AG2_OK;
    if (b.p) {
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaFree(b.p));
        b.p = nullptr;
        b.cap = 0;
    }
    const size_t want = bytes + bytes / 16 + 256;
    CK(cudaMalloc(&b.p, want));
    b.cap = want;
    return AG2_OK;
}

// like reserve(), but the first `keep` bytes survive a reallocation
int reserve_keep(ag2_ctx *ctx, DevBuf &b, size_t bytes, size_t keep)
{
    if (bytes <= b.cap) return AG2_OK;
    const size_t want = bytes + bytes / 8 + 256;
    void *np = nullptr;
    CK(cudaMalloc(&np, want));
    if (b.p) {
        if (keep) CK(cudaMemcpyAsync(np, b.p, keep, cudaMemcpyDeviceToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaFree(b.p));
    }
    b.p = np;
    b.cap = want;
    return AG2_OK;
}

#define RESERVE(buf, bytes)                                   \
    do {                                                      \
        int r_ = reserve(ctx, buf, bytes);                    \
        if (r_ != AG2_OK) return r_;                          \
    } while (0)

int grid_for(int64_t items, int threads, int sm_count)
{
    int64_t g = (items + threads - 1) / threads;
    const int64_t cap = (int64_t)sm_count * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

PackedSeqs seqs_of(const ag2_ctx *ctx)
{
    PackedSeqs s;
    s.ref2 = (const uint32_t *)ctx->ref2.p;
    s.ref_len = ctx->ref_len;
    s.reads2 = (const uint32_t *)ctx->reads2.p;
    s.reads_irr = (const uint32_t *)ctx->reads_irr.p;
    s.read_off = (const int64_t *)ctx->read_off.p;
    s.read_len = (const int32_t *)ctx->read_len.p;
    return s;
}

} // namespace

extern "C" {

const char *ag2_version(void) { return "aligngraph2_b200 0.1 (sm_100a)"; }

const char *ag2_last_error(const ag2_ctx *ctx) { return ctx ? ctx->err.c_str() : "no context"; }

int ag2_device_count(int *count)
{
    if (!count) return AG2_EINVAL;
    *count = 0;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) return AG2_ENODEV;
    *count = n;
    return AG2_OK;
}

int ag2_ctx_create(int device, ag2_ctx **out)
{
    if (!out) return AG2_EINVAL;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return AG2_ENODEV;
    if (device < 0 || device >= count) return AG2_EINVAL;
    ag2_ctx *ctx = new ag2_ctx();
    ctx->device = device;
    cudaDeviceProp prop;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        delete ctx;
        return AG2_ENODEV;
    }
    ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithPriority(&ctx->side_stream, cudaStreamNonBlocking, -1) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->side_done, cudaEventDisableTiming) != cudaSuccess ||
        cudaStreamCreateWithPriority(&ctx->in_stream, cudaStreamNonBlocking, -1) != cudaSuccess ||   // pack kernels go ahead of pending work
        cudaStreamCreateWithPriority(&ctx->post_stream, cudaStreamNonBlocking, -1) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->post_done, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->chunk_done_ev, cudaEventDisableTiming) != cudaSuccess ||
        cudaHostAlloc((void **)&ctx->chunk_flag, kMaxStreamChunks * sizeof(int), cudaHostAllocMapped) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->chunk_done, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess ||
        cudaMalloc(&ctx->scalars.p, sizeof(Scalars)) != cudaSuccess) {
        delete ctx;
        return AG2_ECUDA;
    }
    ctx->scalars.cap = sizeof(Scalars);
    bool carve_ok = true;
    // The small kernels that must run BESIDE the resident pair kernel (packing of the pieces still arriving, finalize / scan /
    // assemble of the finished chunks) need the pair kernel's shared-memory carve-out: an SM changes its carve-out only when
    // idle, so a kernel that asked for another one would wait for the pair kernel to end -- for the pack kernel, whose flags
    // the pair kernel waits for, that would be a deadlock.
    carve_ok = carve_ok && cudaSuccess == (cudaFuncSetAttribute(pack_reads_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    carve_ok = carve_ok && cudaSuccess == (cudaFuncSetAttribute(extend_finalize_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    carve_ok = carve_ok && cudaSuccess == (cudaFuncSetAttribute(scan_chunk_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    carve_ok = carve_ok && cudaSuccess == (cudaFuncSetAttribute(assemble_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    carve_ok = carve_ok && cudaSuccess == (cudaFuncSetAttribute(pack_ops_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));

    if (!carve_ok) {
        ag2_ctx_destroy(ctx);
        return AG2_ECUDA;
    }
    // test knobs: many small chunks / upload pieces on small inputs
    if (const char *e = getenv("AG2_PIECE_BYTES")) ctx->piece_bytes = std::max(1ll, atoll(e));
    *out = ctx;
    return AG2_OK;
}

void ag2_ctx_destroy(ag2_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    DevBuf *all[] = {&ctx->ref_irr, &ctx->ref_ascii, &ctx->ref_offs, &ctx->ix_rcnt, &ctx->ix_cnt, &ctx->ix_off, &ctx->ix_pos,
                     &ctx->ix_fill, &ctx->ix_tiles, &ctx->ix_kcount, &ctx->ix_vote, &ctx->seed_need, &ctx->seed_prefix,
                     &ctx->seed_scratch, &ctx->seed_cands, &ctx->seed_ncand, &ctx->seed_ctl, &ctx->seed_ovf1, &ctx->seed_ovf2, &ctx->seed_pool, &ctx->plan_list, &ctx->rec_pool, &ctx->map_cand, &ctx->map_cand_prefix,
                     &ctx->map_plans, &ctx->map_rescue_n, &ctx->map_rescue_prefix, &ctx->map_out_refs, &ctx->map_nout, &ctx->map_flags,
                     &ctx->map_list, &ctx->map_out_prefix, &ctx->map_out_rec, &ctx->km_table, &ctx->km_hist, &ctx->km_flags, &ctx->km_offs,
                     &ctx->km_out, &ctx->ascii, &ctx->ref2, &ctx->reads2, &ctx->reads_irr, &ctx->read_off, &ctx->read_len,
                     &ctx->ascii_offs, &ctx->cand, &ctx->geom, &ctx->caps, &ctx->prefix, &ctx->nmeta, &ctx->meta_prefix, &ctx->meta, &ctx->res, &ctx->rec,
                     &ctx->str_begin, &ctx->ok_len, &ctx->dense_off, &ctx->ws_q, &ctx->ws_t, &ctx->out_q,
                     &ctx->out_t, &ctx->out_ops, &ctx->tb, &ctx->tb_wide, &ctx->tb_pair, &ctx->wide_queue, &ctx->lane_queue, &ctx->lane_resume, &ctx->defer_queue, &ctx->defer_resume, &ctx->scalars,
                     &ctx->order_keys, &ctx->order_keys2, &ctx->order_ids, &ctx->order_queue, &ctx->order_tmp, &ctx->tb_stream,
                     &ctx->piece_flag, &ctx->piece_end, &ctx->chunk_count};
    for (DevBuf *b : all)
        if (b->p) cudaFree(b->p);
    for (auto &e : ctx->chain_events) {
        cudaEventDestroy(e.first);
        cudaEventDestroy(e.second);
    }
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->chunk_done) cudaEventDestroy(ctx->chunk_done);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->post_stream) cudaStreamDestroy(ctx->post_stream);
    if (ctx->post_done) cudaEventDestroy(ctx->post_done);
    if (ctx->chunk_done_ev) cudaEventDestroy(ctx->chunk_done_ev);
    if (ctx->chunk_flag) cudaFreeHost(ctx->chunk_flag);
    if (ctx->side_stream) cudaStreamDestroy(ctx->side_stream);
    if (ctx->side_done) cudaEventDestroy(ctx->side_done);
    if (ctx->in_stream) {
        cudaStreamSynchronize(ctx->in_stream);
        cudaStreamDestroy(ctx->in_stream);
    }
    for (cudaEvent_t e : ctx->piece_events) cudaEventDestroy(e);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

void *ag2_ctx_stream(ag2_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int ag2_host_alloc(size_t bytes, void **ptr)
{
    if (!ptr) return AG2_EINVAL;
    *ptr = nullptr;
    const cudaError_t e = cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocPortable);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return e == cudaErrorMemoryAllocation ? AG2_ENOMEM : AG2_ECUDA;
    }
    return AG2_OK;
}

void ag2_host_free(void *ptr)
{
    if (ptr) cudaFreeHost(ptr);
}

int ag2_ref_load(ag2_ctx *ctx, const char *ref, int64_t ref_len)
{
    if (!ctx || !ref || ref_len <= 0 || ref_len > 0xfffffff0ll) return fail(ctx, AG2_EINVAL, "ag2_ref_load: bad argument");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int64_t groups = (ref_len + 31) >> 5;
    RESERVE(ctx->ref_ascii, (size_t)ref_len + 64);
    RESERVE(ctx->ref2, (size_t)(groups * 2 + 8) * 4);
    RESERVE(ctx->ref_irr, (size_t)(groups + 8) * 4);
    RESERVE(ctx->ref_offs, 4 * 8);
    const int64_t h_offs[4] = {0, ref_len, 0, groups << 5}; // ASCII offsets {0, R}, packed offsets {0, end}
    CK(cudaMemcpyAsync(ctx->ref_ascii.p, ref, (size_t)ref_len, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->ref_offs.p, h_offs, sizeof h_offs, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(ctx->ref2.p, 0, (size_t)(groups * 2 + 8) * 4, st));
    CK(cudaMemsetAsync(ctx->ref_irr.p, 0, (size_t)(groups + 8) * 4, st));
    pack_reads_kernel<<<grid_for(groups, 256, ctx->sm_count), 256, 0, st>>>(
        (const char *)ctx->ref_ascii.p, (const int64_t *)ctx->ref_offs.p, (const int64_t *)ctx->ref_offs.p + 2, 1, 0, groups,
        (uint32_t *)ctx->ref2.p, (uint32_t *)ctx->ref_irr.p, nullptr, nullptr);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st)); // h_offs is on the stack
    ctx->ref_len = ref_len;
    ctx->ref_indexed = false;
    ctx->votes_ready = false;
    return AG2_OK;
}

// ag2_reads_load / ag2_reads_load_async.  The ASCII bases go up in pieces of ~256 MB on the `in_stream`, each followed by
// its pack kernel and an event, so that (a) copies and packing overlap and (b) the asynchronous form lets
// ag2_xdrop_extend_batch start on the first reads while the later ones are still in flight.
static int reads_barrier(ag2_ctx *ctx)
{
    if (ctx->reads_pending) {
        CK(cudaStreamSynchronize(ctx->in_stream));
        ctx->reads_pending = false;
    }
    return AG2_OK;
}

static int reads_load_impl(ag2_ctx *ctx, const char *bases, const int64_t *offs, int64_t n, bool async)
{
    if (!ctx || !bases || !offs || n <= 0 || offs[0] != 0) return fail(ctx, AG2_EINVAL, "ag2_reads_load: bad argument");
    CK(cudaSetDevice(ctx->device));
    int rb = reads_barrier(ctx);
    if (rb != AG2_OK) return rb;
    std::vector<int64_t> &poff = ctx->h_poff;
    std::vector<int32_t> &lens = ctx->h_lens;
    std::vector<int64_t> &pend = ctx->h_piece_end;   // pieces: reads [pend[k-1], pend[k]) = up to piece_bytes of ASCII, at least one read
    poff.resize((size_t)n + 1);
    lens.resize((size_t)n);
    pend.clear();
    const int64_t piece_bytes = ctx->piece_bytes;
    for (int64_t r0 = 0; r0 < n;) {
        int64_t r1 = r0 + 1;
        if (offs[r0] > offs[n] || offs[r0 + 1] < offs[r0]) return fail(ctx, AG2_EINVAL, "ag2_reads_load: offsets must not decrease");
        if (offs[n] - offs[r0] <= piece_bytes) r1 = n;
        else r1 = std::max<int64_t>(r0 + 1, std::upper_bound(offs + r0, offs + n + 1, offs[r0] + piece_bytes) - offs - 1);
        pend.push_back(r1);
        r0 = r1;
    }
    int64_t p = 0;
    size_t pk = 0;
    for (int64_t r = 0; r < n; ++r) {
        const int64_t len = offs[r + 1] - offs[r];
        if (len < 0 || len > 0x7fffffff) return fail(ctx, AG2_EINVAL, "ag2_reads_load: read %ld has bad length", (long)r);
        if (r == pend[pk]) {
            // a piece starts on its own 128-byte lines of both packed arrays (1024 bases of the 1-bit mask), one spare line
            // behind the previous piece: a kernel that already runs on the earlier pieces (streamed run) must never pull a
            // line into L1 that a later pack kernel still has to write
            ++pk;
            p = ((p + 1023) & ~(int64_t)1023) + 1024;
        }
        poff[r] = p;
        lens[r] = (int32_t)len;
        p += (len + 31) & ~(int64_t)31;
    }
    poff[n] = p;
    const int64_t total = offs[n], groups = p >> 5;
    RESERVE(ctx->ascii, (size_t)total + 64);
    RESERVE(ctx->ascii_offs, (size_t)(n + 1) * 8);
    RESERVE(ctx->read_off, (size_t)(n + 1) * 8);
    RESERVE(ctx->read_len, (size_t)n * 4);
    RESERVE(ctx->reads2, (size_t)(groups * 2 + 4) * 4);
    RESERVE(ctx->reads_irr, (size_t)(groups + 4) * 4);
    RESERVE(ctx->piece_flag, pend.size() * 8);   // flags, then the pack kernels' CTA counters
    RESERVE(ctx->piece_end, pend.size() * 8);
    cudaStream_t in = ctx->in_stream;
    CK(cudaMemcpyAsync(ctx->ascii_offs.p, offs, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, in));
    CK(cudaMemcpyAsync(ctx->read_off.p, poff.data(), (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, in));
    CK(cudaMemcpyAsync(ctx->read_len.p, lens.data(), (size_t)n * 4, cudaMemcpyHostToDevice, in));
    CK(cudaMemcpyAsync(ctx->piece_end.p, pend.data(), pend.size() * 8, cudaMemcpyHostToDevice, in));
    CK(cudaMemsetAsync(ctx->piece_flag.p, 0, pend.size() * 8, in));
    CK(cudaStreamSynchronize(in));   // small; `offs` is the caller's and may go away after an asynchronous return
    ctx->pieces.clear();
    int64_t r0 = 0;
    for (size_t k = 0; k < pend.size(); ++k) {
        const int64_t r1 = pend[k];
        const int64_t b0 = offs[r0], b1 = offs[r1], g0 = poff[r0] >> 5, g1 = poff[r1] >> 5;
        if (b1 > b0) CK(cudaMemcpyAsync((char *)ctx->ascii.p + b0, bases + b0, (size_t)(b1 - b0), cudaMemcpyHostToDevice, in));
        if (g1 > g0) {
            pack_reads_kernel<<<grid_for(g1 - g0, 128, ctx->sm_count), 128, 0, in>>>(
                (const char *)ctx->ascii.p, (const int64_t *)ctx->ascii_offs.p, (const int64_t *)ctx->read_off.p, n, g0, g1 - g0,
                (uint32_t *)ctx->reads2.p, (uint32_t *)ctx->reads_irr.p, (unsigned int *)ctx->piece_flag.p + pend.size() + k,
                (int *)ctx->piece_flag.p + k);   // raises flag k: what a running streamed extension waits for
            CK(cudaGetLastError());
        } else {
            CK(cudaMemsetAsync((int *)ctx->piece_flag.p + k, 1, 4, in));   // a piece of empty reads: nothing to pack
        }
        if (k >= ctx->piece_events.size()) {
            cudaEvent_t e;
            CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            ctx->piece_events.push_back(e);
        }
        CK(cudaEventRecord(ctx->piece_events[k], in));
        ctx->pieces.push_back({r1, ctx->piece_events[k]});
        r0 = r1;
    }
    ctx->reads_pending = true;
    ctx->n_reads = n;
    ctx->read_bases = total;
    {   // 90th percentile of a sample of the lengths
        std::vector<int32_t> sample;
        const int64_t step = std::max<int64_t>(1, n / 32768);
        for (int64_t r = 0; r < n; r += step) sample.push_back(lens[r]);
        const size_t k = std::min(sample.size() - 1, sample.size() * 9 / 10);
        std::nth_element(sample.begin(), sample.begin() + k, sample.end());
        ctx->read_len_p90 = sample[k];
    }
    {   // which reads enter the read index: the first <= 100 000 while the running length (+1 each) < 1e9 (:277)
        int64_t lenl = 0, kk = 0;
        while (kk < n && kk < 100000 && lenl < 1000000000ll) {
            lenl += (offs[kk + 1] - offs[kk]) + 1;
            ++kk;
        }
        ctx->read_prefix_len = offs[kk];
    }
    return async ? AG2_OK : reads_barrier(ctx);
}

int ag2_reads_load(ag2_ctx *ctx, const char *bases, const int64_t *offs, int64_t n) { return reads_load_impl(ctx, bases, offs, n, false); }
int ag2_reads_load_async(ag2_ctx *ctx, const char *bases, const int64_t *offs, int64_t n) { return reads_load_impl(ctx, bases, offs, n, true); }
int ag2_reads_wait(ag2_ctx *ctx)
{
    if (!ctx) return AG2_EINVAL;
    CK(cudaSetDevice(ctx->device));
    return reads_barrier(ctx);
}

int ag2_extend_upload(ag2_ctx *ctx, const ag2_candidate *cand, int64_t n)
{
    if (!ctx || !cand || n <= 0 || n > 0x3fffffff) return fail(ctx, AG2_EINVAL, "ag2_extend_upload: bad argument");
    if (!ctx->ref_len || !ctx->n_reads) return fail(ctx, AG2_ESTATE, "ag2_extend_upload: load reference and reads first");
    CK(cudaSetDevice(ctx->device));
    RESERVE(ctx->cand, (size_t)n * sizeof(Candidate));
    CK(cudaMemcpyAsync(ctx->cand.p, cand, (size_t)n * sizeof(Candidate), cudaMemcpyHostToDevice, ctx->stream));
    ctx->n_cand = n;
    ctx->ran = false;
    return AG2_OK;
}

// extend_candidate over n device-resident candidates.  Records go to d_rec[0..n); the strings of the ok records are
// appended to the dense string pool at dense_base (which this returns advanced), earlier contents are kept.
struct HostSink { // caller-owned host buffers that receive each chunk's results while the next chunk computes
    ag2_record *rec = nullptr;
    char *q = nullptr, *t = nullptr;
    int64_t cap = 0;
    uint32_t *ops = nullptr;   // packed form: 2-bit ops instead of the two strings (cap = columns); chunks start on word boundaries
    bool overflow = false;
};

// A chunk's share of the dense pool [base, base + total) goes home: both strings, or their 2-bit ops.
static int sink_copy_chunk(ag2_ctx *ctx, HostSink *sink, int64_t dense_base, int64_t chunk_total, cudaStream_t compute, cudaStream_t copy,
                           cudaEvent_t ev, int *launches)
{
    if (sink->ops) {
        if (dense_base + chunk_total > sink->cap) {
            sink->overflow = true;
            return AG2_OK;
        }
        const int64_t w_lo = dense_base >> 4, w_hi = (dense_base + chunk_total + 15) >> 4;   // out_ops was sized by the caller
        if (w_hi > w_lo) {
            pack_ops_kernel<<<grid_for(w_hi - w_lo, 128, ctx->sm_count), 128, 0, compute>>>((const char *)ctx->out_q.p, (const char *)ctx->out_t.p, w_lo, w_hi,
                                                                                         dense_base + chunk_total, (uint32_t *)ctx->out_ops.p);
            CK(cudaGetLastError());
            ++*launches;
        }
        CK(cudaEventRecord(ev, compute));
        CK(cudaStreamWaitEvent(copy, ev, 0));
        if (w_hi > w_lo)
            CK(cudaMemcpyAsync(sink->ops + w_lo, (uint32_t *)ctx->out_ops.p + w_lo, (size_t)(w_hi - w_lo) * 4, cudaMemcpyDeviceToHost, copy));
        return AG2_OK;
    }
    CK(cudaEventRecord(ev, compute));
    CK(cudaStreamWaitEvent(copy, ev, 0));
    if (sink->q && sink->t && dense_base + chunk_total <= sink->cap) {
        CK(cudaMemcpyAsync(sink->q + dense_base, (char *)ctx->out_q.p + dense_base, (size_t)chunk_total, cudaMemcpyDeviceToHost, copy));
        CK(cudaMemcpyAsync(sink->t + dense_base, (char *)ctx->out_t.p + dense_base, (size_t)chunk_total, cudaMemcpyDeviceToHost, copy));
    } else if (sink->q || sink->t) {
        sink->overflow = true;
    }
    return AG2_OK;
}

static int extend_batch(ag2_ctx *ctx, const Candidate *d_cand, int64_t n, Record *d_rec, int64_t dense_base_in, int64_t *dense_base_out,
                        bool fresh_stats, HostSink *sink = nullptr, const ag2_candidate *h_cand = nullptr)
{
    cudaStream_t st = ctx->stream;
    int launches = 0;
    // AG2_TRACE: host clock at the points where this function waits for the stream anyway (no extra synchronisation)
    static const bool trace = getenv("AG2_TRACE") != nullptr;
    const auto t_begin = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (trace) fprintf(stderr, "[ag2 trace] extend_batch %-28s %8.3f ms\n", what,
                           std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count());
    };
    if (!h_cand) { // without the host's copy of the candidates nothing says which reads a chunk needs: wait for all of them
        int rb = reads_barrier(ctx);
        if (rb != AG2_OK) return rb;
    }
    RESERVE(ctx->geom, (size_t)n * sizeof(ExtGeom));
    RESERVE(ctx->caps, (size_t)n * 8);
    RESERVE(ctx->prefix, (size_t)(n + 1) * 8);
    RESERVE(ctx->nmeta, (size_t)n * 8);
    RESERVE(ctx->meta_prefix, (size_t)(n + 1) * 8);
    RESERVE(ctx->res, (size_t)n * 2 * sizeof(ChainResult));
    RESERVE(ctx->str_begin, (size_t)n * 8);
    RESERVE(ctx->ok_len, (size_t)n * 8);
    RESERVE(ctx->dense_off, (size_t)(n + 1) * 8);
    RESERVE(ctx->wide_queue, (size_t)n * 2 * 4);
    RESERVE(ctx->lane_queue, ((size_t)n * 2 + 1024) * 4);   // + one unpublished ticket per consumer warp
    RESERVE(ctx->lane_resume, (size_t)n * 2 * sizeof(LaneResume));
    // Opt-in experiment (AG2_DEFER=1): long last blocks set aside and run together at the end of the launch.  Measured on
    // configs[1]: no gain (pair kernel 385 -> 389 ms, step +17 ms) -- the longest-first queue already gives a warp
    // directions of equal length, so their last blocks fall into the same round (profiles/defer_r02.md).
    static const bool defer_on = getenv("AG2_DEFER") != nullptr;
    if (defer_on) {
        RESERVE(ctx->defer_queue, (size_t)n * 2 * 4);
        RESERVE(ctx->defer_resume, (size_t)n * 2 * sizeof(LaneResume));
    }

    // pair kernel: the band window of 128 directions per CTA in shared memory
    int pocc = 0;
    const size_t pair_smem = sizeof(PairSmem);
    CK(cudaFuncSetAttribute(xdrop_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pair_smem));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pocc, xdrop_pair_kernel, kPairThreads, pair_smem));
    if (pocc < 1) pocc = 1;
    const int pair_grid = ctx->sm_count * pocc;
    RESERVE(ctx->tb_pair, kPairCtaScratch * (size_t)pair_grid);
    // lane kernel: as many CTAs per SM as shared memory allows (band + target block per thread)
    int occ = 0;
    const size_t lane_smem = sizeof(LaneSmem);
    CK(cudaFuncSetAttribute(xdrop_lane_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lane_smem));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, xdrop_lane_kernel, kLaneThreads, lane_smem));
    if (occ < 1) occ = 1;
    const int lane_grid = ctx->sm_count * occ;
    RESERVE(ctx->tb, (size_t)kLaneScratch * lane_grid * kLaneThreads);
    int wocc = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&wocc, xdrop_chains_kernel<kWideK, kWideWarps>, kWideWarps * 32, 0));
    if (wocc < 1) wocc = 1;
    if (wocc > 8) wocc = 8;
    const int wide_grid = ctx->sm_count * wocc;
    // what the pair kernel hands over goes straight to the wide kernel (one WARP per direction, restarted from its first
    // block: a few ms of latency) unless there is more of it than the wide grid absorbs in a few rounds; then the lane
    // kernel (one THREAD per direction, resumed at the failing block) has the higher throughput
    const unsigned lane_threshold = (unsigned)wide_grid * kWideWarps * 4;
    const size_t tbw_stride = (size_t)(kMaxBlk + 2) * TbLayout<kWideK>::kRowBytes;
    RESERVE(ctx->tb_wide, tbw_stride * wide_grid * kWideWarps);
    // same shared-memory carve-out as the pair kernel, or an SM that holds a consumer CTA could not take pair CTAs at all
    CK(cudaFuncSetAttribute(xdrop_stream_kernel<kWideK, kWideWarps>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CK(cudaFuncSetAttribute(xdrop_pair_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    int stream_want = 16;
    if (const char *e = getenv("AG2_STREAM_GRID")) stream_want = atoi(e);   // tuning knob
    // 0 = no concurrent consumer, everything handed over waits for the post-pass (needed under ncu, which serialises kernels:
    // a consumer launched first would poll for producers that never start)
    const int stream_grid = stream_want <= 0 ? 0 : std::max(1, std::min(std::min(48, stream_want), ctx->sm_count / 3));
    static_assert(48 * kWideWarps <= 1024, "the hand-over queue keeps one spare entry per consumer warp");   // consumer CTAs beside the pair kernel: each takes one pair CTA's place
    RESERVE(ctx->tb_stream, tbw_stride * std::max(1, stream_grid) * kWideWarps);

    const PackedSeqs sq = seqs_of(ctx);
    Scalars *sc = (Scalars *)ctx->scalars.p;
    if (fresh_stats) {
        CK(cudaMemsetAsync(sc, 0, sizeof(Scalars), st));
        ctx->stats_lane_chains = 0;
        ctx->stats_direct_wide = 0;
    }
    extend_setup_kernel<<<grid_for(n, 256, ctx->sm_count), 256, 0, st>>>(d_cand, n, sq, ctx->n_reads,
                                                                         (ExtGeom *)ctx->geom.p, (int64_t *)ctx->caps.p,
                                                                         (int64_t *)ctx->nmeta.p);
    exclusive_scan_i64<<<1, 1024, 0, st>>>((const int64_t *)ctx->caps.p, n, (int64_t *)ctx->prefix.p);
    exclusive_scan_i64<<<1, 1024, 0, st>>>((const int64_t *)ctx->nmeta.p, n, (int64_t *)ctx->meta_prefix.p);
    launches += 3;
    CK(cudaGetLastError());
    ctx->h_prefix.resize((size_t)n + 1);
    ctx->h_meta_prefix.resize((size_t)n + 1);
    CK(cudaMemcpyAsync(ctx->h_prefix.data(), ctx->prefix.p, (size_t)(n + 1) * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(ctx->h_meta_prefix.data(), ctx->meta_prefix.p, (size_t)(n + 1) * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    lap("setup, scans, prefixes home");
    const std::vector<int64_t> &pf = ctx->h_prefix, &mf = ctx->h_meta_prefix;

    // candidates are processed in chunks whose workspace strings fit ws_limit
    size_t max_chunk = 0, max_meta = 0;
    std::vector<std::pair<int64_t, int64_t>> chunks;
    {   // as few chunks as the limit allows, of equal size (a small last chunk would still cost a whole launch)
        const size_t limit = std::max<size_t>(1, sink ? ctx->ws_call : ctx->ws_limit);
        const int64_t nch = std::max<int64_t>(1, (int64_t)(((size_t)pf[n] + limit - 1) / limit));
        for (int64_t k = 0, lo = 0; lo < n; ++k) {
            int64_t hi = n;
            if (k + 1 < nch) {
                const int64_t end = (int64_t)((double)pf[n] * (double)(k + 1) / (double)nch);
                hi = std::upper_bound(pf.begin() + lo + 1, pf.begin() + n + 1, end) - pf.begin() - 1;
                hi = std::min<int64_t>(n, std::max<int64_t>(lo + 1, hi));
            }
            chunks.push_back({lo, hi});
            max_chunk = std::max(max_chunk, (size_t)(pf[hi] - pf[lo]));
            max_meta = std::max(max_meta, (size_t)(mf[hi] - mf[lo]));
            lo = hi;
        }
    }
    RESERVE(ctx->ws_q, max_chunk + 64);
    RESERVE(ctx->ws_t, max_chunk + 64);
    RESERVE(ctx->meta, (max_meta + 16) * 4);
    int64_t max_cn = 0;
    for (auto &c : chunks) max_cn = std::max(max_cn, c.second - c.first);
    size_t order_tmp_bytes = 0;
    CK(cub::DeviceRadixSort::SortPairsDescending(nullptr, order_tmp_bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr,
                                                 (const int32_t *)nullptr, (int32_t *)nullptr, (int)(2 * max_cn), 0, 24, st));
    RESERVE(ctx->order_keys, (size_t)max_cn * 2 * 4 + 16);
    RESERVE(ctx->order_keys2, (size_t)max_cn * 2 * 4 + 16);
    RESERVE(ctx->order_ids, (size_t)max_cn * 2 * 4 + 16);
    RESERVE(ctx->order_queue, (size_t)max_cn * 2 * 4 + 16);
    RESERVE(ctx->order_tmp, order_tmp_bytes + 16);
    // upper bound of the dense strings: every column consumes a base of the read or of its window
    {
        const size_t pad = 64 + 16 * chunks.size();   // packed output rounds every chunk up to 16 columns
        int rk = reserve_keep(ctx, ctx->out_q, (size_t)(dense_base_in + pf[n]) + pad, (size_t)dense_base_in);
        if (rk != AG2_OK) return rk;
        rk = reserve_keep(ctx, ctx->out_t, (size_t)(dense_base_in + pf[n]) + pad, (size_t)dense_base_in);
        if (rk != AG2_OK) return rk;
    }
    // packed output: one word per 16 columns, every chunk rounded up to a word
    if (sink && sink->ops) RESERVE(ctx->out_ops, ((size_t)(dense_base_in + pf[n]) / 16 + chunks.size() + 8) * 4);
    while (ctx->chain_events.size() < chunks.size()) {
        cudaEvent_t a, b;
        CK(cudaEventCreate(&a));
        CK(cudaEventCreate(&b));
        ctx->chain_events.push_back({a, b});
    }

    // per chunk: lane kernel -> (wide rerun) -> finalize -> scan of the ok lengths -> assemble into the
    // dense output at the running base
    int64_t dense_base = dense_base_in;
    for (size_t ci = 0; ci < chunks.size(); ++ci) {
        const int64_t lo = chunks[ci].first, cn = chunks[ci].second - chunks[ci].first;
        if (ctx->reads_pending && h_cand) { // an asynchronous read load is in flight: this chunk waits for the piece that holds its last read
            int64_t last = -1;
            for (int64_t i = lo; i < lo + cn; ++i) last = std::max<int64_t>(last, h_cand[i].read);
            for (const auto &pc : ctx->pieces)
                if (pc.end_read > last || &pc == &ctx->pieces.back()) {
                    CK(cudaStreamWaitEvent(st, pc.ready, 0));
                    break;
                }
        }
        set_slots_kernel<<<grid_for(cn, 256, ctx->sm_count), 256, 0, st>>>((ExtGeom *)ctx->geom.p, (const int64_t *)ctx->prefix.p,
                                                                           (const int64_t *)ctx->meta_prefix.p, lo, cn);
        CK(cudaMemsetAsync(&sc->next_fast, 0, 4 * sizeof(unsigned long long) + 10 * sizeof(unsigned int), st));
        CK(cudaMemsetAsync(ctx->lane_queue.p, 0xff, ((size_t)cn * 2 + 1024) * 4, st));   // -1 = not published
        LaneArgs a = {};
        a.seqs = sq;
        a.cand = d_cand + lo;
        a.geom = (const ExtGeom *)ctx->geom.p + lo;
        a.res = (ChainResult *)ctx->res.p + 2 * lo;
        a.meta = (uint32_t *)ctx->meta.p;
        a.ws_q = (char *)ctx->ws_q.p;
        a.ws_t = (char *)ctx->ws_t.p;
        a.counters = &sc->ctr;
        // A handful of candidates (rescue extensions, the second pass of a few reads): a direction is a chain of ~20
        // sequentially dependent block DPs, and on the pair kernel, built for throughput, a lone chain takes ~25 ms.  One
        // WARP per direction (row-parallel kernel) brings that to a few ms.
        const bool small_batch = 2 * cn <= (int64_t)wide_grid * kWideWarps * 4 && !getenv("AG2_NO_SMALL_BATCH");   // up to four waves
        if (small_batch) {
            ChainArgs w = {};
            w.seqs = sq;
            w.cand = a.cand;
            w.geom = a.geom;
            w.res = a.res;
            w.ws_q = a.ws_q;
            w.ws_t = a.ws_t;
            w.tb = (uint8_t *)ctx->tb_wide.p;
            w.tb_stride = tbw_stride;
            w.n_chains = 2 * cn;
            w.queue = nullptr;
            w.next = &sc->next_post;
            w.counters = &sc->ctr;
            CK(cudaEventRecord(ctx->chain_events[ci].first, st));
            CK(cudaEventRecord(ctx->chain_events[ci].second, st));
            // first with 4 columns per lane (a 128-column window: ~5 x shorter rows than the 736-column form); the directions
            // whose band leaves that window are rerun by the full-width form
            w.wide_queue = (int32_t *)ctx->wide_queue.p;
            w.wide_count = &sc->wide_count;
            const int sgrid = (int)std::min<int64_t>(wide_grid, (2 * cn + kWideWarps - 1) / kWideWarps);
            xdrop_chains_kernel<kNarrowK, kWideWarps><<<sgrid, kWideWarps * 32, 0, st>>>(w);
            CK(cudaGetLastError());
            ++launches;
            unsigned int n_again = 0;
            CK(cudaMemcpyAsync(&n_again, &sc->wide_count, sizeof n_again, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            if (n_again > 0) {
                w.n_chains = n_again;
                w.queue = (const int32_t *)ctx->wide_queue.p;
                w.next = &sc->next_wide;
                w.wide_queue = nullptr;
                w.wide_count = nullptr;
                xdrop_chains_kernel<kWideK, kWideWarps><<<wide_grid, kWideWarps * 32, 0, st>>>(w);
                CK(cudaGetLastError());
                ++launches;
            }
            ctx->stats_direct_wide += 2 * cn;
        }
        // pair kernel over all directions of the chunk; what it hands over goes to the lane queue
        LaneArgs pa = a;
        if (!small_batch) {
            pa.scratch = (uint8_t *)ctx->tb_pair.p;
            pa.n_chains = 2 * cn;
            // longest directions first (chain_keys_kernel): no long direction is left running alone when the queue runs dry
            chain_keys_kernel<<<grid_for(2 * cn, 256, ctx->sm_count), 256, 0, st>>>((const ExtGeom *)ctx->geom.p + lo, 2 * cn,
                                                                                    (uint32_t *)ctx->order_keys.p, (int32_t *)ctx->order_ids.p);
            {
                size_t tmp = ctx->order_tmp.cap;
                CK(cub::DeviceRadixSort::SortPairsDescending(ctx->order_tmp.p, tmp, (const uint32_t *)ctx->order_keys.p, (uint32_t *)ctx->order_keys2.p,
                                                             (const int32_t *)ctx->order_ids.p, (int32_t *)ctx->order_queue.p, (int)(2 * cn), 0, 24, st));
            }
            launches += 2;
            pa.queue = (const int32_t *)ctx->order_queue.p;
            pa.next = &sc->next_pair;
            pa.wide_queue = (int32_t *)ctx->lane_queue.p;
            pa.wide_count = &sc->lane_count;
            pa.resume = (LaneResume *)ctx->lane_resume.p;
            pa.done_ctas = &sc->pair_done;
            if (defer_on) {   // long last blocks are set aside and run together at the end of the launch (xdrop_pair.cuh)
                CK(cudaMemsetAsync(ctx->defer_queue.p, 0xff, (size_t)cn * 2 * 4, st));
                pa.defer_queue = (int32_t *)ctx->defer_queue.p;
                pa.defer_resume = (LaneResume *)ctx->defer_resume.p;
                pa.defer_ctl = sc->defer_ctl;
            }
            // the consumer of the hand-overs (xdrop_stream_kernel) goes first, on its own high-priority stream, once the chunk's
            // inputs are in place: it must be resident before the pair kernel fills the SMs
            ChainArgs cw = {};
            cw.seqs = sq;
            cw.cand = a.cand;
            cw.geom = a.geom;
            cw.res = a.res;
            cw.ws_q = a.ws_q;
            cw.ws_t = a.ws_t;
            cw.tb = (uint8_t *)ctx->tb_stream.p;
            cw.tb_stride = tbw_stride;
            cw.queue = (const int32_t *)ctx->lane_queue.p;
            cw.next = &sc->next_wide;
            cw.counters = &sc->ctr;
            cw.try_narrow = 1;
            cw.resume = (const LaneResume *)ctx->lane_resume.p;   // hand-overs are continued at the block they stopped at
            cw.meta = a.meta;
            CK(cudaStreamSynchronize(st));
            lap("chunk inputs in place");
            if (stream_grid > 0) xdrop_stream_kernel<kWideK, kWideWarps><<<stream_grid, kWideWarps * 32, 0, ctx->side_stream>>>(cw, &sc->pair_done, (unsigned)pair_grid);
            CK(cudaGetLastError());
            CK(cudaEventRecord(ctx->side_done, ctx->side_stream));
            CK(cudaEventRecord(ctx->chain_events[ci].first, st));
            xdrop_pair_kernel<<<pair_grid, kPairThreads, pair_smem, st>>>(pa);
            CK(cudaEventRecord(ctx->chain_events[ci].second, st));
            CK(cudaGetLastError());
            CK(cudaStreamWaitEvent(st, ctx->side_done, 0));
            unsigned int n_handed = 0, n_wide = 0;
            unsigned long long taken = 0;
            CK(cudaMemcpyAsync(&n_handed, &sc->lane_count, sizeof n_handed, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(&taken, &sc->next_wide, sizeof taken, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            lap("pair kernel + consumer done");
            launches += 3;
            ctx->stats_lane_chains += n_handed;
            // [first, n_handed) was published after the consumer's last ticket: few -> wide kernel, many -> lane kernel
            const unsigned int first = (unsigned int)std::min<unsigned long long>(taken, n_handed);
            const unsigned int n_lane = n_handed - first;
            ctx->stats_direct_wide += first;
            const bool lane_path = n_lane > lane_threshold;
            const int32_t *post_queue = (const int32_t *)ctx->lane_queue.p + first;
            if (n_lane > 0 && !lane_path) {
                n_wide = n_lane;
                ctx->stats_direct_wide += n_lane;
            }
            if (n_lane > 0 && lane_path) {
                // one THREAD per direction, resumed at the block the pair window could not hold
                a.scratch = (uint8_t *)ctx->tb.p;
                a.n_chains = n_lane;
                a.queue = post_queue;
                a.resume = (LaneResume *)ctx->lane_resume.p + first;
                a.next = &sc->next_fast;
                a.wide_queue = (int32_t *)ctx->wide_queue.p;
                a.wide_count = &sc->wide_count;
                xdrop_lane_kernel<<<lane_grid, kLaneThreads, lane_smem, st>>>(a);
                CK(cudaGetLastError());
                // directions that left the lane path too (band > 120 columns, or reservation exceeded): wide kernel
                CK(cudaMemcpyAsync(&n_wide, &sc->wide_count, sizeof n_wide, cudaMemcpyDeviceToHost, st));
                CK(cudaStreamSynchronize(st));
                post_queue = (const int32_t *)ctx->wide_queue.p;
                ++launches;
            }
            if (n_wide > 0) {
                ChainArgs w = {};
                w.seqs = sq;
                w.cand = a.cand;
                w.geom = a.geom;
                w.res = a.res;
                w.ws_q = a.ws_q;
                w.ws_t = a.ws_t;
                w.tb = (uint8_t *)ctx->tb_wide.p;
                w.tb_stride = tbw_stride;
                w.n_chains = n_wide;
                w.queue = post_queue;
                w.next = &sc->next_post;
                w.wide_queue = nullptr;
                w.wide_count = nullptr;
                w.counters = &sc->ctr;
                w.try_narrow = 1;
                if (!lane_path) {   // straight from the hand-over queue: the states are there
                    w.resume = (const LaneResume *)ctx->lane_resume.p + first;
                    w.meta = a.meta;
                }
                xdrop_chains_kernel<kWideK, kWideWarps><<<wide_grid, kWideWarps * 32, 0, st>>>(w);
                CK(cudaGetLastError());
                ++launches;
            }

        }
        extend_finalize_kernel<<<grid_for(cn, 256, ctx->sm_count), 256, 0, st>>>(
            d_cand, (const ExtGeom *)ctx->geom.p, (const ChainResult *)ctx->res.p,
            (const int32_t *)ctx->read_len.p, lo, cn, d_rec, (int64_t *)ctx->str_begin.p,
            (int64_t *)ctx->ok_len.p);
        exclusive_scan_i64<<<1, 1024, 0, st>>>((const int64_t *)ctx->ok_len.p + lo, cn, (int64_t *)ctx->dense_off.p + lo);
        launches += 2;
        int64_t chunk_total = 0;
        CK(cudaMemcpyAsync(&chunk_total, (int64_t *)ctx->dense_off.p + lo + cn, 8, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        lap("post-pass, finalize, scan");
        assemble_kernel<<<grid_for(cn * 32, 256, ctx->sm_count), 256, 0, st>>>(
            d_rec, (const ExtGeom *)ctx->geom.p, (const ChainResult *)ctx->res.p, (const uint32_t *)ctx->meta.p,
            (const int64_t *)ctx->dense_off.p, dense_base, lo, cn, (const char *)ctx->ws_q.p, (const char *)ctx->ws_t.p,
            (char *)ctx->out_q.p, (char *)ctx->out_t.p, &sc->aligned, &sc->columns);
        ++launches;
        CK(cudaGetLastError());
        if (sink) { // this chunk's records and strings (or their ops) go home on the copy stream while the next chunk computes
            const int rc = sink_copy_chunk(ctx, sink, dense_base, chunk_total, st, ctx->copy_stream, ctx->chunk_done, &launches);
            if (rc != AG2_OK) return rc;
            CK(cudaMemcpyAsync(sink->rec + lo, d_rec + lo, (size_t)cn * sizeof(Record), cudaMemcpyDeviceToHost, ctx->copy_stream));
        }
        dense_base += chunk_total;
        if (sink && sink->ops) dense_base = (dense_base + 15) & ~(int64_t)15;   // the next chunk's ops start on a word of their own
    }
    CK(cudaStreamSynchronize(st));
    if (sink) CK(cudaStreamSynchronize(ctx->copy_stream));
    lap("assemble (and copies home)");
    *dense_base_out = dense_base;

    Scalars hs;
    CK(cudaMemcpy(&hs, sc, sizeof hs, cudaMemcpyDeviceToHost));
    ag2_extend_stats &s = ctx->stats;
    s.cells = (int64_t)hs.ctr.cells;
    s.rows = (int64_t)hs.ctr.rows;
    s.blocks = (int64_t)hs.ctr.blocks;
    s.interior = (int64_t)hs.ctr.interior;
    s.slots = (int64_t)hs.ctr.slots;
    s.wide_chains = (int64_t)hs.ctr.wide + ctx->stats_direct_wide;
    s.aligned = (int64_t)hs.aligned;
    s.columns = (int64_t)hs.columns;
    s.launches = (fresh_stats ? 0 : s.launches) + launches;
    if (fresh_stats) s.kernel_ms = 0;
    s.lane_chains = ctx->stats_lane_chains;
    for (size_t ci = 0; ci < chunks.size(); ++ci) {
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, ctx->chain_events[ci].first, ctx->chain_events[ci].second));
        s.kernel_ms += ms;
    }
    return AG2_OK;
}

// Which form ag2_xdrop_extend_batch takes when AG2_E2E_PATH does not say (see there).
constexpr bool kStreamedByDefault = true;
constexpr int kStreamPairCtasPerSm = 6;   // streamed form: pair CTAs per SM (8 fit), see extend_batch_streamed
// extend_batch_streamed asks for the chunked form instead (positive: not an error of the call)
constexpr int kStreamStalled = 1;         // in-kernel waits for the reads ran into their limit, results void
constexpr int kStreamTooBig = 2;          // the whole batch's workspace does not fit beside what is resident; nothing ran

// extend_candidate over n device-resident candidates, results streamed to the caller's host buffers (`sink`).
// ONE launch of the pair kernel covers every direction (a direction is a chain of sequentially dependent block DPs: cutting
// the batch into separately launched chunks would make each launch as long as its longest direction).  The candidates are
// cut into output chunks instead: the queue order (stream_keys_kernel) finishes them one after the other, the kernel raises
// a host-mapped flag per finished chunk, and this thread finalises, assembles and copies that chunk home on side streams
// while the kernel goes on.  Reads still on their way up (ag2_reads_load_async) are waited for inside the kernel, per piece.
static int extend_batch_streamed(ag2_ctx *ctx, const Candidate *d_cand, int64_t n, Record *d_rec, int64_t *dense_total, HostSink *sink)
{
    cudaStream_t st = ctx->stream, post = ctx->post_stream;
    int launches = 0;
    static const bool trace = getenv("AG2_TRACE") != nullptr;   // per-chunk host timeline on stderr
    const auto t_begin = std::chrono::steady_clock::now();
    auto now_ms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count(); };
    std::vector<double> tr;
    RESERVE(ctx->geom, (size_t)n * sizeof(ExtGeom));
    RESERVE(ctx->caps, (size_t)n * 8);
    RESERVE(ctx->prefix, (size_t)(n + 1) * 8);
    RESERVE(ctx->nmeta, (size_t)n * 8);
    RESERVE(ctx->meta_prefix, (size_t)(n + 1) * 8);
    RESERVE(ctx->res, (size_t)n * 2 * sizeof(ChainResult));
    RESERVE(ctx->str_begin, (size_t)n * 8);
    RESERVE(ctx->ok_len, (size_t)n * 8);
    RESERVE(ctx->dense_off, (size_t)(n + 2 * kMaxStreamChunks + 1) * 8);
    RESERVE(ctx->wide_queue, (size_t)n * 2 * 4);
    RESERVE(ctx->lane_queue, ((size_t)n * 2 + 1024) * 4);
    RESERVE(ctx->lane_resume, (size_t)n * 2 * sizeof(LaneResume));
    RESERVE(ctx->chunk_count, (size_t)kMaxStreamChunks * 4);

    // The small kernels of this form (packing of the pieces still arriving, finalize / scan / assemble of the finished
    // chunks) must run BESIDE the resident pair kernel.  At its full occupancy they did not get onto the SMs before pair
    // CTAs retired (configs[1]: the first chunk's finalize waited 250 ms, in-kernel waits for the pack kernels ran into their
    // limit), so this launch asks for more dynamic shared memory than it uses until only kStreamPairCtasPerSm fit: every SM
    // keeps a free slot for them.  Registers are not what is missing: with the 96-register build of the kernel
    // (AG2_STREAM_LEAN_KERNEL=1) all 8 CTAs fit beside 16 K free registers, and one run in eight still stalled.
    // An SM has ONE shared-memory carve-out at a time: a kernel whose launch picks another one cannot join the pair CTAs on
    // their SMs at all and waits for an idle SM -- which a persistent launch does not leave.  The small kernels ask for the
    // pair kernel's carve-out, as the consumer kernel does.
    CK(cudaFuncSetAttribute(pack_reads_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CK(cudaFuncSetAttribute(pack_ops_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CK(cudaFuncSetAttribute(extend_finalize_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CK(cudaFuncSetAttribute(scan_chunk_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CK(cudaFuncSetAttribute(assemble_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    int pocc = 0, want_occ = kStreamPairCtasPerSm;
    if (const char *e = getenv("AG2_STREAM_CTAS_PER_SM")) want_occ = std::max(1, atoi(e));   // tuning knob
    static const bool lean_kernel = getenv("AG2_STREAM_LEAN_KERNEL") != nullptr;            // tuning knob: the 96-register build
    auto pair_kernel = lean_kernel ? xdrop_pair_kernel_lean : xdrop_pair_kernel;
    size_t pair_smem = sizeof(PairSmem);
    CK(cudaFuncSetAttribute(pair_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    for (;;) {
        CK(cudaFuncSetAttribute(pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pair_smem));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pocc, pair_kernel, kPairThreads, pair_smem));
        if (pocc <= want_occ || pair_smem + 1024 > (size_t)200 * 1024) break;
        pair_smem += 1024;
    }
    if (pocc < 1) pocc = 1;
    const int pair_full = ctx->sm_count * pocc;
    RESERVE(ctx->tb_pair, kPairCtaScratch * (size_t)pair_full);
    int occ = 0;
    const size_t lane_smem = sizeof(LaneSmem);
    CK(cudaFuncSetAttribute(xdrop_lane_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lane_smem));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, xdrop_lane_kernel, kLaneThreads, lane_smem));
    if (occ < 1) occ = 1;
    const int lane_grid = ctx->sm_count * occ;
    RESERVE(ctx->tb, (size_t)kLaneScratch * lane_grid * kLaneThreads);
    int wocc = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&wocc, xdrop_chains_kernel<kWideK, kWideWarps>, kWideWarps * 32, 0));
    wocc = std::max(1, std::min(wocc, 8));
    const int wide_grid = ctx->sm_count * wocc;
    const unsigned lane_threshold = (unsigned)wide_grid * kWideWarps * 4;
    const size_t tbw_stride = (size_t)(kMaxBlk + 2) * TbLayout<kWideK>::kRowBytes;
    RESERVE(ctx->tb_wide, tbw_stride * wide_grid * kWideWarps);
    CK(cudaFuncSetAttribute(xdrop_stream_kernel<kWideK, kWideWarps>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    int stream_want = 16;
    if (const char *e = getenv("AG2_STREAM_GRID")) stream_want = atoi(e);
    const int stream_grid = stream_want <= 0 ? 0 : std::max(1, std::min(std::min(48, stream_want), ctx->sm_count / 3));
    RESERVE(ctx->tb_stream, tbw_stride * std::max(1, stream_grid) * kWideWarps);
    // every pair CTA resident from the start (a consumer CTA takes one pair CTA's place): nothing of this launch is left
    // pending in front of the small kernels that have to run beside it
    // AG2_STREAM_FREE_CTAS (tuning knob): CTAs left out of the grid on top of the consumer's, so that as many SMs keep a free
    // slot for the small kernels without every SM giving one up (with AG2_STREAM_CTAS_PER_SM=7)
    int free_ctas = 0;
    if (const char *e = getenv("AG2_STREAM_FREE_CTAS")) free_ctas = std::max(0, atoi(e));
    const int pair_grid = std::max(1, pair_full - stream_grid - free_ctas);

    const PackedSeqs sq = seqs_of(ctx);
    Scalars *sc = (Scalars *)ctx->scalars.p;
    CK(cudaMemsetAsync(sc, 0, sizeof(Scalars), st));
    ctx->stats_lane_chains = 0;
    ctx->stats_direct_wide = 0;
    extend_setup_kernel<<<grid_for(n, 256, ctx->sm_count), 256, 0, st>>>(d_cand, n, sq, ctx->n_reads, (ExtGeom *)ctx->geom.p,
                                                                         (int64_t *)ctx->caps.p, (int64_t *)ctx->nmeta.p);
    exclusive_scan_i64<<<1, 1024, 0, st>>>((const int64_t *)ctx->caps.p, n, (int64_t *)ctx->prefix.p);
    exclusive_scan_i64<<<1, 1024, 0, st>>>((const int64_t *)ctx->nmeta.p, n, (int64_t *)ctx->meta_prefix.p);
    launches += 3;
    CK(cudaGetLastError());
    int64_t ws_total = 0, meta_total = 0;
    CK(cudaMemcpyAsync(&ws_total, (int64_t *)ctx->prefix.p + n, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&meta_total, (int64_t *)ctx->meta_prefix.p + n, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    {   // one launch needs the workspace strings of the WHOLE batch (the chunked form only those of a chunk)
        size_t free_b = 0, total_b = 0, grow = 0;
        CK(cudaMemGetInfo(&free_b, &total_b));
        for (const DevBuf *b : {&ctx->ws_q, &ctx->ws_t})
            if (b->cap < (size_t)ws_total + 64) grow += (size_t)ws_total + 64;
        if (grow > 0 && grow + ((size_t)meta_total + 16) * 4 > free_b - std::min(free_b, (size_t)1 << 30)) return kStreamTooBig;
    }
    RESERVE(ctx->ws_q, (size_t)ws_total + 64);
    RESERVE(ctx->ws_t, (size_t)ws_total + 64);
    RESERVE(ctx->meta, ((size_t)meta_total + 16) * 4);
    RESERVE(ctx->out_q, (size_t)ws_total + 64 + 16 * kMaxStreamChunks);   // upper bound of the dense strings: every column consumes a base of the read or of its window
    RESERVE(ctx->out_t, (size_t)ws_total + 64 + 16 * kMaxStreamChunks);
    if (sink->ops) RESERVE(ctx->out_ops, ((size_t)ws_total / 16 + kMaxStreamChunks + 8) * 4);

    // output chunks: equal candidate counts, about ws_limit_streamed bytes of workspace string each
    int64_t want = (int64_t)std::min<size_t>(kMaxStreamChunks, std::max<size_t>(1, ((size_t)ws_total + ctx->ws_call - 1) / ctx->ws_call));
    const int64_t chunk_cn = std::max<int64_t>(1, (n + want - 1) / want);
    const int n_chunks = (int)((n + chunk_cn - 1) / chunk_cn);
    if (chunk_cn > 0x7fffffff) return fail(ctx, AG2_EINVAL, "ag2_xdrop_extend_batch: too many candidates per chunk");
    const int64_t slots = (int64_t)pair_grid * 2 * kPairThreads;
    const int64_t delta = std::max<int64_t>(1, ws_total / std::max<int64_t>(1, slots * n_chunks));

    set_slots_kernel<<<grid_for(n, 256, ctx->sm_count), 256, 0, st>>>((ExtGeom *)ctx->geom.p, (const int64_t *)ctx->prefix.p,
                                                                      (const int64_t *)ctx->meta_prefix.p, 0, n);
    CK(cudaMemsetAsync(ctx->lane_queue.p, 0xff, ((size_t)n * 2 + 1024) * 4, st));   // -1 = not published
    CK(cudaMemsetAsync(ctx->chunk_count.p, 0, (size_t)kMaxStreamChunks * 4, st));
    for (int c = 0; c < n_chunks; ++c) reinterpret_cast<volatile int *>(ctx->chunk_flag)[c] = 0;
    size_t order_tmp_bytes = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, order_tmp_bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr, (const int32_t *)nullptr,
                                       (int32_t *)nullptr, (int)(2 * n), 0, 32, st));
    RESERVE(ctx->order_keys, (size_t)n * 2 * 4 + 16);
    RESERVE(ctx->order_keys2, (size_t)n * 2 * 4 + 16);
    RESERVE(ctx->order_ids, (size_t)n * 2 * 4 + 16);
    RESERVE(ctx->order_queue, (size_t)n * 2 * 4 + 16);
    RESERVE(ctx->order_tmp, order_tmp_bytes + 16);
    stream_keys_kernel<<<grid_for(2 * n, 256, ctx->sm_count), 256, 0, st>>>((const ExtGeom *)ctx->geom.p, 2 * n, (int32_t)chunk_cn, delta,
                                                                            (uint32_t *)ctx->order_keys.p, (int32_t *)ctx->order_ids.p);
    {
        size_t tmp = ctx->order_tmp.cap;
        CK(cub::DeviceRadixSort::SortPairs(ctx->order_tmp.p, tmp, (const uint32_t *)ctx->order_keys.p, (uint32_t *)ctx->order_keys2.p,
                                           (const int32_t *)ctx->order_ids.p, (int32_t *)ctx->order_queue.p, (int)(2 * n), 0, 32, st));
    }
    launches += 3;

    StreamSignal sig = {};
    sig.chunk_done = (unsigned int *)ctx->chunk_count.p;
    CK(cudaHostGetDevicePointer((void **)&sig.chunk_flag, ctx->chunk_flag, 0));
    sig.n_cand = n;
    sig.chunk_cn = (int32_t)chunk_cn;
    sig.error = &sc->stream_error;
    // An asynchronous read load in flight: the kernel starts at once and every direction waits for the piece that holds its
    // read (wait_for_read) -- the pack kernels of the later pieces run in the slot every SM keeps free (see above; with all 7
    // pair CTAs per SM they did not, and the waits ran into their limit: kStreamStalled).  AG2_STREAM_WAIT_HOST=1 makes the
    // launch wait for the last piece instead (configs[1]: 571 ms per step instead of 519).
    const bool wait_on_host = getenv("AG2_STREAM_WAIT_HOST") != nullptr;
    if (ctx->reads_pending && wait_on_host) {
        CK(cudaStreamWaitEvent(st, ctx->pieces.back().ready, 0));
    } else if (ctx->reads_pending) {   // an asynchronous read load is in flight: directions wait for their piece inside the kernel
        sig.piece_flag = (const int *)ctx->piece_flag.p;
        sig.piece_end = (const int64_t *)ctx->piece_end.p;
        sig.n_pieces = (int32_t)ctx->h_piece_end.size();
    }
    LaneArgs a = {};
    a.sig = sig;
    a.seqs = sq;
    a.cand = d_cand;
    a.geom = (const ExtGeom *)ctx->geom.p;
    a.res = (ChainResult *)ctx->res.p;
    a.meta = (uint32_t *)ctx->meta.p;
    a.ws_q = (char *)ctx->ws_q.p;
    a.ws_t = (char *)ctx->ws_t.p;
    a.counters = &sc->ctr;
    LaneArgs pa = a;
    pa.scratch = (uint8_t *)ctx->tb_pair.p;
    pa.n_chains = 2 * n;
    pa.queue = (const int32_t *)ctx->order_queue.p;
    pa.next = &sc->next_pair;
    pa.wide_queue = (int32_t *)ctx->lane_queue.p;
    pa.wide_count = &sc->lane_count;
    pa.resume = (LaneResume *)ctx->lane_resume.p;
    pa.done_ctas = &sc->pair_done;
    ChainArgs cw = {};
    cw.sig = sig;
    cw.sig.piece_flag = nullptr;   // a handed-over direction was started by the pair kernel: its read is there
    cw.seqs = sq;
    cw.cand = a.cand;
    cw.geom = a.geom;
    cw.res = a.res;
    cw.ws_q = a.ws_q;
    cw.ws_t = a.ws_t;
    cw.tb = (uint8_t *)ctx->tb_stream.p;
    cw.tb_stride = tbw_stride;
    cw.queue = (const int32_t *)ctx->lane_queue.p;
    cw.next = &sc->next_wide;
    cw.counters = &sc->ctr;
    cw.try_narrow = 1;
    cw.resume = (const LaneResume *)ctx->lane_resume.p;
    cw.meta = pa.meta;
    CK(cudaStreamSynchronize(st));
    if (stream_grid > 0) xdrop_stream_kernel<kWideK, kWideWarps><<<stream_grid, kWideWarps * 32, 0, ctx->side_stream>>>(cw, &sc->pair_done, (unsigned)pair_grid);
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->side_done, ctx->side_stream));
    while (ctx->chain_events.size() < 1) {
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        ctx->chain_events.push_back({e0, e1});
    }
    const double t_launch = now_ms();
    CK(cudaEventRecord(ctx->chain_events[0].first, st));
    pair_kernel<<<pair_grid, kPairThreads, pair_smem, st>>>(pa);
    CK(cudaEventRecord(ctx->chain_events[0].second, st));
    CK(cudaGetLastError());
    CK(cudaStreamWaitEvent(st, ctx->side_done, 0));
    CK(cudaEventRecord(ctx->chunk_done_ev, st));   // pair kernel and its consumer are through
    launches += 2;

    // What the consumer did not take ([first, n_handed), published after its last ticket) runs after the pair kernel.
    auto post_pass = [&]() -> int {
        unsigned int n_handed = 0, n_wide = 0;
        unsigned long long taken = 0;
        CK(cudaMemcpyAsync(&n_handed, &sc->lane_count, sizeof n_handed, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(&taken, &sc->next_wide, sizeof taken, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        ctx->stats_lane_chains += n_handed;
        const unsigned int first = (unsigned int)std::min<unsigned long long>(taken, n_handed);
        const unsigned int n_lane = n_handed - first;
        ctx->stats_direct_wide += first;
        const bool lane_path = n_lane > lane_threshold;
        const int32_t *post_queue = (const int32_t *)ctx->lane_queue.p + first;
        if (n_lane > 0 && !lane_path) {
            n_wide = n_lane;
            ctx->stats_direct_wide += n_lane;
        }
        if (n_lane > 0 && lane_path) {
            LaneArgs la = a;
            la.sig.piece_flag = nullptr;
            la.scratch = (uint8_t *)ctx->tb.p;
            la.n_chains = n_lane;
            la.queue = post_queue;
            la.resume = (LaneResume *)ctx->lane_resume.p + first;
            la.next = &sc->next_fast;
            la.wide_queue = (int32_t *)ctx->wide_queue.p;
            la.wide_count = &sc->wide_count;
            xdrop_lane_kernel<<<lane_grid, kLaneThreads, lane_smem, st>>>(la);
            CK(cudaGetLastError());
            CK(cudaMemcpyAsync(&n_wide, &sc->wide_count, sizeof n_wide, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            post_queue = (const int32_t *)ctx->wide_queue.p;
            ++launches;
        }
        if (n_wide > 0) {
            ChainArgs w = cw;
            w.tb = (uint8_t *)ctx->tb_wide.p;
            w.n_chains = n_wide;
            w.queue = post_queue;
            w.next = &sc->next_post;
            w.resume = lane_path ? nullptr : (const LaneResume *)ctx->lane_resume.p + first;   // parallel to `queue`
            xdrop_chains_kernel<kWideK, kWideWarps><<<wide_grid, kWideWarps * 32, 0, st>>>(w);
            CK(cudaGetLastError());
            ++launches;
        }
        CK(cudaStreamSynchronize(st));
        return AG2_OK;
    };

    // drain: chunk after chunk, as their flags come up
    const volatile int *flags = ctx->chunk_flag;
    bool post_ran = false;
    int64_t dense_base = 0;
    int64_t *d_totals = (int64_t *)ctx->dense_off.p + n + 1;   // per chunk: its dense length
    for (int c = 0; c < n_chunks; ++c) {
        const int64_t lo = (int64_t)c * chunk_cn, cn = std::min<int64_t>(chunk_cn, n - lo);
        for (unsigned spins = 0; !flags[c]; ++spins) {
            if (!post_ran && (spins & 63) == 63) {
                const cudaError_t q = cudaEventQuery(ctx->chunk_done_ev);
                if (q == cudaSuccess) {   // the kernels are through and this chunk is not: it has directions left for the post-pass
                    const int rc = post_pass();
                    if (rc != AG2_OK) return rc;
                    post_ran = true;
                    if (!flags[c]) return fail(ctx, AG2_ECUDA, "ag2_xdrop_extend_batch: chunk %d incomplete after the last kernel", c);
                } else if (q != cudaErrorNotReady) {
                    return fail(ctx, AG2_ECUDA, "ag2_xdrop_extend_batch: %s", cudaGetErrorString(q));
                }
            }
            std::this_thread::sleep_for(std::chrono::microseconds(20));
        }
        if (trace) tr.push_back(now_ms());
        // small CTAs: these kernels must fit beside the resident pair kernel
        extend_finalize_kernel<<<grid_for(cn, 128, ctx->sm_count), 128, 0, post>>>(d_cand, (const ExtGeom *)ctx->geom.p, (const ChainResult *)ctx->res.p,
                                                                                 (const int32_t *)ctx->read_len.p, lo, cn, d_rec,
                                                                                 (int64_t *)ctx->str_begin.p, (int64_t *)ctx->ok_len.p);
        scan_chunk_kernel<<<1, 128, 0, post>>>((const int64_t *)ctx->ok_len.p + lo, cn, (int64_t *)ctx->dense_off.p + lo, d_totals + c);
        launches += 2;
        int64_t chunk_total = 0;
        CK(cudaMemcpyAsync(&chunk_total, d_totals + c, 8, cudaMemcpyDeviceToHost, post));
        CK(cudaStreamSynchronize(post));
        if (trace) tr.push_back(now_ms());
        assemble_kernel<<<grid_for(cn * 32, 128, ctx->sm_count), 128, 0, post>>>(
            d_rec, (const ExtGeom *)ctx->geom.p, (const ChainResult *)ctx->res.p, (const uint32_t *)ctx->meta.p, (const int64_t *)ctx->dense_off.p,
            dense_base, lo, cn, (const char *)ctx->ws_q.p, (const char *)ctx->ws_t.p, (char *)ctx->out_q.p, (char *)ctx->out_t.p, &sc->aligned,
            &sc->columns);
        ++launches;
        CK(cudaGetLastError());
        {
            const int rc = sink_copy_chunk(ctx, sink, dense_base, chunk_total, post, ctx->copy_stream, ctx->post_done, &launches);
            if (rc != AG2_OK) return rc;
        }
        CK(cudaMemcpyAsync(sink->rec + lo, d_rec + lo, (size_t)cn * sizeof(Record), cudaMemcpyDeviceToHost, ctx->copy_stream));
        dense_base += chunk_total;
        if (sink->ops) dense_base = (dense_base + 15) & ~(int64_t)15;
    }
    if (!post_ran) {   // nothing was missing for any chunk; the bookkeeping of the hand-overs is still due
        const int rc = post_pass();
        if (rc != AG2_OK) return rc;
    }
    CK(cudaStreamSynchronize(st));
    const double t_kernels = now_ms();
    CK(cudaStreamSynchronize(post));
    CK(cudaStreamSynchronize(ctx->copy_stream));
    *dense_total = dense_base;
    if (trace) {
        fprintf(stderr, "[ag2 trace] launch at %.1f ms, kernels done %.1f, all home %.1f; %d chunks (flag seen / scanned):", t_launch, t_kernels, now_ms(), n_chunks);
        for (size_t k = 0; k + 1 < tr.size(); k += 2) fprintf(stderr, " %.1f/%.1f", tr[k], tr[k + 1]);
        fprintf(stderr, "\n");
    }

    Scalars hs;
    CK(cudaMemcpy(&hs, sc, sizeof hs, cudaMemcpyDeviceToHost));
    if (hs.stream_error) {   // the caller reruns the batch in the chunked form
        fail(ctx, AG2_ECUDA, "ag2_xdrop_extend_batch: %u directions gave up waiting for their reads (upload stalled)", hs.stream_error);
        return kStreamStalled;
    }
    ag2_extend_stats &s = ctx->stats;
    s.cells = (int64_t)hs.ctr.cells;
    s.rows = (int64_t)hs.ctr.rows;
    s.blocks = (int64_t)hs.ctr.blocks;
    s.interior = (int64_t)hs.ctr.interior;
    s.slots = (int64_t)hs.ctr.slots;
    s.wide_chains = (int64_t)hs.ctr.wide + ctx->stats_direct_wide;
    s.aligned = (int64_t)hs.aligned;
    s.columns = (int64_t)hs.columns;
    s.launches = launches;
    s.lane_chains = ctx->stats_lane_chains;
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ctx->chain_events[0].first, ctx->chain_events[0].second));
    s.kernel_ms = ms;
    return AG2_OK;
}

int ag2_extend_run(ag2_ctx *ctx)
{
    if (!ctx || ctx->n_cand <= 0) return fail(ctx, AG2_ESTATE, "ag2_extend_run: no candidates uploaded");
    CK(cudaSetDevice(ctx->device));
    RESERVE(ctx->rec, (size_t)ctx->n_cand * sizeof(Record));
    int64_t total = 0;
    const int rc = extend_batch(ctx, (const Candidate *)ctx->cand.p, ctx->n_cand, (Record *)ctx->rec.p, 0, &total, true);
    if (rc != AG2_OK) return rc;
    ctx->out_total = total;
    ctx->ran = true;
    return AG2_OK;
}


int ag2_extend_get_stats(ag2_ctx *ctx, ag2_extend_stats *out)
{
    if (!ctx || !out) return AG2_EINVAL;
    if (!ctx->ran) return fail(ctx, AG2_ESTATE, "ag2_extend_get_stats: nothing ran");
    *out = ctx->stats;
    return AG2_OK;
}

int ag2_extend_fetch(ag2_ctx *ctx, ag2_record *rec_out, char *qaln_out, char *saln_out, int64_t aln_cap, int64_t *aln_used)
{
    if (!ctx || !rec_out) return fail(ctx, AG2_EINVAL, "ag2_extend_fetch: bad argument");
    if (!ctx->ran) return fail(ctx, AG2_ESTATE, "ag2_extend_fetch: call ag2_extend_run first");
    CK(cudaSetDevice(ctx->device));
    const int64_t n = ctx->n_cand;
    CK(cudaMemcpyAsync(rec_out, ctx->rec.p, (size_t)n * sizeof(Record), cudaMemcpyDeviceToHost, ctx->stream));
    if (aln_used) *aln_used = ctx->out_total;
    int rc = AG2_OK;
    if (qaln_out && saln_out && aln_cap >= ctx->out_total) {
        CK(cudaMemcpyAsync(qaln_out, ctx->out_q.p, (size_t)ctx->out_total, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(saln_out, ctx->out_t.p, (size_t)ctx->out_total, cudaMemcpyDeviceToHost, ctx->stream));
    } else if (qaln_out || saln_out) {
        rc = fail(ctx, AG2_ECAP, "ag2_extend_fetch: need %ld bytes per string, have %ld", (long)ctx->out_total, (long)aln_cap);
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return rc;
}

static int extend_batch_host(ag2_ctx *ctx, const ag2_candidate *cand, int64_t n, HostSink &sink, int64_t *aln_used, const char *who)
{
    int r = ag2_extend_upload(ctx, cand, n);
    if (r != AG2_OK) return r;
    if (!sink.rec) return fail(ctx, AG2_EINVAL, "%s: rec_out is required", who);
    CK(cudaSetDevice(ctx->device));
    RESERVE(ctx->rec, (size_t)n * sizeof(Record));
    int64_t total = 0;
    // Two forms of the host-buffer run: "chunked" = one pair-kernel launch per output chunk, the chunk's results copied home
    // while the next chunk computes; "streamed" = ONE launch for the whole batch with flags per output chunk
    // (extend_batch_streamed).  AG2_E2E_PATH picks one per call; without it the faster one at the full configs[1] size on
    // B200 is used (kStreamedByDefault; 519 vs 602 ms per step, profiles/e2e_sweep_r01l.md).
    bool chunked = !kStreamedByDefault;
    if (const char *e = getenv("AG2_E2E_PATH")) chunked = strcmp(e, "streamed") != 0;
    ctx->ws_call = chunked ? ctx->ws_limit_chunked : ctx->ws_limit_streamed;
    if (const char *e = getenv("AG2_WS_STREAMED")) ctx->ws_call = (size_t)std::max(1ll, atoll(e));   // tuning / test knob, per call
    if (!chunked) {
        r = extend_batch_streamed(ctx, (const Candidate *)ctx->cand.p, n, (Record *)ctx->rec.p, &total, &sink);
        if (r == kStreamStalled || r == kStreamTooBig) {   // stalled: nothing of that run is valid, and the reads are up by now
            if (r == kStreamStalled) fprintf(stderr, "[ag2] %s; running the batch again in the chunked form\n", ctx->err.c_str());
            if ((r = reads_barrier(ctx)) != AG2_OK) return r;
            sink.overflow = false;
            total = 0;
            chunked = true;
            ctx->ws_call = ctx->ws_limit_chunked;
        }
    }
    if (chunked) r = extend_batch(ctx, (const Candidate *)ctx->cand.p, n, (Record *)ctx->rec.p, 0, &total, true, &sink, cand);
    if (r != AG2_OK) return r;
    if ((r = reads_barrier(ctx)) != AG2_OK) return r;   // reads no candidate refers to may still be on their way
    ctx->out_total = total;
    ctx->ran = true;
    if (aln_used) *aln_used = total;
    if (sink.overflow) return fail(ctx, AG2_ECAP, "%s: need room for %ld alignment columns, have %ld", who, (long)total, (long)sink.cap);
    return AG2_OK;
}

int ag2_xdrop_extend_batch(ag2_ctx *ctx, const ag2_candidate *cand, int64_t n, ag2_record *rec_out, char *qaln_out,
                           char *saln_out, int64_t aln_cap, int64_t *aln_used)
{
    HostSink sink;
    sink.rec = rec_out;
    sink.q = qaln_out;
    sink.t = saln_out;
    sink.cap = aln_cap;
    return extend_batch_host(ctx, cand, n, sink, aln_used, "ag2_xdrop_extend_batch");
}

int ag2_xdrop_extend_batch_packed(ag2_ctx *ctx, const ag2_candidate *cand, int64_t n, ag2_record *rec_out, uint32_t *ops_out,
                                  int64_t cap_columns, int64_t *columns_used)
{
    if (!ops_out) return fail(ctx, AG2_EINVAL, "ag2_xdrop_extend_batch_packed: ops_out is required");
    HostSink sink;
    sink.rec = rec_out;
    sink.ops = ops_out;
    sink.cap = cap_columns & ~(int64_t)15;
    return extend_batch_host(ctx, cand, n, sink, columns_used, "ag2_xdrop_extend_batch_packed");
}

// The 2-bit ops of the dense pool of the last run (resident results): the packed counterpart of ag2_extend_fetch.
int ag2_extend_fetch_packed(ag2_ctx *ctx, ag2_record *rec_out, uint32_t *ops_out, int64_t cap_columns, int64_t *columns_used)
{
    if (!ctx || !rec_out) return fail(ctx, AG2_EINVAL, "ag2_extend_fetch_packed: bad argument");
    if (!ctx->ran) return fail(ctx, AG2_ESTATE, "ag2_extend_fetch_packed: call ag2_extend_run first");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int64_t n = ctx->n_cand, total = ctx->out_total, words = (total + 15) >> 4;
    CK(cudaMemcpyAsync(rec_out, ctx->rec.p, (size_t)n * sizeof(Record), cudaMemcpyDeviceToHost, st));
    if (columns_used) *columns_used = total;
    int rc = AG2_OK;
    if (ops_out && cap_columns >= total) {
        RESERVE(ctx->out_ops, (size_t)(words + 8) * 4);
        if (words > 0) {
            pack_ops_kernel<<<grid_for(words, 256, ctx->sm_count), 256, 0, st>>>((const char *)ctx->out_q.p, (const char *)ctx->out_t.p, 0, words, total,
                                                                                (uint32_t *)ctx->out_ops.p);
            CK(cudaGetLastError());
            CK(cudaMemcpyAsync(ops_out, ctx->out_ops.p, (size_t)words * 4, cudaMemcpyDeviceToHost, st));
        }
    } else if (ops_out) {
        rc = fail(ctx, AG2_ECAP, "ag2_extend_fetch_packed: need room for %ld columns, have %ld", (long)total, (long)cap_columns);
    }
    CK(cudaStreamSynchronize(st));
    return rc;
}

// ---- host side of the packed output: the two ASCII strings of every record from the ops, the reads and the reference ----
// code_char(get_dna_encode_table(c)) (MC/defs.cpp:3-36 + mecat2ref_aux.cpp:195-197): ACGT / acgt -> "ACGT", anything else -> 'A'
static inline char expand_base(unsigned char c)
{
    switch (c) {
    case 'A': case 'a': return 'A';
    case 'C': case 'c': return 'C';
    case 'G': case 'g': return 'G';
    case 'T': case 't': return 'T';
    default: return 'A';
    }
}
// the read as reference_mapping hands it to extend_candidate on strand R (impl_large.cpp:799-833): reversed, upper-case ACGT
// complemented, everything else as it is
static inline char expand_base_rc(unsigned char c)
{
    switch (c) {
    case 'A': return 'T';
    case 'C': return 'G';
    case 'G': return 'C';
    case 'T': return 'A';
    default: return expand_base(c);
    }
}

int ag2_expand_alignments(const ag2_record *rec, int64_t n, const uint32_t *ops, const char *read_bases, const int64_t *read_offs,
                          const char *ref, char *qaln_out, char *saln_out, int threads)
{
    if (!rec || n < 0 || !ops || !read_bases || !read_offs || !ref || !qaln_out || !saln_out) return AG2_EINVAL;
    if (threads < 1) threads = 1;
    auto work = [&](int64_t lo, int64_t hi) {
        for (int64_t k = lo; k < hi; ++k) {
            const ag2_record &r = rec[k];
            if (!r.ok) continue;
            const char *rd = read_bases + read_offs[r.read];
            const int64_t rlen = read_offs[r.read + 1] - read_offs[r.read];
            int64_t qi = r.qb, ti = r.sb;
            char *q = qaln_out + r.aln_off, *t = saln_out + r.aln_off;
            for (int64_t c = 0; c < r.aln_len; ++c) {
                const int64_t col = r.aln_off + c;
                const unsigned op = (ops[col >> 4] >> (2 * (col & 15))) & 3u;
                if (op == 1) {
                    q[c] = '-';
                } else {
                    q[c] = r.strand ? expand_base_rc((unsigned char)rd[rlen - 1 - qi]) : expand_base((unsigned char)rd[qi]);
                    ++qi;
                }
                if (op == 2) {
                    t[c] = '-';
                } else {
                    t[c] = expand_base((unsigned char)ref[ti]);
                    ++ti;
                }
            }
        }
    };
    if (threads == 1 || n < 2 * threads) {
        work(0, n);
    } else {
        std::vector<std::thread> th;
        for (int w = 0; w < threads; ++w) th.emplace_back(work, n * w / threads, n * (w + 1) / threads);
        for (auto &x : th) x.join();
    }
    return AG2_OK;
}

// ---- index (A2-A4) ------------------------------------------------------------------------------
static int scan_counts(ag2_ctx *ctx, const int32_t *cnt, uint32_t *off, int64_t *total_out)
{
    cudaStream_t st = ctx->stream;
    const int n_tiles = (kNCodes + kScanTile - 1) / kScanTile;
    RESERVE(ctx->ix_tiles, (size_t)(n_tiles + 2) * 4);
    uint32_t *tiles = (uint32_t *)ctx->ix_tiles.p;
    scan_tile_sums_kernel<<<n_tiles, 1024, 0, st>>>(cnt, kNCodes, tiles);
    scan_tiles_kernel<<<1, 1024, 0, st>>>(tiles, n_tiles, tiles + n_tiles);
    scan_apply_kernel<<<n_tiles, 1024, 0, st>>>(cnt, kNCodes, tiles, off);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(off + kNCodes, tiles + n_tiles, 4, cudaMemcpyDeviceToDevice, st));
    uint32_t total = 0;
    CK(cudaMemcpyAsync(&total, tiles + n_tiles, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    *total_out = total;
    return AG2_OK;
}

static int build_ref_csr(ag2_ctx *ctx)
{
    cudaStream_t st = ctx->stream;
    const int64_t R = ctx->ref_len;
    RESERVE(ctx->ix_cnt, (size_t)kNCodes * 4);
    RESERVE(ctx->ix_off, (size_t)(kNCodes + 1) * 4);
    RESERVE(ctx->ix_fill, (size_t)kNCodes * 4);
    int32_t *cnt = (int32_t *)ctx->ix_cnt.p;
    CK(cudaMemsetAsync(cnt, 0, (size_t)kNCodes * 4, st));
    const int g = grid_for(R, 256, ctx->sm_count);
    ref_kmer_kernel<0><<<g, 256, 0, st>>>((const uint32_t *)ctx->ref2.p, (const uint32_t *)ctx->ref_irr.p, R, cnt, nullptr, nullptr,
                                          nullptr, nullptr, nullptr, 1);
    mask_counts_kernel<<<grid_for(kNCodes, 256, ctx->sm_count), 256, 0, st>>>(cnt, kNCodes);
    CK(cudaGetLastError());
    int64_t total = 0;
    int rc = scan_counts(ctx, cnt, (uint32_t *)ctx->ix_off.p, &total);
    if (rc != AG2_OK) return rc;
    ctx->ix_npos = total;
    RESERVE(ctx->ix_pos, (size_t)(total + 4) * 4);
    CK(cudaMemsetAsync(ctx->ix_fill.p, 0, (size_t)kNCodes * 4, st));
    ref_kmer_kernel<1><<<g, 256, 0, st>>>((const uint32_t *)ctx->ref2.p, (const uint32_t *)ctx->ref_irr.p, R, cnt,
                                          (const uint32_t *)ctx->ix_off.p, (int32_t *)ctx->ix_fill.p, (uint32_t *)ctx->ix_pos.p,
                                          nullptr, nullptr, 1);
    sort_buckets_kernel<<<grid_for(kNCodes, 256, ctx->sm_count), 256, 0, st>>>(cnt, (const uint32_t *)ctx->ix_off.p,
                                                                               (uint32_t *)ctx->ix_pos.p, kNCodes);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    ctx->ref_indexed = true;
    return AG2_OK;
}

int ag2_index_build(ag2_ctx *ctx, int cbl, double alpha, double beta)
{
    if (!ctx || cbl <= 0) return fail(ctx, AG2_EINVAL, "ag2_index_build: bad argument");
    if (!ctx->ref_len || !ctx->n_reads) return fail(ctx, AG2_ESTATE, "ag2_index_build: load reference and reads first");
    CK(cudaSetDevice(ctx->device));
    {
        const int rb_ = reads_barrier(ctx);   // an asynchronous read load must have landed
        if (rb_ != AG2_OK) return rb_;
    }
    cudaStream_t st = ctx->stream;
    if (!ctx->ref_indexed) {
        int rc = build_ref_csr(ctx);
        if (rc != AG2_OK) return rc;
    }
    // A2: masked 13-mer counts of the concatenated read prefix (the ASCII of the batch is still resident)
    RESERVE(ctx->ix_rcnt, (size_t)kNCodes * 4);
    int32_t *rcnt = (int32_t *)ctx->ix_rcnt.p;
    CK(cudaMemsetAsync(rcnt, 0, (size_t)kNCodes * 4, st));
    ascii_kmer_hist_kernel<<<grid_for(ctx->read_prefix_len, 256, ctx->sm_count), 256, 0, st>>>((const char *)ctx->ascii.p,
                                                                                                ctx->read_prefix_len, rcnt);
    mask_counts_kernel<<<grid_for(kNCodes, 256, ctx->sm_count), 256, 0, st>>>(rcnt, kNCodes);
    // A3 (second half): per similarity block, the read counts of its reference k-mers; A4: votes
    const int64_t nblk = ctx->ref_len / cbl + 1;
    RESERVE(ctx->ix_kcount, (size_t)(nblk + 10) * 4);
    RESERVE(ctx->ix_vote, (size_t)(nblk + 10) * 4);
    CK(cudaMemsetAsync(ctx->ix_kcount.p, 0, (size_t)(nblk + 10) * 4, st));
    Scalars *sc = (Scalars *)ctx->scalars.p;
    CK(cudaMemsetAsync(&sc->aligned, 0, 8, st)); // reused as the k_count total
    ref_kmer_kernel<2><<<grid_for(ctx->ref_len, 256, ctx->sm_count), 256, 0, st>>>(
        (const uint32_t *)ctx->ref2.p, (const uint32_t *)ctx->ref_irr.p, ctx->ref_len, nullptr, nullptr, nullptr, nullptr, rcnt,
        (int32_t *)ctx->ix_kcount.p, cbl);
    sum_kcount_kernel<<<grid_for(nblk, 256, ctx->sm_count), 256, 0, st>>>((const int32_t *)ctx->ix_kcount.p, nblk, &sc->aligned);
    vote_kernel<<<grid_for(nblk + 10, 256, ctx->sm_count), 256, 0, st>>>((const int32_t *)ctx->ix_kcount.p, nblk, &sc->aligned, alpha,
                                                                         beta, (float *)ctx->ix_vote.p);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    ctx->ix_nblk = nblk;
    ctx->ix_cbl = cbl;
    ctx->votes_ready = true;
    return AG2_OK;
}

int ag2_index_fetch(ag2_ctx *ctx, int32_t *rcnt, int32_t *cnt, uint32_t *off, uint32_t *pos, int64_t pos_cap, int64_t *n_pos,
                    int32_t *kcount, float *vote, int64_t *nblk)
{
    if (!ctx) return AG2_EINVAL;
    if (!ctx->votes_ready) return fail(ctx, AG2_ESTATE, "ag2_index_fetch: call ag2_index_build first");
    CK(cudaSetDevice(ctx->device));
    if (n_pos) *n_pos = ctx->ix_npos;
    if (nblk) *nblk = ctx->ix_nblk;
    if (rcnt) CK(cudaMemcpy(rcnt, ctx->ix_rcnt.p, (size_t)kNCodes * 4, cudaMemcpyDeviceToHost));
    if (cnt) CK(cudaMemcpy(cnt, ctx->ix_cnt.p, (size_t)kNCodes * 4, cudaMemcpyDeviceToHost));
    if (off) CK(cudaMemcpy(off, ctx->ix_off.p, (size_t)(kNCodes + 1) * 4, cudaMemcpyDeviceToHost));
    if (pos) {
        if (pos_cap < ctx->ix_npos) return fail(ctx, AG2_ECAP, "ag2_index_fetch: need %ld positions", (long)ctx->ix_npos);
        CK(cudaMemcpy(pos, ctx->ix_pos.p, (size_t)ctx->ix_npos * 4, cudaMemcpyDeviceToHost));
    }
    if (kcount) CK(cudaMemcpy(kcount, ctx->ix_kcount.p, (size_t)(ctx->ix_nblk + 10) * 4, cudaMemcpyDeviceToHost));
    if (vote) CK(cudaMemcpy(vote, ctx->ix_vote.p, (size_t)(ctx->ix_nblk + 10) * 4, cudaMemcpyDeviceToHost));
    return AG2_OK;
}

// ---- seeding + candidate scoring (A5-A7) --------------------------------------------------------
// Scratch for the per-read block tables of one seeding launch.  The table of a read grows with its index hits (250 Mb
// reference: 2 600 hits per strand, 8 192 slots of 96 B = 786 KB per read), and a launch runs as many reads at once as
// fit: with the former fixed 4 GiB that was 5 000 reads -- 34 threads per SM, 1.7 s per 100 k reads
// (profiles/bench_r01o_250mb_100k.json).  Now a third of the free HBM, between 4 and 48 GiB.
static size_t seed_limit(ag2_ctx *ctx)
{
    if (const char *e = getenv("AG2_SEED_SCRATCH")) return (size_t)std::max(1ll, atoll(e));   // test knob: many small launches
    if (ctx->seed_scratch_limit) return ctx->seed_scratch_limit;
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return (size_t)4 << 30;
    const size_t third = (free_b + ctx->seed_scratch.cap) / 3;
    return std::min<size_t>((size_t)48 << 30, std::max<size_t>((size_t)4 << 30, third));
}

// Control words of the seeding / rescue-planning stage (device): work counters and overflow list lengths.
struct SeedCtl {
    unsigned next1, ovf1, next2, ovf2;            // seeding: CTA launch 1, its overflow, CTA launch 2, its overflow
    unsigned plan_n, pnext1, povf1, pnext2, povf2; // planning: listed items, then as above
    unsigned snext[3], sovf[3];                   // seeding with three CTA launches (seed_stage)
    unsigned pad[1];
};

constexpr int kSeedCapMax = 13824;   // index hits per strand that fit the 227 KB of shared memory of one CTA

// Events per strand the standard CTA launch is sized for: seeds of a read at the 90th percentile of the lengths x (index
// positions per bucket + share of exact seeds of a 15 %-error read), with a margin; what exceeds it goes to the launch with
// kSeedCapMax, then to the thread path.
static int seed_cap(const ag2_ctx *ctx, int pass, int tier)
{
    if (tier == 0)
        if (const char *e = getenv("AG2_SEED_CAP")) return std::max(64, std::min(kSeedCapMax, atoi(e) & ~63));   // test knob
    if (tier == 1) {
        if (const char *e = getenv("AG2_SEED_CAP2")) return std::max(64, std::min(kSeedCapMax, atoi(e) & ~63));
        return kSeedCapMax;
    }
    const double len = ctx->n_reads ? (double)ctx->read_len_p90 : 10000.0;   // nine reads of ten are not longer
    const double bc = pass == 0 ? std::min(20.0, 5.0 + len / 1000.0) : 5.0;
    const double density = (double)ctx->ix_npos / (double)kNCodes;
    const double ev = (len / bc + 1.0) * (density + 0.2);
    const int cap = ((int)(ev * 1.25) + 128 + 127) & ~127;
    return std::max(256, std::min(kSeedCapMax, cap));
}

// A tighter table for the FIRST launch (6 % + 32 over the expected number of events) where that fits one more CTA on an SM
// and still leaves the SM ~4 KB of L1: the kernel is latency-bound (3 -> 4 CTAs per SM at 250 Mb: 75 -> 65 ms per 250 k
// reads; with no L1 left it was 82 ms), and the few reads the smaller table turns away go to the next launch.  0 = none.
static int seed_cap_tight(const ag2_ctx *ctx, const void *kernel, int pass, int cap)
{
    if (getenv("AG2_SEED_CAP")) return 0;
    const double len = ctx->n_reads ? (double)ctx->read_len_p90 : 10000.0;
    const double bc = pass == 0 ? std::min(20.0, 5.0 + len / 1000.0) : 5.0;
    const double ev = (len / bc + 1.0) * ((double)ctx->ix_npos / (double)kNCodes + 0.2);
    const int floor_cap = std::max(256, ((int)(ev * 1.06) + 32 + 63) & ~63);
    int occ0 = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ0, kernel, kSeedCtaThreads, seed_cta_smem_bytes(cap)) != cudaSuccess) return 0;
    for (int c = cap - 64; c >= floor_cap; c -= 64) {
        int occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kSeedCtaThreads, seed_cta_smem_bytes(c) + 1024) != cudaSuccess) break;
        if (occ > occ0) return c;
    }
    return 0;
}

static int seed_cta_config(ag2_ctx *ctx, const void *kernel, int cap, size_t *smem_out, int *grid_out)
{
    const size_t smem = seed_cta_smem_bytes(cap);
    // the limit of the function, not of this launch: it is per-device state that the contexts of other host threads set too
    // (the executables run one context per thread, possibly several on one GPU), so it is always the largest any launch asks for
    CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)seed_cta_smem_bytes(kSeedCapMax)));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kSeedCtaThreads, smem));
    if (occ < 1) return fail(ctx, AG2_ECUDA, "seeding kernel does not fit an SM (cap %d)", cap);
    *smem_out = smem;
    *grid_out = ctx->sm_count * occ;
    return AG2_OK;
}

// Scratch layout of a one-thread-per-read launch over work[0..n_work): per-entry table bytes -> prefix (host) -> chunks
// that fit the scratch limit.  Rare path: only what overflowed both CTA launches comes here.
static int thread_path_scratch(ag2_ctx *ctx, const RefIndex &ix, int pass, const int32_t *d_reads, const int32_t *d_work, int64_t n_work,
                               std::vector<std::pair<int64_t, int64_t>> &chunks)
{
    cudaStream_t st = ctx->stream;
    const PackedSeqs sq = seqs_of(ctx);
    RESERVE(ctx->seed_need, (size_t)n_work * 8);
    RESERVE(ctx->seed_prefix, (size_t)(n_work + 1) * 8);
    seed_need_sub_kernel<<<grid_for(n_work, 128, ctx->sm_count), 128, 0, st>>>(ix, sq.reads2, sq.reads_irr, sq.read_off, sq.read_len, d_reads,
                                                                            d_work, n_work, pass, (int64_t *)ctx->seed_need.p);
    exclusive_scan_i64<<<1, 1024, 0, st>>>((const int64_t *)ctx->seed_need.p, n_work, (int64_t *)ctx->seed_prefix.p);
    CK(cudaGetLastError());
    std::vector<int64_t> pf((size_t)n_work + 1);
    CK(cudaMemcpyAsync(pf.data(), ctx->seed_prefix.p, (size_t)(n_work + 1) * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    size_t max_chunk = 0;
    const size_t scratch_limit = seed_limit(ctx);
    chunks.clear();
    for (int64_t lo = 0; lo < n_work;) {
        int64_t hi = lo + 1;
        while (hi < n_work && (size_t)(pf[hi + 1] - pf[lo]) <= scratch_limit) ++hi;
        chunks.push_back({lo, hi});
        max_chunk = std::max(max_chunk, (size_t)(pf[hi] - pf[lo]));
        lo = hi;
    }
    RESERVE(ctx->seed_scratch, max_chunk + 64);
    return AG2_OK;
}

// Seeding + candidate scoring (A5-A7) of n items (item k = read d_reads[k], or read k): candidates into seed_cands /
// seed_ncand.  Two CTA-per-read launches (the second takes the first one's overflow list straight from device memory),
// then one small read-back that says whether anything is left for the thread path.
static int seed_stage(ag2_ctx *ctx, int pass, int maxc, const int32_t *d_reads, int64_t n)
{
    cudaStream_t st = ctx->stream;
    const PackedSeqs sq = seqs_of(ctx);
    RefIndex ix = {ctx->ref_len, (const int32_t *)ctx->ix_cnt.p, (const uint32_t *)ctx->ix_off.p, (const uint32_t *)ctx->ix_pos.p,
                   (const float *)ctx->ix_vote.p, ctx->ix_cbl};
    RESERVE(ctx->seed_cands, (size_t)n * maxc * sizeof(SeedCand));
    RESERVE(ctx->seed_ncand, (size_t)n * 4);
    RESERVE(ctx->seed_ctl, sizeof(SeedCtl));
    RESERVE(ctx->seed_ovf1, (size_t)n * 4);
    RESERVE(ctx->seed_ovf2, (size_t)n * 4);
    SeedCtl *ctl = (SeedCtl *)ctx->seed_ctl.p;
    CK(cudaMemsetAsync(ctl, 0, sizeof(SeedCtl), st));
    CK(cudaFuncSetAttribute((const void *)seed_cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)seed_cta_smem_bytes(kSeedCapMax)));
    int caps[3], n_tiers = 0;
    {
        const int standard = seed_cap(ctx, pass, 0), tight = seed_cap_tight(ctx, (const void *)seed_cta_kernel, pass, standard);
        if (tight > 0) caps[n_tiers++] = tight;
        caps[n_tiers++] = standard;
        caps[n_tiers++] = seed_cap(ctx, pass, 1);
    }
    DevBuf *lists[2] = {&ctx->seed_ovf1, &ctx->seed_ovf2};   // launch t reads the list launch t - 1 wrote, and writes the other one
    for (int tier = 0; tier < n_tiers; ++tier) {
        size_t smem = 0;
        int grid = 0;
        int rc = seed_cta_config(ctx, (const void *)seed_cta_kernel, caps[tier], &smem, &grid);
        if (rc != AG2_OK) return rc;
        if (tier == 0) grid = (int)std::min<int64_t>(grid, n);
        RESERVE(ctx->seed_pool, (size_t)grid * (caps[tier] / (kSM + 1) + 1) * kHeavyWords * 4);
        SeedCtaArgs a = {};
        a.ix = ix;
        a.reads2 = sq.reads2;
        a.irr = sq.reads_irr;
        a.read_off = sq.read_off;
        a.read_len = sq.read_len;
        a.reads = d_reads;
        a.work = tier == 0 ? nullptr : (const int32_t *)lists[(tier - 1) & 1]->p;
        a.n_work_dev = tier == 0 ? nullptr : &ctl->sovf[tier - 1];
        a.n_work = (unsigned)n;
        a.pass = pass;
        a.maxc = maxc;
        a.cap = caps[tier];
        a.next = &ctl->snext[tier];
        a.cands = (SeedCand *)ctx->seed_cands.p;
        a.ncand = (int32_t *)ctx->seed_ncand.p;
        a.ovf = (int32_t *)lists[tier & 1]->p;
        a.ovf_count = &ctl->sovf[tier];
        a.heavy_pool = (uint32_t *)ctx->seed_pool.p;
        seed_cta_kernel<<<grid, kSeedCtaThreads, smem, st>>>(a);
        CK(cudaGetLastError());
    }
    SeedCtl h;
    CK(cudaMemcpyAsync(&h, ctl, sizeof h, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    ctx->seed_overflow[0] = h.sovf[0];
    ctx->seed_overflow[1] = h.sovf[n_tiers - 1];
    const int32_t *left_list = (const int32_t *)lists[(n_tiers - 1) & 1]->p;
    const unsigned n_left = h.sovf[n_tiers - 1];
    if (n_left > 0) {   // one thread per read, table in global memory
        std::vector<std::pair<int64_t, int64_t>> chunks;
        int rc = thread_path_scratch(ctx, ix, pass, d_reads, left_list, n_left, chunks);
        if (rc != AG2_OK) return rc;
        for (auto &c : chunks) {
            const int64_t cn = c.second - c.first;
            seed_map_sub_kernel<<<grid_for(cn, 128, ctx->sm_count), 128, 0, st>>>(
                ix, sq.reads2, sq.reads_irr, sq.read_off, sq.read_len, d_reads, left_list, c.first, cn, pass, maxc,
                (const int64_t *)ctx->seed_prefix.p, (uint8_t *)ctx->seed_scratch.p, (SeedCand *)ctx->seed_cands.p, (int32_t *)ctx->seed_ncand.p);
        }
        CK(cudaGetLastError());
    }
    return AG2_OK;
}

int ag2_seed_candidates(ag2_ctx *ctx, int pass, int maxc, ag2_seed_candidate *out, int32_t *ncand_out)
{
    static_assert(sizeof(ag2_seed_candidate) == sizeof(SeedCand), "ag2_seed_candidate layout");
    if (!ctx || pass < 0 || pass > 1 || maxc < 1 || maxc > kMaxCand) return fail(ctx, AG2_EINVAL, "ag2_seed_candidates: bad argument");
    if (!ctx->votes_ready) return fail(ctx, AG2_ESTATE, "ag2_seed_candidates: call ag2_index_build first");
    CK(cudaSetDevice(ctx->device));
    {
        const int rb_ = reads_barrier(ctx);   // an asynchronous read load must have landed
        if (rb_ != AG2_OK) return rb_;
    }
    cudaStream_t st = ctx->stream;
    const int64_t n = ctx->n_reads;
    const int rc = seed_stage(ctx, pass, maxc, nullptr, n);
    if (rc != AG2_OK) return rc;
    if (out) CK(cudaMemcpyAsync(out, ctx->seed_cands.p, (size_t)n * maxc * sizeof(SeedCand), cudaMemcpyDeviceToHost, st));
    if (ncand_out) CK(cudaMemcpyAsync(ncand_out, ctx->seed_ncand.p, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return AG2_OK;
}

// The seed candidates of the last ag2_seed_candidates call become the extension candidates, on the device.
int ag2_extend_upload_from_seeds(ag2_ctx *ctx, int maxc, int64_t *n_out)
{
    if (!ctx || maxc < 1 || maxc > kMaxCand) return fail(ctx, AG2_EINVAL, "ag2_extend_upload_from_seeds: bad argument");
    if (!ctx->seed_ncand.p || !ctx->votes_ready) return fail(ctx, AG2_ESTATE, "ag2_extend_upload_from_seeds: call ag2_seed_candidates first");
    CK(cudaSetDevice(ctx->device));
    {
        const int rb_ = reads_barrier(ctx);   // an asynchronous read load must have landed
        if (rb_ != AG2_OK) return rb_;
    }
    cudaStream_t st = ctx->stream;
    const int64_t n = ctx->n_reads;
    RESERVE(ctx->seed_need, (size_t)n * 8);
    RESERVE(ctx->seed_prefix, (size_t)(n + 1) * 8);
    widen_i32_kernel<<<grid_for(n, 256, ctx->sm_count), 256, 0, st>>>((const int32_t *)ctx->seed_ncand.p, n, (int64_t *)ctx->seed_need.p);
    exclusive_scan_i64<<<1, 1024, 0, st>>>((const int64_t *)ctx->seed_need.p, n, (int64_t *)ctx->seed_prefix.p);
    int64_t total = 0;
    CK(cudaMemcpyAsync(&total, (int64_t *)ctx->seed_prefix.p + n, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (n_out) *n_out = total;
    ctx->n_cand = 0;
    ctx->ran = false;
    if (total == 0) return AG2_OK;
    RESERVE(ctx->cand, (size_t)total * sizeof(Candidate));
    seeds_to_candidates_kernel<<<grid_for(n, 256, ctx->sm_count), 256, 0, st>>>((const SeedCand *)ctx->seed_cands.p,
                                                                                (const int32_t *)ctx->seed_ncand.p,
                                                                                (const int64_t *)ctx->seed_prefix.p, n, maxc,
                                                                                (Candidate *)ctx->cand.p);
    CK(cudaGetLastError());
    ctx->n_cand = total;
    return AG2_OK;
}

// ---- the whole per-read path: reference_mapping()'s loop body for a batch (impl_large.cpp:776-1316) ----
// One pass (pass 0, or the reference's second pass over `reads`, a device list of read indices):
// seed -> extend every candidate -> plan rescue -> extend the rescue candidates -> link + choose the output.
static int map_pass(ag2_ctx *ctx, int pass, int maxc, int num_output, const int32_t *d_reads, int64_t n, int64_t *pool_n, int64_t *dense_base,
                    bool first_batch)
{
    cudaStream_t st = ctx->stream;
    const PackedSeqs sq = seqs_of(ctx);
    RefIndex ix = {ctx->ref_len, (const int32_t *)ctx->ix_cnt.p, (const uint32_t *)ctx->ix_off.p, (const uint32_t *)ctx->ix_pos.p,
                   (const float *)ctx->ix_vote.p, ctx->ix_cbl};
    const int g = grid_for(n, 128, ctx->sm_count);
    ag2_map_stats &ms = ctx->map_stats;
    auto t_prev = std::chrono::steady_clock::now();
    auto lap = [&](double &acc) {   // host wall time since the last lap; every stage below ends at a host synchronisation
        const auto t = std::chrono::steady_clock::now();
        acc += std::chrono::duration<double, std::milli>(t - t_prev).count();
        t_prev = t;
    };
    // seeding + candidates
    {
        const int rc = seed_stage(ctx, pass, maxc, d_reads, n);
        if (rc != AG2_OK) return rc;
        ms.seed_overflow1 += ctx->seed_overflow[0];
        ms.seed_overflow2 += ctx->seed_overflow[1];
    }
    lap(pass == 0 ? ms.seed_ms : ms.pass2_ms);
    // candidates of all reads, in order
    RESERVE(ctx->map_cand_prefix, (size_t)(n + 1) * 8);
    RESERVE(ctx->map_rescue_n, (size_t)n * 8);
    widen_i32_kernel<<<grid_for(n, 256, ctx->sm_count), 256, 0, st>>>((const int32_t *)ctx->seed_ncand.p, n, (int64_t *)ctx->map_rescue_n.p);
    exclusive_scan_i64<<<1, 1024, 0, st>>>((const int64_t *)ctx->map_rescue_n.p, n, (int64_t *)ctx->map_cand_prefix.p);
    int64_t n_cand = 0;
    CK(cudaMemcpyAsync(&n_cand, (int64_t *)ctx->map_cand_prefix.p + n, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const int64_t base0 = *pool_n;
    {
        int rk = reserve_keep(ctx, ctx->rec_pool, (size_t)(base0 + n_cand + 1) * sizeof(Record), (size_t)base0 * sizeof(Record));
        if (rk != AG2_OK) return rk;
    }
    if (n_cand > 0) {
        RESERVE(ctx->map_cand, (size_t)n_cand * sizeof(Candidate));
        seeds_to_candidates_sub_kernel<<<grid_for(n, 256, ctx->sm_count), 256, 0, st>>>(
            (const SeedCand *)ctx->seed_cands.p, (const int32_t *)ctx->seed_ncand.p, (const int64_t *)ctx->map_cand_prefix.p, d_reads, n, maxc,
            (Candidate *)ctx->map_cand.p);
        CK(cudaGetLastError());
        int rc = extend_batch(ctx, (const Candidate *)ctx->map_cand.p, n_cand, (Record *)ctx->rec_pool.p + base0, *dense_base, dense_base,
                              first_batch);
        if (rc != AG2_OK) return rc;
    }
    *pool_n = base0 + n_cand;
    ms.n_candidates += n_cand;
    lap(pass == 0 ? ms.extend_ms : ms.pass2_ms);
    // rescue planning: the alignment lists of all items, then the seeding tables of the few that look clipped
    RESERVE(ctx->map_plans, (size_t)n * sizeof(ReadPlan));
    RESERVE(ctx->map_rescue_prefix, (size_t)(n + 1) * 8);
    RESERVE(ctx->plan_list, (size_t)n * 4);
    SeedCtl *ctl = (SeedCtl *)ctx->seed_ctl.p;
    plan_alns_kernel<<<g, 128, 0, st>>>(sq.read_len, d_reads, n, (const int32_t *)ctx->seed_ncand.p, (const int64_t *)ctx->map_cand_prefix.p,
                                        (const Record *)ctx->rec_pool.p, base0, (ReadPlan *)ctx->map_plans.p, (int64_t *)ctx->map_rescue_n.p,
                                        (int32_t *)ctx->plan_list.p, &ctl->plan_n);
    CK(cudaGetLastError());
    {
        const int caps[2] = {seed_cap(ctx, pass, 0), seed_cap(ctx, pass, 1)};
        for (int tier = 0; tier < 2; ++tier) {
            size_t smem = 0;
            int grid = 0;
            int rc = seed_cta_config(ctx, (const void *)plan_cta_kernel, caps[tier], &smem, &grid);
            if (rc != AG2_OK) return rc;
            if (tier == 0) grid = (int)std::min<int64_t>(grid, n);
            RESERVE(ctx->seed_pool, (size_t)grid * (caps[tier] / (kSM + 1) + 1) * kHeavyWords * 4);
            PlanCtaArgs a = {};
            a.ix = ix;
            a.reads2 = sq.reads2;
            a.irr = sq.reads_irr;
            a.read_off = sq.read_off;
            a.read_len = sq.read_len;
            a.reads = d_reads;
            a.work = (const int32_t *)(tier == 0 ? ctx->plan_list.p : ctx->seed_ovf1.p);
            a.n_work_dev = tier == 0 ? &ctl->plan_n : &ctl->povf1;
            a.pass = pass;
            a.cap = caps[tier];
            a.next = tier == 0 ? &ctl->pnext1 : &ctl->pnext2;
            a.plans = (ReadPlan *)ctx->map_plans.p;
            a.n_rescue = (int64_t *)ctx->map_rescue_n.p;
            a.ovf = (int32_t *)(tier == 0 ? ctx->seed_ovf1.p : ctx->seed_ovf2.p);
            a.ovf_count = tier == 0 ? &ctl->povf1 : &ctl->povf2;
            a.heavy_pool = (uint32_t *)ctx->seed_pool.p;
            plan_cta_kernel<<<grid, kSeedCtaThreads, smem, st>>>(a);
            CK(cudaGetLastError());
        }
        SeedCtl h;
        CK(cudaMemcpyAsync(&h, ctl, sizeof h, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (h.povf2 > 0) {
            std::vector<std::pair<int64_t, int64_t>> chunks;
            int rc = thread_path_scratch(ctx, ix, pass, d_reads, (const int32_t *)ctx->seed_ovf2.p, h.povf2, chunks);
            if (rc != AG2_OK) return rc;
            for (auto &c : chunks) {
                const int64_t cn = c.second - c.first;
                plan_search_sub_kernel<<<grid_for(cn, 128, ctx->sm_count), 128, 0, st>>>(
                    ix, sq.reads2, sq.reads_irr, sq.read_off, sq.read_len, d_reads, (const int32_t *)ctx->seed_ovf2.p, c.first, cn, pass,
                    (const int64_t *)ctx->seed_prefix.p, (uint8_t *)ctx->seed_scratch.p, (ReadPlan *)ctx->map_plans.p,
                    (int64_t *)ctx->map_rescue_n.p);
            }
            CK(cudaGetLastError());
        }
    }
    exclusive_scan_i64<<<1, 1024, 0, st>>>((const int64_t *)ctx->map_rescue_n.p, n, (int64_t *)ctx->map_rescue_prefix.p);
    CK(cudaGetLastError());
    int64_t n_resc = 0;
    CK(cudaMemcpyAsync(&n_resc, (int64_t *)ctx->map_rescue_prefix.p + n, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    ms.n_rescue += n_resc;
    lap(pass == 0 ? ms.plan_ms : ms.pass2_ms);
    const int64_t base1 = *pool_n;
    if (n_resc > 0) {
        int rk = reserve_keep(ctx, ctx->rec_pool, (size_t)(base1 + n_resc + 1) * sizeof(Record), (size_t)base1 * sizeof(Record));
        if (rk != AG2_OK) return rk;
        RESERVE(ctx->map_cand, (size_t)n_resc * sizeof(Candidate));
        rescue_gather_kernel<<<grid_for(n, 256, ctx->sm_count), 256, 0, st>>>((const ReadPlan *)ctx->map_plans.p,
                                                                              (const int64_t *)ctx->map_rescue_prefix.p, d_reads, n,
                                                                              (Candidate *)ctx->map_cand.p);
        CK(cudaGetLastError());
        int rc = extend_batch(ctx, (const Candidate *)ctx->map_cand.p, n_resc, (Record *)ctx->rec_pool.p + base1, *dense_base, dense_base, false);
        if (rc != AG2_OK) return rc;
        *pool_n = base1 + n_resc;
    }
    finish_kernel<<<g, 128, 0, st>>>((ReadPlan *)ctx->map_plans.p, (const int64_t *)ctx->map_rescue_prefix.p, d_reads, n, sq.read_len,
                                     (const Record *)ctx->rec_pool.p, base1, num_output, (int64_t *)ctx->map_out_refs.p,
                                     (int32_t *)ctx->map_nout.p, pass == 0 ? (int32_t *)ctx->map_flags.p : nullptr);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    lap(pass == 0 ? ms.rescue_ms : ms.pass2_ms);
    return AG2_OK;
}

int ag2_map_reads(ag2_ctx *ctx, int maxc, int num_output, int64_t *n_records)
{
    if (!ctx || maxc < 1 || maxc > kMaxCand || num_output < 1) return fail(ctx, AG2_EINVAL, "ag2_map_reads: bad argument");
    if (!ctx->votes_ready) return fail(ctx, AG2_ESTATE, "ag2_map_reads: call ag2_index_build first");
    CK(cudaSetDevice(ctx->device));
    {
        const int rb_ = reads_barrier(ctx);   // an asynchronous read load must have landed
        if (rb_ != AG2_OK) return rb_;
    }
    cudaStream_t st = ctx->stream;
    const int64_t n = ctx->n_reads;
    if (num_output > maxc) num_output = maxc; // mecat2ref.cpp:196-201
    RESERVE(ctx->map_out_refs, (size_t)n * kOutCap * 8);
    RESERVE(ctx->map_nout, (size_t)n * 4);
    RESERVE(ctx->map_flags, (size_t)n * 4);
    RESERVE(ctx->map_list, (size_t)n * 4);
    RESERVE(ctx->map_out_prefix, (size_t)(n + 1) * 8);
    int64_t pool_n = 0, dense = 0;
    ctx->map_stats = ag2_map_stats{};
    const auto t_begin = std::chrono::steady_clock::now();
    int rc = map_pass(ctx, 0, maxc, num_output, nullptr, n, &pool_n, &dense, true);
    if (rc != AG2_OK) return rc;
    // second pass for the reads none of whose candidates extended (:1049)
    flags_to_i64_kernel<<<grid_for(n, 256, ctx->sm_count), 256, 0, st>>>((const int32_t *)ctx->map_flags.p, n, (int64_t *)ctx->map_rescue_n.p);
    exclusive_scan_i64<<<1, 1024, 0, st>>>((const int64_t *)ctx->map_rescue_n.p, n, (int64_t *)ctx->map_out_prefix.p);
    compact_reads_kernel<<<grid_for(n, 256, ctx->sm_count), 256, 0, st>>>((const int32_t *)ctx->map_flags.p, (const int64_t *)ctx->map_out_prefix.p, n,
                                                                          (int32_t *)ctx->map_list.p);
    int64_t n2 = 0;
    CK(cudaMemcpyAsync(&n2, (int64_t *)ctx->map_out_prefix.p + n, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    ctx->map_stats.n_pass2_reads = n2;
    if (n2 > 0) {
        rc = map_pass(ctx, 1, maxc, num_output, (const int32_t *)ctx->map_list.p, n2, &pool_n, &dense, false);
        if (rc != AG2_OK) return rc;
    }
    // records in thread-file order
    widen_i32_kernel<<<grid_for(n, 256, ctx->sm_count), 256, 0, st>>>((const int32_t *)ctx->map_nout.p, n, (int64_t *)ctx->map_rescue_n.p);
    exclusive_scan_i64<<<1, 1024, 0, st>>>((const int64_t *)ctx->map_rescue_n.p, n, (int64_t *)ctx->map_out_prefix.p);
    int64_t n_out = 0;
    CK(cudaMemcpyAsync(&n_out, (int64_t *)ctx->map_out_prefix.p + n, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    RESERVE(ctx->map_out_rec, (size_t)(n_out + 1) * sizeof(Record));
    gather_output_kernel<<<grid_for(n, 256, ctx->sm_count), 256, 0, st>>>((const int64_t *)ctx->map_out_refs.p, (const int32_t *)ctx->map_nout.p,
                                                                          (const int64_t *)ctx->map_out_prefix.p, n, (const Record *)ctx->rec_pool.p,
                                                                          (Record *)ctx->map_out_rec.p);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    ctx->map_n_out = n_out;
    ctx->out_total = dense;
    ctx->mapped = true;
    ctx->stats.aligned = 0; // per-call totals of the extension batches are in cells / rows / blocks; aligned counts every extension
    ctx->ran = true;
    ctx->map_stats.n_reads = n;
    ctx->map_stats.n_records = n_out;
    ctx->map_stats.cells = ctx->stats.cells;
    ctx->map_stats.pair_kernel_ms = ctx->stats.kernel_ms;
    ctx->map_stats.launches = ctx->stats.launches;
    ctx->map_stats.total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
    if (n_records) *n_records = n_out;
    return AG2_OK;
}

int ag2_map_get_stats(ag2_ctx *ctx, ag2_map_stats *out)
{
    if (!ctx || !out) return AG2_EINVAL;
    if (!ctx->mapped) return fail(ctx, AG2_ESTATE, "ag2_map_get_stats: call ag2_map_reads first");
    *out = ctx->map_stats;
    return AG2_OK;
}

int ag2_map_fetch(ag2_ctx *ctx, ag2_record *rec_out, char *qaln_out, char *saln_out, int64_t aln_cap, int64_t *aln_used)
{
    if (!ctx || !rec_out) return fail(ctx, AG2_EINVAL, "ag2_map_fetch: bad argument");
    if (!ctx->mapped) return fail(ctx, AG2_ESTATE, "ag2_map_fetch: call ag2_map_reads first");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(rec_out, ctx->map_out_rec.p, (size_t)ctx->map_n_out * sizeof(Record), cudaMemcpyDeviceToHost, ctx->stream));
    if (aln_used) *aln_used = ctx->out_total;
    int rc = AG2_OK;
    if (qaln_out && saln_out && aln_cap >= ctx->out_total) {
        CK(cudaMemcpyAsync(qaln_out, ctx->out_q.p, (size_t)ctx->out_total, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(saln_out, ctx->out_t.p, (size_t)ctx->out_total, cudaMemcpyDeviceToHost, ctx->stream));
    } else if (qaln_out || saln_out) {
        rc = fail(ctx, AG2_ECAP, "ag2_map_fetch: need %ld bytes per string, have %ld", (long)ctx->out_total, (long)aln_cap);
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return rc;
}

// ag2_map_fetch with the alignments as 2-bit ops (see ag2_xdrop_extend_batch_packed)
int ag2_map_fetch_packed(ag2_ctx *ctx, ag2_record *rec_out, uint32_t *ops_out, int64_t cap_columns, int64_t *columns_used)
{
    if (!ctx || !rec_out) return fail(ctx, AG2_EINVAL, "ag2_map_fetch_packed: bad argument");
    if (!ctx->mapped) return fail(ctx, AG2_ESTATE, "ag2_map_fetch_packed: call ag2_map_reads first");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int64_t total = ctx->out_total, words = (total + 15) >> 4;
    CK(cudaMemcpyAsync(rec_out, ctx->map_out_rec.p, (size_t)ctx->map_n_out * sizeof(Record), cudaMemcpyDeviceToHost, st));
    if (columns_used) *columns_used = total;
    int rc = AG2_OK;
    if (ops_out && cap_columns >= total) {
        RESERVE(ctx->out_ops, (size_t)(words + 8) * 4);
        if (words > 0) {
            pack_ops_kernel<<<grid_for(words, 256, ctx->sm_count), 256, 0, st>>>((const char *)ctx->out_q.p, (const char *)ctx->out_t.p, 0, words, total,
                                                                                (uint32_t *)ctx->out_ops.p);
            CK(cudaGetLastError());
            CK(cudaMemcpyAsync(ops_out, ctx->out_ops.p, (size_t)words * 4, cudaMemcpyDeviceToHost, st));
        }
    } else if (ops_out) {
        rc = fail(ctx, AG2_ECAP, "ag2_map_fetch_packed: need room for %ld columns, have %ld", (long)total, (long)cap_columns);
    }
    CK(cudaStreamSynchronize(st));
    return rc;
}

// ---- PAGraph kmer_counter (B1) -------------------------------------------------------------------
int ag2_kmer_begin(ag2_ctx *ctx, int k)
{
    if (!ctx || k < 1 || k > 15) return fail(ctx, AG2_EINVAL, "ag2_kmer_begin: k must be in 1..15");
    CK(cudaSetDevice(ctx->device));
    const int64_t nbins = 1ll << (2 * k);
    RESERVE(ctx->km_table, (size_t)nbins * 4);
    CK(cudaMemsetAsync(ctx->km_table.p, 0, (size_t)nbins * 4, ctx->stream));
    ctx->km_k = k;
    ctx->km_n_solid = 0;
    return AG2_OK;
}

int ag2_kmer_add_reads(ag2_ctx *ctx)
{
    if (!ctx || !ctx->km_k) return fail(ctx, AG2_ESTATE, "ag2_kmer_add_reads: call ag2_kmer_begin first");
    if (!ctx->n_reads) return fail(ctx, AG2_ESTATE, "ag2_kmer_add_reads: no reads loaded");
    CK(cudaSetDevice(ctx->device));
    {
        const int rb_ = reads_barrier(ctx);   // an asynchronous read load must have landed
        if (rb_ != AG2_OK) return rb_;
    }
    std::vector<int64_t> last(1);
    CK(cudaMemcpy(last.data(), (const int64_t *)ctx->read_off.p + ctx->n_reads, 8, cudaMemcpyDeviceToHost));
    const int64_t groups = last[0] >> 5;
    if (groups > 0) {
        kmer_abundance_kernel<<<grid_for(groups * 32, 256, ctx->sm_count), 256, 0, ctx->stream>>>(
            (const uint32_t *)ctx->reads2.p, (const int64_t *)ctx->read_off.p, (const int32_t *)ctx->read_len.p, ctx->n_reads, groups,
            ctx->km_k, (uint32_t *)ctx->km_table.p);
        CK(cudaGetLastError());
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return AG2_OK;
}

// The histogram merge of the multi-GPU kmer_counter: every GPU counts the k-mers of its read batches into
// its own 4^k table; dst += src, the kernel on dst's GPU reading src's table over NVLink peer memory (through a staging
// copy when the two devices cannot access each other, a plain read when both contexts share a GPU).
int ag2_kmer_merge(ag2_ctx *dst, ag2_ctx *src)
{
    ag2_ctx *ctx = dst;
    if (!dst || !src || dst == src) return fail(ctx, AG2_EINVAL, "ag2_kmer_merge: bad argument");
    if (!dst->km_k || dst->km_k != src->km_k) return fail(ctx, AG2_ESTATE, "ag2_kmer_merge: both contexts need ag2_kmer_begin with the same k");
    const int64_t nbins = 1ll << (2 * dst->km_k);
    CK(cudaSetDevice(src->device));
    CK(cudaStreamSynchronize(src->stream));
    CK(cudaSetDevice(dst->device));
    const uint32_t *from = (const uint32_t *)src->km_table.p;
    DevBuf stage;
    if (src->device != dst->device) {
        int can = 0;
        CK(cudaDeviceCanAccessPeer(&can, dst->device, src->device));
        if (can) {
            const cudaError_t e = cudaDeviceEnablePeerAccess(src->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(ctx, AG2_ECUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
            cudaGetLastError();
        } else {
            CK(cudaMalloc(&stage.p, (size_t)nbins * 4));
            CK(cudaMemcpyPeerAsync(stage.p, dst->device, src->km_table.p, src->device, (size_t)nbins * 4, dst->stream));
            from = (const uint32_t *)stage.p;
        }
    }
    kmer_table_add_kernel<<<grid_for(nbins / 4, 256, dst->sm_count), 256, 0, dst->stream>>>((uint32_t *)dst->km_table.p, from, nbins);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(dst->stream));
    if (stage.p) cudaFree(stage.p);
    return AG2_OK;
}

int ag2_kmer_solid(ag2_ctx *ctx, double threshold, int64_t *min_abundance, int64_t *n_solid)
{
    if (!ctx || !ctx->km_k) return fail(ctx, AG2_ESTATE, "ag2_kmer_solid: call ag2_kmer_begin first");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int64_t nbins = 1ll << (2 * ctx->km_k);
    RESERVE(ctx->km_hist, (size_t)(kAbWindow + 2) * 8);
    unsigned long long *hist = (unsigned long long *)ctx->km_hist.p;
    unsigned int *d_max = (unsigned int *)(hist + kAbWindow);
    // the abundance cut (:58-77): ascending abundances until 1 - bins_so_far / 4^k <= threshold
    std::vector<unsigned long long> h((size_t)kAbWindow + 1);
    unsigned long long sum = 0;
    int64_t cut = 0;
    bool found = false;
    for (uint32_t lo = 0; !found; lo += kAbWindow) {
        CK(cudaMemsetAsync(hist, 0, (size_t)(kAbWindow + 2) * 8, st));
        abundance_hist_kernel<<<ctx->sm_count * 4, 512, 0, st>>>((const uint32_t *)ctx->km_table.p, nbins, lo, hist, d_max);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(h.data(), hist, (size_t)(kAbWindow + 1) * 8, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        const unsigned int mx = (unsigned int)h[(size_t)kAbWindow];
        for (int a = 0; a < kAbWindow; ++a) {
            if (!h[(size_t)a]) continue;
            sum += h[(size_t)a];
            if (1 - sum * 1.0 / (double)nbins <= threshold) {
                cut = (int64_t)lo + a;
                found = true;
                break;
            }
        }
        if (!found && (uint64_t)lo + kAbWindow > mx) break; // no abundance qualified: the reference keeps minAbundance = 0
    }
    // solid k-mers, ascending: tiles of the 4^13-element scan (the flags reuse the scan of the index build)
    RESERVE(ctx->km_flags, (size_t)kNCodes * 4);
    RESERVE(ctx->km_offs, (size_t)(kNCodes + 1) * 4);
    int64_t total = 0;
    // pass 1: count, pass 2: scatter
    std::vector<int64_t> slab_count;
    const int64_t slab = nbins < kNCodes ? nbins : kNCodes;
    for (int pass = 0; pass < 2; ++pass) {
        int64_t base = 0;
        if (pass == 1) RESERVE(ctx->km_out, (size_t)(total + 1) * 8);
        for (int64_t first = 0, si = 0; first < nbins; first += slab, ++si) {
            solid_flags_kernel<<<grid_for(slab, 256, ctx->sm_count), 256, 0, st>>>((const uint32_t *)ctx->km_table.p + first, slab, (uint32_t)cut,
                                                                                   (int32_t *)ctx->km_flags.p);
            if (slab < kNCodes) CK(cudaMemsetAsync((int32_t *)ctx->km_flags.p + slab, 0, (size_t)(kNCodes - slab) * 4, st));
            int64_t cnt = 0;
            int rc = scan_counts(ctx, (const int32_t *)ctx->km_flags.p, (uint32_t *)ctx->km_offs.p, &cnt);
            if (rc != AG2_OK) return rc;
            if (pass == 0) {
                total += cnt;
            } else {
                solid_scatter_kernel<<<grid_for(slab, 256, ctx->sm_count), 256, 0, st>>>((const int32_t *)ctx->km_flags.p,
                                                                                        (const uint32_t *)ctx->km_offs.p, first, slab,
                                                                                        (uint64_t)base, (uint64_t *)ctx->km_out.p);
                CK(cudaGetLastError());
                base += cnt;
            }
        }
    }
    CK(cudaStreamSynchronize(st));
    ctx->km_n_solid = total;
    if (min_abundance) *min_abundance = cut;
    if (n_solid) *n_solid = total;
    return AG2_OK;
}

int ag2_kmer_fetch(ag2_ctx *ctx, uint64_t *codes_out, int64_t cap)
{
    if (!ctx || !codes_out) return fail(ctx, AG2_EINVAL, "ag2_kmer_fetch: bad argument");
    if (cap < ctx->km_n_solid) return fail(ctx, AG2_ECAP, "ag2_kmer_fetch: need room for %ld codes", (long)ctx->km_n_solid);
    CK(cudaSetDevice(ctx->device));
    if (ctx->km_n_solid) CK(cudaMemcpy(codes_out, ctx->km_out.p, (size_t)ctx->km_n_solid * 8, cudaMemcpyDeviceToHost));
    return AG2_OK;
}

} // extern "C"
