// index_kernels.cuh -- kernels of the mecat2ref+ index build and the launch
// wrappers of the seeding stage (A5-A7, device code in seed_device.cuh).
//
//   A2 build_read_index (M2R/mecat2ref_impl_large.cpp:258-399): 13-mer histogram of the concatenated read
//      prefix (k-mers span read boundaries, anything but upper-case ACGT restarts the window), counts > 128 -> 0
//   A3 creat_ref_index (:402-566): the same histogram over the reference, mask, exclusive scan -> CSR offsets,
//      fill with 1-based k-mer starts, ASCENDING inside a bucket (the seeding loop consumes hits in that
//      order), and per similarity block the sum of the read counts of its k-mers
//   A4 get_vote (:568-608)
//
// All of it is HBM-bound integer work: one thread per k-mer end position, coalesced sequence reads, scattered
// atomics into the 4^13-bin tables.  Lanes of a warp that hit the same bin (low-complexity sequence) or the same
// similarity block (adjacent positions: almost always) are merged with __match_any_sync before the atomic.
#pragma once

#include "seed_device.cuh"
#include "xdrop_device.cuh"

namespace ag2 {

constexpr int kNCodes = 1 << (2 * kSeedLen);

// one atomicAdd per distinct key among the calling lanes; `on` lanes contribute `val` to table[key]
__device__ __forceinline__ void warp_aggregated_add(int32_t *table, int key, int val, bool on)
{
    const unsigned active = __activemask();
    const int lane = threadIdx.x & 31;
    const unsigned peers = __match_any_sync(active, on ? key : -1 - lane);
    if (!on) return;
    int sum = val;
    // peers is small in practice (1, or a run of adjacent lanes); fold with shuffles over the peer set
    unsigned rest = peers & ~(1u << lane);
    const int leader = __ffs(peers) - 1;
    if (rest) {
        // all peers must take part in the same shuffles: iterate over the peer mask in lock step
        sum = 0;
        for (unsigned m = peers; m; m &= m - 1) {
            const int src = __ffs(m) - 1;
            sum += __shfl_sync(peers, val, src);
        }
    }
    if (lane == leader) atomicAdd(&table[key], sum);
}

// A2: histogram over concatenated ASCII (the read batch as loaded, prefix of length n)
__global__ void ascii_kmer_hist_kernel(const char *__restrict__ s, int64_t n, int32_t *__restrict__ cnt)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t rounds = (n + stride - 1) / stride;
    for (int64_t r = 0; r < rounds; ++r) { // whole warps stay in the loop: the aggregation uses warp collectives
        const int64_t i = r * stride + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
        bool ok = i >= kSeedLen - 1 && i < n;
        int code = 0;
        if (ok) {
#pragma unroll
            for (int q = 0; q < kSeedLen; ++q) {
                int t;
                switch (s[i - (kSeedLen - 1) + q]) {
                case 'A': t = 0; break;
                case 'T': t = 1; break;
                case 'C': t = 2; break;
                case 'G': t = 3; break;
                default: t = 4; break;
                }
                if (t == 4) ok = false;
                code = (code << 2) | (t & 3);
            }
        }
        warp_aggregated_add(cnt, code, 1, ok);
    }
}

// 13-mer ending at base i of a packed sequence, in atcttrans code; false if the window holds an irregular base
__device__ __forceinline__ bool packed_kmer(const uint32_t *seq2, const uint32_t *irr, int64_t i, int &code)
{
    const int64_t s = i - (kSeedLen - 1);
    const int64_t iw = s >> 5;
    const uint64_t ib = ((uint64_t)irr[iw + 1] << 32 | irr[iw]) >> (s & 31);
    if (ib & ((1u << kSeedLen) - 1)) return false;
    const int64_t w = s >> 4;
    const uint64_t bits = ((uint64_t)seq2[w + 1] << 32 | seq2[w]) >> (2 * (s & 15));
    int c = 0;
#pragma unroll
    for (int q = 0; q < kSeedLen; ++q) c = (c << 2) | atct_of_code((int)(bits >> (2 * q)) & 3);
    code = c;
    return true;
}

// A3, three passes over the reference k-mers: MODE 0 count, 1 fill the CSR, 2 per-block read-count sums
template <int MODE>
__global__ void ref_kmer_kernel(const uint32_t *__restrict__ seq2, const uint32_t *__restrict__ irr, int64_t n,
                                int32_t *__restrict__ cnt, const uint32_t *__restrict__ off, int32_t *__restrict__ fill,
                                uint32_t *__restrict__ pos, const int32_t *__restrict__ rcnt, int32_t *__restrict__ kcount, int cbl)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t rounds = (n + stride - 1) / stride;
    for (int64_t r = 0; r < rounds; ++r) {
        const int64_t i = r * stride + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
        int code = 0;
        const bool ok = i >= kSeedLen - 1 && i < n && packed_kmer(seq2, irr, i, code);
        if (MODE == 0) {
            warp_aggregated_add(cnt, code, 1, ok);
        } else if (MODE == 1) {
            if (ok && cnt[code] > 0) pos[off[code] + (uint32_t)atomicAdd(&fill[code], 1)] = (uint32_t)(i + 2 - kSeedLen);
        } else {
            const int add = ok ? rcnt[code] : 0;
            warp_aggregated_add(kcount, (int)((i + 2 - kSeedLen) / cbl), add, ok && add > 0);
        }
    }
}

__global__ void mask_counts_kernel(int32_t *cnt, int64_t n) // sumvalue_x (:84-93)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if (cnt[i] > 128) cnt[i] = 0;
}

// exclusive scan of n int32 -> uint32 in three coalesced passes; tile = 1024 threads x 16 elements
constexpr int kScanTile = 1024 * 16;
__global__ void scan_tile_sums_kernel(const int32_t *in, int64_t n, uint32_t *tile_sums)
{
    __shared__ uint32_t red[32];
    const int64_t base = (int64_t)blockIdx.x * kScanTile;
    uint32_t s = 0;
    for (int k = threadIdx.x; k < kScanTile; k += 1024)
        if (base + k < n) s += (uint32_t)in[base + k];
    for (int d = 16; d; d >>= 1) s += __shfl_down_sync(0xffffffffu, s, d);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = red[threadIdx.x];
        for (int d = 16; d; d >>= 1) s += __shfl_down_sync(0xffffffffu, s, d);
        if (threadIdx.x == 0) tile_sums[blockIdx.x] = s;
    }
}
__global__ void scan_tiles_kernel(uint32_t *tile_sums, int n_tiles, uint32_t *total) // single block, n_tiles <= 8192
{
    __shared__ uint32_t part[1024];
    const int per = (n_tiles + 1023) / 1024;
    const int lo = min(n_tiles, (int)threadIdx.x * per), hi = min(n_tiles, lo + per);
    uint32_t s = 0;
    for (int i = lo; i < hi; ++i) s += tile_sums[i];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t acc = 0;
        for (int i = 0; i < 1024; ++i) {
            const uint32_t v = part[i];
            part[i] = acc;
            acc += v;
        }
        *total = acc;
    }
    __syncthreads();
    s = part[threadIdx.x];
    for (int i = lo; i < hi; ++i) {
        const uint32_t v = tile_sums[i];
        tile_sums[i] = s;
        s += v;
    }
}
__global__ void scan_apply_kernel(const int32_t *in, int64_t n, const uint32_t *tile_offs, uint32_t *out)
{
    // each warp owns 512 consecutive elements of the tile: lane-strided reads, shuffle scan per 32
    const int64_t base = (int64_t)blockIdx.x * kScanTile;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __shared__ uint32_t wsum[32];
    uint32_t vals[16], tot = 0;
    const int64_t wbase = base + warp * 512;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int64_t idx = wbase + k * 32 + lane;
        vals[k] = idx < n ? (uint32_t)in[idx] : 0u;
        tot += vals[k];
    }
    uint32_t wt = tot;
    for (int d = 16; d; d >>= 1) wt += __shfl_down_sync(0xffffffffu, wt, d);
    if (lane == 0) wsum[warp] = wt;
    __syncthreads();
    uint32_t wo = tile_offs[blockIdx.x];
    for (int w = 0; w < warp; ++w) wo += wsum[w];
    uint32_t run = wo;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        uint32_t v = vals[k], inc = v;
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += o;
        }
        const int64_t idx = wbase + k * 32 + lane;
        if (idx < n) out[idx] = run + inc - v;
        run += __shfl_sync(0xffffffffu, inc, 31);
    }
}

// the fill kernel places a bucket's positions in arrival order; the reference's are ascending (:516-541)
__global__ void sort_buckets_kernel(const int32_t *cnt, const uint32_t *off, uint32_t *pos, int64_t n)
{
    for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < n; c += (int64_t)gridDim.x * blockDim.x) {
        const int k = cnt[c];
        if (k < 2) continue;
        uint32_t *p = pos + off[c];
        for (int i = 1; i < k; ++i) {
            const uint32_t v = p[i];
            int j = i;
            while (j > 0 && p[j - 1] > v) {
                p[j] = p[j - 1];
                --j;
            }
            p[j] = v;
        }
    }
}

__global__ void sum_kcount_kernel(const int32_t *kcount, int64_t nblk, unsigned long long *total)
{
    unsigned long long s = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nblk; i += (int64_t)gridDim.x * blockDim.x) s += (unsigned)kcount[i];
    for (int d = 16; d; d >>= 1) s += __shfl_down_sync(0xffffffffu, s, d);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(total, s);
}

// get_vote (:568-608): ave = (float)(total / nblk) with the integer division first; the comparisons against
// 2*alpha and beta are done in double as in the reference
__global__ void vote_kernel(const int32_t *kcount, int64_t nblk, const unsigned long long *total, double alpha, double beta, float *vote)
{
    const float ave = (float)((long long)*total / (long long)nblk);
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < nblk + 10; j += (int64_t)gridDim.x * blockDim.x) {
        float v = 0.0f; // the 10 slack entries stay 0 (:455-459)
        if (j < nblk) {
            if (ave == 0 || kcount[j] == 0) {
                v = 1.0f;
            } else {
                const float dev = fdiv_rn((float)kcount[j], ave);
                v = ((double)dev < alpha * 2 || (double)dev > beta) ? dev : 1.0f;
            }
        }
        vote[j] = v;
    }
}

// ---- seeding stage (A5-A7) ----------------------------------------------------------------------
__global__ void seeds_to_candidates_kernel(const SeedCand *cands, const int32_t *ncand, const int64_t *prefix, int64_t n_reads,
                                           int maxc, Candidate *out)
{
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n_reads; r += (int64_t)gridDim.x * blockDim.x) {
        const int64_t base = prefix[r];
        for (int i = 0; i < ncand[r]; ++i) {
            const SeedCand &c = cands[r * maxc + i];
            Candidate o;
            o.read = (int32_t)r;
            o.strand = c.chain == 'F' ? 0 : 1;
            o.loc1 = c.loc1;
            o.loc2 = (int32_t)c.loc2;
            o.score = c.score;
            out[base + i] = o;
        }
    }
}

__global__ void widen_i32_kernel(const int32_t *in, int64_t n, int64_t *out)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = in[i];
}

} // namespace ag2
