// pagraph_kernels.cuh -- device side of the A-Bruijn graph build, sm_100a.
//
// All of it is integer / byte work bound by HBM (streams) or by L2 (the solid-k-mer bitmap, the contig position table);
// nothing here is a contraction, so no tensor cores.  Reference paths are relative to PAGraph/src/tools/.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace ag2pg {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

// ---- read packing: CompressedSeq (seq/CompressedSeq.cpp:7-40): C/c 1, G/g 2, T/t 3, anything else 0 -----------------
// 32 bases per uint64, base b of a read at bits 2*(b%32) of word (off + b)/32; every read starts on a word boundary.
__global__ void pg_pack_reads_kernel(const char* __restrict__ ascii, const int64_t* __restrict__ offs,
                                     const int64_t* __restrict__ poff, int64_t n_reads, int64_t n_words,
                                     unsigned long long* __restrict__ out)
{
    for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < n_words; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t pb = g << 5;
        int64_t lo = 0, hi = n_reads - 1;   // last read with poff[r] <= pb
        while (lo < hi) {
            const int64_t mid = (lo + hi + 1) >> 1;
            if (poff[mid] <= pb) lo = mid;
            else hi = mid - 1;
        }
        const int64_t len = offs[lo + 1] - offs[lo];
        const int64_t local = pb - poff[lo];
        const char* src = ascii + offs[lo] + local;
        unsigned long long w = 0;
#pragma unroll 8
        for (int i = 0; i < 32; ++i) {
            if (local + i < len) {
                const unsigned c = (unsigned char)src[i] | 0x20u;   // fold case
                const unsigned long long code = c == 'c' ? 1 : c == 'g' ? 2 : c == 't' ? 3 : 0;
                w |= code << (2 * i);
            }
        }
        out[g] = w;
    }
}

// ---- B2: solid k-mer set -> dense vertex index ------------------------------------------------------------------------
// searchDenseIndex (graph/PABruijnGraph.cpp:98-104) is an unordered_map lookup per read base.  Here: a 4^k-bit bitmap
// (32 MB at k = 14: L2 resident on B200) + a per-word rank; dense index = rank of the code in the sorted unique set.
struct VertexSet {
    const unsigned long long* bitmap;   // 4^k bits, or nullptr -> binary search (k > 16)
    const uint32_t* rank;               // set bits before word w
    const unsigned long long* codes;    // sorted unique words of the solid file
    int64_t n;
    int k;
};

__device__ __forceinline__ bool vertex_lookup(const VertexSet& vs, unsigned long long code, uint32_t& v)
{
    if (vs.bitmap) {
        const unsigned long long w = __ldg(vs.bitmap + (code >> 6));
        const unsigned b = (unsigned)(code & 63);
        if (!((w >> b) & 1ull)) return false;
        v = __ldg(vs.rank + (code >> 6)) + (uint32_t)__popcll(w & ((1ull << b) - 1ull));
        return true;
    }
    int64_t lo = 0, hi = vs.n;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (vs.codes[mid] < code) lo = mid + 1;
        else hi = mid;
    }
    if (lo < vs.n && vs.codes[lo] == code) { v = (uint32_t)lo; return true; }
    return false;
}

__global__ void pg_bitmap_fill_kernel(const unsigned long long* __restrict__ codes, int64_t n, unsigned long long limit,
                                      unsigned long long* __restrict__ bitmap)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long c = codes[i];
        if (c < limit) atomicOr(bitmap + (c >> 6), 1ull << (c & 63));
    }
}

__global__ void pg_popc_kernel(const unsigned long long* __restrict__ bitmap, int64_t n_words, uint32_t* __restrict__ out)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_words; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (uint32_t)__popcll(bitmap[i]);
}

__global__ void pg_flag_unique_kernel(const unsigned long long* __restrict__ sorted, int64_t n, uint32_t* __restrict__ flag)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        flag[i] = (i == 0 || sorted[i] != sorted[i - 1]) ? 1u : 0u;
}

__global__ void pg_scatter_unique_kernel(const unsigned long long* __restrict__ sorted, const uint32_t* __restrict__ flag,
                                         const uint32_t* __restrict__ pos, int64_t n, unsigned long long* __restrict__ out)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if (flag[i]) out[pos[i]] = sorted[i];
}

// ---- B3 + B5: alignment columns -> target position of every query base -------------------------------------------------
// One segment = one (alignment, read strand) pair that passed the filters of Aligner::parseToCtg / parseToRef
// (align/Aligner.tcc:24-171).  ParseAlignTools::parseDiff (ParseAlignTools.cpp:8-26) classifies a column from the two
// text lines, exactAlign (ParseAlignTools.tcc:46-70) walks the columns: a column with a query base calls the functor
// with (curQuery, curRef); the target cursor advances unless the column is a gap in the target.
struct Segment {
    int64_t q_off;       // query line in the text buffer
    int64_t t_off;       // target line
    int64_t tpos_off;    // first entry in the tpos array
    uint64_t tb;         // target cursor at the start of the walk
    uint32_t rb;         // read position (on the segment's strand) of the first query base
    int32_t ncols;
    int32_t target;      // contig or reference index
    int32_t backward;    // walk the columns from the last to the first (exactAlign forward = false)
    int32_t negative;    // contig used in reverse orientation: PositionMapper::dualToSingle gets -(idx + 1)
};

// one warp per segment, 32 columns per step; the two running counts are ballots + popc (no shared memory)
__global__ void pg_walk_kernel(const Segment* __restrict__ segs, int64_t n_segs, const char* __restrict__ text,
                               uint32_t* __restrict__ tpos, int32_t* __restrict__ nq_out)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const unsigned below = (1u << lane) - 1u;
    for (int64_t s = warp; s < n_segs; s += n_warps) {
        const Segment sg = segs[s];
        const char* q = text + sg.q_off;
        const char* t = text + sg.t_off;
        uint32_t emitted = 0;
        uint64_t adv = 0;
        for (int j0 = 0; j0 < sg.ncols; j0 += 32) {
            const int j = j0 + lane;
            bool emit = false, ra = false;
            if (j < sg.ncols) {
                const int c = sg.backward ? sg.ncols - 1 - j : j;
                const char qc = q[c], tc = t[c];
                emit = qc != '-';
                ra = !(emit && tc == '-');
            }
            const unsigned me = __ballot_sync(kFull, emit), ma = __ballot_sync(kFull, ra);
            if (emit) tpos[sg.tpos_off + emitted + __popc(me & below)] = (uint32_t)(sg.tb + adv + __popc(ma & below));
            emitted += __popc(me);
            adv += __popc(ma);
        }
        if (lane == 0) nq_out[s] = (int32_t)emitted;
    }
}

// ---- B4 table: contig base -> list of packed reference positions (AlignReference, align/AlignReference.cpp) -----------
struct CtgTable {
    const int64_t* ctg_base;    // [n_ctg + 1] first base slot of every contig
    const uint32_t* base_off;   // [bases + 1] first entry of every base
    const uint32_t* entry;      // packed reference positions (PositionMapper::dualToSingle), 0 = (0,0)
    const int64_t* ctg_len;
    const uint64_t* ctg_start;  // PositionMapper::_startPos of the contig database
    const uint64_t* ref_start;  // ... of the reference database
};

// ---- B5-B7: one warp per (read, strand) lane ---------------------------------------------------------------------------
struct Lane {
    int32_t read;
    int32_t strand;      // 0 = read as given, 1 = reverse complement (SeqInf::toString(false))
    int32_t seg_begin, seg_end;
};

struct LaneCounts {      // per lane: samples, tuples, edges
    unsigned long long samples, tuples, edges;
};

__device__ __forceinline__ unsigned long long revpairs64(unsigned long long x)
{
    x = __brevll(x);
    return ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);
}

// k-mer code (KmerHelper::kmer2Code, kmer/KmerHelper.cpp:7-25: first base most significant) of the k-mer starting at
// position i of the read (strand 0) or of its reverse complement (strand 1)
__device__ __forceinline__ unsigned long long kmer_code(const unsigned long long* __restrict__ words, int64_t off, int len,
                                                        int k, int strand, int i)
{
    const int64_t p = off + (strand ? len - k - i : i);
    const unsigned sh = (unsigned)(p & 31) * 2u;
    const unsigned long long lo = __ldg(words + (p >> 5));
    unsigned long long win = lo >> sh;
    if (sh) win |= __ldg(words + (p >> 5) + 1) << (64 - sh);
    const unsigned long long mask = k >= 32 ? ~0ull : ((1ull << (2 * k)) - 1ull);
    if (strand) return ~win & mask;                 // window read LSB-first is the reversed k-mer; ~ complements
    return revpairs64(win) >> (64 - 2 * k);
}

struct ExtractArgs {
    const Lane* lanes;
    int64_t n_lanes;
    const Segment* segs;
    const int32_t* seg_nq;
    const uint32_t* tpos;
    const unsigned long long* read_words;
    const int64_t* read_off;     // packed base offset of every read
    const int32_t* read_len;
    VertexSet vs;
    CtgTable tab;
    int phase;                   // 0 read->contig (lists come from the table), 1 read->reference (one entry per base)
    int outer;                   // outerSample
    LaneCounts* counts;          // COUNT pass output / EMIT pass: exclusive prefix over lanes
    uint32_t *t_vertex, *t_ctg, *t_ref;   // tuple stream
    uint32_t *e_from, *e_to;              // edge stream
    int32_t* e_step;
    unsigned long long tuple_base, edge_base;   // where this phase starts in the streams
};

// number of DualPos the read base at position i of this segment contributes (queryContig / parseToRef functor)
__device__ __forceinline__ uint32_t seg_count(const ExtractArgs& a, const Segment& sg, int nq, int i, uint32_t& tp)
{
    if ((uint32_t)i < sg.rb || (uint32_t)i - sg.rb >= (uint32_t)nq) return 0;
    tp = __ldg(a.tpos + sg.tpos_off + ((uint32_t)i - sg.rb));
    if (a.phase) return 1;
    if ((int64_t)tp >= a.tab.ctg_len[sg.target]) return 0;
    const int64_t slot = a.tab.ctg_base[sg.target] + tp;
    return __ldg(a.tab.base_off + slot + 1) - __ldg(a.tab.base_off + slot);
}

// PABruijnGraph::addPositionAndEdge (graph/PABruijnGraph.cpp:238-257) with sampleSequence (PABruijnGraph.tcc:6-27):
// positions with a non-empty list and a solid k-mer, at least `outer` apart (greedy from the left), append their list to
// the vertex and an edge (previous sample -> this sample, distance).  32 positions per step: the list sizes and the
// solid test are per thread, the greedy pick runs on the ballot of candidates (identical in every thread of the warp).
template <bool EMIT>
__global__ void __launch_bounds__(256) pg_extract_kernel(ExtractArgs a)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const unsigned below = (1u << lane) - 1u;
    for (int64_t L = warp; L < a.n_lanes; L += n_warps) {
        const Lane ln = a.lanes[L];
        const int len = a.read_len[ln.read];
        const int64_t off = a.read_off[ln.read];
        const int npos = len >= a.vs.k ? len - a.vs.k + 1 : 0;
        unsigned long long n_samples = 0, n_tuples = 0;
        long long next_free = 0;            // first position the next sample may take
        uint32_t last_v = 0;
        int last_i = 0;
        unsigned long long t_at = 0, s_at = 0, e_base = 0;
        if (EMIT) {
            const LaneCounts c = a.counts[L];
            t_at = a.tuple_base + c.tuples;
            s_at = 0;
            e_base = a.edge_base + c.edges;
        }
        for (int i0 = 0; i0 < npos; i0 += 32) {
            const int i = i0 + lane;
            uint32_t cnt = 0, v = 0;
            bool cand = false;
            if (i < npos) {
                for (int s = ln.seg_begin; s < ln.seg_end; ++s) {
                    uint32_t tp;
                    cnt += seg_count(a, a.segs[s], a.seg_nq[s], i, tp);
                }
                if (cnt) cand = vertex_lookup(a.vs, kmer_code(a.read_words, off, len, a.vs.k, ln.strand, i), v);
            }
            unsigned m = __ballot_sync(kFull, cand);
            unsigned picked = 0;
            {
                const long long rel = next_free - i0;
                unsigned mm = rel <= 0 ? m : (rel < 32 ? m & (kFull << rel) : 0u);
                while (mm) {
                    const int b = __ffs(mm) - 1;
                    picked |= 1u << b;
                    next_free = (long long)i0 + b + a.outer;
                    const int sh = b + a.outer;
                    mm = sh < 32 ? mm & (kFull << sh) : 0u;
                }
            }
            const bool samp = (picked >> lane) & 1u;
            // exclusive prefix of the list sizes of the sampled positions
            uint32_t x = samp ? cnt : 0u, incl = x;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t y = __shfl_up_sync(kFull, incl, d);
                if (lane >= d) incl += y;
            }
            const uint32_t total = __shfl_sync(kFull, incl, 31);
            if (EMIT) {
                const unsigned before = picked & below;
                const int src = before ? 31 - __clz(before) : 0;
                const uint32_t pv = __shfl_sync(kFull, v, src);
                if (samp) {
                    const unsigned long long sidx = s_at + __popc(before);
                    if (sidx > 0) {
                        const unsigned long long e = e_base + sidx - 1;
                        a.e_from[e] = before ? pv : last_v;
                        a.e_to[e] = v;
                        a.e_step[e] = i - (before ? i0 + src : last_i);
                    }
                    unsigned long long w = t_at + (incl - x);
                    for (int s = ln.seg_begin; s < ln.seg_end; ++s) {
                        const Segment sg = a.segs[s];
                        uint32_t tp;
                        const uint32_t c = seg_count(a, sg, a.seg_nq[s], i, tp);
                        if (!c) continue;
                        if (a.phase) {
                            a.t_vertex[w] = v;
                            a.t_ctg[w] = 0;
                            a.t_ref[w] = (uint32_t)(a.tab.ref_start[sg.target] + tp);
                            ++w;
                        } else {
                            const uint32_t cp = (uint32_t)(a.tab.ctg_start[sg.target] +
                                                           (sg.negative ? 2ull * (uint64_t)a.tab.ctg_len[sg.target] : 0ull) + tp);
                            const int64_t slot = a.tab.ctg_base[sg.target] + tp;
                            const uint32_t e0 = a.tab.base_off[slot];
                            for (uint32_t e = 0; e < c; ++e, ++w) {
                                a.t_vertex[w] = v;
                                a.t_ctg[w] = cp;
                                a.t_ref[w] = a.tab.entry[e0 + e];
                            }
                        }
                    }
                }
                if (picked) {
                    const int lastl = 31 - __clz(picked);
                    last_v = __shfl_sync(kFull, v, lastl);
                    last_i = i0 + lastl;
                }
                t_at += total;
                s_at += __popc(picked);
            } else {
                n_samples += __popc(picked);
                n_tuples += total;
            }
        }
        if (!EMIT && lane == 0) {
            LaneCounts c;
            c.samples = n_samples;
            c.tuples = n_tuples;
            c.edges = n_samples ? n_samples - 1 : 0;
            a.counts[L] = c;
        }
    }
}

// exclusive scan of the per-lane counts (three running sums); single CTA, plumbing
__global__ void pg_scan_counts_kernel(LaneCounts* c, int64_t n, LaneCounts* total)
{
    __shared__ unsigned long long part[3][1024];
    const int t = threadIdx.x;
    const int64_t per = (n + blockDim.x - 1) / blockDim.x;
    const int64_t lo = min(n, t * per), hi = min(n, lo + per);
    unsigned long long s0 = 0, s1 = 0, s2 = 0;
    for (int64_t i = lo; i < hi; ++i) { s0 += c[i].samples; s1 += c[i].tuples; s2 += c[i].edges; }
    part[0][t] = s0; part[1][t] = s1; part[2][t] = s2;
    __syncthreads();
    if (t < 3) {
        unsigned long long acc = 0;
        for (unsigned i = 0; i < blockDim.x; ++i) { const unsigned long long v = part[t][i]; part[t][i] = acc; acc += v; }
        if (t == 0) total->samples = acc;
        if (t == 1) total->tuples = acc;
        if (t == 2) total->edges = acc;
    }
    __syncthreads();
    s0 = part[0][t]; s1 = part[1][t]; s2 = part[2][t];
    for (int64_t i = lo; i < hi; ++i) {
        const LaneCounts v = c[i];
        c[i].samples = s0; c[i].tuples = s1; c[i].edges = s2;
        s0 += v.samples; s1 += v.tuples; s2 += v.edges;
    }
}

// ---- B8: per-vertex epsilon join ----------------------------------------------------------------------------------------
__global__ void pg_hist_kernel(const uint32_t* __restrict__ key, int64_t n, uint32_t* __restrict__ cnt)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(cnt + key[i], 1u);
}

__global__ void pg_pack_pairs_kernel(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, int64_t n,
                                     unsigned long long* __restrict__ out)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (unsigned long long)a[i] | ((unsigned long long)b[i] << 32);
}

// isPosSimilar (graph/PABruijnGraph.cpp:379-383) + the predicate of mergeKmerPosition (:259-274), per coordinate
__device__ __forceinline__ bool coord_similar(uint32_t l, uint32_t r, uint32_t dev)
{
    if (l == 0 || r == 0) return l == 0 && r == 0;
    return (l > r ? l - r : r - l) <= dev;
}

constexpr int kRepSmem = 256;     // representatives kept in shared memory per warp; the rest live in the scratch arrays
constexpr int kJoinWarps = 8;

// KMerAdjNode::cluster (node/KMerAdjNode.tcc:74-111): greedy first-fit in insertion order -- an item joins the FIRST
// representative it is similar to, else becomes the next representative; counts are uint16 and wrap.  Then
// sortWithCount (:115-136) by (ctg, ref).  One warp per vertex: the items go one by one, the representatives are
// compared 32 at a time and __ffs of the ballot is the first fit.  Running the reference's two merges (after phase 1
// and after phase 2) equals one pass over the concatenated list, because representatives are pairwise dissimilar.
__global__ void __launch_bounds__(kJoinWarps * 32)
pg_join_kernel(const uint32_t* __restrict__ seg_off, int64_t n_vertices, const unsigned long long* __restrict__ items /* ctg | ref<<32 */,
               uint32_t eps, unsigned long long* __restrict__ rep_scratch, uint32_t* __restrict__ cnt_scratch,
               uint32_t* __restrict__ out_ctg, uint32_t* __restrict__ out_ref, uint16_t* __restrict__ out_cnt,
               uint32_t* __restrict__ nrep, const uint32_t* __restrict__ list, const uint32_t* __restrict__ n_list)
{
    if (list) n_vertices = *n_list;   // the vertices pg_join_small_kernel left for a warp each
    __shared__ unsigned long long s_rep[kJoinWarps][kRepSmem];
    __shared__ uint32_t s_cnt[kJoinWarps][kRepSmem];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t warp = blockIdx.x * (int64_t)kJoinWarps + wib;
    const int64_t n_warps = (int64_t)gridDim.x * kJoinWarps;
    unsigned long long* rep = s_rep[wib];
    uint32_t* cnt = s_cnt[wib];
    for (int64_t w = warp; w < n_vertices; w += n_warps) {
        const int64_t v = list ? (int64_t)list[w] : w;
        const uint32_t b = seg_off[v], e = seg_off[v + 1];
        if (b == e) {
            if (lane == 0) nrep[v] = 0;
            continue;
        }
        unsigned long long* g_rep = rep_scratch + b;     // representative j >= kRepSmem at g_rep[j]
        uint32_t* g_cnt = cnt_scratch + b;
        uint32_t p = 0;
        for (uint32_t i = b; i < e; ++i) {
            const unsigned long long it = __ldg(items + i);
            const uint32_t ic = (uint32_t)it, ir = (uint32_t)(it >> 32);
            int hit = -1;
            for (uint32_t j0 = 0; j0 < p && hit < 0; j0 += 32) {
                const uint32_t j = j0 + lane;
                bool sim = false;
                if (j < p) {
                    const unsigned long long r = j < kRepSmem ? rep[j] : __ldcg(g_rep + j);
                    sim = coord_similar(ic, (uint32_t)r, eps) && coord_similar(ir, (uint32_t)(r >> 32), eps);
                }
                const unsigned m = __ballot_sync(kFull, sim);
                if (m) hit = (int)j0 + __ffs(m) - 1;
            }
            if (lane == 0) {
                if (hit >= 0) {
                    if (hit < kRepSmem) cnt[hit] = (cnt[hit] + 1u) & 0xffffu;
                    else __stcg(g_cnt + hit, (__ldcg(g_cnt + hit) + 1u) & 0xffffu);
                } else {
                    if (p < kRepSmem) { rep[p] = it; cnt[p] = 1u; }
                    else { __stcg(g_rep + p, it); __stcg(g_cnt + p, 1u); }
                }
            }
            if (hit < 0) ++p;
            __syncwarp();
        }
        // rank sort: representatives are distinct, so rank = number of smaller ones
        for (uint32_t j = lane; j < p; j += 32) {
            const unsigned long long r = j < kRepSmem ? rep[j] : __ldcg(g_rep + j);
            const uint32_t c = j < kRepSmem ? cnt[j] : __ldcg(g_cnt + j);
            const unsigned long long key = ((r & 0xffffffffull) << 32) | (r >> 32);   // (ctg, ref) lexicographic
            uint32_t rank = 0;
            for (uint32_t q = 0; q < p; ++q) {
                const unsigned long long o = q < kRepSmem ? rep[q] : __ldcg(g_rep + q);
                const unsigned long long ok = ((o & 0xffffffffull) << 32) | (o >> 32);
                rank += ok < key;
            }
            out_ctg[b + rank] = (uint32_t)r;
            out_ref[b + rank] = (uint32_t)(r >> 32);
            out_cnt[b + rank] = (uint16_t)c;
        }
        if (lane == 0) nrep[v] = p;
        __syncwarp();
    }
}

// The same clustering for the many vertices that hold only a few items (at k = 14 on noisy reads: 100 M vertices, two
// items each on average): one THREAD per vertex, representatives in registers; a vertex with more than kSmallJoin items
// goes to `list` for pg_join_kernel (one warp each).
constexpr int kSmallJoin = 8;
__global__ void __launch_bounds__(256)
pg_join_small_kernel(const uint32_t* __restrict__ seg_off, int64_t n_vertices, const unsigned long long* __restrict__ items, uint32_t eps,
                     uint32_t* __restrict__ out_ctg, uint32_t* __restrict__ out_ref, uint16_t* __restrict__ out_cnt, uint32_t* __restrict__ nrep,
                     uint32_t* __restrict__ list, uint32_t* __restrict__ n_list)
{
    for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n_vertices; v += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t b = seg_off[v], e = seg_off[v + 1];
        const uint32_t n = e - b;
        if (n == 0) {
            nrep[v] = 0;
            continue;
        }
        if (n > (uint32_t)kSmallJoin) {
            list[atomicAdd(n_list, 1u)] = (uint32_t)v;
            continue;
        }
        unsigned long long rep[kSmallJoin];
        uint32_t cnt[kSmallJoin];
        uint32_t p = 0;
        for (uint32_t i = 0; i < n; ++i) {
            const unsigned long long it = __ldg(items + b + i);
            const uint32_t ic = (uint32_t)it, ir = (uint32_t)(it >> 32);
            int hit = -1;
#pragma unroll
            for (int j = kSmallJoin - 1; j >= 0; --j)      // downwards: the lowest matching j stays
                if ((uint32_t)j < p && coord_similar(ic, (uint32_t)rep[j], eps) && coord_similar(ir, (uint32_t)(rep[j] >> 32), eps)) hit = j;
#pragma unroll
            for (int j = 0; j < kSmallJoin; ++j) {
                if (j == hit) cnt[j] = (cnt[j] + 1u) & 0xffffu;
                if (hit < 0 && (uint32_t)j == p) {
                    rep[j] = it;
                    cnt[j] = 1u;
                }
            }
            if (hit < 0) ++p;
        }
#pragma unroll
        for (int j = 0; j < kSmallJoin; ++j) {
            if ((uint32_t)j >= p) continue;
            const unsigned long long key = ((rep[j] & 0xffffffffull) << 32) | (rep[j] >> 32);   // (ctg, ref) lexicographic
            uint32_t rank = 0;
#pragma unroll
            for (int q = 0; q < kSmallJoin; ++q) {
                const unsigned long long ok = ((rep[q] & 0xffffffffull) << 32) | (rep[q] >> 32);
                rank += ((uint32_t)q < p && ok < key) ? 1u : 0u;
            }
            out_ctg[b + rank] = (uint32_t)rep[j];
            out_ref[b + rank] = (uint32_t)(rep[j] >> 32);
            out_cnt[b + rank] = (uint16_t)cnt[j];
        }
        nrep[v] = p;
    }
}

// compaction of the per-vertex results, one thread per slot of the sorted stream: slot i of vertex v = key[i] survives if
// it is one of the first nrep[v] of its segment
__global__ void pg_compact_kernel(const uint32_t* __restrict__ key, int64_t n_items, const uint32_t* __restrict__ seg_off,
                                  const uint32_t* __restrict__ nrep, const unsigned long long* __restrict__ pos_off,
                                  const uint32_t* __restrict__ in_ctg, const uint32_t* __restrict__ in_ref,
                                  const uint16_t* __restrict__ in_cnt, uint32_t* __restrict__ ctg, uint32_t* __restrict__ ref,
                                  uint16_t* __restrict__ cnt)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_items; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t v = key[i];
        const uint32_t r = (uint32_t)i - seg_off[v];
        if (r >= nrep[v]) continue;
        const unsigned long long o = pos_off[v] + r;
        ctg[o] = in_ctg[i];
        ref[o] = in_ref[i];
        cnt[o] = in_cnt[i];
    }
}

__global__ void pg_widen_kernel(const uint32_t* __restrict__ in, int64_t n, unsigned long long* __restrict__ out)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = in[i];
}

// ---- B8 edges: removeDuplicate (node/KMerAdjNode.tcc:46-70) = sort by (to, step) + unique, per `from` vertex ----------
__global__ void pg_edge_flag_kernel(const unsigned long long* __restrict__ from_to, const uint32_t* __restrict__ step, int64_t n,
                                    uint32_t* __restrict__ flag)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        flag[i] = (i == 0 || from_to[i] != from_to[i - 1] || step[i] != step[i - 1]) ? 1u : 0u;
}

__global__ void pg_edge_scatter_kernel(const unsigned long long* __restrict__ from_to, const uint32_t* __restrict__ step,
                                       const uint32_t* __restrict__ flag, const uint32_t* __restrict__ pos, int64_t n,
                                       uint32_t* __restrict__ to_out, int32_t* __restrict__ step_out, uint32_t* __restrict__ per_vertex)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (!flag[i]) continue;
        const unsigned long long ft = from_to[i];
        to_out[pos[i]] = (uint32_t)ft;
        step_out[pos[i]] = (int32_t)step[i];
        atomicAdd(per_vertex + (uint32_t)(ft >> 32), 1u);
    }
}

// ---- multi-GPU: owner of a vertex, gathers ------------------------------------------------------------------------------
// acc[i] += add[i] (offset arrays of the per-GPU tables: element-wise sums are the merged offsets)
__global__ void pg_add_u64_kernel(unsigned long long* __restrict__ acc, const unsigned long long* __restrict__ add, int64_t n)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) acc[i] += add[i];
}

__global__ void pg_owner_kernel(const uint32_t* __restrict__ vertex, int64_t n, uint32_t per_owner, uint32_t* __restrict__ owner,
                                uint32_t* __restrict__ index, unsigned long long* __restrict__ hist)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t o = vertex[i] / per_owner;
        owner[i] = o;
        index[i] = (uint32_t)i;
        atomicAdd(hist + o, 1ull);
    }
}

__global__ void pg_gather3_kernel(const uint32_t* __restrict__ index, int64_t n, const uint32_t* __restrict__ a, const uint32_t* __restrict__ b,
                                  const uint32_t* __restrict__ c, uint32_t* __restrict__ ao, uint32_t* __restrict__ bo, uint32_t* __restrict__ co)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t s = index[i];
        ao[i] = a[s];
        bo[i] = b[s];
        co[i] = c[s];
    }
}

}  // namespace ag2pg
