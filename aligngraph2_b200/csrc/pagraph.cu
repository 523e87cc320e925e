// pagraph.cu -- the C ABI of include/ag2_pagraph.h: PAGraph's A-Bruijn graph build on sm_100a.
//
// Host side of this file = what PositionProcessor / Aligner do per config block around the per-base work (score sorts,
// filters, the contig->reference position table); device side = pagraph_kernels.cuh.  No CPU fallback: every entry
// point needs a live CUDA device.  Reference paths are relative to PAGraph/src/tools/ (PGM = ../main).
#include "../../include/ag2_b200.h"
#include "../../include/ag2_pagraph.h"
#include "pagraph_kernels.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

using namespace ag2pg;

namespace {

// Device allocations of a handle are recycled: a build allocates and frees ~25 arrays of up to several GB, and cudaMalloc /
// cudaFree of that much is slower than the kernels that use it (40 k reads: ~160 of 284 ms per build).  A block goes back to
// the pool of the handle whose entry point is running (PoolScope) and is handed out again for a request it fits within
// 25 %; blocks of a handle are only ever used on that handle's stream, so stream order keeps reuse safe.
struct DevPool {
    struct Block {
        void* p;
        size_t bytes;
        int device;
    };
    std::vector<Block> free_;
    size_t cached = 0;
    static constexpr size_t kCap = (size_t)64 << 30;   // bytes kept at most
    void* take(size_t bytes, int device, size_t* got)
    {
        size_t best = free_.size();
        for (size_t i = 0; i < free_.size(); ++i)
            if (free_[i].device == device && free_[i].bytes >= bytes && free_[i].bytes <= bytes + bytes / 4 + ((size_t)1 << 20) &&
                (best == free_.size() || free_[i].bytes < free_[best].bytes))
                best = i;
        if (best == free_.size()) return nullptr;
        void* p = free_[best].p;
        *got = free_[best].bytes;
        cached -= free_[best].bytes;
        free_.erase(free_.begin() + (long)best);
        return p;
    }
    bool give(void* p, size_t bytes, int device)
    {
        if (cached + bytes > kCap) return false;
        free_.push_back({p, bytes, device});
        cached += bytes;
        return true;
    }
    void flush()
    {
        int cur = 0;
        cudaGetDevice(&cur);
        for (Block& b : free_) {
            cudaSetDevice(b.device);
            cudaFree(b.p);
        }
        cudaSetDevice(cur);
        free_.clear();
        cached = 0;
    }
};
thread_local DevPool* tl_pool = nullptr;
struct PoolScope {
    DevPool* prev;
    explicit PoolScope(DevPool* p) : prev(tl_pool) { tl_pool = p; }
    ~PoolScope() { tl_pool = prev; }
};

template <class T>
struct Dev {
    T* p = nullptr;
    int64_t n = 0;
    size_t bytes = 0;   // of the underlying block
    int device = -1;
    Dev() = default;
    Dev(const Dev&) = delete;
    Dev& operator=(const Dev&) = delete;
    ~Dev() { release(); }
    void release()
    {
        if (p && !(tl_pool && device >= 0 && tl_pool->give(p, bytes, device))) cudaFree(p);
        p = nullptr;
        n = 0;
        bytes = 0;
    }
    cudaError_t alloc(int64_t count)
    {
        release();
        if (count <= 0) count = 1;
        const size_t want = (((size_t)count * sizeof(T)) + 511) & ~(size_t)511;
        cudaGetDevice(&device);
        if (tl_pool) {
            size_t got = 0;
            if (void* q = tl_pool->take(want, device, &got)) {
                p = (T*)q;
                n = count;
                bytes = got;
                return cudaSuccess;
            }
        }
        cudaError_t e = cudaMalloc((void**)&p, want);
        if (e != cudaSuccess && tl_pool && tl_pool->cached) {   // out of memory with blocks parked in the pool: give them up
            cudaGetLastError();
            tl_pool->flush();
            e = cudaMalloc((void**)&p, want);
        }
        if (e == cudaSuccess) {
            n = count;
            bytes = want;
        } else {
            p = nullptr;
        }
        return e;
    }
    void swap(Dev& o)
    {
        std::swap(p, o.p);
        std::swap(n, o.n);
        std::swap(bytes, o.bytes);
        std::swap(device, o.device);
    }
};

// AG2_PG_TRACE=1: host wall clock between the marks of a build, on stderr
struct Lap {
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    bool on = getenv("AG2_PG_TRACE") != nullptr;
    void mark(const char* what)
    {
        if (!on) return;
        const auto n = std::chrono::steady_clock::now();
        fprintf(stderr, "[ag2_pg trace] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(n - t).count());
        t = n;
    }
};

struct AlnSet {
    std::vector<ag2_pg_aln> rec;                 // file order
    std::vector<std::vector<int32_t>> by_query;  // per query: record indices in the reference's processing order
    const char* h_text = nullptr;                // caller-owned text (contig->reference only: walked on the host)
    Dev<char> text;                              // device text (read->contig, read->reference)
    int64_t text_len = 0;
    bool set = false;
};

}  // namespace

struct ag2_pg {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    std::string err;

    // B2
    int k = 0;
    int64_t n_vertices = 0;
    Dev<unsigned long long> codes, bitmap;
    Dev<uint32_t> rank;
    bool use_bitmap = false;

    // B6
    std::vector<int64_t> ctg_len, ref_len;
    std::vector<uint64_t> ctg_start, ref_start;
    Dev<int64_t> d_ctg_len;
    Dev<uint64_t> d_ctg_start, d_ref_start;

    // reads
    int64_t n_reads = 0, first_read = 0;
    std::vector<int64_t> read_len;
    Dev<unsigned long long> read_words;
    Dev<int64_t> d_read_off;
    Dev<int32_t> d_read_len;

    AlnSet aln[3];
    std::vector<uint8_t> ref_flag, ctg_flag, ctg_fwd;

    // B4 table
    Dev<int64_t> d_ctg_base;
    Dev<uint32_t> d_base_off, d_entry;

    // streams
    int64_t n_tuples = 0, n_edges = 0;
    Dev<uint32_t> t_vertex, t_ctg, t_ref, e_from, e_to;
    Dev<int32_t> e_step;
    bool have_streams = false;

    // graph
    Dev<unsigned long long> g_pos_off, g_edge_off;
    Dev<uint32_t> g_ctg, g_ref, g_edge_to;
    Dev<uint16_t> g_cnt;
    Dev<int32_t> g_edge_step;
    int64_t g_npos = 0, g_nedge = 0;
    bool have_graph = false;

    Dev<unsigned char> cub_tmp;
    ag2_pg_stats stats{};
    DevPool pool;   // recycled device blocks of this handle (see DevPool)
};

namespace {

int fail(ag2_pg* pg, int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (pg) pg->err = buf;
    return code;
}

#define PG_CUDA(call)                                                                                             \
    do {                                                                                                          \
        cudaError_t e_ = (call);                                                                                  \
        if (e_ != cudaSuccess)                                                                                    \
            return fail(pg, e_ == cudaErrorMemoryAllocation ? AG2_ENOMEM : AG2_ECUDA, "%s: %s (%s:%d)", #call,     \
                        cudaGetErrorString(e_), __FILE__, __LINE__);                                              \
    } while (0)

#define PG_TRY(call)                \
    do {                            \
        int r_ = (call);            \
        if (r_ != AG2_OK) return r_; \
    } while (0)

inline int grid_for(int64_t n, int block = 256)
{
    int64_t g = (n + block - 1) / block;
    const int64_t cap = 148 * 16;   // a few waves of the 148 SMs; kernels are grid-stride
    return (int)std::max<int64_t>(1, std::min<int64_t>(g, cap));
}

int cub_tmp(ag2_pg* pg, size_t bytes)
{
    if ((int64_t)bytes > pg->cub_tmp.n) PG_CUDA(pg->cub_tmp.alloc((int64_t)bytes + (1 << 20)));
    return AG2_OK;
}

template <typename T>
int exclusive_sum(ag2_pg* pg, const T* in, T* out, int64_t n)
{
    size_t bytes = 0;
    PG_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, n, pg->stream));
    PG_TRY(cub_tmp(pg, bytes));
    PG_CUDA(cub::DeviceScan::ExclusiveSum(pg->cub_tmp.p, bytes, in, out, n, pg->stream));
    ++pg->stats.launches;
    return AG2_OK;
}

template <typename K, typename V>
int sort_pairs(ag2_pg* pg, const K* kin, K* kout, const V* vin, V* vout, int64_t n, int begin_bit, int end_bit)
{
    size_t bytes = 0;
    PG_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, kin, kout, vin, vout, n, begin_bit, end_bit, pg->stream));
    PG_TRY(cub_tmp(pg, bytes));
    PG_CUDA(cub::DeviceRadixSort::SortPairs(pg->cub_tmp.p, bytes, kin, kout, vin, vout, n, begin_bit, end_bit, pg->stream));
    pg->stats.launches += (end_bit - begin_bit + 7) / 8 + 1;
    return AG2_OK;
}

int bits_for(uint64_t max_value)
{
    int b = 1;
    while (b < 64 && (max_value >> b)) ++b;
    return b;
}

// PositionMapper::generateStartPosHelper, position/PositionMapper.cpp:16-32
std::vector<uint64_t> mapper_starts(const std::vector<int64_t>& len)
{
    std::vector<uint64_t> s;
    if (len.empty()) return s;
    s.push_back((uint64_t)len[0]);
    for (size_t i = 1; i < len.size(); ++i) s.push_back(s.back() + 3 * (uint64_t)len[i - 1] + (uint64_t)std::max(len[i - 1], len[i]));
    s.push_back(s.back() + 4 * (uint64_t)len.back());
    return s;
}

// the two score sorts of the reference: the whole database (MecatAlignDatabase.cpp:19, MummerAlignDatabaseV2.cpp:48),
// then every query's list (Aligner::mergeAlignInfHelper, align/Aligner.cpp:32-56).  std::sort is not stable; the
// permutation it yields depends only on the comparison results, so sorting indices with libstdc++'s std::sort and the
// reference's comparator (score descending, align/AlignInf.cpp:31-33) reproduces the reference's order.
void group_alignments(AlnSet& s, int64_t first_query, int64_t n_query, int64_t n_target)
{
    std::vector<int32_t> order(s.rec.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = (int32_t)i;
    const ag2_pg_aln* r = s.rec.data();
    auto by_score = [r](int32_t a, int32_t b) { return r[a].score > r[b].score; };
    std::sort(order.begin(), order.end(), by_score);
    s.by_query.assign((size_t)n_query, {});
    for (int32_t i : order)
        if (r[i].query >= first_query && r[i].query < first_query + n_query && r[i].target >= 0 && r[i].target < n_target)
            s.by_query[r[i].query - first_query].push_back(i);
    for (auto& v : s.by_query) std::sort(v.begin(), v.end(), by_score);
}

inline void flip(uint64_t& l, uint64_t& r, uint64_t len)   // Aligner::flipPosition, align/Aligner.cpp:235-239
{
    const uint64_t t = l;
    l = len - r;
    r = len - t;
}

// B4: Aligner::simpleAlign (align/Aligner.cpp:96-201) + addExtraPosition (:203-209) as a CSR over contig bases holding
// packed reference positions (transformPosition, position/PositionProcessor.cpp:37-55, applied here once per table
// entry instead of once per read base).  Only the orientation each contig is used in (ctg_fwd) is ever queried
// (Aligner.tcc:63: (ii == 0) == _ctgFilterForward), so one list per base suffices.  Host work: O(aligned contig bases).
int build_ctg_table(ag2_pg* pg)
{
    const size_t n_ctg = pg->ctg_len.size();
    std::vector<int64_t> ctg_base(n_ctg + 1, 0);
    for (size_t c = 0; c < n_ctg; ++c) ctg_base[c + 1] = ctg_base[c] + (pg->ctg_flag[c] ? pg->ctg_len[c] : 0);
    const int64_t n_slots = ctg_base[n_ctg];
    // two passes over the same walk, count then fill (a std::vector per contig base was 78 ms of allocations at 1.6 Mb of
    // contigs): the entries of a base keep the order the reference appends them in (records in by_query order, columns in order)
    AlnSet& s = pg->aln[AG2_PG_CTG_TO_REF];
    std::vector<uint32_t> base_off((size_t)n_slots + 1, 0), entry, cursor;
    auto walk = [&](auto&& visit) {
        for (size_t c = 0; c < n_ctg; ++c) {
            if (!pg->ctg_flag[c]) continue;
            const uint64_t len = (uint64_t)pg->ctg_len[c];
            if (c >= s.by_query.size()) continue;
            for (int32_t ai : s.by_query[c]) {
                const ag2_pg_aln& a = s.rec[ai];
                if (!pg->ref_flag[a.target]) continue;
                const bool forward = a.forward != 0;
                if ((pg->ctg_fwd[c] != 0) != forward) continue;
                uint64_t cb = (uint64_t)a.qb, ce = (uint64_t)a.qe;
                if (!forward) flip(cb, ce, len);
                // exactAlign(ctgBegin, refBegin, true, ...): the reference position of every contig base of the record
                const char* q = s.h_text + a.q_off;
                const char* t = s.h_text + a.t_off;
                uint64_t cur_ref = (uint64_t)a.tb, at = cb;
                const uint64_t start = pg->ref_start[a.target];
                for (int32_t j = 0; j < a.ncols; ++j) {
                    const bool emit = q[j] != '-';
                    if (emit) {
                        if (at >= ce) break;
                        if (at < len) visit((size_t)(ctg_base[c] + (int64_t)at), (uint32_t)(start + cur_ref));
                        ++at;
                    }
                    if (!(emit && t[j] == '-')) ++cur_ref;
                }
            }
        }
    };
    walk([&](size_t slot, uint32_t) { ++base_off[slot + 1]; });
    for (int64_t i = 0; i < n_slots; ++i) {
        if (base_off[(size_t)i + 1] == 0) base_off[(size_t)i + 1] = 1;   // (0,0): PositionMapper::dualToSingle(0, 0) == 0
        base_off[(size_t)i + 1] += base_off[(size_t)i];
    }
    entry.assign(base_off[(size_t)n_slots], 0u);
    cursor.assign(base_off.begin(), base_off.end() - 1);
    walk([&](size_t slot, uint32_t v) { entry[cursor[slot]++] = v; });
    PG_CUDA(pg->d_ctg_base.alloc((int64_t)n_ctg + 1));
    PG_CUDA(pg->d_base_off.alloc(n_slots + 1));
    PG_CUDA(pg->d_entry.alloc((int64_t)entry.size()));
    PG_CUDA(cudaMemcpyAsync(pg->d_ctg_base.p, ctg_base.data(), (n_ctg + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, pg->stream));
    PG_CUDA(cudaMemcpyAsync(pg->d_base_off.p, base_off.data(), base_off.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, pg->stream));
    if (!entry.empty())
        PG_CUDA(cudaMemcpyAsync(pg->d_entry.p, entry.data(), entry.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, pg->stream));
    PG_CUDA(cudaStreamSynchronize(pg->stream));
    return AG2_OK;
}

struct PhasePlan {
    std::vector<Segment> segs;
    std::vector<Lane> lanes;
    int64_t tpos_total = 0, columns = 0;
};

// the per-alignment filters of Aligner::parseToCtg (align/Aligner.tcc:24-103) -> segments, grouped into (read, strand) lanes
void plan_phase0(ag2_pg* pg, const ag2_pg_params& P, PhasePlan& plan)
{
    AlnSet& s = pg->aln[AG2_PG_READ_TO_CTG];
    std::vector<Segment> per_strand[2];
    for (int64_t r = 0; r < pg->n_reads && r < (int64_t)s.by_query.size(); ++r) {
        const uint64_t read_len = (uint64_t)pg->read_len[r];
        per_strand[0].clear();
        per_strand[1].clear();
        int topk = 0;
        for (int32_t ai : s.by_query[r]) {
            if (P.read_to_ctg_topk >= 0 && topk >= P.read_to_ctg_topk) break;
            const ag2_pg_aln& a = s.rec[ai];
            const int c = a.target;
            if (!pg->ctg_flag[c]) continue;
            uint64_t rb = (uint64_t)a.qb, re = (uint64_t)a.qe;
            if ((double)(re - rb) * 1.0 / (double)read_len < P.read_to_ctg_ratio) continue;
            bool is_fwd = a.forward != 0;
            uint64_t cb = (uint64_t)a.tb, ce = (uint64_t)a.te;
            const uint64_t clen = (uint64_t)pg->ctg_len[c];
            if (ce >= clen || cb >= clen) continue;
            if (!is_fwd) flip(rb, re, read_len);
            if (!pg->ctg_fwd[c]) {           // ii == 1: strand, read range and contig range flipped once more
                is_fwd = !is_fwd;
                flip(rb, re, read_len);
                flip(cb, ce, clen);
            }
            ++topk;
            if (rb >= read_len) continue;    // no read base can pass `curRead < positions.size()`
            Segment sg{};
            sg.q_off = a.q_off;
            sg.t_off = a.t_off;
            sg.tb = cb;
            sg.rb = (uint32_t)rb;
            sg.ncols = a.ncols;
            sg.target = c;
            sg.backward = pg->ctg_fwd[c] ? 0 : 1;
            sg.negative = pg->ctg_fwd[c] ? 0 : 1;
            per_strand[is_fwd ? 0 : 1].push_back(sg);
        }
        for (int st = 0; st < 2; ++st) {
            if (per_strand[st].empty()) continue;
            Lane ln{(int32_t)r, st, (int32_t)plan.segs.size(), 0};
            for (auto& sg : per_strand[st]) {
                sg.tpos_off = plan.tpos_total;
                plan.tpos_total += sg.ncols;
                plan.columns += sg.ncols;
                plan.segs.push_back(sg);
            }
            ln.seg_end = (int32_t)plan.segs.size();
            plan.lanes.push_back(ln);
        }
    }
}

// Aligner::parseToRef (align/Aligner.tcc:106-171).  covInfHelper (align/Aligner.cpp:58-87) SORTS every reference's
// coverage array, and parseToRef takes the maximum over [refBegin, refEnd) of the sorted array, i.e. its element
// min(refEnd, len) - 1 (replicated, not fixed).
void plan_phase1(ag2_pg* pg, const ag2_pg_params& P, PhasePlan& plan)
{
    AlnSet& s = pg->aln[AG2_PG_READ_TO_REF];
    const size_t n_ref = pg->ref_len.size();
    std::vector<std::vector<uint32_t>> cov(n_ref);
    for (size_t f = 0; f < n_ref; ++f) cov[f].assign((size_t)pg->ref_len[f] + 1, 0);
    for (const ag2_pg_aln& a : s.rec) {       // every record whose reference name is known, whatever its query
        if (a.target < 0 || (size_t)a.target >= n_ref) continue;
        const uint64_t len = (uint64_t)pg->ref_len[a.target];
        const uint64_t b = std::min((uint64_t)a.tb, len), e = std::min((uint64_t)a.te, len);
        if (b < e) { ++cov[a.target][b]; --cov[a.target][e]; }
    }
    for (size_t f = 0; f < n_ref; ++f) {
        uint32_t run = 0;
        for (size_t i = 0; i + 1 < cov[f].size(); ++i) { run += cov[f][i]; cov[f][i] = run; }
        cov[f].pop_back();
        std::sort(cov[f].begin(), cov[f].end());
    }
    std::vector<Segment> per_strand[2];
    for (int64_t r = 0; r < pg->n_reads && r < (int64_t)s.by_query.size(); ++r) {
        const uint64_t read_len = (uint64_t)pg->read_len[r];
        per_strand[0].clear();
        per_strand[1].clear();
        int topk = 0;
        for (int32_t ai : s.by_query[r]) {
            if (P.read_to_ref_topk >= 0 && topk >= P.read_to_ref_topk) break;
            const ag2_pg_aln& a = s.rec[ai];
            const int f = a.target;
            if (!pg->ref_flag[f]) continue;
            uint64_t rb = (uint64_t)a.qb, re = (uint64_t)a.qe;
            if ((double)(re - rb) * 1.0 / (double)read_len < P.read_to_ref_ratio) continue;
            const bool is_fwd = a.forward != 0;
            uint64_t max_cov = 0;
            {
                const uint64_t len = (uint64_t)cov[f].size();
                const uint64_t e = std::min((uint64_t)a.te, len);
                if ((uint64_t)a.tb < e) max_cov = cov[f][e - 1];
            }
            if (max_cov < (uint64_t)P.cov_filter) continue;
            if (!is_fwd) flip(rb, re, read_len);
            ++topk;
            if (rb >= read_len) continue;
            Segment sg{};
            sg.q_off = a.q_off;
            sg.t_off = a.t_off;
            sg.tb = (uint64_t)a.tb;
            sg.rb = (uint32_t)rb;
            sg.ncols = a.ncols;
            sg.target = f;
            per_strand[is_fwd ? 0 : 1].push_back(sg);
        }
        for (int st = 0; st < 2; ++st) {
            if (per_strand[st].empty()) continue;
            Lane ln{(int32_t)r, st, (int32_t)plan.segs.size(), 0};
            for (auto& sg : per_strand[st]) {
                sg.tpos_off = plan.tpos_total;
                plan.tpos_total += sg.ncols;
                plan.columns += sg.ncols;
                plan.segs.push_back(sg);
            }
            ln.seg_end = (int32_t)plan.segs.size();
            plan.lanes.push_back(ln);
        }
    }
}

struct PhaseDev {
    Dev<Segment> segs;
    Dev<Lane> lanes;
    Dev<int32_t> nq;
    Dev<uint32_t> tpos;
    Dev<LaneCounts> counts;   // [n_lanes] + 1 total
    LaneCounts total{};
    ExtractArgs args{};
};

int run_walk_and_count(ag2_pg* pg, int phase, const ag2_pg_params& P, const PhasePlan& plan, PhaseDev& d)
{
    const int64_t n_segs = (int64_t)plan.segs.size(), n_lanes = (int64_t)plan.lanes.size();
    PG_CUDA(d.segs.alloc(n_segs));
    PG_CUDA(d.lanes.alloc(n_lanes));
    PG_CUDA(d.nq.alloc(n_segs));
    PG_CUDA(d.tpos.alloc(plan.tpos_total));
    PG_CUDA(d.counts.alloc(n_lanes + 1));
    d.total = LaneCounts{0, 0, 0};
    if (n_lanes == 0) return AG2_OK;
    PG_CUDA(cudaMemcpyAsync(d.segs.p, plan.segs.data(), n_segs * sizeof(Segment), cudaMemcpyHostToDevice, pg->stream));
    PG_CUDA(cudaMemcpyAsync(d.lanes.p, plan.lanes.data(), n_lanes * sizeof(Lane), cudaMemcpyHostToDevice, pg->stream));
    AlnSet& s = pg->aln[phase == 0 ? AG2_PG_READ_TO_CTG : AG2_PG_READ_TO_REF];
    pg_walk_kernel<<<grid_for(n_segs * 32), 256, 0, pg->stream>>>(d.segs.p, n_segs, s.text.p, d.tpos.p, d.nq.p);
    ExtractArgs& a = d.args;
    a.lanes = d.lanes.p;
    a.n_lanes = n_lanes;
    a.segs = d.segs.p;
    a.seg_nq = d.nq.p;
    a.tpos = d.tpos.p;
    a.read_words = pg->read_words.p;
    a.read_off = pg->d_read_off.p;
    a.read_len = pg->d_read_len.p;
    a.vs = VertexSet{pg->use_bitmap ? pg->bitmap.p : nullptr, pg->rank.p, pg->codes.p, pg->n_vertices, pg->k};
    a.tab = CtgTable{pg->d_ctg_base.p, pg->d_base_off.p, pg->d_entry.p, pg->d_ctg_len.p, pg->d_ctg_start.p, pg->d_ref_start.p};
    a.phase = phase;
    a.outer = P.outer_sample;
    a.counts = d.counts.p;
    pg_extract_kernel<false><<<grid_for(n_lanes * 32), 256, 0, pg->stream>>>(a);
    pg_scan_counts_kernel<<<1, 1024, 0, pg->stream>>>(d.counts.p, n_lanes, d.counts.p + n_lanes);
    pg->stats.launches += 3;
    PG_CUDA(cudaMemcpyAsync(&d.total, d.counts.p + n_lanes, sizeof(LaneCounts), cudaMemcpyDeviceToHost, pg->stream));
    PG_CUDA(cudaStreamSynchronize(pg->stream));
    return AG2_OK;
}

int check_ready(ag2_pg* pg)
{
    if (!pg->n_vertices && !pg->codes.p) return fail(pg, AG2_ESTATE, "ag2_pg_set_kmers has not been called");
    if (pg->ctg_start.empty() && pg->ref_start.empty()) return fail(pg, AG2_ESTATE, "ag2_pg_set_targets has not been called");
    if (!pg->d_read_off.p) return fail(pg, AG2_ESTATE, "ag2_pg_set_reads has not been called");
    for (int w = 0; w < 3; ++w)
        if (!pg->aln[w].set) return fail(pg, AG2_ESTATE, "ag2_pg_set_alignments(%d) has not been called", w);
    if (pg->ref_flag.size() != pg->ref_len.size() || pg->ctg_flag.size() != pg->ctg_len.size())
        return fail(pg, AG2_ESTATE, "ag2_pg_set_filters has not been called");
    return AG2_OK;
}

}  // namespace

extern "C" {

int ag2_pg_create(int device, ag2_pg** out)
{
    if (!out) return AG2_EINVAL;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return AG2_ENODEV;
    if (cudaSetDevice(device) != cudaSuccess) return AG2_ENODEV;
    ag2_pg* pg = new (std::nothrow) ag2_pg();
    if (!pg) return AG2_ENOMEM;
    pg->device = device;
    if (cudaStreamCreateWithFlags(&pg->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete pg;
        return AG2_ECUDA;
    }
    for (auto& e : pg->ev) cudaEventCreate(&e);
    *out = pg;
    return AG2_OK;
}

void ag2_pg_destroy(ag2_pg* pg)
{
    if (!pg) return;
    cudaSetDevice(pg->device);
    cudaStreamSynchronize(pg->stream);
    for (auto& e : pg->ev)
        if (e) cudaEventDestroy(e);
    cudaStreamDestroy(pg->stream);
    pg->pool.flush();
    delete pg;
}

const char* ag2_pg_last_error(const ag2_pg* pg) { return pg ? pg->err.c_str() : "null handle"; }

void ag2_pg_params_default(ag2_pg_params* p)
{
    if (!p) return;
    p->outer_sample = 3;
    p->read_to_ctg_topk = -1;
    p->read_to_ref_topk = -1;
    p->read_to_ctg_ratio = 0.35;
    p->read_to_ref_ratio = 0.10;
    p->epsilon = 10;
    p->cov_filter = 1;
}

void* ag2_pg_stream(ag2_pg* pg) { return pg ? (void*)pg->stream : nullptr; }

int ag2_pg_set_kmers(ag2_pg* pg, const uint64_t* words, int64_t n_words, int64_t* n_vertices)
{
    PoolScope pool_scope(pg ? &pg->pool : nullptr);
    if (!pg || !words || n_words < 1) return fail(pg, AG2_EINVAL, "ag2_pg_set_kmers: bad arguments");
    PG_CUDA(cudaSetDevice(pg->device));
    const uint64_t k = words[0];
    if (k < 1 || k > 32) return fail(pg, AG2_EINVAL, "ag2_pg_set_kmers: k = %llu out of range", (unsigned long long)k);
    pg->k = (int)k;
    Dev<unsigned long long> in, sorted;
    Dev<uint32_t> flag, pos;
    PG_CUDA(in.alloc(n_words));
    PG_CUDA(sorted.alloc(n_words));
    PG_CUDA(flag.alloc(n_words + 1));
    PG_CUDA(pos.alloc(n_words + 1));
    PG_CUDA(cudaMemcpyAsync(in.p, words, (size_t)n_words * 8, cudaMemcpyHostToDevice, pg->stream));
    {
        size_t bytes = 0;
        PG_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, bytes, in.p, sorted.p, n_words, 0, 64, pg->stream));
        PG_TRY(cub_tmp(pg, bytes));
        PG_CUDA(cub::DeviceRadixSort::SortKeys(pg->cub_tmp.p, bytes, in.p, sorted.p, n_words, 0, 64, pg->stream));
    }
    pg_flag_unique_kernel<<<grid_for(n_words), 256, 0, pg->stream>>>(sorted.p, n_words, flag.p);
    PG_CUDA(cudaMemsetAsync(flag.p + n_words, 0, 4, pg->stream));
    PG_TRY(exclusive_sum(pg, flag.p, pos.p, n_words + 1));
    uint32_t nv = 0;
    PG_CUDA(cudaMemcpyAsync(&nv, pos.p + n_words, 4, cudaMemcpyDeviceToHost, pg->stream));
    PG_CUDA(cudaStreamSynchronize(pg->stream));
    pg->n_vertices = nv;
    PG_CUDA(pg->codes.alloc(nv));
    pg_scatter_unique_kernel<<<grid_for(n_words), 256, 0, pg->stream>>>(sorted.p, flag.p, pos.p, n_words, pg->codes.p);
    pg->use_bitmap = pg->k <= 16;
    if (pg->use_bitmap) {
        const unsigned long long limit = 1ull << (2 * pg->k);
        const int64_t n_bm = (int64_t)std::max<unsigned long long>(1, limit >> 6);
        Dev<uint32_t> pc;
        PG_CUDA(pg->bitmap.alloc(n_bm));
        PG_CUDA(pg->rank.alloc(n_bm));
        PG_CUDA(pc.alloc(n_bm));
        PG_CUDA(cudaMemsetAsync(pg->bitmap.p, 0, (size_t)n_bm * 8, pg->stream));
        pg_bitmap_fill_kernel<<<grid_for(nv), 256, 0, pg->stream>>>(pg->codes.p, nv, limit, pg->bitmap.p);
        pg_popc_kernel<<<grid_for(n_bm), 256, 0, pg->stream>>>(pg->bitmap.p, n_bm, pc.p);
        PG_TRY(exclusive_sum(pg, pc.p, pg->rank.p, n_bm));
        PG_CUDA(cudaStreamSynchronize(pg->stream));
    }
    PG_CUDA(cudaStreamSynchronize(pg->stream));
    PG_CUDA(cudaGetLastError());
    pg->stats.n_vertices = nv;
    pg->have_graph = pg->have_streams = false;
    if (n_vertices) *n_vertices = nv;
    return AG2_OK;
}

int ag2_pg_fetch_codes(ag2_pg* pg, uint64_t* out, int64_t cap)
{
    if (!pg || !out) return fail(pg, AG2_EINVAL, "ag2_pg_fetch_codes: bad arguments");
    if (cap < pg->n_vertices) return fail(pg, AG2_ECAP, "ag2_pg_fetch_codes: need %lld entries", (long long)pg->n_vertices);
    PG_CUDA(cudaSetDevice(pg->device));
    PG_CUDA(cudaMemcpy(out, pg->codes.p, (size_t)pg->n_vertices * 8, cudaMemcpyDeviceToHost));
    return AG2_OK;
}

int ag2_pg_set_targets(ag2_pg* pg, const int64_t* ctg_len, int64_t n_ctg, const int64_t* ref_len, int64_t n_ref)
{
    PoolScope pool_scope(pg ? &pg->pool : nullptr);
    if (!pg || n_ctg < 0 || n_ref < 0 || (n_ctg && !ctg_len) || (n_ref && !ref_len)) return fail(pg, AG2_EINVAL, "ag2_pg_set_targets: bad arguments");
    PG_CUDA(cudaSetDevice(pg->device));
    pg->ctg_len.assign(ctg_len, ctg_len + n_ctg);
    pg->ref_len.assign(ref_len, ref_len + n_ref);
    pg->ctg_start = mapper_starts(pg->ctg_len);
    pg->ref_start = mapper_starts(pg->ref_len);
    if (!pg->ctg_start.empty() && pg->ctg_start.back() > 0xffffffffull)
        fprintf(stderr, "ag2_pg: contig positions exceed 32 bits and wrap, as PABruijnGraph::PosType does\n");
    PG_CUDA(pg->d_ctg_len.alloc(n_ctg));
    PG_CUDA(pg->d_ctg_start.alloc(n_ctg + 1));
    PG_CUDA(pg->d_ref_start.alloc(n_ref + 1));
    if (n_ctg) {
        PG_CUDA(cudaMemcpy(pg->d_ctg_len.p, ctg_len, (size_t)n_ctg * 8, cudaMemcpyHostToDevice));
        PG_CUDA(cudaMemcpy(pg->d_ctg_start.p, pg->ctg_start.data(), pg->ctg_start.size() * 8, cudaMemcpyHostToDevice));
    }
    if (n_ref) PG_CUDA(cudaMemcpy(pg->d_ref_start.p, pg->ref_start.data(), pg->ref_start.size() * 8, cudaMemcpyHostToDevice));
    pg->ref_flag.clear();
    pg->ctg_flag.clear();
    pg->ctg_fwd.clear();
    return AG2_OK;
}

int ag2_pg_set_reads(ag2_pg* pg, const char* bases, const int64_t* offs, int64_t n_reads, int64_t first_read)
{
    PoolScope pool_scope(pg ? &pg->pool : nullptr);
    if (!pg || n_reads < 0 || first_read < 0 || !offs || (n_reads && !bases && offs[n_reads] > 0)) return fail(pg, AG2_EINVAL, "ag2_pg_set_reads: bad arguments");
    PG_CUDA(cudaSetDevice(pg->device));
    pg->n_reads = n_reads;
    pg->first_read = first_read;
    pg->read_len.resize((size_t)n_reads);
    std::vector<int64_t> poff((size_t)n_reads + 1, 0);
    std::vector<int32_t> len32((size_t)n_reads);
    for (int64_t r = 0; r < n_reads; ++r) {
        const int64_t len = offs[r + 1] - offs[r];
        if (len < 0 || len > 0x7fffffff) return fail(pg, AG2_EINVAL, "ag2_pg_set_reads: read %lld has length %lld", (long long)r, (long long)len);
        pg->read_len[r] = len;
        len32[r] = (int32_t)len;
        poff[r + 1] = poff[r] + ((len + 31) & ~31ll);
    }
    const int64_t n_words = poff[n_reads] >> 5;
    Dev<char> ascii;
    Dev<int64_t> d_offs;
    PG_CUDA(ascii.alloc(offs[n_reads]));
    PG_CUDA(d_offs.alloc(n_reads + 1));
    PG_CUDA(pg->read_words.alloc(n_words + 2));
    PG_CUDA(pg->d_read_off.alloc(n_reads + 1));
    PG_CUDA(pg->d_read_len.alloc(n_reads));
    PG_CUDA(cudaMemsetAsync(pg->read_words.p, 0, (size_t)(n_words + 2) * 8, pg->stream));
    if (offs[n_reads] > 0) PG_CUDA(cudaMemcpyAsync(ascii.p, bases, (size_t)offs[n_reads], cudaMemcpyHostToDevice, pg->stream));
    PG_CUDA(cudaMemcpyAsync(d_offs.p, offs, (size_t)(n_reads + 1) * 8, cudaMemcpyHostToDevice, pg->stream));
    PG_CUDA(cudaMemcpyAsync(pg->d_read_off.p, poff.data(), (size_t)(n_reads + 1) * 8, cudaMemcpyHostToDevice, pg->stream));
    if (n_reads) PG_CUDA(cudaMemcpyAsync(pg->d_read_len.p, len32.data(), (size_t)n_reads * 4, cudaMemcpyHostToDevice, pg->stream));
    if (n_words > 0)
        pg_pack_reads_kernel<<<grid_for(n_words), 256, 0, pg->stream>>>(ascii.p, d_offs.p, pg->d_read_off.p, n_reads, n_words, pg->read_words.p);
    PG_CUDA(cudaStreamSynchronize(pg->stream));
    PG_CUDA(cudaGetLastError());
    pg->aln[0].set = pg->aln[1].set = false;   // read indices changed
    pg->have_graph = pg->have_streams = false;
    return AG2_OK;
}

int ag2_pg_set_alignments(ag2_pg* pg, int which, const ag2_pg_aln* alns, int64_t n, const char* text, int64_t text_len)
{
    PoolScope pool_scope(pg ? &pg->pool : nullptr);
    if (!pg || which < 0 || which > 2 || n < 0 || (n && !alns) || text_len < 0 || (text_len && !text))
        return fail(pg, AG2_EINVAL, "ag2_pg_set_alignments: bad arguments");
    PG_CUDA(cudaSetDevice(pg->device));
    for (int64_t i = 0; i < n; ++i)
        if (alns[i].ncols < 0 || alns[i].q_off < 0 || alns[i].t_off < 0 || alns[i].q_off + alns[i].ncols > text_len ||
            alns[i].t_off + alns[i].ncols > text_len)
            return fail(pg, AG2_EINVAL, "ag2_pg_set_alignments: record %lld points outside the text buffer", (long long)i);
    AlnSet& s = pg->aln[which];
    s.rec.assign(alns, alns + n);
    s.text_len = text_len;
    const int64_t n_query = which == AG2_PG_CTG_TO_REF ? (int64_t)pg->ctg_len.size() : pg->n_reads;
    const int64_t n_target = which == AG2_PG_READ_TO_CTG ? (int64_t)pg->ctg_len.size() : (int64_t)pg->ref_len.size();
    group_alignments(s, which == AG2_PG_CTG_TO_REF ? 0 : pg->first_read, n_query, n_target);
    if (which == AG2_PG_CTG_TO_REF) {
        s.h_text = text;   // walked on the host inside ag2_pg_build / ag2_pg_extract: must stay valid until then
    } else {
        PG_CUDA(s.text.alloc(text_len));
        if (text_len) {
            PG_CUDA(cudaMemcpyAsync(s.text.p, text, (size_t)text_len, cudaMemcpyHostToDevice, pg->stream));
            PG_CUDA(cudaStreamSynchronize(pg->stream));
        }
    }
    s.set = true;
    pg->have_graph = pg->have_streams = false;
    return AG2_OK;
}

int ag2_pg_set_filters(ag2_pg* pg, const uint8_t* ref_flag, const uint8_t* ctg_flag, const uint8_t* ctg_forward)
{
    if (!pg || (!ref_flag && !pg->ref_len.empty()) || ((!ctg_flag || !ctg_forward) && !pg->ctg_len.empty()))
        return fail(pg, AG2_EINVAL, "ag2_pg_set_filters: bad arguments");
    pg->ref_flag.assign(ref_flag, ref_flag + pg->ref_len.size());
    pg->ctg_flag.assign(ctg_flag, ctg_flag + pg->ctg_len.size());
    pg->ctg_fwd.assign(ctg_forward, ctg_forward + pg->ctg_len.size());
    pg->have_graph = pg->have_streams = false;
    return AG2_OK;
}

int ag2_pg_extract(ag2_pg* pg, const ag2_pg_params* params)
{
    PoolScope pool_scope(pg ? &pg->pool : nullptr);
    if (!pg || !params) return fail(pg, AG2_EINVAL, "ag2_pg_extract: bad arguments");
    PG_TRY(check_ready(pg));
    if (params->outer_sample < 1) return fail(pg, AG2_EINVAL, "ag2_pg_extract: outer_sample must be >= 1");
    PG_CUDA(cudaSetDevice(pg->device));
    const ag2_pg_params& P = *params;
    const int64_t nv = pg->stats.n_vertices;
    pg->stats = ag2_pg_stats{};
    pg->stats.n_vertices = nv;
    pg->have_graph = pg->have_streams = false;
    Lap lap;
    PG_TRY(build_ctg_table(pg));
    lap.mark("extract: ctg table");

    PhasePlan plan[2];
    plan_phase0(pg, P, plan[0]);
    plan_phase1(pg, P, plan[1]);
    lap.mark("extract: plans (host)");
    PG_CUDA(cudaEventRecord(pg->ev[0], pg->stream));
    PhaseDev dev[2];
    for (int ph = 0; ph < 2; ++ph) PG_TRY(run_walk_and_count(pg, ph, P, plan[ph], dev[ph]));
    lap.mark("extract: walk + count");
    const unsigned long long nt = dev[0].total.tuples + dev[1].total.tuples, ne = dev[0].total.edges + dev[1].total.edges;
    if (nt >= 0xffffffffull || ne >= 0xffffffffull) return fail(pg, AG2_ECAP, "ag2_pg_extract: %llu tuples / %llu edges exceed the 32-bit stream index", nt, ne);
    PG_CUDA(pg->t_vertex.alloc((int64_t)nt));
    PG_CUDA(pg->t_ctg.alloc((int64_t)nt));
    PG_CUDA(pg->t_ref.alloc((int64_t)nt));
    PG_CUDA(pg->e_from.alloc((int64_t)ne));
    PG_CUDA(pg->e_to.alloc((int64_t)ne));
    PG_CUDA(pg->e_step.alloc((int64_t)ne));
    lap.mark("extract: stream allocation");
    for (int ph = 0; ph < 2; ++ph) {
        const int64_t n_lanes = (int64_t)plan[ph].lanes.size();
        pg->stats.lanes[ph] = n_lanes;
        pg->stats.columns[ph] = plan[ph].columns;
        pg->stats.samples[ph] = (int64_t)dev[ph].total.samples;
        pg->stats.tuples[ph] = (int64_t)dev[ph].total.tuples;
        pg->stats.edges_raw[ph] = (int64_t)dev[ph].total.edges;
        if (!n_lanes) continue;
        ExtractArgs a = dev[ph].args;
        a.t_vertex = pg->t_vertex.p;
        a.t_ctg = pg->t_ctg.p;
        a.t_ref = pg->t_ref.p;
        a.e_from = pg->e_from.p;
        a.e_to = pg->e_to.p;
        a.e_step = pg->e_step.p;
        a.tuple_base = ph ? dev[0].total.tuples : 0;
        a.edge_base = ph ? dev[0].total.edges : 0;
        pg_extract_kernel<true><<<grid_for(n_lanes * 32), 256, 0, pg->stream>>>(a);
        ++pg->stats.launches;
    }
    PG_CUDA(cudaEventRecord(pg->ev[1], pg->stream));
    PG_CUDA(cudaStreamSynchronize(pg->stream));
    PG_CUDA(cudaGetLastError());
    float ms = 0;
    cudaEventElapsedTime(&ms, pg->ev[0], pg->ev[1]);
    pg->stats.extract_ms = ms;
    lap.mark("extract: emit kernels");
    pg->n_tuples = (int64_t)nt;
    pg->n_edges = (int64_t)ne;
    pg->have_streams = true;
    return AG2_OK;
}

int ag2_pg_partition(ag2_pg* pg, int n_owners, int64_t* counts)
{
    PoolScope pool_scope(pg ? &pg->pool : nullptr);
    if (!pg || n_owners < 1 || !counts) return fail(pg, AG2_EINVAL, "ag2_pg_partition: bad arguments");
    if (!pg->have_streams) return fail(pg, AG2_ESTATE, "ag2_pg_partition: no streams (call ag2_pg_extract)");
    PG_CUDA(cudaSetDevice(pg->device));
    const uint32_t per = (uint32_t)std::max<int64_t>(1, (pg->n_vertices + n_owners - 1) / n_owners);
    const int bits = bits_for((uint64_t)n_owners);
    Dev<unsigned long long> hist;
    PG_CUDA(hist.alloc(2 * (int64_t)n_owners));
    PG_CUDA(cudaMemsetAsync(hist.p, 0, 2 * (size_t)n_owners * 8, pg->stream));
    for (int which = 0; which < 2; ++which) {
        const int64_t n = which ? pg->n_edges : pg->n_tuples;
        if (n == 0) continue;
        Dev<uint32_t> owner, owner2, index, index2, a, b, c;
        PG_CUDA(owner.alloc(n));
        PG_CUDA(owner2.alloc(n));
        PG_CUDA(index.alloc(n));
        PG_CUDA(index2.alloc(n));
        PG_CUDA(a.alloc(n));
        PG_CUDA(b.alloc(n));
        PG_CUDA(c.alloc(n));
        const uint32_t* key = which ? pg->e_from.p : pg->t_vertex.p;
        pg_owner_kernel<<<grid_for(n), 256, 0, pg->stream>>>(key, n, per, owner.p, index.p, hist.p + which * n_owners);
        PG_TRY(sort_pairs(pg, owner.p, owner2.p, index.p, index2.p, n, 0, bits));
        if (which == 0) {
            pg_gather3_kernel<<<grid_for(n), 256, 0, pg->stream>>>(index2.p, n, pg->t_vertex.p, pg->t_ctg.p, pg->t_ref.p, a.p, b.p, c.p);
            PG_CUDA(cudaStreamSynchronize(pg->stream));
            pg->t_vertex.swap(a);
            pg->t_ctg.swap(b);
            pg->t_ref.swap(c);
        } else {
            pg_gather3_kernel<<<grid_for(n), 256, 0, pg->stream>>>(index2.p, n, pg->e_from.p, pg->e_to.p, (const uint32_t*)pg->e_step.p, a.p, b.p, c.p);
            PG_CUDA(cudaStreamSynchronize(pg->stream));
            pg->e_from.swap(a);
            pg->e_to.swap(b);
            std::swap(*(uint32_t**)&pg->e_step.p, c.p);
            std::swap(pg->e_step.n, c.n);
            std::swap(pg->e_step.bytes, c.bytes);
            std::swap(pg->e_step.device, c.device);
        }
        pg->stats.launches += 2;
    }
    std::vector<unsigned long long> h(2 * (size_t)n_owners);
    PG_CUDA(cudaMemcpyAsync(h.data(), hist.p, h.size() * 8, cudaMemcpyDeviceToHost, pg->stream));
    PG_CUDA(cudaStreamSynchronize(pg->stream));
    PG_CUDA(cudaGetLastError());
    for (size_t i = 0; i < h.size(); ++i) counts[i] = (int64_t)h[i];
    return AG2_OK;
}

int ag2_pg_stream_dev(ag2_pg* pg, int64_t* n_tuples, void** tuple_dev3, int64_t* n_edges, void** edge_dev3)
{
    if (!pg) return AG2_EINVAL;
    if (!pg->have_streams) return fail(pg, AG2_ESTATE, "ag2_pg_stream_dev: no streams (call ag2_pg_extract)");
    if (n_tuples) *n_tuples = pg->n_tuples;
    if (n_edges) *n_edges = pg->n_edges;
    if (tuple_dev3) { tuple_dev3[0] = pg->t_vertex.p; tuple_dev3[1] = pg->t_ctg.p; tuple_dev3[2] = pg->t_ref.p; }
    if (edge_dev3) { edge_dev3[0] = pg->e_from.p; edge_dev3[1] = pg->e_to.p; edge_dev3[2] = pg->e_step.p; }
    return AG2_OK;
}

int ag2_pg_import_dev(ag2_pg* pg, int64_t n_tuples, void* const* tuple_dev3, int64_t n_edges, void* const* edge_dev3)
{
    PoolScope pool_scope(pg ? &pg->pool : nullptr);
    if (!pg || n_tuples < 0 || n_edges < 0 || (n_tuples && !tuple_dev3) || (n_edges && !edge_dev3)) return fail(pg, AG2_EINVAL, "ag2_pg_import_dev: bad arguments");
    if (n_tuples >= 0xffffffffll || n_edges >= 0xffffffffll) return fail(pg, AG2_ECAP, "ag2_pg_import_dev: streams exceed the 32-bit index");
    PG_CUDA(cudaSetDevice(pg->device));
    Dev<uint32_t> a, b, c, d, e;
    Dev<int32_t> f;
    PG_CUDA(a.alloc(n_tuples));
    PG_CUDA(b.alloc(n_tuples));
    PG_CUDA(c.alloc(n_tuples));
    PG_CUDA(d.alloc(n_edges));
    PG_CUDA(e.alloc(n_edges));
    PG_CUDA(f.alloc(n_edges));
    if (n_tuples) {
        PG_CUDA(cudaMemcpyAsync(a.p, tuple_dev3[0], (size_t)n_tuples * 4, cudaMemcpyDeviceToDevice, pg->stream));
        PG_CUDA(cudaMemcpyAsync(b.p, tuple_dev3[1], (size_t)n_tuples * 4, cudaMemcpyDeviceToDevice, pg->stream));
        PG_CUDA(cudaMemcpyAsync(c.p, tuple_dev3[2], (size_t)n_tuples * 4, cudaMemcpyDeviceToDevice, pg->stream));
    }
    if (n_edges) {
        PG_CUDA(cudaMemcpyAsync(d.p, edge_dev3[0], (size_t)n_edges * 4, cudaMemcpyDeviceToDevice, pg->stream));
        PG_CUDA(cudaMemcpyAsync(e.p, edge_dev3[1], (size_t)n_edges * 4, cudaMemcpyDeviceToDevice, pg->stream));
        PG_CUDA(cudaMemcpyAsync(f.p, edge_dev3[2], (size_t)n_edges * 4, cudaMemcpyDeviceToDevice, pg->stream));
    }
    PG_CUDA(cudaStreamSynchronize(pg->stream));
    pg->t_vertex.swap(a);
    pg->t_ctg.swap(b);
    pg->t_ref.swap(c);
    pg->e_from.swap(d);
    pg->e_to.swap(e);
    pg->e_step.swap(f);
    pg->n_tuples = n_tuples;
    pg->n_edges = n_edges;
    pg->have_streams = true;
    pg->have_graph = false;
    return AG2_OK;
}

// The one exchange step of the graph build between the GPUs of ONE process: every handle has extracted the
// tuples and edges of its contiguous read range; each stream is partitioned stably by the owner of its vertex and every
// (rank, owner) segment is copied straight into the owner's receive buffer at its place in rank order -- peer copies over
// NVLink (cudaMemcpyPeerAsync; a plain device copy when two handles share a GPU), no host staging, no collective library.
// Rank order is the global read order, so the owner's first-fit join sees the items as the one-GPU build does.
int ag2_pg_group_exchange(ag2_pg* const* pgs, int n)
{
    if (!pgs || n < 1) return AG2_EINVAL;
    for (int r = 0; r < n; ++r)
        if (!pgs[r] || !pgs[r]->have_streams) return fail(pgs[r], AG2_ESTATE, "ag2_pg_group_exchange: handle %d has no streams (call ag2_pg_extract)", r);
    if (n == 1) return AG2_OK;
    ag2_pg* pg = pgs[0];   // errors are reported on the first handle
    std::vector<std::vector<int64_t>> counts((size_t)n, std::vector<int64_t>(2 * (size_t)n));
    for (int r = 0; r < n; ++r) {
        const int rc = ag2_pg_partition(pgs[r], n, counts[(size_t)r].data());
        if (rc != AG2_OK) {
            if (r) fail(pg, rc, "ag2_pg_group_exchange: partition on handle %d: %s", r, pgs[r]->err.c_str());
            return rc;
        }
    }
    for (int a = 0; a < n; ++a)       // peer access both ways between distinct devices
        for (int b = 0; b < n; ++b) {
            if (pgs[a]->device == pgs[b]->device) continue;
            int can = 0;
            PG_CUDA(cudaDeviceCanAccessPeer(&can, pgs[a]->device, pgs[b]->device));
            if (!can) continue;       // cudaMemcpyPeerAsync then stages through the host by itself
            PG_CUDA(cudaSetDevice(pgs[a]->device));
            const cudaError_t e = cudaDeviceEnablePeerAccess(pgs[b]->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(pg, AG2_ECUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
            cudaGetLastError();
        }
    struct Recv {
        Dev<uint32_t> a, b, c, d, e;
        Dev<int32_t> f;
        int64_t nt = 0, ne = 0;
    };
    std::vector<Recv> recv((size_t)n);
    for (int o = 0; o < n; ++o) {
        Recv& R = recv[(size_t)o];
        for (int r = 0; r < n; ++r) {
            R.nt += counts[(size_t)r][(size_t)o];
            R.ne += counts[(size_t)r][(size_t)(n + o)];
        }
        if (R.nt >= 0xffffffffll || R.ne >= 0xffffffffll) return fail(pg, AG2_ECAP, "ag2_pg_group_exchange: streams exceed the 32-bit index");
        PG_CUDA(cudaSetDevice(pgs[o]->device));
        PG_CUDA(R.a.alloc(R.nt));
        PG_CUDA(R.b.alloc(R.nt));
        PG_CUDA(R.c.alloc(R.nt));
        PG_CUDA(R.d.alloc(R.ne));
        PG_CUDA(R.e.alloc(R.ne));
        PG_CUDA(R.f.alloc(R.ne));
    }
    for (int r = 0; r < n; ++r) {     // rank r pushes its segments, on its own stream
        ag2_pg* S = pgs[r];
        PG_CUDA(cudaSetDevice(S->device));
        int64_t st = 0, se = 0;       // start of owner o's segment in r's partitioned streams
        for (int o = 0; o < n; ++o) {
            const int64_t ct = counts[(size_t)r][(size_t)o], ce = counts[(size_t)r][(size_t)(n + o)];
            int64_t dt = 0, de = 0;   // its place in the owner's buffers: behind the lower ranks
            for (int q = 0; q < r; ++q) {
                dt += counts[(size_t)q][(size_t)o];
                de += counts[(size_t)q][(size_t)(n + o)];
            }
            Recv& R = recv[(size_t)o];
            const int dd = pgs[o]->device, sd = S->device;
            if (ct) {
                PG_CUDA(cudaMemcpyPeerAsync(R.a.p + dt, dd, S->t_vertex.p + st, sd, (size_t)ct * 4, S->stream));
                PG_CUDA(cudaMemcpyPeerAsync(R.b.p + dt, dd, S->t_ctg.p + st, sd, (size_t)ct * 4, S->stream));
                PG_CUDA(cudaMemcpyPeerAsync(R.c.p + dt, dd, S->t_ref.p + st, sd, (size_t)ct * 4, S->stream));
            }
            if (ce) {
                PG_CUDA(cudaMemcpyPeerAsync(R.d.p + de, dd, S->e_from.p + se, sd, (size_t)ce * 4, S->stream));
                PG_CUDA(cudaMemcpyPeerAsync(R.e.p + de, dd, S->e_to.p + se, sd, (size_t)ce * 4, S->stream));
                PG_CUDA(cudaMemcpyPeerAsync(R.f.p + de, dd, S->e_step.p + se, sd, (size_t)ce * 4, S->stream));
            }
            st += ct;
            se += ce;
        }
    }
    for (int r = 0; r < n; ++r) {
        PG_CUDA(cudaSetDevice(pgs[r]->device));
        PG_CUDA(cudaStreamSynchronize(pgs[r]->stream));
    }
    for (int o = 0; o < n; ++o) {     // the received streams replace the extracted ones
        ag2_pg* D = pgs[o];
        Recv& R = recv[(size_t)o];
        PG_CUDA(cudaSetDevice(D->device));
        D->t_vertex.swap(R.a);
        D->t_ctg.swap(R.b);
        D->t_ref.swap(R.c);
        D->e_from.swap(R.d);
        D->e_to.swap(R.e);
        D->e_step.swap(R.f);
        D->n_tuples = R.nt;
        D->n_edges = R.ne;
        D->have_graph = false;
        R.a.release(); R.b.release(); R.c.release(); R.d.release(); R.e.release(); R.f.release();   // on D's device
    }
    return AG2_OK;
}

// The merge of the per-GPU vertex tables before the traversal, into handle 0: every handle holds the CSR of the
// vertices it owns (the others empty), owner ranges ascend with the handle index, so the merged payload is the
// concatenation in handle order and the merged offsets are the element-wise sums -- peer copies over NVLink plus one add
// kernel per array, no host staging.  Handle 0 then answers ag2_pg_graph_fetch / ag2_pg_job_travel / ag2_pg_job_dump for
// the whole graph.
int ag2_pg_group_gather(ag2_pg* const* pgs, int n)
{
    if (!pgs || n < 1 || !pgs[0]) return AG2_EINVAL;
    ag2_pg* pg = pgs[0];
    for (int r = 0; r < n; ++r)
        if (!pgs[r] || !pgs[r]->have_graph) return fail(pg, AG2_ESTATE, "ag2_pg_group_gather: handle %d has no graph (call ag2_pg_join)", r);
    if (n == 1) return AG2_OK;
    const int64_t nv = pg->n_vertices;
    int64_t npos = 0, nedge = 0;
    for (int r = 0; r < n; ++r) {
        if (pgs[r]->n_vertices != nv) return fail(pg, AG2_EINVAL, "ag2_pg_group_gather: handles differ in their vertex sets");
        npos += pgs[r]->g_npos;
        nedge += pgs[r]->g_nedge;
    }
    for (int r = 0; r < n; ++r) {   // the payload copies read what the join of every handle wrote
        PG_CUDA(cudaSetDevice(pgs[r]->device));
        PG_CUDA(cudaStreamSynchronize(pgs[r]->stream));
    }
    PG_CUDA(cudaSetDevice(pg->device));
    Dev<uint32_t> ctg, ref, eto;
    Dev<uint16_t> cnt;
    Dev<int32_t> estep;
    Dev<unsigned long long> tmp;
    PG_CUDA(ctg.alloc(npos));
    PG_CUDA(ref.alloc(npos));
    PG_CUDA(cnt.alloc(npos));
    PG_CUDA(eto.alloc(nedge));
    PG_CUDA(estep.alloc(nedge));
    PG_CUDA(tmp.alloc(nv + 1));
    int64_t bp = 0, be = 0;
    for (int r = 0; r < n; ++r) {
        ag2_pg* S = pgs[r];
        const int sd = S->device, dd = pg->device;
        if (S->g_npos) {
            PG_CUDA(cudaMemcpyPeerAsync(ctg.p + bp, dd, S->g_ctg.p, sd, (size_t)S->g_npos * 4, pg->stream));
            PG_CUDA(cudaMemcpyPeerAsync(ref.p + bp, dd, S->g_ref.p, sd, (size_t)S->g_npos * 4, pg->stream));
            PG_CUDA(cudaMemcpyPeerAsync(cnt.p + bp, dd, S->g_cnt.p, sd, (size_t)S->g_npos * 2, pg->stream));
        }
        if (S->g_nedge) {
            PG_CUDA(cudaMemcpyPeerAsync(eto.p + be, dd, S->g_edge_to.p, sd, (size_t)S->g_nedge * 4, pg->stream));
            PG_CUDA(cudaMemcpyPeerAsync(estep.p + be, dd, S->g_edge_step.p, sd, (size_t)S->g_nedge * 4, pg->stream));
        }
        bp += S->g_npos;
        be += S->g_nedge;
        if (r > 0) {
            PG_CUDA(cudaMemcpyPeerAsync(tmp.p, dd, S->g_pos_off.p, sd, (size_t)(nv + 1) * 8, pg->stream));
            pg_add_u64_kernel<<<grid_for(nv + 1), 256, 0, pg->stream>>>(pg->g_pos_off.p, tmp.p, nv + 1);
            PG_CUDA(cudaMemcpyPeerAsync(tmp.p, dd, S->g_edge_off.p, sd, (size_t)(nv + 1) * 8, pg->stream));
            pg_add_u64_kernel<<<grid_for(nv + 1), 256, 0, pg->stream>>>(pg->g_edge_off.p, tmp.p, nv + 1);
        }
    }
    PG_CUDA(cudaStreamSynchronize(pg->stream));
    PG_CUDA(cudaGetLastError());
    pg->g_ctg.swap(ctg);
    pg->g_ref.swap(ref);
    pg->g_cnt.swap(cnt);
    pg->g_edge_to.swap(eto);
    pg->g_edge_step.swap(estep);
    pg->g_npos = npos;
    pg->g_nedge = nedge;
    pg->stats.positions = npos;
    pg->stats.edges = nedge;
    return AG2_OK;
}

int ag2_pg_join(ag2_pg* pg, const ag2_pg_params* params)
{
    PoolScope pool_scope(pg ? &pg->pool : nullptr);
    if (!pg || !params) return fail(pg, AG2_EINVAL, "ag2_pg_join: bad arguments");
    if (!pg->have_streams) return fail(pg, AG2_ESTATE, "ag2_pg_join: no streams (call ag2_pg_extract)");
    if (params->epsilon < 0) return fail(pg, AG2_EINVAL, "ag2_pg_join: epsilon < 0");
    PG_CUDA(cudaSetDevice(pg->device));
    const int64_t nv = pg->n_vertices, nt = pg->n_tuples, ne = pg->n_edges;
    const uint32_t eps = (uint32_t)std::min<int64_t>(params->epsilon, 0xffffffffll);
    const int vbits = bits_for((uint64_t)std::max<int64_t>(nv, 1));
    Lap lap;
    PG_CUDA(cudaEventRecord(pg->ev[2], pg->stream));

    // ---- positions: stable sort by vertex, per-vertex first-fit clustering, (ctg, ref) order, compaction
    Dev<uint32_t> seg_cnt, seg_off, nrep, key_sorted, o_ctg, o_ref, cnt_scratch;
    Dev<unsigned long long> items, items_sorted, rep_scratch, nrep64;
    Dev<uint16_t> o_cnt;
    PG_CUDA(seg_cnt.alloc(nv + 1));
    PG_CUDA(seg_off.alloc(nv + 1));
    PG_CUDA(nrep.alloc(nv + 1));
    PG_CUDA(nrep64.alloc(nv + 1));
    PG_CUDA(pg->g_pos_off.alloc(nv + 1));
    PG_CUDA(cudaMemsetAsync(seg_cnt.p, 0, (size_t)(nv + 1) * 4, pg->stream));
    PG_CUDA(cudaMemsetAsync(nrep.p, 0, (size_t)(nv + 1) * 4, pg->stream));
    PG_CUDA(items.alloc(nt));
    PG_CUDA(items_sorted.alloc(nt));
    PG_CUDA(key_sorted.alloc(nt));
    PG_CUDA(rep_scratch.alloc(nt));
    PG_CUDA(cnt_scratch.alloc(nt));
    PG_CUDA(o_ctg.alloc(nt));
    PG_CUDA(o_ref.alloc(nt));
    PG_CUDA(o_cnt.alloc(nt));
    lap.mark("join: position allocations");
    if (nt) {
        pg_hist_kernel<<<grid_for(nt), 256, 0, pg->stream>>>(pg->t_vertex.p, nt, seg_cnt.p);
        pg_pack_pairs_kernel<<<grid_for(nt), 256, 0, pg->stream>>>(pg->t_ctg.p, pg->t_ref.p, nt, items.p);
        PG_TRY(sort_pairs(pg, pg->t_vertex.p, key_sorted.p, items.p, items_sorted.p, nt, 0, vbits));
        pg->stats.launches += 2;
    }
    PG_TRY(exclusive_sum(pg, seg_cnt.p, seg_off.p, nv + 1));
    Dev<uint32_t> big_list, big_n;
    PG_CUDA(big_n.alloc(1));
    PG_CUDA(cudaMemsetAsync(big_n.p, 0, 4, pg->stream));
    PG_CUDA(cudaEventRecord(pg->ev[0], pg->stream));          // sorted: the clustering starts
    if (nt) {
        // a vertex with more than kSmallJoin items has at least kSmallJoin + 1 of the nt items
        PG_CUDA(big_list.alloc(nt / (kSmallJoin + 1) + 1));
        pg_join_small_kernel<<<grid_for(nv), 256, 0, pg->stream>>>(seg_off.p, nv, items_sorted.p, eps, o_ctg.p, o_ref.p, o_cnt.p, nrep.p,
                                                                   big_list.p, big_n.p);
        pg_join_kernel<<<148 * 8, kJoinWarps * 32, 0, pg->stream>>>(seg_off.p, nv, items_sorted.p, eps, rep_scratch.p, cnt_scratch.p,
                                                                    o_ctg.p, o_ref.p, o_cnt.p, nrep.p, big_list.p, big_n.p);
        pg->stats.launches += 2;
    }
    pg_widen_kernel<<<grid_for(nv + 1), 256, 0, pg->stream>>>(nrep.p, nv + 1, nrep64.p);
    PG_TRY(exclusive_sum(pg, nrep64.p, pg->g_pos_off.p, nv + 1));
    unsigned long long npos = 0;
    PG_CUDA(cudaMemcpyAsync(&npos, pg->g_pos_off.p + nv, 8, cudaMemcpyDeviceToHost, pg->stream));
    PG_CUDA(cudaStreamSynchronize(pg->stream));
    PG_CUDA(pg->g_ctg.alloc((int64_t)npos));
    PG_CUDA(pg->g_ref.alloc((int64_t)npos));
    PG_CUDA(pg->g_cnt.alloc((int64_t)npos));
    if (npos) {
        pg_compact_kernel<<<grid_for(nt), 256, 0, pg->stream>>>(key_sorted.p, nt, seg_off.p, nrep.p, pg->g_pos_off.p, o_ctg.p, o_ref.p, o_cnt.p,
                                                                 pg->g_ctg.p, pg->g_ref.p, pg->g_cnt.p);
        ++pg->stats.launches;
    }
    pg->g_npos = (int64_t)npos;
    lap.mark("join: positions (kernels)");
    PG_CUDA(cudaEventRecord(pg->ev[1], pg->stream));          // positions done: the edges start

    // ---- edges: sort by (from, to, step), unique
    Dev<unsigned long long> ft, ft2, e_cnt64;
    Dev<uint32_t> st, st2, flag, epos, per_vertex;
    PG_CUDA(ft.alloc(ne));
    PG_CUDA(ft2.alloc(ne));
    PG_CUDA(st.alloc(ne));
    PG_CUDA(st2.alloc(ne));
    PG_CUDA(flag.alloc(ne + 1));
    PG_CUDA(epos.alloc(ne + 1));
    PG_CUDA(per_vertex.alloc(nv + 1));
    PG_CUDA(e_cnt64.alloc(nv + 1));
    PG_CUDA(pg->g_edge_off.alloc(nv + 1));
    PG_CUDA(cudaMemsetAsync(per_vertex.p, 0, (size_t)(nv + 1) * 4, pg->stream));
    lap.mark("join: edge allocations");
    unsigned long long nedge = 0;
    if (ne) {
        pg_pack_pairs_kernel<<<grid_for(ne), 256, 0, pg->stream>>>(pg->e_to.p, pg->e_from.p, ne, ft.p);   // to | from << 32
        // LSD: step first, then (from, to); both passes stable
        PG_TRY(sort_pairs(pg, (const uint32_t*)pg->e_step.p, st2.p, ft.p, ft2.p, ne, 0, 32));
        PG_TRY(sort_pairs(pg, ft2.p, ft.p, st2.p, st.p, ne, 0, 32 + vbits));
        pg_edge_flag_kernel<<<grid_for(ne), 256, 0, pg->stream>>>(ft.p, st.p, ne, flag.p);
        PG_CUDA(cudaMemsetAsync(flag.p + ne, 0, 4, pg->stream));
        PG_TRY(exclusive_sum(pg, flag.p, epos.p, ne + 1));
        uint32_t n32 = 0;
        PG_CUDA(cudaMemcpyAsync(&n32, epos.p + ne, 4, cudaMemcpyDeviceToHost, pg->stream));
        PG_CUDA(cudaStreamSynchronize(pg->stream));
        nedge = n32;
        PG_CUDA(pg->g_edge_to.alloc((int64_t)nedge));
        PG_CUDA(pg->g_edge_step.alloc((int64_t)nedge));
        pg_edge_scatter_kernel<<<grid_for(ne), 256, 0, pg->stream>>>(ft.p, st.p, flag.p, epos.p, ne, pg->g_edge_to.p, pg->g_edge_step.p, per_vertex.p);
        pg->stats.launches += 3;
    } else {
        PG_CUDA(pg->g_edge_to.alloc(0));
        PG_CUDA(pg->g_edge_step.alloc(0));
    }
    pg_widen_kernel<<<grid_for(nv + 1), 256, 0, pg->stream>>>(per_vertex.p, nv + 1, e_cnt64.p);
    PG_TRY(exclusive_sum(pg, e_cnt64.p, pg->g_edge_off.p, nv + 1));
    pg->stats.launches += 2;
    pg->g_nedge = (int64_t)nedge;
    PG_CUDA(cudaEventRecord(pg->ev[3], pg->stream));
    PG_CUDA(cudaStreamSynchronize(pg->stream));
    PG_CUDA(cudaGetLastError());
    float ms = 0;
    cudaEventElapsedTime(&ms, pg->ev[2], pg->ev[3]);
    pg->stats.join_ms = ms;
    cudaEventElapsedTime(&ms, pg->ev[2], pg->ev[0]);
    pg->stats.join_sort_ms = ms;
    cudaEventElapsedTime(&ms, pg->ev[0], pg->ev[1]);
    pg->stats.join_cluster_ms = ms;
    cudaEventElapsedTime(&ms, pg->ev[1], pg->ev[3]);
    pg->stats.join_edges_ms = ms;
    lap.mark("join: edges (kernels)");
    pg->stats.positions = pg->g_npos;
    pg->stats.edges = pg->g_nedge;
    pg->have_graph = true;
    return AG2_OK;
}

int ag2_pg_build(ag2_pg* pg, const ag2_pg_params* params)
{
    PG_TRY(ag2_pg_extract(pg, params));
    return ag2_pg_join(pg, params);
}

int ag2_pg_get_stats(ag2_pg* pg, ag2_pg_stats* out)
{
    if (!pg || !out) return AG2_EINVAL;
    *out = pg->stats;
    return AG2_OK;
}

int ag2_pg_graph_fetch(ag2_pg* pg, int64_t* pos_off, uint32_t* ctg, uint32_t* ref, uint16_t* count, int64_t pos_cap,
                       int64_t* edge_off, uint32_t* edge_to, int32_t* edge_step, int64_t edge_cap)
{
    if (!pg) return AG2_EINVAL;
    if (!pg->have_graph) return fail(pg, AG2_ESTATE, "ag2_pg_graph_fetch: no graph (call ag2_pg_build)");
    PG_CUDA(cudaSetDevice(pg->device));
    const int64_t nv = pg->n_vertices;
    if ((ctg || ref || count) && pos_cap < pg->g_npos) return fail(pg, AG2_ECAP, "ag2_pg_graph_fetch: %lld positions", (long long)pg->g_npos);
    if ((edge_to || edge_step) && edge_cap < pg->g_nedge) return fail(pg, AG2_ECAP, "ag2_pg_graph_fetch: %lld edges", (long long)pg->g_nedge);
    if (pos_off) PG_CUDA(cudaMemcpy(pos_off, pg->g_pos_off.p, (size_t)(nv + 1) * 8, cudaMemcpyDeviceToHost));
    if (edge_off) PG_CUDA(cudaMemcpy(edge_off, pg->g_edge_off.p, (size_t)(nv + 1) * 8, cudaMemcpyDeviceToHost));
    if (ctg && pg->g_npos) PG_CUDA(cudaMemcpy(ctg, pg->g_ctg.p, (size_t)pg->g_npos * 4, cudaMemcpyDeviceToHost));
    if (ref && pg->g_npos) PG_CUDA(cudaMemcpy(ref, pg->g_ref.p, (size_t)pg->g_npos * 4, cudaMemcpyDeviceToHost));
    if (count && pg->g_npos) PG_CUDA(cudaMemcpy(count, pg->g_cnt.p, (size_t)pg->g_npos * 2, cudaMemcpyDeviceToHost));
    if (edge_to && pg->g_nedge) PG_CUDA(cudaMemcpy(edge_to, pg->g_edge_to.p, (size_t)pg->g_nedge * 4, cudaMemcpyDeviceToHost));
    if (edge_step && pg->g_nedge) PG_CUDA(cudaMemcpy(edge_step, pg->g_edge_step.p, (size_t)pg->g_nedge * 4, cudaMemcpyDeviceToHost));
    return AG2_OK;
}

}  // extern "C"
