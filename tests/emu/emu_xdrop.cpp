// emu_xdrop.cpp -- runs the product's X-drop device code (xdrop_device.cuh) under the one-warp CPU
// emulation of warp_emu.h.  TEST INFRASTRUCTURE ONLY: built by tests/emu/build.sh into
// tests/emu/libemu_xdrop.so and loaded by tests/test_emu_xdrop.py.
#define AG2_EMU 1
#include "../../aligngraph2_b200/csrc/xdrop_device.cuh"
#include "../../aligngraph2_b200/csrc/xdrop_lane.cuh"
#include "../../aligngraph2_b200/csrc/xdrop_pair.cuh"

#include <vector>

using namespace ag2;

namespace {

template <int K>
void block_impl(const uint8_t *A, int M, const uint8_t *B, int N, int *ae, int *be, uint8_t *ops, int *nops,
                long *cells, int *overflow, long *interior)
{
    static WarpSmem sm;
    std::vector<uint8_t> tb((size_t)(kMaxBlk + 2) * TbLayout<K>::kRowBytes);
    memcpy(sm.A, A, M);
    memcpy(sm.B, B, N);
    int r_ae = 0, r_be = 0, r_n = 0, r_over = 0;
    ChainCounters ctr = {0, 0, 0, 0, 0};
    warp_emu::run_warp([&]() {
        const int lane = warp_emu::st().cur;
        ChainCounters lc = {0, 0, 0, 0, 0};
        int a = 0, b = 0;
        const int over = dp_block<K>(sm.A, M, sm.B, N, tb.data(), lane, a, b, lc);
        __syncwarp();
        if (lane == 0) {
            r_over = over;
            ctr = lc;
            if (!over) {
                int q, t, ac, m, w;
                r_n = walk_block<K>(tb.data(), a, b, sm, q, t, ac, m, w);
                r_ae = a;
                r_be = b;
            }
        }
    });
    *ae = r_ae;
    *be = r_be;
    *nops = r_n;
    *overflow = r_over;
    *cells = (long)ctr.cells;
    *interior = (long)ctr.interior;
    static const uint8_t map[3] = {3, 0, 6}; // kOpSub/kOpGapA/kOpGapB -> reference op codes
    for (int i = 0; i < r_n; ++i) ops[i] = map[sm.ops[i]];
}

void pack2(const char *s, int64_t n, std::vector<uint32_t> &w2, std::vector<uint32_t> &irr, int64_t base)
{
    for (int64_t i = 0; i < n; ++i) {
        int code = 0, ir = 1;
        switch (s[i]) {
        case 'A': code = 0; ir = 0; break;
        case 'C': code = 1; ir = 0; break;
        case 'G': code = 2; ir = 0; break;
        case 'T': code = 3; ir = 0; break;
        case 'a': code = 0; break;
        case 'c': code = 1; break;
        case 'g': code = 2; break;
        case 't': code = 3; break;
        default: code = 0; break;
        }
        const int64_t p = base + i;
        w2[p >> 4] |= (uint32_t)code << (2 * (p & 15));
        if (ir) irr[p >> 5] |= 1u << (p & 31);
    }
}

} // namespace

extern "C" {

void emu_dp_block(int K, const uint8_t *A, int M, const uint8_t *B, int N, int *ae, int *be, uint8_t *ops,
                  int *nops, long *cells, int *overflow, long *interior)
{
    if (K == 2) block_impl<2>(A, M, B, N, ae, be, ops, nops, cells, overflow, interior);
    else if (K == 3) block_impl<3>(A, M, B, N, ae, be, ops, nops, cells, overflow, interior);
    else if (K == 4) block_impl<4>(A, M, B, N, ae, be, ops, nops, cells, overflow, interior);
    else if (K == 8) block_impl<8>(A, M, B, N, ae, be, ops, nops, cells, overflow, interior);
    else block_impl<23>(A, M, B, N, ae, be, ops, nops, cells, overflow, interior);
}

// extend_candidate on raw ASCII (read as given + strand flag).  Returns ok; rec = qb qe sb se aln_len
// cells wide; strings into qaln/taln.
int emu_extend(int K, const char *ref, long ref_len, const char *read, int read_len, int strand, long loc1,
               int loc2, long *rec, char *qaln, char *taln)
{
    std::vector<uint32_t> ref2((ref_len >> 4) + 2, 0), dummy((ref_len >> 5) + 2, 0);
    pack2(ref, ref_len, ref2, dummy, 0);
    std::vector<uint32_t> rd2((read_len >> 4) + 2, 0), irr((read_len >> 5) + 2, 0);
    pack2(read, read_len, rd2, irr, 0);
    int64_t roff = 0;
    int32_t rlen = read_len;
    PackedSeqs sq = {ref2.data(), ref_len, rd2.data(), irr.data(), &roff, &rlen};
    Candidate c = {0, strand, loc1, loc2, 7};
    ExtGeom g;
    int64_t nm_;
    const int64_t cap = setup_one(c, sq, 1, g, nm_);
    if (!g.valid) return -1;
    std::vector<char> wq(cap + 1, '?'), wt(cap + 1, '?');
    ChainResult res[2];
    ChainCounters ctr = {0, 0, 0, 0, 0};
    std::vector<uint8_t> tb((size_t)(kMaxBlk + 2) * TbLayout<23>::kRowBytes);
    static WarpSmem sm;
    ChainArgs a = {};
    a.seqs = sq;
    a.cand = &c;
    a.geom = &g;
    a.res = res;
    a.ws_q = wq.data();
    a.ws_t = wt.data();
    a.n_chains = 2;
    int wide = 0;
    warp_emu::run_warp([&]() {
        const int lane = warp_emu::st().cur;
        ChainCounters lc = {0, 0, 0, 0, 0};
        for (int chain = 0; chain < 2; ++chain) {
            bool ok;
            if (K == 2) ok = run_chain<2>(a, chain, sm, tb.data(), lane, lc);
            else if (K == 3) ok = run_chain<3>(a, chain, sm, tb.data(), lane, lc);
            else if (K == 4) ok = run_chain<4>(a, chain, sm, tb.data(), lane, lc);
            else ok = run_chain<23>(a, chain, sm, tb.data(), lane, lc);
            if (!ok) {
                if (lane == 0) ++wide;
                run_chain<23>(a, chain, sm, tb.data(), lane, lc);
            }
        }
        if (lane == 0) ctr = lc;
    });
    Record o;
    int64_t sb;
    finalize_one(c, g, res[0], res[1], rlen, o, sb);
    rec[0] = o.qb;
    rec[1] = o.qe;
    rec[2] = o.sb;
    rec[3] = o.se;
    rec[4] = o.aln_len;
    rec[5] = (long)ctr.cells;
    rec[6] = wide;
    rec[7] = (long)ctr.interior;
    memcpy(qaln, wq.data() + sb, o.aln_len);
    memcpy(taln, wt.data() + sb, o.aln_len);
    return o.ok;
}


// The whole lane path for a batch of candidates over one reference, as ONE emulated warp:
// setup_one -> lane_kernel_body (+ wide rerun of handed-over directions) -> finalize_one -> assemble_record.
// reads: concatenated ASCII with offs[n+1]; cand: n x (read, strand, loc1, loc2).  Outputs per candidate:
// rec[8] = ok qb qe sb se aln_len mode_left mode_right; strings concatenated into qaln/taln at aoff[i].
static int g_emu_defer = 0;
static int g_emu_row_resume = 0;   // handed-over directions continued by the row-parallel form (as xdrop_stream_kernel does), not by the lane kernel
static int batch_impl(bool pair, const char *ref, long ref_len, const char *reads, const long *offs, int n_reads, const long *cand,
                      int n, long *rec, long *aoff, char *qaln, char *taln, long *stats /* cells wide handed */)
{
    std::vector<uint32_t> ref2((ref_len >> 4) + 8, 0), dummy((ref_len >> 5) + 8, 0);
    pack2(ref, ref_len, ref2, dummy, 0);
    std::vector<int64_t> roff(n_reads + 1);
    std::vector<int32_t> rlen(n_reads);
    int64_t p = 0;
    for (int r = 0; r < n_reads; ++r) {
        roff[r] = p;
        rlen[r] = (int32_t)(offs[r + 1] - offs[r]);
        p += (rlen[r] + 31) & ~31;
    }
    std::vector<uint32_t> rd2((p >> 4) + 8, 0), irr((p >> 5) + 8, 0);
    for (int r = 0; r < n_reads; ++r) pack2(reads + offs[r], rlen[r], rd2, irr, roff[r]);
    PackedSeqs sq = {ref2.data(), ref_len, rd2.data(), irr.data(), roff.data(), rlen.data()};
    std::vector<Candidate> cs(n);
    std::vector<ExtGeom> ge(n);
    int64_t ws = 0, nm = 0;
    for (int i = 0; i < n; ++i) {
        cs[i] = Candidate{(int32_t)cand[4 * i], (int32_t)cand[4 * i + 1], cand[4 * i + 2], (int32_t)cand[4 * i + 3], 3};
        int64_t m;
        const int64_t cap = setup_one(cs[i], sq, n_reads, ge[i], m);
        ge[i].slot = ws;
        ge[i].meta = nm;
        ws += cap;
        nm += m;
    }
    std::vector<char> wq(ws + 64, '?'), wt(ws + 64, '?');
    std::vector<uint32_t> meta(nm + 16, 0);
    std::vector<ChainResult> res(2 * n);
    std::vector<uint8_t> scratch((size_t)kLaneScratch * 32 + 64);
    std::vector<int32_t> wide_queue(2 * n + 1);
    unsigned long long next = 0;
    unsigned int wide_count = 0;
    ChainCounters ctr = {0, 0, 0, 0, 0};
    static LaneSmem lsm;
    LaneArgs a = {};
    a.seqs = sq;
    a.cand = cs.data();
    a.geom = ge.data();
    a.res = res.data();
    a.meta = meta.data();
    a.ws_q = wq.data();
    a.ws_t = wt.data();
    a.scratch = (uint8_t *)(((uintptr_t)scratch.data() + 15) & ~(uintptr_t)15);
    a.n_chains = 2 * n;
    a.next = &next;
    a.wide_queue = wide_queue.data();
    a.wide_count = &wide_count;
    a.counters = &ctr;
    unsigned int handed = 0;
    if (pair) {
        // pair path first (two directions per lane); what it hands over is restarted by the lane kernel body
        static PairSmem psm;
        std::vector<uint8_t> pscratch(kPairCtaScratch + 64);
        std::vector<int32_t> lane_queue(2 * n + 1);
        std::vector<LaneResume> lane_resume(2 * n + 1);
        LaneArgs pa = a;
        std::vector<int32_t> defer_queue(2 * n + 1, -1);
        std::vector<LaneResume> defer_resume(2 * n + 1);
        unsigned int defer_ctl[4] = {0, 0, 0, 0};
        if (g_emu_defer) {   // long last blocks set aside and run at the end (one emulated warp = the whole launch)
            pa.defer_queue = defer_queue.data();
            pa.defer_resume = defer_resume.data();
            pa.defer_ctl = defer_ctl;
        }
        pa.scratch = (uint8_t *)(((uintptr_t)pscratch.data() + 15) & ~(uintptr_t)15);
        pa.wide_queue = lane_queue.data();
        pa.wide_count = &handed;
        pa.resume = lane_resume.data();
        warp_emu::run_warp([&]() {
            const int lane = warp_emu::st().cur;
            pair_kernel_body(pa, psm, lane, pa.scratch);
        });
        if (handed && g_emu_row_resume) {
            // the consumer's way: continue_handed_over (128 columns per lane, then 736), a restart on the plain form if that gives up
            static WarpSmem sm;
            std::vector<uint8_t> tb((size_t)(kMaxBlk + 2) * TbLayout<23>::kRowBytes);
            ChainArgs w = {};
            w.seqs = sq;
            w.cand = cs.data();
            w.geom = ge.data();
            w.res = res.data();
            w.ws_q = wq.data();
            w.ws_t = wt.data();
            w.try_narrow = 1;
            w.resume = lane_resume.data();
            w.meta = meta.data();
            warp_emu::run_warp([&]() {
                const int lane = warp_emu::st().cur;
                ChainCounters lc = {0, 0, 0, 0, 0};
                for (unsigned t = 0; t < handed; ++t) {
                    bool ok = continue_handed_over<23>(w, lane_queue[t], t, sm, tb.data(), lane, lc);
                    if (!ok) ok = run_chain<4>(w, lane_queue[t], sm, tb.data(), lane, lc);
                    if (!ok) run_chain<23>(w, lane_queue[t], sm, tb.data(), lane, lc);
                }
                if (lane == 0) ctr.cells += lc.cells;
            });
        } else if (handed) {
            next = 0;
            a.queue = lane_queue.data();
            a.resume = lane_resume.data();
            a.n_chains = handed;
            warp_emu::run_warp([&]() {
                const int lane = warp_emu::st().cur;
                lane_kernel_body(a, lsm, lane, a.scratch + (size_t)lane * kLaneScratch);
            });
        }
    } else {
        warp_emu::run_warp([&]() {
            const int lane = warp_emu::st().cur;
            lane_kernel_body(a, lsm, lane, a.scratch + (size_t)lane * kLaneScratch);
        });
    }
    // wide rerun
    if (wide_count) {
        static WarpSmem sm;
        std::vector<uint8_t> tb((size_t)(kMaxBlk + 2) * TbLayout<23>::kRowBytes);
        ChainArgs w = {};
        w.seqs = sq;
        w.cand = cs.data();
        w.geom = ge.data();
        w.res = res.data();
        w.ws_q = wq.data();
        w.ws_t = wt.data();
        warp_emu::run_warp([&]() {
            const int lane = warp_emu::st().cur;
            ChainCounters lc = {0, 0, 0, 0, 0};
            for (unsigned k = 0; k < wide_count; ++k) run_chain<23>(w, wide_queue[k], sm, tb.data(), lane, lc);
            if (lane == 0) {
                ctr.cells += lc.cells;
            }
        });
    }
    long o = 0;
    for (int i = 0; i < n; ++i) {
        Record r;
        int64_t sb;
        finalize_one(cs[i], ge[i], res[2 * i], res[2 * i + 1], ge[i].valid ? rlen[cs[i].read] : 0, r, sb);
        rec[8 * i + 0] = r.ok;
        rec[8 * i + 1] = r.qb;
        rec[8 * i + 2] = r.qe;
        rec[8 * i + 3] = r.sb;
        rec[8 * i + 4] = r.se;
        rec[8 * i + 5] = r.aln_len;
        rec[8 * i + 6] = res[2 * i].mode;
        rec[8 * i + 7] = res[2 * i + 1].mode;
        aoff[i] = o;
        if (r.ok) {
            const ExtGeom g2 = ge[i];
            const ChainResult l = res[2 * i], rr = res[2 * i + 1];
            char *oq = qaln + o, *ot = taln + o;
            warp_emu::run_warp([&]() { assemble_record(g2, l, rr, meta.data(), wq.data(), wt.data(), oq, ot, warp_emu::st().cur); });
            o += r.aln_len;
        }
    }
    stats[0] = (long)ctr.cells;
    stats[1] = (long)wide_count;
    stats[2] = (long)handed;
    return 0;
}

void emu_set_defer(int on) { g_emu_defer = on; }
void emu_set_row_resume(int on) { g_emu_row_resume = on; }

int emu_lane_batch(const char *ref, long ref_len, const char *reads, const long *offs, int n_reads, const long *cand,
                   int n, long *rec, long *aoff, char *qaln, char *taln, long *stats)
{
    return batch_impl(false, ref, ref_len, reads, offs, n_reads, cand, n, rec, aoff, qaln, taln, stats);
}

// The pair path (xdrop_pair.cuh) for a batch: pair kernel body -> lane kernel body for what it hands over -> wide rerun.
int emu_pair_batch(const char *ref, long ref_len, const char *reads, const long *offs, int n_reads, const long *cand,
                   int n, long *rec, long *aoff, char *qaln, char *taln, long *stats)
{
    return batch_impl(true, ref, ref_len, reads, offs, n_reads, cand, n, rec, aoff, qaln, taln, stats);
}

} // extern "C"
#ifdef AG2_EMU_STATS
extern "C" void emu_pair_stats(long *out)
{
    memcpy(out, &ag2::g_pes, sizeof(ag2::g_pes));
    memset(&ag2::g_pes, 0, sizeof(ag2::g_pes));
}
#endif
