// xdrop_lane.cuh -- the "lane" path of the X-drop extension: ONE THREAD owns one extension direction
// and runs the reference's loops literally (xdrop_align, MC/xdrop_gapalign.cpp:11-213; align_ex
// :263-357) -- row-major order, the running best, the undecayed horizontal gap across pruned cells
// and the MIN_SCORE arithmetic are the reference's own statements, so there is nothing to prove.
//
// What makes it a GPU kernel is the layout around those loops:
//   * 32 chains advance in lock step per warp (SIMT): one DP row at a time, the inner column loop
//     runs until the widest of the 32 bands is done; chains of different length are refilled from a
//     global queue at block boundaries so lanes stay busy;
//   * the band state (16 bits per column, 128-column circular window) and the block's target codes
//     (2 bit) live in shared memory laid out [slot][thread]: lane i always hits bank i -- conflict
//     free whatever column each lane is at;
//   * query codes of the block are normalised once per block into a per-thread scratch line and read
//     one 32-bit word per 16 rows;
//   * traceback cells are 4 bits, accumulated in a register and stored one 32-bit word per 8 cells
//     into the thread's 64-byte row slot (column b at nibble b mod 128);
//   * the walk writes its columns in walk order into the direction's workspace area;
//     assemble_record() (one warp per record, coalesced) puts them into final order.
// A chain whose band exceeds 120 columns, or that outgrows its workspace reservation, is handed to
// the wide row-parallel kernel (xdrop_device.cuh), which has no such limits.
#pragma once

#include "xdrop_device.cuh"

namespace ag2 {

constexpr int kLaneThreads = 64;    // threads per CTA of the lane kernel
constexpr int kBandSlots = 128;     // circular band window (columns)
constexpr int kBandMax = 120;       // widest band a lane may hold: a row's first and last column must not
                                    // share a 32-bit traceback word of the circular 128-column row
constexpr int kSeqWords = 46;       // 736 two-bit codes
constexpr int kNeg16 = -16000;      // MIN_SCORE on the lane path: same arithmetic, fits int16 storage
constexpr int kLaneTbBytes = (kMaxBlk + 2) * 64;
constexpr int kLaneScratch = kLaneTbBytes + 256; // + normalised query codes of the block

// Band state of one column in 16 bits: (x << 1) | dead with x = e + 1.  A live column always has
// e = h - 1 (every statement that writes both sets them so: :58-59, :122/:135, :148-149), so x = h;
// a pruned column that stays in the band has h = MIN_SCORE and keeps its e (:111); the sentinel is
// (MIN_SCORE, MIN_SCORE) (:161-162).  Two consecutive columns of a thread share one 32-bit word
// so that lane i only ever touches bank i.
struct LaneSmem {
    int16_t band[kBandSlots / 2][kLaneThreads][2];
    uint32_t tseq[kSeqWords][kLaneThreads];  // target block, 16 codes per word, extension order
};

__device__ __forceinline__ int16_t *band_slot(LaneSmem &sm, int tid, int b)
{
    const int slot = b & (kBandSlots - 1);
    return &sm.band[slot >> 1][tid][slot & 1];
}
__device__ __forceinline__ int16_t band_live(int h) { return (int16_t)(h * 2); }
constexpr int16_t kBandSentinel = (int16_t)(((kNeg16 + 1) * 2) | 1);

// State of a direction at a block boundary, written by the pair kernel when it hands the direction over: the lane kernel
// continues from that block instead of restarting the direction.
struct LaneResume {
    int32_t qidx, tidx, ncols, qcons, tcons, last_op, nblocks, pad;
    int64_t seg;
    unsigned long long cells, rows, blocks;
};

struct LaneArgs {
    StreamSignal sig;        // streamed runs: piece flags to wait for, chunk counters to raise (xdrop_device.cuh)
    PackedSeqs seqs;
    const Candidate *cand;
    const ExtGeom *geom;
    ChainResult *res;        // [2 * n]
    uint32_t *meta;          // block metadata words
    char *ws_q, *ws_t;
    uint8_t *scratch;        // kLaneScratch bytes per resident thread
    int64_t n_chains;
    const int32_t *queue;    // nullptr: chains 0..n_chains-1; else chain ids to run
    LaneResume *resume;      // parallel to `queue` (lane kernel: read) / to `wide_queue` (pair kernel: written); may be null
    unsigned long long *next;
    int32_t *wide_queue;
    unsigned int *wide_count;
    ChainCounters *counters;
    unsigned int *done_ctas; // pair kernel: every CTA adds 1 when it has published all its hand-overs (may be null); done_ctas[1] is raised when the first CTA starts
    // pair kernel: directions whose LAST block is longer than a normal one are set aside and run together at the end
    // (xdrop_pair.cuh); all null = off
    int32_t *defer_queue;    // preset to -1
    LaneResume *defer_resume;
    unsigned int *defer_ctl; // [0] slots reserved, [1] tickets taken, [2] warps in their main phase
};

// Hand a direction over: the payload first, then -- fenced -- the queue entry, which a concurrently running consumer
// (xdrop_stream_kernel) polls; the queue is pre-set to -1.
__device__ __forceinline__ void publish_chain(int32_t *queue, unsigned slot, int64_t chain)
{
    __threadfence();
    *reinterpret_cast<volatile int32_t *>(queue + slot) = (int32_t)chain;
}

// 16 two-bit codes starting at base position p (forward) of a packed sequence
__device__ __forceinline__ uint32_t load16(const uint32_t *seq, int64_t p)
{
    const int64_t w = p >> 4;
    const int sh = 2 * (int)(p & 15);
    const uint32_t lo = seq[w];
    if (sh == 0) return lo;
    const uint32_t hi = seq[w + 1];
    return (lo >> sh) | (hi << (32 - sh));
}

// reverse the order of the 16 two-bit codes in a word
__device__ __forceinline__ uint32_t rev16(uint32_t x)
{
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
    x = ((x >> 8) & 0x00FF00FFu) | ((x & 0x00FF00FFu) << 8);
    return (x >> 16) | (x << 16);
}

// 16 codes of a sequence in extension order: positions p0, p0+dir, ... (dir = +1 / -1), clipped to
// [0, len) (codes outside are 0 and never used)
__device__ __forceinline__ uint32_t fetch16(const uint32_t *seq, int64_t off, int64_t len, int64_t p0, int dir)
{
    if (dir > 0) {
        if (p0 >= len) return 0;
        return load16(seq, off + p0);
    }
    // backward: codes p0, p0-1, ..., p0-15 = reverse of the forward word starting at p0-15
    int64_t s = p0 - 15;
    int drop = 0;
    if (s < 0) {
        drop = (int)(-s);
        s = 0;
    }
    if (p0 < 0) return 0;
    uint32_t w = load16(seq, off + s);     // codes s .. s+15
    w = rev16(w);                          // codes s+15 .. s
    return drop ? (w >> (2 * drop)) : w;   // start at p0 = s + 15 - drop
}

__device__ __forceinline__ uint32_t spread_bits16(uint32_t m) // 16 bits -> 16 two-bit groups 0b11 / 0b00
{
    uint32_t x = m & 0xffffu;
    x = (x | (x << 8)) & 0x00FF00FFu;
    x = (x | (x << 4)) & 0x0F0F0F0Fu;
    x = (x | (x << 2)) & 0x33333333u;
    x = (x | (x << 1)) & 0x55555555u;
    return x | (x << 1);
}

// 16 "not upper-case ACGT" bits in extension order (for the reverse strand's complement): bit i = position p0 + dir * i,
// positions below 0 read as 0
__device__ __forceinline__ uint32_t load16_bits(const uint32_t *bits, int64_t s) // bit j = position s + j
{
    const int64_t w = s >> 5;
    const int sh = (int)(s & 31);
    uint32_t r = bits[w] >> sh;
    if (sh > 16) r |= bits[w + 1] << (32 - sh);
    return r & 0xffffu;
}
__device__ __forceinline__ uint32_t fetch16_bits(const uint32_t *bits, int64_t off, int64_t p0, int dir)
{
    if (dir > 0) {
        if (p0 >= 0) return load16_bits(bits, off + p0);
        if (p0 <= -16) return 0;
        return load16_bits(bits, off) << (int)(-p0) & 0xffffu;
    }
    if (p0 < 0) return 0;
    int64_t s = p0 - 15;
    int drop = 0;
    if (s < 0) {
        drop = (int)(-s);
        s = 0;
    }
    const uint32_t x = load16_bits(bits, off + s);        // bit j = position s + j
    uint32_t r = x;                                        // reverse the 16 bits
    r = ((r >> 1) & 0x5555u) | ((r & 0x5555u) << 1);
    r = ((r >> 2) & 0x3333u) | ((r & 0x3333u) << 2);
    r = ((r >> 4) & 0x0f0fu) | ((r & 0x0f0fu) << 4);
    r = ((r >> 8) & 0x00ffu) | ((r & 0x00ffu) << 8);      // bit k = position s + 15 - k
    return r >> drop;                                      // bit i = position p0 - i
}

struct LaneBlock {
    int M, N;
    int ae, be;
};

// xdrop_align forward pass, one thread.  Returns 0, or 1 if the band outgrew the window.
__device__ int lane_dp(LaneSmem &sm, int tid, const uint32_t *__restrict__ qcodes, uint8_t *__restrict__ tb, int M, int N,
                       int &ae_out, int &be_out, unsigned &cells_out, unsigned &rows_out)
{
    const int xd = kXdrop;
    int bsize = min(N, xd) + 1, first = 0, best = 0, ae = 0, be = 0;
    unsigned cells = 0, rows = 0;
    // row 0 (:53-67)
    for (int i = 0; i < bsize; ++i) *band_slot(sm, tid, i) = band_live(-i);

    uint32_t aw = 0;
    for (int a = 1; a <= M; ++a) {
        if (((a - 1) & 15) == 0) aw = qcodes[(a - 1) >> 4];
        const int ac = (int)(aw & 3u);
        aw >>= 2;
        uint32_t *trow = reinterpret_cast<uint32_t *>(tb + (size_t)a * 64);
        int diag = kNeg16, hgap = kNeg16;
        int thr = best - xd;
        cells += (unsigned)(bsize - first);
        ++rows;
        uint32_t tbw = 0;
        int sh = 4 * (first & 7);
        uint32_t bw = sm.tseq[first >> 4][tid] >> (2 * (first & 15));
        int b;
        for (b = first; b < bsize; ++b) { // (:84-140)
            int16_t *cell = band_slot(sm, tid, b);
            const int st = *cell;
            const int x = st >> 1;
            const int e = x - 1;
            const int h = (st & 1) ? kNeg16 : x;
            const int next_diag = h + ((int)(bw & 3u) == ac ? 1 : -1);
            bw >>= 2;
            if (((b + 1) & 15) == 0) bw = sm.tseq[(b + 1) >> 4][tid];
            int sc = diag;
            int nib = kOpSub;
            if (sc < e) { sc = e; nib = kOpGapB; }
            if (sc < hgap) { sc = hgap; nib = kOpGapA; }
            if (sc < thr) { // best - sc > x_dropoff (:109)
                if (first == b) ++first;         // dropped from the band: the column is dead
                else *cell = (int16_t)(st | 1); // h = MIN_SCORE, e unchanged (:111)
            } else {
                if (sc > best) { best = sc; thr = sc - xd; ae = a; be = b; }
                if (e >= sc) nib |= kExtA;     // e - ge >= sc - goe (:121-126)
                if (hgap >= sc) nib |= kExtB;  // (:129-133)
                hgap = sc - 1;
                *cell = band_live(sc);         // h = sc, e = sc - 1
            }
            diag = next_diag;
            tbw |= (uint32_t)nib << sh;
            sh += 4;
            if (sh == 32) {
                trow[(b & (kBandSlots - 1)) >> 3] = tbw;
                tbw = 0;
                sh = 0;
            }
        }
        if (first == bsize) break; // (:142)
        // last_b_index (:113): the last cell of this row that was not pruned.  Every band cell was just rewritten, and
        // a pruned one carries the dead bit, so walk back from the end instead of tracking it per cell.
        int last = bsize - 1;
        while (*band_slot(sm, tid, last) & 1) --last;
        if (last < bsize - 1) {
            if (sh != 0) trow[((b - 1) & (kBandSlots - 1)) >> 3] = tbw;
            bsize = last + 1; // (:144-145)
        } else {
            while (hgap >= best - xd && bsize < N) { // (:147-153)
                *band_slot(sm, tid, bsize) = band_live(hgap);
                hgap -= 1;
                tbw |= (uint32_t)kOpGapA << sh;
                sh += 4;
                if (sh == 32) {
                    trow[(bsize & (kBandSlots - 1)) >> 3] = tbw;
                    tbw = 0;
                    sh = 0;
                }
                ++bsize;
            }
            if (sh != 0) trow[((bsize - 1) & (kBandSlots - 1)) >> 3] = tbw;
        }
        if (bsize < N) { // (:160-164)
            *band_slot(sm, tid, bsize) = kBandSentinel;
            ++bsize;
        }
        if (bsize - first > kBandMax) return 1;
    }
    ae_out = ae;
    be_out = be;
    cells_out = cells;
    rows_out = rows;
    return 0;
}

// Traceback (:170-210) + script_to_aligned_string in walk order + trim_mismatch_end, one thread.
// Writes ASCII columns for walk steps 0..n-1 to wq/wt.  Returns n.
__device__ int lane_walk(const LaneSmem &sm, int tid, const uint32_t *__restrict__ qcodes, const uint8_t *__restrict__ tb,
                         int ae, int be, char *wq, char *wt, int cap, int &qcnt, int &tcnt, int &acnt, bool &trim_ok,
                         int &first_op, int &op_after_trim)
{
    int a = ae, b = be, n = 0, cur = kOpSub;
    int m = 0, q = 0, t = 0, ac = 0;
    bool scanning = true;
    first_op = -1;
    op_after_trim = -1;
    int qi = -1;
    uint32_t qwv = 0;
    while ((a > 0 || b > 0) && n < cap) {
        int cell = kOpGapA; // row 0 is all SCRIPT_GAP_IN_A (:61)
        if (a > 0) {
            const unsigned slot = (unsigned)b & 127u;
            const unsigned byte = tb[(size_t)a * 64 + (slot >> 1)];
            cell = (int)((byte >> ((slot & 1u) * 4)) & 15u);
        }
        int nxt = cell & 3;
        if (cur == kOpGapA && (cell & kExtA)) nxt = kOpGapA;
        if (cur == kOpGapB && (cell & kExtB)) nxt = kOpGapB;
        cur = nxt;
        if (cur != kOpGapA) --a;
        if (cur != kOpGapB) --b;
        int qc = 4, tc = 4;
        if (cur != kOpGapA) {
            if ((a >> 4) != qi) {
                qi = a >> 4;
                qwv = qcodes[qi];
            }
            qc = (int)((qwv >> (2 * (a & 15))) & 3u);
        }
        if (cur != kOpGapB) tc = (int)((sm.tseq[b >> 4][tid] >> (2 * (b & 15))) & 3u);
        if (n == 0) first_op = cur;
        if (!scanning && op_after_trim < 0) op_after_trim = cur;
        if (scanning) { // trim_mismatch_end scans from the END of the block's alignment = walk start
            ++ac;
            if (cur != kOpGapA) ++q;
            if (cur != kOpGapB) ++t;
            m = (qc == tc) ? m + 1 : 0;
            if (m == kTailMatch) scanning = false;
        }
        wq[n] = code_char(qc);
        wt[n] = code_char(tc);
        ++n;
    }
    qcnt = q;
    tcnt = t;
    acnt = ac;
    // k = n-1-w counts down; true iff 4 matches were found and k > 0 after the final --k (gapalign.cpp:67)
    trim_ok = !scanning && (n - 1 - ac) > 0;
    return n;
}

struct LaneChain { // registers of the chain a lane is running
    int64_t chain;   // -1: none
    Candidate c;
    ExtGeom ge;
    int rlen;
    int64_t roff;
    int qsize, tsize, q0, inc;
    int64_t t0;
    int qidx, tidx;
    int ncols, qcons, tcons, last_op, nblocks;
    int64_t seg;     // next free column of the direction's workspace area
    int64_t seg_end;
    int64_t meta, meta_end;
    unsigned long long cells, rows, blocks;
};

template <typename Args>   // LaneArgs or ChainArgs
__device__ __forceinline__ void lane_start_chain(const Args &g, int64_t chain, LaneChain &s)
{
    s.chain = chain;
    const int64_t ci = chain >> 1;
    const bool forward = (chain & 1) != 0;
    s.c = g.cand[ci];
    s.ge = g.geom[ci];
    s.qidx = s.tidx = 0;
    s.ncols = s.qcons = s.tcons = 0;
    s.last_op = -1;
    s.nblocks = 0;
    s.cells = s.rows = s.blocks = 0;
    if (!s.ge.valid) return;
    s.rlen = g.seqs.read_len[s.c.read];
    s.roff = g.seqs.read_off[s.c.read];
    const int read_start = s.c.loc2;
    const int64_t ref_start = s.c.loc1 - 1;
    s.qsize = forward ? s.rlen - read_start : read_start;
    s.tsize = forward ? s.ge.right : s.ge.left;
    s.q0 = forward ? read_start : read_start - 1;
    s.t0 = forward ? ref_start : ref_start - 1;
    s.inc = forward ? 1 : -1;
    s.seg = forward ? s.ge.slot + s.ge.capL : s.ge.slot;
    s.seg_end = s.seg + (forward ? s.ge.capR : s.ge.capL);
    s.meta = forward ? s.ge.meta + s.ge.nmetaL : s.ge.meta;
    s.meta_end = s.meta + (forward ? s.ge.nmetaR : s.ge.nmetaL);
}

__device__ __forceinline__ void lane_resume(LaneChain &s, const LaneResume &r)
{
    s.qidx = r.qidx;
    s.tidx = r.tidx;
    s.ncols = r.ncols;
    s.qcons = r.qcons;
    s.tcons = r.tcons;
    s.last_op = r.last_op;
    s.nblocks = r.nblocks;
    s.seg = r.seg;
    s.cells = r.cells;
    s.rows = r.rows;
    s.blocks = r.blocks;
}
__device__ __forceinline__ LaneResume lane_save(const LaneChain &s)
{
    LaneResume r;
    r.qidx = s.qidx;
    r.tidx = s.tidx;
    r.ncols = s.ncols;
    r.qcons = s.qcons;
    r.tcons = s.tcons;
    r.last_op = s.last_op;
    r.nblocks = s.nblocks;
    r.pad = 0;
    r.seg = s.seg;
    r.cells = s.cells;
    r.rows = s.rows;
    r.blocks = s.blocks;
    return r;
}

// align_ex's bookkeeping after one block's walk (:334-357): metadata word, consumed counts, next block origin.
// Returns 0 = chain continues, 1 = chain finished.
__device__ __forceinline__ int lane_block_tail(const LaneArgs &g, LaneChain &s, int qblk, int tblk, bool last_block, int ae, int be,
                                               int nops, int qcnt, int tcnt, int acnt, bool trim_ok, int first_op, int op_after)
{
    const bool full_map = (qblk - ae <= kFullMapSlack) || (tblk - be <= kFullMapSlack); // (:334-335)
    const bool stop = !full_map || last_block;
    int skip = 0;
    if (!stop) {
        if (!trim_ok) { // (:349) the block's columns are dropped and the direction ends
            g.meta[s.meta + s.nblocks++] = (uint32_t)nops | ((uint32_t)nops << 16);
            return 1;
        }
        skip = acnt;
    }
    g.meta[s.meta + s.nblocks++] = (uint32_t)nops | ((uint32_t)skip << 16);
    s.seg += nops;
    if (nops - skip > 0) {
        s.last_op = skip ? op_after : first_op;
        s.ncols += nops - skip;
        s.qcons += skip ? ae - qcnt : ae;
        s.tcons += skip ? be - tcnt : be;
    }
    if (stop) return 1;
    s.qidx += ae - qcnt; // (:354-355)
    s.tidx += be - tcnt;
    return 0;
}


// A direction on the row-parallel kernel (one WARP per direction, xdrop_device.cuh) but in THIS path's format -- per block a
// segment of the workspace in walk order and a metadata word -- from the block boundary held in `s` (all lanes hold the same
// `s`).  It is how a direction the pair kernel hands over late is finished: continued at the block it stopped at, not
// restarted (a restart of a 10 kb direction is ~17 ms on one warp, and the launch waited for it).
// Returns 0 = the direction is complete (result written), 2 = this block's band does not fit K columns per lane (`s` is at
// that block's boundary: a wider K continues from it), 3 = a reservation of the format is exceeded (restart on the plain form).
template <int K>
__device__ int run_chain_resumed(const ChainArgs &g, LaneChain &s, WarpSmem &sm, uint8_t *tb, int lane, ChainCounters &lc)
{
    for (int iter = 0; iter < (1 << 14); ++iter) { // the bound only guards against a hang
        // retrieve_next_aln_block (MC/gapalign.cpp:9-45)
        const int qleft = s.qsize - s.qidx, tleft = s.tsize - s.tidx;
        int qblk, tblk;
        bool last_block;
        if (qleft < kBlk + kBlkSlack || tleft < kBlk + kBlkSlack) {
            qblk = min(qleft, stretch_0p2(tleft));
            tblk = min(tleft, stretch_0p2(qleft));
            last_block = true;
        } else {
            qblk = kBlk;
            tblk = kBlk;
            last_block = false;
        }
        // stage the block: one code per byte, extension order (as run_chain)
        __syncwarp();
        for (int i = lane; i < qblk; i += 32) {
            const int p = s.q0 + s.inc * (s.qidx + i);
            int code;
            if (s.c.strand == 0) {
                code = get2(g.seqs.reads2, s.roff + p);
            } else {
                const int64_t fp = s.roff + (s.rlen - 1 - p);
                code = get2(g.seqs.reads2, fp);
                if (!get1(g.seqs.reads_irr, fp)) code ^= 3;
            }
            sm.A[i] = (uint8_t)code;
        }
        for (int i = lane; i < tblk; i += 32) sm.B[i] = (uint8_t)get2(g.seqs.ref2, s.t0 + (int64_t)s.inc * (s.tidx + i));
        __syncwarp();
        int ae = 0, be = 0;
        ChainCounters bc = {0, 0, 0, 0, 0};
        if (qblk > 0 && tblk > 0) {
            if (dp_block<K>(sm.A, qblk, sm.B, tblk, tb, lane, ae, be, bc)) return 2;
        }
        lc.cells += bc.cells;
        lc.rows += bc.rows;
        lc.blocks += bc.blocks;
        lc.interior += bc.interior;
        __syncwarp();
        const int cap = (int)min((int64_t)(2 * kMaxBlk), s.seg_end - s.seg);
        if (ae + be > cap || s.meta + s.nblocks >= s.meta_end) return 3;
        int nops = 0, qcnt = 0, tcnt = 0, acnt = 0, trim_m = 0, trim_w = -1;
        if (lane == 0) nops = walk_block<K>(tb, ae, be, sm, qcnt, tcnt, acnt, trim_m, trim_w);
        nops = __shfl_sync(kFull, nops, 0);
        qcnt = __shfl_sync(kFull, qcnt, 0);
        tcnt = __shfl_sync(kFull, tcnt, 0);
        acnt = __shfl_sync(kFull, acnt, 0);
        trim_m = __shfl_sync(kFull, trim_m, 0);
        trim_w = __shfl_sync(kFull, trim_w, 0);
        __syncwarp();
        // the segment: walk step i (from the block's end cell towards its origin) at seg + i
        int qi = 0, ti = 0;
        for (int base = 0; base < nops; base += 32) {
            const int i = base + lane;
            const bool on = i < nops;
            const int op = on ? sm.ops[i] : kOpGapA;
            const unsigned qm = __ballot_sync(kFull, on && op != kOpGapA);
            const unsigned tm = __ballot_sync(kFull, on && op != kOpGapB);
            const unsigned lt = (1u << lane) - 1u;
            if (on) {
                g.ws_q[s.seg + i] = op != kOpGapA ? code_char(sm.A[ae - 1 - (qi + __popc(qm & lt))]) : '-';
                g.ws_t[s.seg + i] = op != kOpGapB ? code_char(sm.B[be - 1 - (ti + __popc(tm & lt))]) : '-';
            }
            qi += __popc(qm);
            ti += __popc(tm);
        }
        // lane_block_tail, by all lanes alike; lane 0 writes the metadata word
        const bool full_map = (qblk - ae <= kFullMapSlack) || (tblk - be <= kFullMapSlack); // (:334-335)
        const bool stop = !full_map || last_block;
        const bool trim_ok = trim_m == kTailMatch && (nops - 2 - trim_w) > 0;
        int skip = 0;
        if (!stop) {
            if (!trim_ok) { // (:349) the block's columns are dropped and the direction ends
                if (lane == 0) g.meta[s.meta + s.nblocks] = (uint32_t)nops | ((uint32_t)nops << 16);
                ++s.nblocks;
                break;
            }
            skip = acnt;
        }
        if (lane == 0) g.meta[s.meta + s.nblocks] = (uint32_t)nops | ((uint32_t)skip << 16);
        ++s.nblocks;
        s.seg += nops;
        if (nops - skip > 0) {
            s.last_op = sm.ops[skip];                  // first op after the trimmed ones (the walk's first op without a trim)
            s.ncols += nops - skip;
            s.qcons += skip ? ae - qcnt : ae;
            s.tcons += skip ? be - tcnt : be;
        }
        if (stop) break;
        s.qidx += ae - qcnt; // (:354-355)
        s.tidx += be - tcnt;
    }
    __syncwarp(); // the segments were written by all lanes
    if (lane == 0) {
        const ChainResult out = {s.ncols, s.qcons, s.tcons, s.last_op, s.nblocks, 0, 0, 0};
        g.res[s.chain] = out;
        signal_direction_done(g.sig, s.chain);
    }
    return 0;
}

// Entry t of the hand-over queue: continued from the block the pair kernel stopped at (run_chain_resumed: 128 columns per
// lane first, the kernel's K from the block that does not fit them).  false = not possible (no state, or a reservation of
// the pair / lane format is exceeded): the caller restarts the direction on the plain form.
template <int K>
__device__ bool continue_handed_over(const ChainArgs &g, int64_t chain, unsigned long long t, WarpSmem &sm, uint8_t *tb, int lane,
                                     ChainCounters &ctr)
{
    if (!g.resume) return false;
    LaneChain s;
    lane_start_chain(g, chain, s);
    if (!s.ge.valid) return false;
    lane_resume(s, g.resume[t]);
    ChainCounters lc = {0, 0, 0, 0, 0};
    int rc = 2;
    if (K > kNarrowK && g.try_narrow) rc = run_chain_resumed<kNarrowK>(g, s, sm, tb, lane, lc);
    if (rc == 2) rc = run_chain_resumed<K>(g, s, sm, tb, lane, lc);
    if (rc != 0) return false;
    ctr.cells += lc.cells + s.cells;      // s.cells / rows / blocks: what the pair kernel did before it handed over
    ctr.rows += lc.rows + s.rows;
    ctr.blocks += lc.blocks + s.blocks;
    ctr.interior += lc.interior;
    return true;
}

// One block of align_ex for the lane's chain.  Returns 0 = chain continues, 1 = chain finished,
// 2 = hand the chain to the wide path.
__device__ int lane_block(const LaneArgs &g, LaneChain &s, LaneSmem &sm, int tid, uint8_t *scratch)
{
    // retrieve_next_aln_block (MC/gapalign.cpp:9-45)
    const int qleft = s.qsize - s.qidx, tleft = s.tsize - s.tidx;
    int qblk, tblk;
    bool last_block;
    if (qleft < kBlk + kBlkSlack || tleft < kBlk + kBlkSlack) {
        qblk = min(qleft, stretch_0p2(tleft));
        tblk = min(tleft, stretch_0p2(qleft));
        last_block = true;
    } else {
        qblk = kBlk;
        tblk = kBlk;
        last_block = false;
    }
    uint32_t *qcodes = reinterpret_cast<uint32_t *>(scratch + kLaneTbBytes);
    uint8_t *tb = scratch;
    // stage the block in extension order: query -> scratch line, target -> shared memory
    {
        const int qw = (qblk + 15) >> 4, tw = (tblk + 16) >> 4; // one spare target code: column N reads B[N] (:85)
        const int qdir = s.c.strand == 0 ? s.inc : -s.inc;
        for (int w = 0; w < qw; ++w) {
            const int p = s.q0 + s.inc * (s.qidx + 16 * w);           // oriented read position of the word's first code
            const int64_t fp = s.c.strand == 0 ? p : (int64_t)s.rlen - 1 - p; // position in the read as stored
            uint32_t v = fetch16(g.seqs.reads2, s.roff, s.rlen, fp, qdir);
            if (s.c.strand != 0) v ^= ~spread_bits16(fetch16_bits(g.seqs.reads_irr, s.roff, fp, qdir)); // complement ACGT only
            qcodes[w] = v;
        }
        for (int w = 0; w < tw && w < kSeqWords; ++w)
            sm.tseq[w][tid] = fetch16(g.seqs.ref2, 0, g.seqs.ref_len, s.t0 + (int64_t)s.inc * (s.tidx + 16 * w), s.inc);
    }
    int ae = 0, be = 0;
    unsigned cells = 0, rows = 0;
    if (qblk > 0 && tblk > 0) {
        if (lane_dp(sm, tid, qcodes, tb, qblk, tblk, ae, be, cells, rows)) return 2;
        s.blocks += 1;
    }
    s.cells += cells;
    s.rows += rows;
    const int cap = (int)min((int64_t)(2 * kMaxBlk), s.seg_end - s.seg);
    if (ae + be > cap || s.meta + s.nblocks >= s.meta_end) return 2; // reservation too small: wide path
    int qcnt, tcnt, acnt, first_op, op_after;
    bool trim_ok;
    const int nops = lane_walk(sm, tid, qcodes, tb, ae, be, g.ws_q + s.seg, g.ws_t + s.seg, cap, qcnt, tcnt, acnt, trim_ok,
                               first_op, op_after);
    return lane_block_tail(g, s, qblk, tblk, last_block, ae, be, nops, qcnt, tcnt, acnt, trim_ok, first_op, op_after);
}

// Body of xdrop_lane_kernel: a thread keeps pulling directions from the queue and advances its
// current one by one block per round; the warp leaves when the queue is drained and every lane is idle.
__device__ void lane_kernel_body(const LaneArgs &g, LaneSmem &sm, int tid, uint8_t *scratch)
{
    LaneChain s;
    s.chain = -1;
    unsigned long long cells = 0, rows = 0, blocks = 0, wide = 0;
    bool drained = false;
    for (;;) {
        if (s.chain < 0 && !drained) {
            const unsigned long long t = atomicAdd(g.next, 1ull);
            if ((int64_t)t < g.n_chains) {
                lane_start_chain(g, g.queue ? (int64_t)g.queue[t] : (int64_t)t, s);
                if (g.queue && g.resume && s.ge.valid) lane_resume(s, g.resume[t]);
                if (!s.ge.valid) {
                    const ChainResult out = {0, 0, 0, -1, 0, 0, 0, 0};
                    g.res[s.chain] = out;
                    signal_direction_done(g.sig, s.chain);
                    s.chain = -1;
                }
            } else {
                drained = true;
            }
        }
        const bool active = s.chain >= 0;
        if (!__any_sync(kFull, active || !drained)) break;
        if (active) {
            const int rc = lane_block(g, s, sm, tid, scratch);
            if (rc == 1) {
                const ChainResult out = {s.ncols, s.qcons, s.tcons, s.last_op, s.nblocks, 0, 0, 0};
                g.res[s.chain] = out;
                signal_direction_done(g.sig, s.chain);
                cells += s.cells;
                rows += s.rows;
                blocks += s.blocks;
                s.chain = -1;
            } else if (rc == 2) {
                g.wide_queue[atomicAdd(g.wide_count, 1u)] = (int32_t)s.chain;
                ++wide;
                s.chain = -1;
            }
        }
    }
    atomicAdd(&g.counters->cells, cells);
    atomicAdd(&g.counters->rows, rows);
    atomicAdd(&g.counters->blocks, blocks);
    atomicAdd(&g.counters->wide, wide);
}

// dst[i] = src[i] (forward) or src[-i] (reverse) for i < n, by the 32 lanes of a warp: 4 bytes per lane and step through
// ALIGNED destination words (two aligned source words, a funnel shift, and a byte reversal when the order flips); the up to
// 3 + 3 bytes around them singly.  Sources lie inside the workspace strings, which are allocated with slack on both sides of
// what the shifted word loads touch.
__device__ __forceinline__ void warp_copy_bytes(char *dst, const char *src, int n, bool reverse, int lane)
{
#ifdef AG2_EMU
    for (int i = lane; i < n; i += 32) dst[i] = reverse ? src[-i] : src[i];
#else
    int head = (int)((4 - (reinterpret_cast<uintptr_t>(dst) & 3)) & 3);
    if (head > n) head = n;
    if (lane < head) dst[lane] = reverse ? src[-lane] : src[lane];
    const int nw = (n - head) >> 2;
    uint32_t *d = reinterpret_cast<uint32_t *>(dst + head);
    for (int w = lane; w < nw; w += 32) {
        const char *first = reverse ? src - (head + 4 * w) - 3 : src + head + 4 * w;   // lowest address of the 4 source bytes
        const uintptr_t a = reinterpret_cast<uintptr_t>(first) & ~(uintptr_t)3;
        const unsigned sh = (unsigned)(reinterpret_cast<uintptr_t>(first) & 3) * 8;
        const uint32_t lo = *reinterpret_cast<const uint32_t *>(a);
        const uint32_t hi = sh ? *reinterpret_cast<const uint32_t *>(a + 4) : 0u;
        uint32_t x = __funnelshift_r(lo, hi, sh);
        if (reverse) x = __byte_perm(x, 0, 0x0123);
        d[w] = x;
    }
    const int done = head + 4 * nw;
    if (lane < n - done) dst[done + lane] = reverse ? src[-(done + lane)] : src[done + lane];
#endif
}

// Dense strings of one record from the two directions' workspace areas.  Executed by one warp.
//   left, lane path : blocks last -> first, each in walk order, minus the very first column (:401-402)
//   right, lane path: blocks first -> last, each in reversed walk order
//   wide path       : already in final order (see ExtGeom)
__device__ void assemble_record(const ExtGeom &ge, const ChainResult &l, const ChainResult &r, const uint32_t *meta,
                                const char *ws_q, const char *ws_t, char *out_q, char *out_t, int lane)
{
    int64_t o = 0;
    // ---- left ----
    if (l.ncols > 1) {
        if (l.mode == 1) {
            const int64_t src = ge.slot + ge.capL - l.ncols + 1;
            warp_copy_bytes(out_q, ws_q + src, l.ncols - 1, false, lane);
            warp_copy_bytes(out_t, ws_t + src, l.ncols - 1, false, lane);
            o = l.ncols - 1;
        } else {
            int64_t end = ge.slot; // one past the last block's segment
            for (int k = 0; k < l.nblocks; ++k) end += meta[ge.meta + k] & 0xffffu;
            bool drop = true;
            for (int k = l.nblocks - 1; k >= 0; --k) {
                const uint32_t w = meta[ge.meta + k];
                const int n = (int)(w & 0xffffu), skip = (int)(w >> 16);
                const int64_t seg = end - n;
                end = seg;
                int lo = skip;
                if (drop && n - skip > 0) {
                    ++lo;
                    drop = false;
                }
                warp_copy_bytes(out_q + o, ws_q + seg + lo, n - lo, false, lane);
                warp_copy_bytes(out_t + o, ws_t + seg + lo, n - lo, false, lane);
                o += n - lo;
            }
        }
    }
    // ---- right ----
    if (r.mode == 1) {
        const int64_t src = ge.slot + ge.capL;
        warp_copy_bytes(out_q + o, ws_q + src, r.ncols, false, lane);
        warp_copy_bytes(out_t + o, ws_t + src, r.ncols, false, lane);
    } else {
        int64_t seg = ge.slot + ge.capL;
        for (int k = 0; k < r.nblocks; ++k) {
            const uint32_t w = meta[ge.meta + ge.nmetaL + k];
            const int n = (int)(w & 0xffffu), skip = (int)(w >> 16);
            const int e = n - skip;
            warp_copy_bytes(out_q + o, ws_q + seg + n - 1, e, true, lane);
            warp_copy_bytes(out_t + o, ws_t + seg + n - 1, e, true, lane);
            o += e;
            seg += n;
        }
    }
}

} // namespace ag2
