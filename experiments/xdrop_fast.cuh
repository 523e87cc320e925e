// xdrop_fast.cuh -- the fast path of the X-drop extension: same results as dp_block<K> /
// run_chain<K> of xdrop_device.cuh (and therefore as MC/xdrop_gapalign.cpp), ~4x fewer instructions.
//
// Row-parallel X-drop in packed fp16 (exact small integers, see h2ops.cuh):
//
//   * lane l owns the 4 adjacent columns base+4l .. base+4l+3 of a 128-column window that slides
//     with the band (whole-lane rebase by warp shuffle when the band's first column leaves lane 0);
//   * column state h (best), e (best_gap) of every column OUTSIDE the band is -inf.  This makes band
//     membership implicit: the cell right of the band sees h = -inf from the sentinel column
//     (xdrop_gapalign.cpp:160-164) and e = -inf, so its local score L is -inf and it only gets a
//     value from the horizontal gap -- which is exactly the reference's band-growth loop (:147-153).
//     No per-cell range tests remain;
//   * with gap_open = 0 / gap_extend = 1 the row recurrence collapses to two prefix maxima over the
//     row (DESIGN.md "Row-parallel X-drop" has the proof):
//         s(b)       = max_{b' <= b} (L(b') + b') - b            (cell score incl. horizontal gap)
//         pruned(b) <=> s(b) < max(best, max_{b' <= b} L(b')) - X   (row-major running best, :109)
//     both computed by ONE warp scan over a packed (T, L) pair;
//   * a pruned cell between unpruned ones keeps a stale e and needs the reference's undecayed
//     horizontal gap for its traceback op (:104-112); that case (~3 % of rows) takes a warp-uniform
//     slow branch.
//
// Traceback cells are 4 bits, one 64-byte row per DP row, written coalesced; the walk stages 32 rows
// at a time in shared memory.
#pragma once

#include "h2ops.cuh"
#include "xdrop_device.cuh"

namespace ag2 {

constexpr int kFastCols = 128;
constexpr int kLutCols = kMaxBlk + kFastCols + 16;
constexpr uint32_t kNegInf2 = 0xFC00FC00u; // (-inf, -inf)
constexpr int kStageRows = 32;

struct alignas(16) FastSmem {
    uint8_t A[kMaxBlk];            // query block codes, extension order
    uint8_t B[kMaxBlk];            // target block codes
    // lut[c][b] = high byte of fp16(+1) if B[b-1] == c else of fp16(-1): the substitution score of
    // column b for query code c, ready to be dropped into a half2 by one PRMT; fp16(-inf) for the
    // columns the band may never reach (b >= N, or b > N when N <= 30)
    uint8_t lut[4][kLutCols];
    uint8_t ops[2 * kMaxBlk];      // traceback ops, walk order
    uint8_t stage[kStageRows * 64];// traceback rows staged for the walk
};

__device__ __forceinline__ uint32_t sel32(uint32_t mask, uint32_t a, uint32_t b) { return (a & mask) | (b & ~mask); }

// 0xFFFF-per-half masks of two pairs -> 4 bits (bit j = column cb + j)
__device__ __forceinline__ unsigned mask4(uint32_t m0, uint32_t m1)
{
    return (m0 & 1u) | ((m0 >> 15) & 2u) | ((m1 & 1u) << 2) | ((m1 >> 13) & 8u);
}

// Forward pass of xdrop_align in packed fp16.  Returns 0, or 1 if the band outgrew the window.
__device__ int dp_block_h2(const FastSmem &sm, int M, int N, uint8_t *tb, int lane, int &ae_out, int &be_out,
                           ChainCounters &ctr)
{
    const uint32_t kM30 = 0xCF80CF80u; // (-30, -30)
    const uint32_t kM1 = 0xBC00BC00u;  // (-1, -1)
    int base = 0, cb = 4 * lane;
    int first = 0, bsize = min(N, kXdrop) + 1, best = 0, ae = 0, be = 0;
    const int nlim = N <= kXdrop ? N + 1 : N; // columns a cell may ever occupy: b < nlim
    uint32_t C0 = h2_from_ints(cb, cb + 1), C1 = h2_from_ints(cb + 2, cb + 3);
    // row 0 (:53-67)
    uint32_t H0, H1, E0, E1;
    {
        const uint32_t in0 = hlt2_mask(C0, h2_from_ints(bsize, bsize)), in1 = hlt2_mask(C1, h2_from_ints(bsize, bsize));
        const uint32_t z = 0;
        H0 = sel32(in0, hsub2(z, C0), kNegInf2);
        H1 = sel32(in1, hsub2(z, C1), kNegInf2);
        E0 = sel32(in0, hadd2(hsub2(z, C0), kM1), kNegInf2);
        E1 = sel32(in1, hadd2(hsub2(z, C1), kM1), kNegInf2);
    }
    const uint32_t nl2 = h2_from_ints(nlim, nlim);
    uint32_t LN0 = hlt2_mask(C0, nl2), LN1 = hlt2_mask(C1, nl2);
    const uint32_t notlane0 = lane == 0 ? 0u : 0xffffffffu;
    uint32_t bestpack = h2_from_ints(0, 0) | 0x0000FC00u; // (-inf, best)
    bestpack = (bestpack & 0xffff0000u) | 0xFC00u;
    unsigned long long cells = 0, rows = 0, interior_rows = 0;

    for (int a = 1; a <= M; ++a) {
        const int ac = sm.A[a - 1];
        const uint32_t mw = *reinterpret_cast<const uint32_t *>(&sm.lut[ac][cb]);
        cells += (unsigned)(bsize - first);
        ++rows;
        const uint32_t bs2 = h2_from_ints(bsize, bsize);
        // diagonal inputs: h of the column to the left
        uint32_t hp3 = __shfl_up_sync(kFull, H1, 1);
        hp3 = sel32(notlane0, hp3, kNegInf2);
        const uint32_t HP0 = prmt(hp3, H0, 0x5432), HP1 = prmt(H0, H1, 0x5432);
        const uint32_t D0 = hadd2(HP0, prmt(mw, 0, 0x1404)), D1 = hadd2(HP1, prmt(mw, 0, 0x3424));
        const uint32_t L0 = hmax2(D0, E0), L1 = hmax2(D1, E1);
        const uint32_t OB0 = hlt2_mask(D0, E0), OB1 = hlt2_mask(D1, E1);
        const uint32_t T0 = hadd2(L0, C0), T1 = hadd2(L1, C1);
        // lane-local inclusive prefix maxima of T and L
        uint32_t q0 = hmax2(T0, prmt(T0, kNegInf2, 0x1054));
        uint32_t q1 = hmax2(T1, prmt(T1, kNegInf2, 0x1054));
        q1 = hmax2(q1, prmt(q0, q0, 0x3232));
        uint32_t r0 = hmax2(L0, prmt(L0, kNegInf2, 0x1054));
        uint32_t r1 = hmax2(L1, prmt(L1, kNegInf2, 0x1054));
        r1 = hmax2(r1, prmt(r0, r0, 0x3232));
        // one warp scan over (max T, max L) of the lanes
        uint32_t tot = prmt(q1, r1, 0x7632);
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) tot = hmax2(tot, __shfl_up_sync(kFull, tot, d));
        uint32_t exc = __shfl_up_sync(kFull, tot, 1);
        exc = hmax2(sel32(notlane0, exc, kNegInf2), bestpack);
        const int rowmax = h2_hi_int(__shfl_sync(kFull, tot, 31));
        const uint32_t cT = prmt(exc, exc, 0x1010), cL = prmt(exc, exc, 0x3232);
        q0 = hmax2(q0, cT);
        q1 = hmax2(q1, cT);
        r0 = hmax2(r0, cL);
        r1 = hmax2(r1, cL);
        // scores, pruning
        const uint32_t S0 = hsub2(q0, C0), S1 = hsub2(q1, C1);
        const uint32_t U0 = hge2_mask(S0, hadd2(r0, kM30)) & LN0, U1 = hge2_mask(S1, hadd2(r1, kM30)) & LN1;
        uint32_t OA0 = hgt2_mask(q0, T0), OA1 = hgt2_mask(q1, T1);
        const uint32_t ex0 = prmt(cT, q0, 0x5410), ex1 = prmt(q0, q1, 0x5432);
        const uint32_t FB0 = (OA0 | heq2_mask(ex0, T0)) & hlt2_mask(C0, bs2);
        const uint32_t FB1 = (OA1 | heq2_mask(ex1, T1)) & hlt2_mask(C1, bs2);
        const uint32_t XA0 = heq2_mask(E0, S0), XA1 = heq2_mask(E1, S1);

        const unsigned u4 = mask4(U0, U1);
        const int lmin = u4 ? cb + __ffs((int)u4) - 1 : 0x7fffffff;
        const int lmax = u4 ? cb + 31 - __clz((int)u4) : -1;
        const int fmin = __reduce_min_sync(kFull, lmin);
        if (fmin == 0x7fffffff) break; // every cell pruned (:142)
        const int lastu = __reduce_max_sync(kFull, lmax);
        const int cnt = __reduce_add_sync(kFull, __popc(u4));
        if (rowmax > best) {
            const uint32_t rm2 = h2_from_ints(rowmax, rowmax);
            const unsigned e4 = mask4(heq2_mask(L0, rm2), heq2_mask(L1, rm2));
            be = __reduce_min_sync(kFull, e4 ? cb + __ffs((int)e4) - 1 : 0x7fffffff);
            ae = a;
            best = rowmax;
            bestpack = (rm2 & 0xffff0000u) | 0xFC00u;
        }

        uint32_t KE0 = 0, KE1 = 0; // pruned cells that keep their stale e
        if (cnt != lastu - fmin + 1) {
            // interior pruned cells (see xdrop_device.cuh dp_block): exact op from the undecayed gap
            ++interior_rows;
            int L[4], s[4], k[4];
            L[0] = h2_lo_int(hmax2(L0, h2_from_ints(-4000, -4000)));
            L[1] = h2_hi_int(hmax2(L0, h2_from_ints(-4000, -4000)));
            L[2] = h2_lo_int(hmax2(L1, h2_from_ints(-4000, -4000)));
            L[3] = h2_hi_int(hmax2(L1, h2_from_ints(-4000, -4000)));
            s[0] = h2_lo_int(hmax2(S0, h2_from_ints(-2047, -2047)));
            s[1] = h2_hi_int(hmax2(S0, h2_from_ints(-2047, -2047)));
            s[2] = h2_lo_int(hmax2(S1, h2_from_ints(-2047, -2047)));
            s[3] = h2_hi_int(hmax2(S1, h2_from_ints(-2047, -2047)));
            int runk = -1;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                runk = max(runk, (u4 >> j & 1u) ? (((cb + j) << 12) | (s[j] + 2048)) : -1);
                k[j] = runk;
            }
            int sc = runk;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) sc = max(sc, __shfl_up_sync(kFull, sc, d));
            int ck = __shfl_up_sync(kFull, sc, 1);
            if (lane == 0) ck = -1;
            unsigned fix_on = 0, fix_a = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int b = cb + j;
                const int kk = j == 0 ? ck : max(ck, k[j - 1]);
                if (!(u4 >> j & 1u) && b > fmin && b < lastu) {
                    fix_on |= 1u << j;
                    const int realf = ((kk & 4095) - 2048) - 1;
                    if (L[j] < realf) fix_a |= 1u << j;
                }
            }
            const uint32_t on0 = ((fix_on & 1u) ? 0xffffu : 0u) | ((fix_on & 2u) ? 0xffff0000u : 0u);
            const uint32_t on1 = ((fix_on & 4u) ? 0xffffu : 0u) | ((fix_on & 8u) ? 0xffff0000u : 0u);
            const uint32_t fa0 = ((fix_a & 1u) ? 0xffffu : 0u) | ((fix_a & 2u) ? 0xffff0000u : 0u);
            const uint32_t fa1 = ((fix_a & 4u) ? 0xffffu : 0u) | ((fix_a & 8u) ? 0xffff0000u : 0u);
            OA0 = sel32(on0, fa0, OA0);
            OA1 = sel32(on1, fa1, OA1);
            KE0 = on0;
            KE1 = on1;
        }

        // traceback cells of this row
        const uint32_t n0 = (OA0 & 0x00010001u) | (~OA0 & OB0 & 0x00020002u) | (U0 & ((XA0 & 0x00040004u) | (FB0 & 0x00080008u)));
        const uint32_t n1 = (OA1 & 0x00010001u) | (~OA1 & OB1 & 0x00020002u) | (U1 & ((XA1 & 0x00040004u) | (FB1 & 0x00080008u)));
        const uint32_t v = n0 | (n1 << 8);
        const uint32_t w16 = (v & 0x0F0Fu) | ((v >> 12) & 0xF0F0u);
        *reinterpret_cast<uint16_t *>(tb + (size_t)a * 64 + (((cb >> 2) & 31) << 1)) = (uint16_t)w16;

        // column state (:109-136 for band cells, :147-153 for the cells the band grows into)
        H0 = sel32(U0, S0, kNegInf2);
        H1 = sel32(U1, S1, kNegInf2);
        E0 = sel32(U0, hadd2(S0, kM1), sel32(KE0, E0, kNegInf2));
        E1 = sel32(U1, hadd2(S1, kM1), sel32(KE1, E1, kNegInf2));

        first = fmin;
        bsize = lastu + 1;
        if (bsize < N) ++bsize; // sentinel column (:160-164); its state is already (-inf, -inf)
        if (bsize > base + kFastCols) return 1; // the band may have grown past the window
        const int sh = (first - base) >> 2;
        if (sh > 0) { // slide the window: whole lanes
            const bool keep = lane + sh < 32;
            const uint32_t a0 = __shfl_down_sync(kFull, H0, sh), a1 = __shfl_down_sync(kFull, H1, sh);
            const uint32_t b0 = __shfl_down_sync(kFull, E0, sh), b1 = __shfl_down_sync(kFull, E1, sh);
            H0 = keep ? a0 : kNegInf2;
            H1 = keep ? a1 : kNegInf2;
            E0 = keep ? b0 : kNegInf2;
            E1 = keep ? b1 : kNegInf2;
            base += 4 * sh;
            cb += 4 * sh;
            C0 = h2_from_ints(cb, cb + 1);
            C1 = h2_from_ints(cb + 2, cb + 3);
            LN0 = hlt2_mask(C0, nl2);
            LN1 = hlt2_mask(C1, nl2);
        }
    }
    ae_out = ae;
    be_out = be;
    ctr.cells += cells;
    ctr.rows += rows;
    ctr.interior += interior_rows;
    ctr.blocks += 1;
    return 0;
}

// Traceback (:170-210): all lanes stage 32 rows (2 KB) of 4-bit cells in shared memory, lane 0 walks
// them.  Returns the op count; trim_mismatch_end's quantities are derived from the first ops.
__device__ int walk_block_fast(const uint8_t *tb, int ae, int be, FastSmem &sm, int lane, int &qcnt, int &tcnt,
                               int &acnt, int &trim_m, int &trim_w)
{
    int a = ae, b = be, n = 0, cur = kOpSub;
    int a_top = ae;
    while (a_top > 0) {
        const int a_lo = max(1, a_top - (kStageRows - 1));
        {
            const uint4 *src = reinterpret_cast<const uint4 *>(tb + (size_t)a_lo * 64);
            uint4 *dst = reinterpret_cast<uint4 *>(sm.stage);
            const int n16 = (a_top - a_lo + 1) * 4;
            for (int i = lane; i < n16; i += 32) dst[i] = __ldcg(src + i);
        }
        __syncwarp();
        if (lane == 0) {
            while (a >= a_lo && n < 2 * kMaxBlk) {
                const unsigned slot = (unsigned)b & 127u;
                const unsigned byte = sm.stage[(a - a_lo) * 64 + (slot >> 1)];
                const unsigned cell = (byte >> ((slot & 1u) * 4)) & 15u;
                int nxt = cell & 3u;
                if (cur == kOpGapA && (cell & kExtA)) nxt = kOpGapA;
                if (cur == kOpGapB && (cell & kExtB)) nxt = kOpGapB;
                cur = nxt;
                sm.ops[n++] = (uint8_t)cur;
                a -= (cur != kOpGapA);
                b -= (cur != kOpGapB);
            }
            if (n >= 2 * kMaxBlk) a = b = 0; // cannot happen with a consistent traceback; never hang
        }
        a_top = __shfl_sync(kFull, a, 0);
        __syncwarp();
    }
    qcnt = tcnt = acnt = 0;
    trim_m = 0;
    trim_w = -1;
    if (lane == 0) {
        while (b > 0 && n < 2 * kMaxBlk) { // row 0 is all SCRIPT_GAP_IN_A (:61)
            sm.ops[n++] = kOpGapA;
            --b;
        }
        // trim_mismatch_end (MC/gapalign.cpp:47-68) looks at the END of the block's alignment = walk start
        int ra = ae, rb = be, m = 0, q = 0, t = 0, w = 0;
        for (; w < n && m < kTailMatch; ++w) {
            const int op = sm.ops[w];
            bool match = false;
            if (op == kOpGapA) {
                --rb;
                ++t;
            } else if (op == kOpGapB) {
                --ra;
                ++q;
            } else {
                --ra;
                --rb;
                ++q;
                ++t;
                match = sm.A[ra] == sm.B[rb];
            }
            m = match ? m + 1 : 0;
        }
        qcnt = q;
        tcnt = t;
        acnt = w;
        trim_m = m;
        trim_w = m == kTailMatch ? w - 1 : -1;
    }
    return n;
}

// align_ex for one extension direction on the fast path.  false = band overflow, rerun on the wide path.
__device__ bool run_chain_fast(const ChainArgs &g, int64_t chain, FastSmem &sm, uint8_t *tb, int lane, ChainCounters &ctr)
{
    const int64_t ci = chain >> 1;
    const bool forward = (chain & 1) != 0;
    const Candidate c = g.cand[ci];
    const ExtGeom ge = g.geom[ci];
    ChainResult out = {0, 0, 0, -1, 0, 1, 0, 0};
    if (!ge.valid) {
        if (lane == 0) g.res[chain] = out;
        return true;
    }
    const int rlen = g.seqs.read_len[c.read];
    const int64_t roff = g.seqs.read_off[c.read];
    const int read_start = c.loc2;
    const int64_t ref_start = c.loc1 - 1;
    const int qsize = forward ? rlen - read_start : read_start;
    const int tsize = forward ? ge.right : ge.left;
    const int q0 = forward ? read_start : read_start - 1;
    const int64_t t0 = forward ? ref_start : ref_start - 1;
    const int inc = forward ? 1 : -1;
    const int64_t mid = ge.slot + ge.capL;
    int qidx = 0, tidx = 0;
    int ncols = 0, qcons = 0, tcons = 0, last_op = -1;
    ChainCounters lc = {0, 0, 0, 0, 0};

    for (int iter = 0; iter < (1 << 14); ++iter) {
        const int qleft = qsize - qidx, tleft = tsize - tidx;
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
        __syncwarp();
        for (int i = lane; i < qblk; i += 32) {
            const int p = q0 + inc * (qidx + i);
            int code;
            if (c.strand == 0) {
                code = get2(g.seqs.reads2, roff + p);
            } else {
                const int64_t fp = roff + (rlen - 1 - p);
                code = get2(g.seqs.reads2, fp);
                if (!get1(g.seqs.reads_irr, fp)) code ^= 3;
            }
            sm.A[i] = (uint8_t)code;
        }
        // target block: codes for the emit step, and the substitution-score bytes per query code,
        // indexed by DP column (column b compares B[b-1]); everything outside the block mismatches
        const int lut_n = min(tblk + kFastCols + 8, kLutCols);
        const int nlim = tblk <= kXdrop ? tblk + 1 : tblk; // the band never holds a column >= nlim (:147,:160)
        for (int i = lane; i < lut_n; i += 32) {
            int code = -1;
            if (i >= 1 && i <= tblk) {
                code = get2(g.seqs.ref2, t0 + (int64_t)inc * (tidx + i - 1));
                sm.B[i - 1] = (uint8_t)code;
            }
            // +1 / -1 as fp16 high bytes; -inf beyond the last column so such a cell can never score
#pragma unroll
            for (int k = 0; k < 4; ++k) sm.lut[k][i] = i >= nlim ? 0xFC : (code == k ? 0x3C : 0xBC);
        }
        __syncwarp();

        int ae = 0, be = 0;
        if (qblk > 0 && tblk > 0) {
            if (dp_block_h2(sm, qblk, tblk, tb, lane, ae, be, lc)) return false;
        }
        __syncwarp();
        int qcnt, tcnt, acnt, trim_m, trim_w;
        int nops = walk_block_fast(tb, ae, be, sm, lane, qcnt, tcnt, acnt, trim_m, trim_w);
        nops = __shfl_sync(kFull, nops, 0);
        qcnt = __shfl_sync(kFull, qcnt, 0);
        tcnt = __shfl_sync(kFull, tcnt, 0);
        acnt = __shfl_sync(kFull, acnt, 0);
        trim_m = __shfl_sync(kFull, trim_m, 0);
        trim_w = __shfl_sync(kFull, trim_w, 0);
        __syncwarp();

        const bool full_map = (qblk - ae <= kFullMapSlack) || (tblk - be <= kFullMapSlack);
        const bool stop = !full_map || last_block;
        int emit = nops;
        if (!stop) {
            const bool trim = trim_m == kTailMatch && (nops - 2 - trim_w) > 0;
            if (!trim) break;
            emit = nops - acnt;
        }
        int qi = 0, ti = 0;
        for (int bs = 0; bs < emit; bs += 32) {
            const int col = bs + lane;
            const bool on = col < emit;
            const int op = on ? sm.ops[nops - 1 - col] : kOpGapA;
            const unsigned qm = __ballot_sync(kFull, on && op != kOpGapA);
            const unsigned tm = __ballot_sync(kFull, on && op != kOpGapB);
            const unsigned lt = (1u << lane) - 1u;
            if (on) {
                const char qc = op != kOpGapA ? "ACGT"[sm.A[qi + __popc(qm & lt)]] : '-';
                const char tc = op != kOpGapB ? "ACGT"[sm.B[ti + __popc(tm & lt)]] : '-';
                const int64_t pos = forward ? mid + ncols + col : mid - 1 - (ncols + col);
                g.ws_q[pos] = qc;
                g.ws_t[pos] = tc;
            }
            qi += __popc(qm);
            ti += __popc(tm);
        }
        if (emit > 0) {
            last_op = sm.ops[nops - emit];
            ncols += emit;
            qcons += qi;
            tcons += ti;
        }
        if (stop) break;
        qidx += ae - qcnt;
        tidx += be - tcnt;
    }
    out.ncols = ncols;
    out.qcons = qcons;
    out.tcons = tcons;
    out.last_op = last_op;
    if (lane == 0) g.res[chain] = out;
    ctr.cells += lc.cells;
    ctr.rows += lc.rows;
    ctr.blocks += lc.blocks;
    ctr.interior += lc.interior;
    return true;
}

} // namespace ag2
