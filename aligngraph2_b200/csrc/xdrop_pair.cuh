// xdrop_pair.cuh -- the "pair" path of the X-drop extension: every thread runs TWO extension directions at once, one
// in each 16-bit half of its registers, and all 64 directions of a warp advance through the DP in lock step without a
// single data-dependent branch inside a row.
//
// What is computed is xdrop_align (MC/xdrop_gapalign.cpp:11-213) cell by cell, in the reference's row-major order:
// the running best, the X-drop test against it, the un-decayed horizontal gap across pruned cells (:109-112) and the
// stale vertical gap of a pruned interior cell (:111) are kept as they are.  What is designed is the form:
//
//   * scores are binary16 integers offset by +1024 (exact up to 2048; a block scores <= 720): HADD2 on the FMA pipe,
//     HMNMX2 / HSET2 (compare -> 16-bit mask) / LOP3 on the ALU pipe, two directions per instruction (h2ops.cuh);
//   * one 16-bit word v per column holds the column state: v > 0 live (h = v, e = v - 1, the only case the reference's
//     parameters produce, :58-59, :122/:135, :148-149), v < 0 pruned inside the band (h = MIN, e = |v| - 1 kept), v = 0
//     outside the band or the sentinel column (:161-162).  Everything at or below zero is "MIN": real scores are
//     >= 1024 - 62, and MIN only ever takes part in comparisons it loses;
//   * the band lives in a window of kPairSlots columns [base, base + kPairSlots) in shared memory, base a multiple of
//     8 that follows the band's first column (checked every 4 rows); a row is a loop over groups of 8 slots, every slot
//     the same straight-line code.  Cells left of the band, right of it, the row-end extension (:147-153) and the
//     sentinel are not special cases: a slot is a "standard" cell when it is below nD = last live slot of the previous
//     row + 2, an "extension" cell (no diagonal, no vertical gap, live only if its left neighbour is) above it, and the
//     leading cells that leave the band (:110) are cleared by a running mask;
//   * the reference never lets the band reach column N (:147 `b_size < N`): columns >= N can become live here, but the
//     target codes from position N - 1 on are forced to mismatch, so those cells stay strictly below the best score,
//     feed nothing back (the recurrence only looks left and up) and are clipped from the band end every row;
//   * traceback: 4 bits per cell as on the lane path, one 32-bit word per 8 slots, rows relative to the window base;
//     the rows at which the base moved are logged for the walk.
//
// A direction this path cannot hold (window overflow, a target block shorter than 32, reservation exceeded) is handed
// over with its state at that block: the consumer kernel beside this one continues it there on the row-parallel DP
// (continue_handed_over, xdrop_lane.cuh), the lane kernel does when there are many; behind both stands the wide kernel.
#pragma once

#include "h2ops.cuh"
#include "xdrop_lane.cuh"

namespace ag2 {

constexpr int kPairThreads = 64;                   // threads per CTA (128 directions)
constexpr int kPairSlots = 96;                     // window width in columns (band + lead of CLR reads: 89 at most in 6500 sampled blocks)
constexpr int kPairGroups = kPairSlots / 8;
constexpr int kPairQuads = (kPairGroups + 3) / 4;  // a traceback row = kPairQuads pieces of 16 bytes (4 groups = 32 columns each)
constexpr int kPairQuadStride = 2 * kPairThreads * 16; // the CTA's 128 directions side by side: a warp stores 512 contiguous bytes
constexpr int kPairRowStride = kPairQuads * kPairQuadStride;
constexpr size_t kPairTbCta = (size_t)(kMaxBlk + 2) * kPairRowStride; // traceback of one CTA
constexpr int kPairLogBytes = 256;                 // rows at which the window base moved (u16), <= 92 entries
constexpr int kPairSeqBytes = 192;                 // kSeqWords (46) words, padded
constexpr int kPairAux = kPairLogBytes + 2 * kPairSeqBytes; // per direction
constexpr size_t kPairCtaScratch = kPairTbCta + (size_t)2 * kPairThreads * kPairAux;
#ifndef AG2_PAIR_SHIFT_MASK
#define AG2_PAIR_SHIFT_MASK 3
#endif
constexpr int kPairShiftMask = AG2_PAIR_SHIFT_MASK;   // the window base is checked every (mask + 1) rows
constexpr int kPairOff = 1024;                     // score offset
constexpr int kPairMinN = 32;                      // shortest target block this path takes

struct PairSmem {
    uint32_t v[kPairSlots / 4][kPairThreads][4];   // slot j of both directions: v[j >> 2][tid][j & 3]
    uint32_t tg[kPairGroups][kPairThreads];        // 8 target codes of group g (2 bits each), per half
    uint32_t wgroups[kPairThreads / 32];           // per warp: groups executed (statistics; a counter here costs no register)
};
// 27 KB: eight CTAs (16 warps) per SM.  The forced-mismatch bits of the columns from N - 1 on are not stored: they are
// needed in the last rows of a block only, and computed there (pair_dp).

// Views into the CTA's scratch for one direction (hh = 0 / 1: low / high half of thread tid).  Traceback: piece q of row a
// is the 16 bytes at tb + (a * kPairQuads + q) * kPairQuadStride -- the directions of a warp are adjacent, so the lock-step DP
// writes whole 512-byte runs.
struct PairScratch {
    uint8_t *tb;
    uint16_t *log;
    uint32_t *qcodes;
    uint32_t *tcodes;
};
__device__ __forceinline__ PairScratch pair_scratch(uint8_t *cta, int tid, int hh)
{
    PairScratch s;
    const int col = hh * kPairThreads + tid;
    s.tb = cta + (size_t)col * 16;
    uint8_t *aux = cta + kPairTbCta + (size_t)col * kPairAux;
    s.log = reinterpret_cast<uint16_t *>(aux);
    s.qcodes = reinterpret_cast<uint32_t *>(aux + kPairLogBytes);
    s.tcodes = reinterpret_cast<uint32_t *>(aux + kPairLogBytes + kPairSeqBytes);
    return s;
}

// 8 two-bit target codes at positions p0 .. p0+7 (p0 a multiple of 8), and the forced-mismatch bits for p >= limit
__device__ __forceinline__ uint32_t pair_tg16(const uint32_t *tcodes, int p0)
{
    const int w = p0 >> 4;
    if (w >= kSeqWords) return 0;
    return (tcodes[w] >> (16 * ((p0 >> 3) & 1))) & 0xffffu;
}
__device__ __forceinline__ uint32_t pair_ph16(int p0, int limit)
{
    const int n = min(max(limit - p0, 0), 8);
    return (0x5555u << (2 * n)) & 0xffffu;
}

// Shared-memory loads the compiler must leave where they are written (volatile): the next group's operands are requested
// before the current group's arithmetic, a whole group ahead of their use.
__device__ __forceinline__ void pair_lds4(const uint32_t *p, uint32_t (&d)[4])
{
#ifdef AG2_EMU
    d[0] = p[0]; d[1] = p[1]; d[2] = p[2]; d[3] = p[3];
#else
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3])
                 : "r"((unsigned)__cvta_generic_to_shared(p)));
#endif
}
__device__ __forceinline__ void pair_sts4(uint32_t *p, const uint32_t (&d)[4])
{
#ifdef AG2_EMU
    p[0] = d[0]; p[1] = d[1]; p[2] = d[2]; p[3] = d[3];
#else
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"((unsigned)__cvta_generic_to_shared(p)), "r"(d[0]), "r"(d[1]), "r"(d[2]), "r"(d[3])
                 : "memory");
#endif
}
__device__ __forceinline__ uint32_t pair_lds1(const uint32_t *p)
{
#ifdef AG2_EMU
    return *p;
#else
    uint32_t d;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(d) : "r"((unsigned)__cvta_generic_to_shared(p)));
    return d;
#endif
}

__device__ __forceinline__ void pair_store16(uint8_t *p, const uint32_t (&w)[4])
{
#ifdef AG2_EMU
    memcpy(p, w, 16);
#else
    *reinterpret_cast<uint4 *>(p) = make_uint4(w[0], w[1], w[2], w[3]);
#endif
}

struct PairCarry {
    h2 diag, hgap, best, thr, be, lastj;
    uint32_t Ms, Mlp, Mlead, cnt;
    uint32_t one; // binary16 1.0 in both halves, held in a register
};

// One DP cell of both directions.  K = slot within the group.  END = some running direction of the warp has its band end in
// the group (a slot of the group, or the one after it, is not a standard cell): the boundary masks are needed.  LEAD = some
// running direction is still in its run of leading pruned cells.  Neither: the group is interior.
// Instruction budget (the kernel is bound by the ALU pipe, where HSET2 / HMNMX2 / LOP3 issue, at one warp instruction per
// two cycles; HADD2 / HFMA2 / IMAD issue on the FMA pipe): selects keyed on a compare are done as g * (a - b) + b with the
// compare's 1.0 / 0.0 form wherever the operands stay exact, and the traceback nibble is summed, not assembled.
// Nibble: bit 0 "gap in A beats the rest", bit 1 "gap in B beats the diagonal" (the walk gives bit 0 priority), bit 2 kExtA,
// bit 3 kExtB; slot k of a group sits at bits 16 * (k / 4) + 4 * (3 - k % 4) of the group's word.
template <int K, bool END, bool LEAD>
__device__ __forceinline__ void pair_slot(uint32_t &v, PairCarry &c, const h2 nDg, const uint32_t mmw, uint32_t &acc, h2 &npend)
{
    constexpr bool IN = !END && !LEAD;
    const h2 mt = and_or(mmw << (15 - 2 * K), kH2Sign, c.one); // +1 match, -1 mismatch
    const h2 nd = h2_add(v, mt);                           // h(a-1, b) + match: the next slot's diagonal
    h2 e = h2_add(h2_abs(v), H2C(-1));                     // e(a-1, b)
    h2 dg = c.diag;                                        // h(a-1, b-1) + match
    uint32_t Ms = 0xffffffffu;
    if (END) {
        Ms = c.Ms;                                         // this slot is a standard cell (column < nD)
        const uint32_t Me = h2_gt(nDg, H2C(K + 1));        // the next one is: this slot's vertical gap is valid
        c.Ms = Me;
        e &= Me;
        dg &= Ms;
    }
    const h2 m1 = h2_max(dg, e);
    const h2 sc = h2_max(m1, c.hgap);                      // (:95-107)
    uint32_t Ml = h2_ge(sc, c.thr);                        // not pruned (:109)
    if (END) Ml &= Ms | c.Mlp;                             // extension cells need a live left neighbour
    const h2 gl = Ml & c.one;                              // 1.0 where live
    const h2 gls = END ? (Ml & Ms & c.one) : gl;           // 1.0 where a live standard cell
    // traceback nibble (+1024 to read it off the mantissa)
    const h2 P1 = h2_ltf(dg, e), P2 = h2_ltf(m1, c.hgap);
    const h2 PA = h2_eqf(e, sc), PB = h2_eqf(c.hgap, sc);  // kExtA (:121-126), kExtB (:129-133): unpruned cells only
    const h2 fl = h2_mul(h2_fma(PB, H2C(2), PA), gls);
    if ((K & 1) == 0) {                                    // even slot: its nibble waits, 16-fold, for the odd one's
        npend = h2_fma(fl, H2C(64), h2_fma(P1, H2C(32), h2_mul(P2, H2C(16))));
    } else {
        const h2 nib = h2_fma(fl, H2C(4), h2_fma(P1, H2C(2), h2_add(P2, H2C(1024))));
        acc = acc * 256u + (h2_add(nib, npend) & 0x00ff00ffu);
    }
    // running best (:114-118) and band bookkeeping
    const h2 gnew = h2_gtf(sc, c.best);
    c.best = h2_max(c.best, sc);
    c.thr = h2_add(c.best, H2C(-kXdrop));
    c.be = h2_fma(gnew, h2_sub(H2C(K), c.be), c.be);       // be, lastj: relative to the group's first column
    c.lastj = h2_sel(Ml, H2C(K), c.lastj);
    // horizontal gap: not decayed across pruned cells.  A select (ALU pipe) in interior groups, g * (a - b) + b (FMA pipe)
    // in the others, whose masks already fill the ALU pipe
    if (IN) c.hgap = h2_sel(Ml, h2_add(sc, H2C(-1)), c.hgap);
    else c.hgap = h2_fma(gl, h2_sub(h2_add(sc, H2C(-1)), c.hgap), c.hgap);
    uint32_t dead = v | kH2Sign;                           // pruned inside the band: h = MIN, e kept
    if (LEAD) {
        c.Mlead &= ~Ml;                                    // still in the run of leading pruned cells (:110)
        c.cnt = vadd2(c.cnt, c.Mlead);
        dead &= ~c.Mlead;                                  // left the band
    }
    v = h2_sel(Ml, sc, dead);
    c.diag = nd;
    c.Mlp = Ml;
}

// One group of 8 slots: va = slots 0-3, vb = slots 4-7, computed in place.  Where the stores and loads of the window stand
// matters: a store's source registers cannot be overwritten until the store has left the memory queue, and a register move
// into them right behind it stalls the warp for that long (with the next group's operands prefetched into a second set of
// registers and moved over at the end of the loop that was 5 % of the kernel, twice).  So there is no second set: each half
// is stored as soon as it is computed, and the same registers are then loaded with the NEXT group's half (row `vnext`) --
// memory-queue operations are processed in order -- half a group before anything computes on them again.
template <bool END, bool LEAD>
__device__ __forceinline__ void pair_group(uint32_t (&va)[4], uint32_t (&vb)[4], uint32_t *vrow, const uint32_t *vnext, PairCarry &c,
                                           const h2 nDg, const uint32_t mmw, uint32_t &acc0, uint32_t &acc1)
{
    h2 npend = 0;
    pair_slot<0, END, LEAD>(va[0], c, nDg, mmw, acc0, npend);
    pair_slot<1, END, LEAD>(va[1], c, nDg, mmw, acc0, npend);
    pair_slot<2, END, LEAD>(va[2], c, nDg, mmw, acc0, npend);
    pair_slot<3, END, LEAD>(va[3], c, nDg, mmw, acc0, npend);
    pair_sts4(vrow, va);
    pair_lds4(vnext, va);
    pair_slot<4, END, LEAD>(vb[0], c, nDg, mmw, acc1, npend);
    pair_slot<5, END, LEAD>(vb[1], c, nDg, mmw, acc1, npend);
    pair_slot<6, END, LEAD>(vb[2], c, nDg, mmw, acc1, npend);
    pair_slot<7, END, LEAD>(vb[3], c, nDg, mmw, acc1, npend);
    pair_sts4(vrow + 4 * kPairThreads, vb);
    pair_lds4(vnext + 4 * kPairThreads, vb);
    if (!END) c.Ms = 0xffffffffu;
    c.be = h2_add(c.be, H2C(-8));                          // relative to the next group
    c.lastj = h2_add(c.lastj, H2C(-8));
}

#ifdef AG2_EMU_STATS
struct PairEmuStats { long warp_rows, groups, groups_in, dir_rows, dir_groups_need, shifts, rounds, dir_rounds, walk_cols, gstd_groups, tail_groups; };
static PairEmuStats g_pes;
#define PES(...) __VA_ARGS__
#else
#define PES(...)
#endif

struct PairIO {
    int M[2], N[2];          // block sizes per direction (M = 0: no block)
    // results
    int ae[2], be[2], nshift[2], bail[2];
    unsigned cells[2], rows[2];
};

// xdrop_align forward pass of both directions of the thread; the warp runs max(M) rows.
__device__ void pair_dp(PairSmem &sm, const int tid, const PairScratch &s0, const PairScratch &s1, PairIO &io)
{
    const int M0 = io.M[0], M1 = io.M[1], N0 = io.N[0], N1 = io.N[1];
    // row 0 (:53-67): columns 0..30 hold -j (N >= 32 here); the band ends at column 30, no sentinel yet
    {
        const uint32_t on = (M0 > 0 ? 0xffffu : 0u) | (M1 > 0 ? 0xffff0000u : 0u);
        for (int q = 0; q < kPairSlots / 4; ++q) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int j = 4 * q + i;
                sm.v[q][tid][i] = j <= kXdrop ? h2_from_ints(kPairOff - j, kPairOff - j) & on : 0u;
            }
        }
        for (int g = 0; g < kPairGroups; ++g) {
            sm.tg[g][tid] = pair_tg16(s0.tcodes, 8 * g) | (pair_tg16(s1.tcodes, 8 * g) << 16);
        }
    }
    const int rows_max = __reduce_max_sync(kFull, max(M0, M1));
    const h2 Mh = h2_from_ints(M0, M1);
    int base0 = 0, base1 = 0, nsh0 = 0, nsh1 = 0;
    h2 nrelm1 = h2_from_ints(N0 - 1, N1 - 1);      // N - 1 - base
    h2 nD = h2_from_ints(M0 > 0 ? kXdrop + 1 : 0, M1 > 0 ? kXdrop + 1 : 0);
    h2 frelh = 0, ah = 0, ae = 0;
    uint32_t alive = (M0 > 0 ? 0xffffu : 0u) | (M1 > 0 ? 0xffff0000u : 0u);
    PairCarry c;
    c.best = H2C(kPairOff);
    c.thr = H2C(kPairOff - kXdrop);
    c.be = 0;
    c.one = h2_opaque(H2C(1));
    unsigned cells0 = 0, cells1 = 0, rows2 = 0;
    uint32_t aw0 = 0, aw1 = 0, an0 = s0.qcodes[0], an1 = s1.qcodes[0], bail = 0;
    // target codes just right of the window, held a word ahead: tl = word (base + kPairSlots) >> 4.  A window move takes the
    // new last group from it (a load consumed on the spot stalled the warp for a memory latency at every fourth check)
    static_assert((kPairSlots >> 4) < kSeqWords, "look-ahead word");
    uint32_t tl0 = s0.tcodes[kPairSlots >> 4], tl1 = s1.tcodes[kPairSlots >> 4];

    for (int a = 1; a <= rows_max; ++a) {
        if ((a & kPairShiftMask) == 0) {
            // move the window of a direction whose band start has advanced by 8 columns or more
            const uint32_t Msh = h2_ge(frelh, H2C(8));
            if (__any_sync(kFull, Msh != 0)) {
                PES(if (tid == 0) g_pes.shifts++;)
                for (int q = 0; q < kPairSlots / 4; ++q) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const uint32_t up = q + 2 < kPairSlots / 4 ? sm.v[q + 2][tid][i] : 0u;
                        sm.v[q][tid][i] = h2_sel(Msh, up, sm.v[q][tid][i]);
                    }
                }
                if (Msh & 0xffffu) {
                    base0 += 8;
                    s0.log[nsh0++] = (uint16_t)a;
                }
                if (Msh >> 16) {
                    base1 += 8;
                    s1.log[nsh1++] = (uint16_t)a;
                }
                for (int g = 0; g < kPairGroups; ++g) {
                    uint32_t tn;
                    if (g + 1 < kPairGroups) tn = sm.tg[g + 1][tid];
                    else   // positions base + 8 g ..: the half of the look-ahead word they are in (only a moved direction's half is kept)
                        tn = ((tl0 >> (16 * (((base0 + 8 * g) >> 3) & 1))) & 0xffffu) | ((tl1 >> (16 * (((base1 + 8 * g) >> 3) & 1))) << 16);
                    sm.tg[g][tid] = h2_sel(Msh, tn, sm.tg[g][tid]);
                }
                if ((Msh & 0xffffu) && ((base0 + kPairSlots) & 15) == 0) {
                    const int w = (base0 + kPairSlots) >> 4;
                    tl0 = w < kSeqWords ? s0.tcodes[w] : 0u;
                }
                if ((Msh >> 16) && ((base1 + kPairSlots) & 15) == 0) {
                    const int w = (base1 + kPairSlots) >> 4;
                    tl1 = w < kSeqWords ? s1.tcodes[w] : 0u;
                }
                const h2 m8 = Msh & H2C(-8);
                frelh = h2_add(frelh, m8);
                nD = h2_add(nD, m8);
                nrelm1 = h2_add(nrelm1, m8);
                c.be = h2_add(c.be, m8);
            }
        }
        // row start
        if (((a - 1) & 15) == 0) { // next 16 query codes: fetched 16 rows ago, the ones after them are requested now
            const int w = ((a - 1) >> 4) + 1;
            aw0 = an0;
            aw1 = an1;
            an0 = w < kSeqWords ? s0.qcodes[w] : 0u;
            an1 = w < kSeqWords ? s1.qcodes[w] : 0u;
        }
        const uint32_t acrep = ((aw0 & 3u) * 0x5555u) | ((aw1 & 3u) * 0x55550000u);
        aw0 >>= 2;
        aw1 >>= 2;
        ah = h2_add(ah, H2C(1));
        const uint32_t run = h2_le(ah, Mh) & alive;
        if (!__any_sync(kFull, run != 0)) break;
        {   // counters: cells of this row = band end - band start (:84), rows
            const h2 wdt = h2_add(h2_min(nD, h2_add(nrelm1, H2C(1))), frelh ^ kH2Sign) & run;
            const uint32_t wi = h2_add(wdt, H2C(1024)) & 0x03ff03ffu;
            cells0 += wi & 0xffffu;
            cells1 += wi >> 16;
            rows2 = vadd2(rows2, run & 0x00010001u);
        }
        const h2 nDe = nD & run;
        PES(if (tid == 0) g_pes.warp_rows++;
            g_pes.dir_rows += ((run & 0xffffu) != 0) + ((run >> 16) != 0);)
        c.best = h2_sel(run, c.best, H2C(2047)); // a direction that has stopped can no longer prune in, whatever its masks say
        c.thr = h2_sel(run, c.thr, H2C(2047 - kXdrop));
        const h2 best0 = c.best;
        c.diag = 0;
        c.hgap = 0;
        c.Mlead = 0xffffffffu;
        c.Mlp = 0;
        c.cnt = 0;
        c.lastj = H2C(-1);
        c.Ms = h2_gt(nDe, 0u);
        h2 Jg = 0, nDg = nDe;
        uint8_t *row0 = s0.tb + (size_t)a * kPairRowStride, *row1 = s1.tb + (size_t)a * kPairRowStride;
        uint32_t t0[4] = {0, 0, 0, 0}, t1[4] = {0, 0, 0, 0}; // traceback words of the current piece (4 groups) per direction
        // groups that hold standard cells of some direction of the warp run without asking; after them one group at a
        // time while a direction still has a live cell to extend from (:147-153)
        int g_std;
        {
            const int n0 = h2_lo_int(nDe), n1 = h2_hi_int(nDe);
            g_std = min(__reduce_max_sync(kFull, (max(n0, n1) + 7) >> 3), kPairGroups);
        }
        int g = 0;
        bool more = g_std > 0;
        // operands of the next group are fetched from shared memory while the current one computes
        uint32_t va[4], vb[4], tgw;
        // forced mismatches from column N - 1 on: only when that column is inside the window of some direction of the warp
        const bool lim_in = __any_sync(kFull, (h2_lt(nrelm1, H2C(kPairSlots)) & run) != 0);
        int lim0 = 0, lim1 = 0;
        if (lim_in) {
            lim0 = N0 - 1 - base0;
            lim1 = N1 - 1 - base1;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            va[i] = sm.v[0][tid][i];
            vb[i] = sm.v[1][tid][i];
        }
        tgw = sm.tg[0][tid];
        while (more) {
            const int gn = min(g + 1, kPairGroups - 1);
            const uint32_t ntg = pair_lds1(&sm.tg[gn][tid]);
            const uint32_t x = tgw ^ acrep;
            uint32_t mmw = x | (x >> 1); // bit 2k of each half: slot k mismatches
            if (lim_in) mmw |= pair_ph16(8 * g, lim0) | (pair_ph16(8 * g, lim1) << 16);
            uint32_t acc0 = 0, acc1 = 0;
            // which masks the group needs: has a running direction of the warp its band end in these 8 slots or the one after
            // them; is one still in its leading run.  (Two votes and two branches: a four-way switch on one reduced value
            // became an indexed jump through a table, 6 % of the kernel waiting for the table load.)
            const bool noend = __all_sync(kFull, (h2_ge(nDg, H2C(9)) | ~run) == 0xffffffffu);
            const bool nolead = __all_sync(kFull, (~c.Mlead | ~run) == 0xffffffffu);
            PES(if (tid == 0) { g_pes.groups++; g_pes.groups_in += noend && nolead; if (g < g_std) g_pes.gstd_groups++; else g_pes.tail_groups++; })
            uint32_t *vrow = &sm.v[2 * g][tid][0];
            const uint32_t *vnext = &sm.v[2 * gn][tid][0];
            if (nolead) {
                if (noend) pair_group<false, false>(va, vb, vrow, vnext, c, nDg, mmw, acc0, acc1);
                else pair_group<true, false>(va, vb, vrow, vnext, c, nDg, mmw, acc0, acc1);
            } else {
                pair_group<true, true>(va, vb, vrow, vnext, c, nDg, mmw, acc0, acc1);
            }
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                t0[i] = t0[i + 1];
                t1[i] = t1[i + 1];
            }
            t0[3] = (acc0 & 0xffffu) | (acc1 << 16);
            t1[3] = (acc0 >> 16) | (acc1 & 0xffff0000u);
            ++g;
            if ((g & 3) == 0) { // a piece is complete: one 16-byte store per direction, coalesced over the warp
                pair_store16(row0 + (size_t)((g >> 2) - 1) * kPairQuadStride, t0);
                pair_store16(row1 + (size_t)((g >> 2) - 1) * kPairQuadStride, t1);
            }
            Jg = h2_add(Jg, H2C(8));
            nDg = h2_add(nDg, H2C(-8));
            tgw = ntg;
            if (g < g_std) continue;
            more = __any_sync(kFull, ((c.Ms | c.Mlp) & run) != 0);
            if (g >= kPairGroups) break;
        }
        if (g & 3) { // the row's last, partial piece
            for (int r = g & 3; r < 4; ++r) {
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    t0[i] = t0[i + 1];
                    t1[i] = t1[i + 1];
                }
            }
            pair_store16(row0 + (size_t)(g >> 2) * kPairQuadStride, t0);
            pair_store16(row1 + (size_t)(g >> 2) * kPairQuadStride, t1);
        }
        if (more) bail |= (c.Ms | c.Mlp) & run; // the window is too narrow for this direction
        if ((tid & 31) == 0) sm.wgroups[tid >> 5] += (unsigned)g;
        c.be = h2_add(c.be, Jg);                // back to window columns
        c.lastj = h2_add(c.lastj, Jg);
        PES({ const int l0 = h2_lo_int(c.lastj), l1 = h2_hi_int(c.lastj);
              if (run & 0xffffu) g_pes.dir_groups_need += (max(l0, 0) + 2 + 7) >> 3;
              if (run >> 16) g_pes.dir_groups_need += (max(l1, 0) + 2 + 7) >> 3; })
        // row end (:142-164): the next band ends one past the last live cell (the sentinel), clipped to column N - 1
        const uint32_t any_live = h2_ge(c.lastj, 0u);
        alive &= any_live | ~run;      // every cell pruned: the reference leaves the loop (:142)
        nD = h2_sel(run, h2_add(h2_min(c.lastj, nrelm1), H2C(2)), nD);
        frelh = h2_sel(run, h2_add(vadd2(vadd2(~c.cnt, 0x00010001u), H2C(1024)), H2C(-1024)), frelh); // -cnt leading cells left the band
        ae = h2_sel(h2_ne(c.best, best0), ah, ae);
    }
    io.ae[0] = h2_lo_int(ae);
    io.ae[1] = h2_hi_int(ae);
    io.be[0] = base0 + h2_lo_int(c.be);
    io.be[1] = base1 + h2_hi_int(c.be);
    io.nshift[0] = nsh0;
    io.nshift[1] = nsh1;
    io.bail[0] = (bail & 0xffffu) != 0;
    io.bail[1] = (bail >> 16) != 0;
    io.cells[0] = cells0;
    io.cells[1] = cells1;
    io.rows[0] = rows2 & 0xffffu;
    io.rows[1] = rows2 >> 16;
}

// 16-byte global -> shared copy that does not pass through registers (LDGSTS); per-thread completion
__device__ __forceinline__ void pair_copy16_async(uint32_t *smem_dst, const uint8_t *gmem_src)
{
#ifdef AG2_EMU
    memcpy(smem_dst, gmem_src, 16);
#else
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
#endif
}
__device__ __forceinline__ void pair_copy_commit()
{
#ifndef AG2_EMU
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
// waits until all but the thread's N most recent groups of copies have landed
template <int N>
__device__ __forceinline__ void pair_copy_wait_but()
{
#ifndef AG2_EMU
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
#endif
}
__device__ __forceinline__ void pair_copy_wait()
{
#ifndef AG2_EMU
    asm volatile("cp.async.wait_all;" ::: "memory");
#endif
}

constexpr int kPairWalkRows = (kPairSlots / 4) / kPairQuads;             // traceback rows in flight: what the band window holds
constexpr int kPairMoveStep = kPairShiftMask + 1;                        // the window base can move at multiples of this row
static_assert(kPairGroups == 4 * kPairQuads && kPairWalkRows >= 4 && kPairWalkRows <= 16, "walk ring");
static_assert((kPairMoveStep & (kPairMoveStep - 1)) == 0, "move rows are tested with a mask");
static_assert(kPairGroups >= 12 && (kMaxBlk + 2) / kPairMoveStep < 8 * 32, "sm.tg holds the walk's code words and move bits");
static_assert(kPairLogBytes <= kPairWalkRows * kPairQuads * 16 + 64, "the move log is staged in the row ring");

// 4-byte global -> shared copy (LDGSTS)
__device__ __forceinline__ void pair_copy4_async(uint32_t *smem_dst, const uint32_t *gmem_src)
{
#ifdef AG2_EMU
    *smem_dst = *gmem_src;
#else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
#endif
}

// Traceback (:170-210) over the pair layout: row a holds the nibbles of columns base(a) .. base(a) + kPairSlots - 1, base(a) =
// 8 x (number of logged window moves at rows <= a).  Otherwise lane_walk.
//
// Everything the walk reads lives in HBM (the scratch of all resident directions is far larger than L2), a step depends on
// the previous one, and the 32 lanes of the warp are at different places of their blocks.  Two things follow.
// (1) A load into a register anywhere in the step -- however far ahead of its use for the lane that issues it -- stalls the
// WARP at the next step, when another lane touches that register (the scoreboard is per warp): the first form of this walk,
// which read the code words and the move log one word ahead, ran at one memory latency per step.  So the step loads nothing
// into registers from global memory: all of it arrives in shared memory through asynchronous copies, whose completion is
// counted per thread.
//   * rows: the thread's band window, idle during the walk, is a ring of kPairWalkRows whole rows; when the walk leaves a
//     row, the row kPairWalkRows below it is requested into the slot that has become free;
//   * code words of the block (16 positions each): two-word rings in sm.tg[0..3]; entering a word requests the one below;
//   * window moves: a bit per kPairMoveStep rows in sm.tg[4..11], built once from the log.
// Every step commits one group of copies (possibly empty) and starts by waiting for all but the kPairWalkRows - 2 youngest:
// the row the walk stands on was requested at least that many steps ago, a code word at least 16.
// (2) A branch on the lane's own state (kind of step, word boundary, trimming) makes the warp run every side of it at nearly
// every step.  The step is straight-line code: selects, and side effects under the lane's `walking`.
// Called by the whole warp (`active` = this lane has a direction to walk).
__device__ int pair_walk(PairSmem &sm, int tid, uint8_t *cta_scratch, int hh, bool active, int nshift, int ae, int be, char *wq, char *wt,
                         int cap, int &qcnt, int &tcnt, int &acnt, bool &trim_ok, int &first_op, int &op_after_trim)
{
    const PairScratch ps = pair_scratch(cta_scratch, tid, hh); // derived from register-held values, nothing read from local memory
    int a = ae, b = be, n = 0, cur = kOpSub;
    int m = 0, q = 0, t = 0, ac = 0;
    bool scanning = true;
    first_op = -1;
    op_after_trim = -1;
    bool walking = active && (a > 0 || b > 0) && n < cap;
    // ---- the window moves as a bit set; ns = moves at rows <= a
    int ns = 0;
    {
        const int nchunk = walking ? (nshift + 7) >> 3 : 0;
        for (int c = 0; c < nchunk; ++c) pair_copy16_async(&sm.v[c][tid][0], reinterpret_cast<const uint8_t *>(ps.log) + 16 * c);
        for (int w = 4; w < 12; ++w) sm.tg[w][tid] = 0;
        pair_copy_wait();
        if (walking) {
            for (int i = 0; i < nshift; ++i) {
                const uint32_t pr = sm.v[i >> 3][tid][(i >> 1) & 3];
                const int row = (int)((pr >> (16 * (i & 1))) & 0xffffu);
                const int bit = row / kPairMoveStep;
                sm.tg[4 + (bit >> 5)][tid] |= 1u << (bit & 31);
                ns += row <= a;
            }
        }
    }
    // ---- code words: position p of the block is in word p >> 4; ring slot = word & 1
    int qi = min(max(a - 1, 0) >> 4, kSeqWords - 1), ti = min(max(b - 1, 0) >> 4, kSeqWords - 1);
    if (walking) {
        pair_copy4_async(&sm.tg[qi & 1][tid], ps.qcodes + qi);
        if (qi > 0) pair_copy4_async(&sm.tg[(qi - 1) & 1][tid], ps.qcodes + qi - 1);
        pair_copy4_async(&sm.tg[2 + (ti & 1)][tid], ps.tcodes + ti);
        if (ti > 0) pair_copy4_async(&sm.tg[2 + ((ti - 1) & 1)][tid], ps.tcodes + ti - 1);
    }
    // ---- the row ring: row r in slot r % kPairWalkRows
    int slot_a = a % kPairWalkRows;                 // slot of row a
    int fetch = a;                                  // next row to request
    const uint8_t *fsrc = ps.tb + (size_t)fetch * kPairRowStride;
    {
        int fs = slot_a;
        for (int i = 0; i < kPairWalkRows; ++i) {
            if (walking && fetch >= 1) {
#pragma unroll
                for (int qd = 0; qd < kPairQuads; ++qd) pair_copy16_async(&sm.v[fs * kPairQuads + qd][tid][0], fsrc + (size_t)qd * kPairQuadStride);
            }
            pair_copy_commit();
            --fetch;
            fsrc -= kPairRowStride;
            fs = fs ? fs - 1 : kPairWalkRows - 1;
        }
    }
    // The refill of a slot is issued one step late, behind the next step's read of its cell: the copies sit in the same
    // memory queue as that read and would hold it up, and their address registers would be rewritten right behind them.
    bool refill = false;
    int refill_slot = 0;
    const uint8_t *refill_src = fsrc;
    while (__any_sync(kFull, walking)) {
        pair_copy_wait_but<kPairWalkRows - 2>();
        // the cell: row a (ring slot slot_a), window slot b - 8 ns
        const int slot = b - 8 * ns;
        const int f = slot_a * kPairGroups + min(max(slot >> 3, 0), kPairGroups - 1);
        const uint32_t word = sm.v[f >> 2][tid][f & 3];
        if (refill) {
#pragma unroll
            for (int qd = 0; qd < kPairQuads; ++qd)
                pair_copy16_async(&sm.v[refill_slot * kPairQuads + qd][tid][0], refill_src + (size_t)qd * kPairQuadStride);
        }
        int cell = (int)((word >> (16 * ((slot >> 2) & 1) + 4 * (3 - (slot & 3)))) & 15u);
        cell = a > 0 ? cell : kOpGapA;              // row 0 is all SCRIPT_GAP_IN_A (:61)
        int nxt = (cell & 1) ? kOpGapA : (cell & 2); // kOpGapB == 2, kOpSub == 0
        nxt = (cur == kOpGapA && (cell & kExtA)) ? kOpGapA : nxt;
        nxt = (cur == kOpGapB && (cell & kExtB)) ? kOpGapB : nxt;
        cur = walking ? nxt : cur;
        const bool da = walking && cur != kOpGapA, db = walking && cur != kOpGapB;
        // a window move logged at the row being left no longer counts
        const int mbit = a / kPairMoveStep;
        const uint32_t mword = sm.tg[4 + ((mbit >> 5) & 7)][tid];
        ns -= (da && (a & (kPairMoveStep - 1)) == 0) ? (int)((mword >> (mbit & 31)) & 1u) : 0;
        // the slot of the row being left takes the row kPairWalkRows below it (at the next step)
        refill = da && fetch >= 1;
        refill_slot = slot_a;
        refill_src = fsrc;
        a -= da;
        fetch -= da;
        fsrc -= da ? kPairRowStride : 0;
        slot_a = da ? (slot_a ? slot_a - 1 : kPairWalkRows - 1) : slot_a;
        b -= db;
        // codes of the step; a position in the word below: that word is in the ring, the one below it is requested
        const int qn = da ? a >> 4 : qi, tn = db ? b >> 4 : ti;
        if (qn != qi && qn > 0) pair_copy4_async(&sm.tg[(qn - 1) & 1][tid], ps.qcodes + qn - 1);
        if (tn != ti && tn > 0) pair_copy4_async(&sm.tg[2 + ((tn - 1) & 1)][tid], ps.tcodes + tn - 1);
        qi = qn;
        ti = tn;
        pair_copy_commit();
        const uint32_t qwv = sm.tg[qi & 1][tid], twv = sm.tg[2 + (ti & 1)][tid];
        const int qc = da ? (int)((qwv >> (2 * (a & 15))) & 3u) : 4;
        const int tc = db ? (int)((twv >> (2 * (b & 15))) & 3u) : 4;
        first_op = (walking && n == 0) ? cur : first_op;
        op_after_trim = (walking && !scanning && op_after_trim < 0) ? cur : op_after_trim;
        {   // trim_mismatch_end scans from the END of the block's alignment = walk start
            const bool sc = walking && scanning;
            ac += sc;
            q += sc && da;
            t += sc && db;
            m = sc ? ((qc == tc) ? m + 1 : 0) : m;
            scanning = scanning && !(sc && m == kTailMatch);
        }
        if (walking) {
            wq[n] = code_char(qc);
            wt[n] = code_char(tc);
        }
        n += walking;
        walking = walking && (a > 0 || b > 0) && n < cap;
    }
    pair_copy_wait(); // nothing may land in the window once the next block's DP owns it
    qcnt = q;
    tcnt = t;
    acnt = ac;
    trim_ok = !scanning && (n - 1 - ac) > 0;
    return n;
}

struct PairBlk {
    int qblk, tblk;
    bool last_block;
};

// fetch16 / fetch16_bits (xdrop_lane.cuh) in two halves, without a branch: `issue` computes the address and starts the loads,
// `finish` shifts, reverses and clips.  Staging a block issues four words' loads before it finishes the first (with the
// branchy form every load was consumed on the spot: one memory latency per word, 7 % of the kernel's warp time).
struct PairFetch {
    uint32_t lo, hi;
    int sh, drop;      // drop: codes / bits to drop at the front (backward), or to shift in (bits, forward, p0 < 0)
    bool rev, valid;
};
__device__ __forceinline__ uint32_t pair_ld_ro(const uint32_t *p)
{
#ifdef AG2_EMU
    return *p;
#else
    return __ldg(p);
#endif
}
__device__ __forceinline__ void pair_fetch16_issue(PairFetch &f, const uint32_t *seq, int64_t off, int64_t len, int64_t p0, int dir, bool want)
{
    const bool fwd = dir > 0;
    const int64_t s0 = fwd ? p0 : p0 - 15;
    f.rev = !fwd;
    f.drop = (!fwd && s0 < 0) ? (int)(-s0) : 0;
    f.valid = want && (fwd ? (p0 < len && p0 >= 0) : (p0 >= 0));
    const int64_t pos = f.valid ? off + (s0 < 0 ? 0 : s0) : off;
    f.sh = 2 * (int)(pos & 15);
    const uint32_t *p = seq + (pos >> 4);
    f.lo = f.valid ? pair_ld_ro(p) : 0u;
    f.hi = (f.valid && f.sh) ? pair_ld_ro(p + 1) : 0u;
}
__device__ __forceinline__ uint32_t pair_fetch16_finish(const PairFetch &f)
{
    const uint32_t w = f.sh ? (f.lo >> f.sh) | (f.hi << (32 - f.sh)) : f.lo;
    const uint32_t r = rev16(w) >> (2 * f.drop);
    return f.valid ? (f.rev ? r : w) : 0u;
}
__device__ __forceinline__ void pair_fetch16_bits_issue(PairFetch &f, const uint32_t *bits, int64_t off, int64_t p0, int dir, bool want)
{
    const bool fwd = dir > 0;
    const int64_t s0 = fwd ? p0 : p0 - 15;
    f.rev = !fwd;
    f.drop = s0 < 0 ? (int)(-s0) : 0;
    f.valid = want && (fwd ? p0 > -16 : p0 >= 0);
    const int64_t pos = f.valid ? off + (s0 < 0 ? 0 : s0) : off;
    f.sh = (int)(pos & 31);
    const uint32_t *p = bits + (pos >> 5);
    f.lo = f.valid ? pair_ld_ro(p) : 0u;
    f.hi = (f.valid && f.sh > 16) ? pair_ld_ro(p + 1) : 0u;
}
__device__ __forceinline__ uint32_t pair_fetch16_bits_finish(const PairFetch &f)
{
    uint32_t x = f.lo >> f.sh;
    if (f.sh > 16) x |= f.hi << (32 - f.sh);
    x &= 0xffffu;                                          // bit j = position s + j
    uint32_t r = x;                                        // reverse the 16 bits
    r = ((r >> 1) & 0x5555u) | ((r & 0x5555u) << 1);
    r = ((r >> 2) & 0x3333u) | ((r & 0x3333u) << 2);
    r = ((r >> 4) & 0x0f0fu) | ((r & 0x0f0fu) << 4);
    r = ((r >> 8) & 0x00ffu) | ((r & 0x00ffu) << 8);
    const uint32_t fw = (x << f.drop) & 0xffffu, bw = r >> f.drop;
    return f.valid ? (f.rev ? bw : fw) : 0u;
}

// retrieve_next_aln_block (MC/gapalign.cpp:9-45) + staging of the block's codes in extension order
__device__ void pair_prepare(const LaneArgs &g, const LaneChain &s, const PairScratch &ps, PairBlk &blk)
{
    const int qleft = s.qsize - s.qidx, tleft = s.tsize - s.tidx;
    if (qleft < kBlk + kBlkSlack || tleft < kBlk + kBlkSlack) {
        blk.qblk = min(qleft, stretch_0p2(tleft));
        blk.tblk = min(tleft, stretch_0p2(qleft));
        blk.last_block = true;
    } else {
        blk.qblk = kBlk;
        blk.tblk = kBlk;
        blk.last_block = false;
    }
    const int qw = (blk.qblk + 15) >> 4, tw = (blk.tblk + 16) >> 4;
    const int qdir = s.c.strand == 0 ? s.inc : -s.inc;
    const bool rc = s.c.strand != 0;
    // four words at a time: their loads are in flight together, and leave as one 16-byte store (the arrays are padded to 48 words)
    static_assert(kPairSeqBytes >= 4 * ((kSeqWords + 3) / 4 * 4), "code arrays hold whole groups of four words");
    for (int w0 = 0; w0 < qw; w0 += 4) {
        PairFetch fc[4], fb[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int w = w0 + i;
            const int p = s.q0 + s.inc * (s.qidx + 16 * w);
            const int64_t fp = rc ? (int64_t)s.rlen - 1 - p : p;
            pair_fetch16_issue(fc[i], g.seqs.reads2, s.roff, s.rlen, fp, qdir, w < qw);
            pair_fetch16_bits_issue(fb[i], g.seqs.reads_irr, s.roff, fp, qdir, rc && w < qw);
        }
        uint32_t v4[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint32_t v = pair_fetch16_finish(fc[i]);
            if (rc) v ^= ~spread_bits16(pair_fetch16_bits_finish(fb[i]));
            v4[i] = w0 + i < qw ? v : 0u;
        }
        pair_store16(reinterpret_cast<uint8_t *>(ps.qcodes + w0), v4);
    }
    const int twc = min(tw, kSeqWords);
    for (int w0 = 0; w0 < twc; w0 += 4) {
        PairFetch fc[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            pair_fetch16_issue(fc[i], g.seqs.ref2, 0, g.seqs.ref_len, s.t0 + (int64_t)s.inc * (s.tidx + 16 * (w0 + i)), s.inc, w0 + i < twc);
        }
        uint32_t v4[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) v4[i] = pair_fetch16_finish(fc[i]);
        pair_store16(reinterpret_cast<uint8_t *>(ps.tcodes + w0), v4);
    }
}

__device__ __forceinline__ unsigned pair_vol_u32(const unsigned int *p) { return *reinterpret_cast<const volatile unsigned int *>(p); }
__device__ __forceinline__ void pair_pause()
{
#ifndef AG2_EMU
    __nanosleep(500);
#endif
}

// Body of xdrop_pair_kernel.  Every thread holds two directions; the warp advances all of them by one block per round
// (stage -> pair_dp in lock step -> per direction: walk and align_ex's bookkeeping), refilling finished ones from the queue.
// g.wide_queue / g.wide_count receive the directions handed to the lane kernel.
//
// A round lasts as long as its longest block.  Blocks are 500 rows except the LAST block of a direction, which has up to
// 599 (MC/gapalign.cpp:24-31), and with 64 directions per warp nearly every round held one: ~10 % of the row slots ran
// empty.  So a direction that reaches such a block (g.defer_queue set) is saved and put on a second queue, and the warps
// take that queue when their main work is done: rounds of long last blocks only, one block per direction -- which also
// makes the end of the launch fine-grained.  A warp takes deferred work only once none of its lanes can still produce
// some (no waiting inside a warp on its own lanes), and never blocks on a ticket: what is not published yet is looked at
// again in the next round.
__device__ void pair_kernel_body(const LaneArgs &g, PairSmem &sm, int tid, uint8_t *scratch)
{
    LaneChain s[2];
    s[0].chain = s[1].chain = -1;
    const PairScratch ps[2] = {pair_scratch(scratch, tid, 0), pair_scratch(scratch, tid, 1)}; // scratch = the CTA's
    unsigned long long cells = 0, rows = 0, blocks = 0, handed = 0;
    if ((tid & 31) == 0) sm.wgroups[tid >> 5] = 0;
    bool drained = false, warp_main_done = false;
    unsigned from_defer = 0;            // bit h: slot h runs a deferred direction
    unsigned done2 = g.defer_queue ? 0u : 3u;   // bit h: slot h will get nothing more from the deferred queue
    long long ticket[2] = {-1, -1};     // ticket of the deferred queue not yet served
    if (g.defer_queue && (tid & 31) == 0) atomicAdd(g.defer_ctl + 2, 1u);   // one more warp in its main phase
    for (;;) {
        if (!warp_main_done && g.defer_queue) {
            const bool mine = drained && (s[0].chain < 0 || (from_defer & 1u)) && (s[1].chain < 0 || (from_defer & 2u));
            if (__all_sync(kFull, mine)) {   // this warp will defer nothing more
                warp_main_done = true;
                if ((tid & 31) == 0) {
                    __threadfence();        // its deferrals are published before it is counted out
                    atomicSub(g.defer_ctl + 2, 1u);
                }
            }
        }
        PairIO io;
        PairBlk blk[2];
#pragma unroll 1
        for (int h = 0; h < 2; ++h) { // not unrolled: the chain state is indexed, so it lives in local memory, not in registers
            io.M[h] = io.N[h] = 0;
            io.ae[h] = io.be[h] = io.nshift[h] = io.bail[h] = 0;
            io.cells[h] = io.rows[h] = 0;
            for (int attempt = 0; attempt < 3; ++attempt) {
                if (s[h].chain < 0 && !drained) {
                    const unsigned long long t = atomicAdd(g.next, 1ull);
                    if ((int64_t)t < g.n_chains) {
                        lane_start_chain(g, g.queue ? (int64_t)g.queue[t] : (int64_t)t, s[h]);
                        from_defer &= ~(1u << h);
                        if (s[h].ge.valid && !wait_for_read(g.sig, s[h].c.read)) s[h].ge.valid = 0; // streamed run: its read is still on the way
                        if (!s[h].ge.valid) {
                            const ChainResult out = {0, 0, 0, -1, 0, 0, 0, 0};
                            g.res[s[h].chain] = out;
                            signal_direction_done(g.sig, s[h].chain);
                            s[h].chain = -1;
                        }
                    } else {
                        drained = true;
                    }
                } else if (s[h].chain < 0 && warp_main_done && !(done2 & (1u << h))) {
                    // A ticket of the deferred queue (fetch-and-add: compare-and-swap claims serialised the 130 000 resident
                    // threads on one word -- 385 -> 633 ms).  The ticket is kept until its entry is published or until it is
                    // known to lie behind the last entry: no warp of the launch in its main phase (the main queue is empty
                    // by then -- this warp drained it -- so a CTA that starts later cannot defer anything).
                    if (ticket[h] < 0) ticket[h] = (long long)atomicAdd(g.defer_ctl + 1, 1u);
                    if ((unsigned long long)ticket[h] < pair_vol_u32(g.defer_ctl)) {
                        const int32_t ch = *reinterpret_cast<const volatile int32_t *>(g.defer_queue + ticket[h]);
                        if (ch >= 0) {      // published
                            __threadfence();
                            lane_start_chain(g, ch, s[h]);
                            lane_resume(s[h], g.defer_resume[ticket[h]]);
                            from_defer |= 1u << h;
                            ticket[h] = -1;
                        }
                    } else if (pair_vol_u32(g.defer_ctl + 2) == 0) {
                        __threadfence();
                        if ((unsigned long long)ticket[h] >= pair_vol_u32(g.defer_ctl)) done2 |= 1u << h;
                    }
                }
                if (s[h].chain < 0) break;
                pair_prepare(g, s[h], ps[h], blk[h]);
                if (blk[h].qblk > 0 && blk[h].tblk > 0) {
                    if (blk[h].tblk < kPairMinN) { // the reference's row 0 may reach column N here: lane kernel
                        const unsigned slot = atomicAdd(g.wide_count, 1u);
                        if (g.resume) g.resume[slot] = lane_save(s[h]);
                        publish_chain(g.wide_queue, slot, s[h].chain);
                        ++handed;
                        s[h].chain = -1;
                    } else if (g.defer_queue && !(from_defer & (1u << h)) && blk[h].last_block && blk[h].qblk > kBlk) {
                        const unsigned slot = atomicAdd(g.defer_ctl, 1u);
                        g.defer_resume[slot] = lane_save(s[h]);
                        publish_chain(g.defer_queue, slot, s[h].chain);
                        s[h].chain = -1;
                        continue;           // the slot is free again: take another direction for this round
                    } else {
                        io.M[h] = blk[h].qblk;
                        io.N[h] = blk[h].tblk;
                    }
                }
                break;
            }
        }
        if (!__any_sync(kFull, s[0].chain >= 0 || s[1].chain >= 0 || !drained || done2 != 3u)) break;
        PES(if (tid == 0) g_pes.rounds++;
            g_pes.dir_rounds += (io.M[0] > 0) + (io.M[1] > 0);)
        if (__any_sync(kFull, io.M[0] > 0 || io.M[1] > 0)) pair_dp(sm, tid, ps[0], ps[1], io);
        else if (!__any_sync(kFull, s[0].chain >= 0 || s[1].chain >= 0)) pair_pause();   // waiting for deferred work of other warps
#pragma unroll 1
        for (int h = 0; h < 2; ++h) { // not unrolled: the chain state is indexed, so it lives in local memory, not in registers
            LaneChain &c = s[h];
            const bool have = c.chain >= 0;
            const int ae = io.ae[h], be = io.be[h];
            const int cap = have ? (int)min((int64_t)(2 * kMaxBlk), c.seg_end - c.seg) : 0;
            const bool over = have && (io.bail[h] || ae + be > cap || c.meta + c.nblocks >= c.meta_end);
#ifdef AG2_EXP_NOWALK
            const bool walk = false;
#else
            const bool walk = have && !over;
#endif
            int qcnt, tcnt, acnt, first_op, op_after;
            bool trim_ok;
            const int64_t seg = walk ? c.seg : 0;
            const int nops = pair_walk(sm, tid, scratch, h, walk, io.nshift[h], ae, be, g.ws_q + seg, g.ws_t + seg, cap, qcnt, tcnt, acnt,
                                       trim_ok, first_op, op_after);
            if (!have) continue;
            int rc = 2;
#ifdef AG2_EXP_NOWALK
            if (have && !over) {
#else
            if (walk) {
#endif
                if (io.M[h] > 0) c.blocks += 1;
                c.cells += io.cells[h];
                c.rows += io.rows[h];
                rc = lane_block_tail(g, c, blk[h].qblk, blk[h].tblk, blk[h].last_block, ae, be, nops, qcnt, tcnt, acnt, trim_ok,
                                     first_op, op_after);
            }
            if (rc == 1) {
                const ChainResult out = {c.ncols, c.qcons, c.tcons, c.last_op, c.nblocks, 0, 0, 0};
                g.res[c.chain] = out;
                signal_direction_done(g.sig, c.chain);
                cells += c.cells;
                rows += c.rows;
                blocks += c.blocks;
                c.chain = -1;
            } else if (rc == 2) { // c is still the state at the start of this block
                const unsigned slot = atomicAdd(g.wide_count, 1u);
                if (g.resume) g.resume[slot] = lane_save(c);
                publish_chain(g.wide_queue, slot, c.chain);
                ++handed;
                c.chain = -1;
            }
        }
    }
    atomicAdd(&g.counters->cells, cells);
    atomicAdd(&g.counters->rows, rows);
    atomicAdd(&g.counters->blocks, blocks);
    // every group of a warp evaluates 8 slots of its 64 directions, running or not
    if ((tid & 31) == 0) atomicAdd(&g.counters->slots, (unsigned long long)sm.wgroups[tid >> 5] * (8ull * 64ull));
    (void)handed;
}

} // namespace ag2
