// xdrop_device.cuh -- device code of the X-drop extension.
//
// One warp owns one extension direction ("chain") of one candidate and runs the reference's
// align_ex loop (MC/xdrop_gapalign.cpp:263-357) to completion:
//
//   stage block  : 2-bit packed read / reference -> one code per byte in shared memory, in
//                  extension order (so the DP never sees direction or strand)
//   dp_block<K>  : xdrop_align's forward pass (MC/xdrop_gapalign.cpp:53-165), one DP ROW per step,
//                  the row spread over the 32 lanes (K adjacent columns per lane, columns mapped
//                  circularly: column b lives in lane (b / K) % 32), row-major semantics kept
//                  exactly by two warp-shuffle prefix-max scans (horizontal gap, running best)
//   walk         : traceback (:170-210) over 4-bit cells written coalesced, one row per store
//   emit         : script_to_aligned_string (:215-261) + trim_mismatch_end (MC/gapalign.cpp:47-68),
//                  ASCII columns written coalesced to the per-candidate slot in HBM
//
// Why the scans are exact (the reference prunes against a ROW-MAJOR running maximum and does not
// decay the horizontal gap across pruned cells, :109-112): see DESIGN.md "Row-parallel X-drop".
//
// This header is compiled by nvcc into libag2_b200.so and -- unchanged, with -DAG2_EMU -- by g++
// against tests/emu/warp_emu.h, a lock-step fibre emulation of one warp used by the CPU-side
// tests to exercise this exact code without a GPU.  The emulation is test infrastructure; the
// product has no CPU path.
#pragma once

#ifdef AG2_EMU
#include "warp_emu.h"
#else
#include <cuda_runtime.h>
#endif
#include <stdint.h>

namespace ag2 {

// MC/xdrop_gapalign.h:98-114, MC/xdrop_gapalign.cpp:8, MC/gapalign.cpp:26
constexpr int kNegInf = -100000000;  // MIN_SCORE; arithmetically live, see dp_block
constexpr int kNegBig = -0x3fffffff; // "no value" for the scans, below anything kNegInf can become
constexpr int kXdrop = 30;
constexpr int kBlk = 500;
constexpr int kBlkSlack = 100;
constexpr int kTailMatch = 4;
constexpr int kFullMapSlack = 20;
constexpr int kMaxBlk = 736;         // a block is at most (int)(599 * 1.2) = 718 bases long
constexpr unsigned kFull = 0xffffffffu;

// 4-bit traceback cell: bits 0-1 op, bit 2 "continue gap in A", bit 3 "continue gap in B"
// (SCRIPT_SUB / SCRIPT_GAP_IN_A / SCRIPT_GAP_IN_B / SCRIPT_EXTEND_GAP_A / _B, xdrop_gapalign.h:26-34)
constexpr int kOpSub = 0, kOpGapA = 1, kOpGapB = 2, kExtA = 4, kExtB = 8;

struct PackedSeqs {
    const uint32_t *ref2;      // reference, 16 bases per word, base i at bits 2*(i%16)
    int64_t ref_len;
    const uint32_t *reads2;    // reads, same packing, every read starts on a word boundary
    const uint32_t *reads_irr; // 1 bit per base (32 per word): base is not upper-case ACGT
    const int64_t *read_off;   // base offset of read r in reads2 (multiple of 32)
    const int32_t *read_len;
};

struct Candidate { // == ag2_candidate
    int32_t read, strand;
    int64_t loc1;
    int32_t loc2, score;
};

// Workspace slot of one candidate: [left area: capL columns][right area: capR columns].
//   lane path : every block of a direction appends its columns in WALK order (end -> origin) to the
//               direction's area; per-block (n_ops, skip) words go to `meta`; assemble_record() puts
//               them in final order
//   wide path : columns are written at their final place, the left direction right-to-left ending
//               at slot + capL, the right direction left-to-right from there
struct ExtGeom {      // per candidate, written by extend_setup_kernel
    int64_t slot;     // first column of this candidate's slot in the workspace strings
    int64_t meta;     // first block-metadata word of this candidate (left blocks, then right blocks)
    int32_t left;     // left_ref_size  (mecat2ref_aux.cpp:186)
    int32_t right;    // right_ref_size (:187)
    int32_t valid;
    int32_t capL;     // columns reserved for the left direction
    int32_t capR;
    int32_t nmetaL;   // block-metadata words reserved for the left direction
    int32_t nmetaR;
    int32_t pad;
};

struct ChainResult {
    int32_t ncols;    // columns emitted by this direction
    int32_t qcons;    // query bases consumed by those columns
    int32_t tcons;    // target bases consumed
    int32_t last_op;  // op of the farthest column (for the left direction's dropped column)
    int32_t nblocks;  // lane path: number of block-metadata words written
    int32_t mode;     // 0 = lane path (segments + metadata), 1 = wide path (final positions)
    int32_t pad0, pad1;
};

struct ChainCounters {
    unsigned long long cells, rows, blocks, interior, wide;
    unsigned long long slots;   // pair kernel: window slots evaluated (8 per executed group for each of a warp's 64 directions); last: initialisers stay short
};

// ASCII of a code 0..3 (ACGT) or 4 ('-'), computed: an indexed read of a string literal is a global load
__device__ __forceinline__ char code_char(int c) { return (char)(c < 4 ? (0x54474341u >> (8 * c)) & 0xffu : 0x2du); }

__device__ __forceinline__ int get2(const uint32_t *p, int64_t i)
{
    return (int)((p[i >> 4] >> (2 * (int)(i & 15))) & 3u);
}
__device__ __forceinline__ int get1(const uint32_t *p, int64_t i)
{
    return (int)((p[i >> 5] >> (int)(i & 31)) & 1u);
}

// (long)(x * 1.2) and (int)(x + x * 0.2) with the reference's double arithmetic, no FMA contraction
__device__ __forceinline__ int64_t mul_1p2(int64_t x) { return __double2ll_rz(__dmul_rn((double)x, 1.2)); }
__device__ __forceinline__ int stretch_0p2(int x)
{
    return __double2int_rz(__dadd_rn((double)x, __dmul_rn((double)x, 0.2)));
}

template <int K>
struct TbLayout {
    static constexpr int kCols = 32 * K;
    static constexpr int kLaneBytes = (K <= 4) ? 2 : 4 * ((K + 7) / 8);
    static constexpr int kRowBytes = 32 * kLaneBytes;
};

// Exclusive prefix max over the lanes in ROTATED order (rl = 0 is the band's head lane).
__device__ __forceinline__ int rot_excl_scan_max(int v, int lane, int rl, int init)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_sync(kFull, v, (lane - d) & 31);
        if (rl >= d) v = max(v, o);
    }
    const int p = __shfl_sync(kFull, v, (lane - 1) & 31);
    return rl >= 1 ? max(p, init) : init;
}

// xdrop_align forward pass.  As/Bs: block codes in shared memory (extension order).
// tb: this warp's traceback scratch in global memory, row a at tb + a * kRowBytes.
// Returns 0, or 1 when the band outgrew the 32*K column window (caller retries with a larger K).
template <int K>
__device__ int dp_block(const uint8_t *As, int M, const uint8_t *Bs, int N, uint8_t *tb, int lane,
                        int &ae_out, int &be_out, ChainCounters &ctr)
{
    using TL = TbLayout<K>;
    int h[K], e[K];
    int first = 0, best = 0, ae = 0, be = 0;
    // row 0 (:53-67): columns 0..min(N, 30)
    int bsize = min(N, kXdrop) + 1;
#pragma unroll
    for (int j = 0; j < K; ++j) {
        const int b = lane * K + j;
        h[j] = b == 0 ? 0 : -b;
        e[j] = h[j] - 1;
    }
    unsigned long long cells = 0, rows = 0, interior = 0;

    for (int a = 1; a <= M; ++a) {
        const int ac = As[a - 1];
        const int fk = first / K;
        const int rl = (lane - fk) & 31;
        const int cb = (fk + rl) * K; // absolute column of this lane's slot 0
        const int hp = __shfl_sync(kFull, h[K - 1], (lane - 1) & 31);
        cells += (unsigned)(bsize - first);
        ++rows;

        int L[K], t[K], E[K];
        unsigned opb = 0, valid = 0;
        int run = kNegBig;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const int b = cb + j;
            const bool v = b >= first && b < bsize;
            const int bc = Bs[min(max(b - 1, 0), kMaxBlk - 1)];
            const int hprev = j == 0 ? hp : h[j - 1];
            const int d = (b == first) ? kNegInf : hprev + (ac == bc ? 1 : -1);
            E[j] = e[j];
            L[j] = max(d, E[j]);
            if (d < E[j]) opb |= 1u << j;
            if (v) valid |= 1u << j;
            run = max(run, v ? L[j] + b : kNegBig);
            t[j] = run;
        }
        const int c1 = rot_excl_scan_max(run, lane, rl, kNegBig);

        int s[K], s2[K];
        unsigned opa = 0, fb = 0;
        int run2 = kNegBig;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const int b = cb + j;
            const int ex = j == 0 ? c1 : max(c1, t[j - 1]);
            const int f0 = ex - b;
            s[j] = max(L[j], f0);
            if (L[j] < f0) opa |= 1u << j;
            if (f0 >= L[j]) fb |= 1u << j;
            run2 = max(run2, (valid >> j & 1u) ? s[j] : kNegBig);
            s2[j] = run2;
        }
        const int c2 = rot_excl_scan_max(run2, lane, rl, best);

        unsigned unp = 0;
        int lmin = 0x7fffffff, lmax = -1, lcnt = 0, slast = kNegBig;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const int b = cb + j;
            const int pj = j == 0 ? c2 : max(c2, s2[j - 1]);
            const bool u = (valid >> j & 1u) && !(pj - s[j] > kXdrop);
            if (u) {
                unp |= 1u << j;
                lmin = min(lmin, b);
                lmax = b;
                ++lcnt;
                slast = s[j];
            }
        }
        const int fmin = __reduce_min_sync(kFull, lmin);
        if (fmin == 0x7fffffff) break; // every cell pruned (:142)
        const int lastu = __reduce_max_sync(kFull, lmax);
        const int cnt = __reduce_add_sync(kFull, lcnt);
        const int rowmax = __reduce_max_sync(kFull, run2);
        const int hg0 = __reduce_max_sync(kFull, lmax == lastu ? slast : kNegBig);
        if (rowmax > best) {
            int lbe = 0x7fffffff;
#pragma unroll
            for (int j = K - 1; j >= 0; --j)
                if ((valid >> j & 1u) && s[j] == rowmax) lbe = cb + j;
            be = __reduce_min_sync(kFull, lbe);
            ae = a;
            best = rowmax;
        }

        if (cnt != lastu - fmin + 1) {
            // Rare: a pruned cell sits between unpruned ones.  Its h becomes kNegInf but its stale e
            // stays reachable from the row below, so its traceback op must be the reference's, which
            // compares against the UNDECAYED horizontal gap = score of the last unpruned cell - 1.
            ++interior;
            int k[K], runk = -1;
#pragma unroll
            for (int j = 0; j < K; ++j) {
                runk = max(runk, (unp >> j & 1u) ? (((cb + j) << 12) | (s[j] + 2048)) : -1);
                k[j] = runk;
            }
            const int ck = rot_excl_scan_max(runk, lane, rl, -1);
#pragma unroll
            for (int j = 0; j < K; ++j) {
                const int b = cb + j;
                const int kk = j == 0 ? ck : max(ck, k[j - 1]);
                if ((valid >> j & 1u) && !(unp >> j & 1u) && b > fmin && b < lastu) {
                    const int realf = ((kk & 4095) - 2048) - 1;
                    if (L[j] < realf) opa |= 1u << j;
                    else opa &= ~(1u << j);
                }
            }
        }

        // column state update (:109-136) and this row's traceback cells
        first = fmin;
        int nb = bsize;
        int hg = kNegBig, next = 0;
        if (lastu < bsize - 1) {
            nb = lastu + 1; // (:144-145)
        } else {
            hg = hg0 - 1;   // horizontal gap score leaving the last cell
            next = max(0, min(N - bsize, hg - (best - kXdrop) + 1)); // (:147-153)
        }
        const int ext_lo = nb, ext_hi = nb + next;
        nb = ext_hi;
        const int sentinel = nb < N ? nb : -1; // (:160-164)
        if (sentinel >= 0) ++nb;

        int nibv[K];
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const int b = cb + j;
            int nib = (opa >> j & 1u) ? kOpGapA : ((opb >> j & 1u) ? kOpGapB : kOpSub);
            if (unp >> j & 1u) {
                if (E[j] == s[j]) nib |= kExtA;
                if (fb >> j & 1u) nib |= kExtB;
                h[j] = s[j];
                e[j] = s[j] - 1;
            } else if ((valid >> j & 1u) && b > fmin) {
                h[j] = kNegInf;
            }
            if (b >= ext_lo && b < ext_hi) {
                h[j] = hg - (b - ext_lo);
                e[j] = h[j] - 1;
                nib = kOpGapA;
            }
            if (b == sentinel) {
                h[j] = kNegInf;
                e[j] = kNegInf;
            }
            nibv[j] = nib;
        }
        uint8_t *trow = tb + (size_t)a * TL::kRowBytes + lane * TL::kLaneBytes;
        if (K <= 4) {
            unsigned w = 0;
#pragma unroll
            for (int j = 0; j < K; ++j) w |= (unsigned)nibv[j] << (4 * j);
            *reinterpret_cast<uint16_t *>(trow) = (uint16_t)w;
        } else {
#pragma unroll
            for (int w0 = 0; w0 < K; w0 += 8) {
                unsigned w = 0;
#pragma unroll
                for (int j = w0; j < w0 + 8 && j < K; ++j) w |= (unsigned)nibv[j] << (4 * (j - w0));
                *reinterpret_cast<uint32_t *>(trow + (w0 >> 1)) = w;
            }
        }

        bsize = nb;
        if (bsize > fk * K + TL::kCols) return 1; // band left the register window: nothing counted
    }
    ae_out = ae;
    be_out = be;
    ctr.cells += cells;
    ctr.rows += rows;
    ctr.interior += interior;
    ctr.blocks += 1;
    return 0;
}

// Shared memory of one warp.
struct WarpSmem {
    uint8_t A[kMaxBlk];
    uint8_t B[kMaxBlk];
    uint8_t ops[2 * kMaxBlk]; // traceback ops in walk order (end -> origin), values kOp*
};

// Traceback (:170-210) by one lane; returns the number of ops and the trim_mismatch_end
// quantities (MC/gapalign.cpp:47-68) which only depend on the first ops of the walk.
template <int K>
__device__ int walk_block(const uint8_t *tb, int ae, int be, WarpSmem &sm, int &qcnt, int &tcnt, int &acnt,
                          int &trim_m, int &trim_w)
{
    using TL = TbLayout<K>;
    int a = ae, b = be, n = 0, cur = kOpSub;
    int m = 0, q = 0, t = 0, ac = 0, w_done = -1;
    while ((a > 0 || b > 0) && n < 2 * kMaxBlk) {
        int cell = kOpGapA; // row 0 is all SCRIPT_GAP_IN_A (:61)
        if (a > 0) {
            const int slot = b % TL::kCols;
            const int l = slot / K, j = slot % K;
            const uint8_t byte = tb[(size_t)a * TL::kRowBytes + l * TL::kLaneBytes + (j >> 1)];
            cell = (byte >> (4 * (j & 1))) & 15;
        }
        if (cur == kOpGapA && (cell & kExtA)) cur = kOpGapA;
        else if (cur == kOpGapB && (cell & kExtB)) cur = kOpGapB;
        else cur = cell & 3;
        bool match = false;
        if (cur == kOpGapA) {
            --b;
        } else if (cur == kOpGapB) {
            --a;
        } else {
            --a;
            --b;
            match = sm.A[a] == sm.B[b];
        }
        if (w_done < 0) { // trim_mismatch_end scans from the END of the block's alignment = walk start
            ++ac;
            if (cur != kOpGapA) ++q;
            if (cur != kOpGapB) ++t;
            m = match ? m + 1 : 0;
            if (m == kTailMatch) w_done = n;
        }
        sm.ops[n++] = (uint8_t)cur;
    }
    qcnt = q;
    tcnt = t;
    acnt = ac;
    trim_m = m;
    trim_w = w_done;
    return n;
}

// Streamed runs (ag2_xdrop_extend_batch with host buffers): ONE launch covers every direction of the batch, and the two
// ends of the pipeline are signalled through memory instead of through kernel boundaries.
//   * in:  the reads go up in pieces while the kernel already runs; a direction waits for the flag of the piece that
//     holds its read (set on the upload stream after the piece's pack kernel);
//   * out: the candidates are cut into chunks; every finished direction is counted on its chunk, and the direction that
//     completes a chunk raises the chunk's flag in host-mapped memory -- the host then finalises, assembles and copies
//     that chunk home while the kernel works on the later ones.
// All pointers null (chunk_cn 0): resident run, nothing is signalled or waited for.
struct StreamSignal {
    unsigned int *chunk_done;   // [chunks] finished directions
    int *chunk_flag;            // [chunks] host-mapped
    int64_t n_cand;
    int32_t chunk_cn;           // candidates per chunk
    int32_t n_pieces;
    const int *piece_flag;      // [n_pieces]
    const int64_t *piece_end;   // [n_pieces] reads [piece_end[k-1], piece_end[k]) arrive with piece k
    unsigned int *error;        // raised when a wait ran into its time limit
};

// Called by the thread that wrote the direction's result (its strings, metadata and ChainResult).
__device__ __forceinline__ void signal_direction_done(const StreamSignal &s, int64_t chain)
{
#ifndef AG2_EMU
    if (!s.chunk_cn) return;
    const int64_t ci = chain >> 1;
    const int64_t c = ci / s.chunk_cn;
    const int64_t in_chunk = min((int64_t)s.chunk_cn, s.n_cand - c * s.chunk_cn);
    __threadfence();
    const unsigned int prev = atomicAdd(s.chunk_done + c, 1u);
    if ((int64_t)prev + 1 == 2 * in_chunk) {
        __threadfence();
        *reinterpret_cast<volatile int *>(s.chunk_flag + c) = 1;
        __threadfence_system();
    }
#endif
}

// Blocks until the piece that holds `read` has been packed.  False: gave up after kStreamWaitNs (the upload stream made
// no progress -- e.g. under a profiler that serialises kernels); the caller drops the direction and the host reports it.
constexpr unsigned long long kStreamWaitNs = 3ull * 1000 * 1000 * 1000;
__device__ __forceinline__ bool wait_for_read(const StreamSignal &s, int64_t read)
{
#ifndef AG2_EMU
    if (!s.piece_flag) return true;
    int k = 0;
    while (k + 1 < s.n_pieces && s.piece_end[k] <= read) ++k;
    const volatile int *flag = s.piece_flag + k;
    if (*flag == 0) {
        const volatile unsigned int *err = s.error;
        unsigned long long t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        for (;;) {
            __nanosleep(2000);
            if (*flag) break;
            if (*err) return false;   // somebody else already ran into the limit: the whole run is void, do not wait again
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > kStreamWaitNs) {
                atomicAdd(s.error, 1u);
                return false;
            }
        }
    }
    __threadfence();
#endif
    return true;
}

struct LaneResume;   // xdrop_lane.cuh: state of a direction at a block boundary

struct ChainArgs {
    StreamSignal sig;
    PackedSeqs seqs;
    const Candidate *cand;
    const ExtGeom *geom;
    ChainResult *res;       // [2 * n]: left, right per candidate
    char *ws_q, *ws_t;      // workspace strings
    uint8_t *tb;            // traceback scratch, tb_stride bytes per resident warp
    size_t tb_stride;
    int64_t n_chains;       // 2 * n candidates, or the length of `queue`
    const int32_t *queue;   // nullptr: chains 0..n_chains-1; else chain ids to run
    unsigned long long *next; // work counter
    int32_t *wide_queue;    // chains that outgrew K (written when wide_count != nullptr)
    unsigned int *wide_count;
    ChainCounters *counters;
    int try_narrow;          // run the direction with kNarrowK columns per lane first (the kernel's K only if its band leaves that window)
    // hand-overs of the pair kernel: entry t of `queue` comes with the state at the block it stopped at and is CONTINUED
    // from there in the pair / lane format (xdrop_lane.cuh: run_chain_resumed); null = every direction starts at its origin
    const LaneResume *resume;
    uint32_t *meta;          // block metadata words of that format
};

constexpr int kNarrowK = 4;  // 128 columns: rows ~5 x shorter than with 23 columns per lane (736, any band)

// One extension direction, start to finish (align_ex, MC/xdrop_gapalign.cpp:263-357).
template <int K>
__device__ bool run_chain(const ChainArgs &g, int64_t chain, WarpSmem &sm, uint8_t *tb, int lane, ChainCounters &ctr)
{
    const int64_t ci = chain >> 1;
    const bool forward = (chain & 1) != 0; // 0 = left (backward), 1 = right
    const Candidate c = g.cand[ci];
    const ExtGeom ge = g.geom[ci];
    ChainResult out = {0, 0, 0, -1, 0, 1, 0, 0};
    if (!ge.valid) {
        if (lane == 0) {
            g.res[chain] = out;
            signal_direction_done(g.sig, chain);
        }
        return true;
    }
    const int rlen = g.seqs.read_len[c.read];
    const int64_t roff = g.seqs.read_off[c.read];
    const int read_start = c.loc2;
    const int64_t ref_start = c.loc1 - 1;
    // XdropAligner::go (:364-396): origins and sizes of the two directions
    const int qsize = forward ? rlen - read_start : read_start;
    const int tsize = forward ? ge.right : ge.left;
    const int q0 = forward ? read_start : read_start - 1;      // oriented read position of block-local 0
    const int64_t t0 = forward ? ref_start : ref_start - 1;    // reference position of block-local 0
    const int inc = forward ? 1 : -1;
    const int64_t mid = ge.slot + ge.capL;                      // first column of the right direction
    int qidx = 0, tidx = 0;
    int ncols = 0, qcons = 0, tcons = 0, last_op = -1;
    ChainCounters lc = {0, 0, 0, 0, 0};

    for (int iter = 0; iter < (1 << 14); ++iter) { // the bound only guards against a hang
        // retrieve_next_aln_block (MC/gapalign.cpp:9-45)
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
        // stage the block: one code per byte, extension order
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
        for (int i = lane; i < tblk; i += 32) sm.B[i] = (uint8_t)get2(g.seqs.ref2, t0 + (int64_t)inc * (tidx + i));
        __syncwarp();

        int ae = 0, be = 0;
        if (qblk > 0 && tblk > 0) {
            if (dp_block<K>(sm.A, qblk, sm.B, tblk, tb, lane, ae, be, lc)) return false;
        }
        __syncwarp();
        // traceback + trim bookkeeping by lane 0
        int nops = 0, qcnt = 0, tcnt = 0, acnt = 0, trim_m = 0, trim_w = -1;
        if (lane == 0) nops = walk_block<K>(tb, ae, be, sm, qcnt, tcnt, acnt, trim_m, trim_w);
        nops = __shfl_sync(kFull, nops, 0);
        qcnt = __shfl_sync(kFull, qcnt, 0);
        tcnt = __shfl_sync(kFull, tcnt, 0);
        acnt = __shfl_sync(kFull, acnt, 0);
        trim_m = __shfl_sync(kFull, trim_m, 0);
        trim_w = __shfl_sync(kFull, trim_w, 0);
        __syncwarp();

        const bool full_map = (qblk - ae <= kFullMapSlack) || (tblk - be <= kFullMapSlack); // (:334-335)
        bool stop = !full_map || last_block;
        int emit = nops;
        if (!stop) {
            // trim_mismatch_end: k = nops-1-w counts down; true iff 4 matches found and k > 0 after --k
            const bool trim = trim_m == kTailMatch && (nops - 2 - trim_w) > 0;
            if (!trim) break; // (:349) this block's columns are dropped
            emit = nops - acnt;
        }
        // script_to_aligned_string (:215-261): column c (from the origin) is walk step nops-1-c
        int qi = 0, ti = 0;
        for (int base = 0; base < emit; base += 32) {
            const int col = base + lane;
            const bool on = col < emit;
            const int op = on ? sm.ops[nops - 1 - col] : kOpGapA;
            const unsigned qm = __ballot_sync(kFull, on && op != kOpGapA);
            const unsigned tm = __ballot_sync(kFull, on && op != kOpGapB);
            const unsigned lt = (1u << lane) - 1u;
            if (on) {
                const char qc = op != kOpGapA ? code_char(sm.A[qi + __popc(qm & lt)]) : '-';
                const char tc = op != kOpGapB ? code_char(sm.B[ti + __popc(tm & lt)]) : '-';
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
        qidx += ae - qcnt; // (:354-355)
        tidx += be - tcnt;
    }
    out.ncols = ncols;
    out.qcons = qcons;
    out.tcons = tcons;
    out.last_op = last_op;
    __syncwarp(); // the strings were written by all lanes
    if (lane == 0) {
        g.res[chain] = out;
        signal_direction_done(g.sig, chain);
    }
    ctr.cells += lc.cells;
    ctr.rows += lc.rows;
    ctr.blocks += lc.blocks;
    ctr.interior += lc.interior;
    return true;
}

struct Record { // == ag2_record
    int32_t ok, read, strand, vscore;
    int32_t qb, qe, qs, aln_len;
    int64_t sb, se;
    int64_t aln_off;
};

// extract_sequences (M2R/mecat2ref_aux.cpp:171-208): window sizes around the seed.
// Returns the number of workspace columns this candidate needs (0 if the candidate is malformed);
// n_meta receives the number of block-metadata words.
__device__ __forceinline__ int64_t setup_one(const Candidate &c, const PackedSeqs &sq, int64_t n_reads, ExtGeom &g,
                                             int64_t &n_meta)
{
    g.slot = 0;
    g.meta = 0;
    g.left = g.right = 0;
    g.valid = 0;
    g.capL = g.capR = g.nmetaL = g.nmetaR = g.pad = 0;
    n_meta = 0;
    if (c.read < 0 || c.read >= n_reads) return 0;
    const int64_t rlen = sq.read_len[c.read];
    const int64_t read_start = c.loc2, ref_start = c.loc1 - 1;
    if (read_start < 0 || read_start > rlen || ref_start < 0 || ref_start > sq.ref_len) return 0;
    const int64_t L1 = read_start, R1 = rlen - read_start;
    const int64_t L2 = ref_start, R2 = sq.ref_len - ref_start;
    const int64_t L = min(L1, L2), R = min(R1, R2);
    g.left = (int32_t)min(L2, mul_1p2(L));
    g.right = (int32_t)min(R2, mul_1p2(R));
    g.valid = 1;
    // a block that is not the last one advances by >= ~480 bases on one side; the slack covers the
    // columns trim_mismatch_end hands back to the next block.  A direction that outgrows either
    // reservation is rerun on the wide path, so these are performance, not correctness, bounds.
    g.nmetaL = (int32_t)((L1 + g.left) / 400 + 4);
    g.nmetaR = (int32_t)((R1 + g.right) / 400 + 4);
    g.capL = (int32_t)(L1 + g.left + 64 * g.nmetaL);
    g.capR = (int32_t)(R1 + g.right + 64 * g.nmetaR);
    n_meta = g.nmetaL + g.nmetaR;
    return (int64_t)g.capL + g.capR;
}

// XdropAligner::go's assembly (:398-438) + extend_candidate's record (mecat2ref_aux.cpp:240-251).
// The farthest left column is dropped (:401-402).  assemble_record() copies the columns.
__device__ __forceinline__ void finalize_one(const Candidate &c, const ExtGeom &g, const ChainResult &l,
                                             const ChainResult &r, int rlen, Record &o, int64_t &str_begin)
{
    o.read = c.read;
    o.strand = c.strand;
    o.vscore = c.score;
    o.qs = rlen;
    o.ok = 0;
    o.qb = o.qe = o.aln_len = 0;
    o.sb = o.se = 0;
    o.aln_off = 0;
    str_begin = 0;
    if (!g.valid) return;
    int lcols = l.ncols, li = l.qcons, lj = l.tcons;
    if (lcols > 0) {
        --lcols;
        if (l.last_op != kOpGapA) --li;
        if (l.last_op != kOpGapB) --lj;
    }
    const int read_start = c.loc2;
    const int64_t ref_start = c.loc1 - 1;
    const int qoff = read_start - li, qend = read_start + r.qcons;
    const int toff = g.left - lj, tend = g.left + r.tcons;
    o.qb = qoff;
    o.qe = qend;
    o.sb = ref_start - g.left + toff;
    o.se = ref_start - g.left + tend;
    o.aln_len = lcols + r.ncols;
    o.ok = (qend - qoff >= 1000) ? 1 : 0;
    str_begin = g.slot + g.capL - lcols; // only meaningful when both directions ran on the wide path
}

} // namespace ag2
