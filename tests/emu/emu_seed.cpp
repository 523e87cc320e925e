// emu_seed.cpp -- runs the product's seeding / candidate device code (seed_device.cuh, compiled unchanged
// with -DAG2_EMU) on the CPU against inputs prepared the way the kernels see them.  TEST INFRASTRUCTURE ONLY.
#define AG2_EMU 1
#include "../../aligngraph2_b200/csrc/seed_device.cuh"

#include <vector>

using namespace ag2;

extern "C" {

// cnt/off/pos/vote: an index in the device layout (built by the caller, e.g. from the oracle's index);
// reads: ASCII concatenated.  out: n_reads x maxc x 10 longs (loc1 loc2 left1 left2 right1 right2 score num1 num2 chain),
// ncand: per read.
int emu_seed_batch(long ref_len, const int *cnt, const unsigned *off, const unsigned *pos, const float *vote, int cbl,
                   const char *reads, const long *offs, int n_reads, int pass, int maxc, long *out, int *ncand_out)
{
    RefIndex ix = {ref_len, cnt, off, pos, vote, cbl};
    for (int r = 0; r < n_reads; ++r) {
        const int rlen = (int)(offs[r + 1] - offs[r]);
        std::vector<uint32_t> rd2((rlen >> 4) + 4, 0), irr((rlen >> 5) + 4, 0);
        for (int i = 0; i < rlen; ++i) {
            int code = 0, ir = 1;
            switch (reads[offs[r] + i]) {
            case 'A': code = 0; ir = 0; break;
            case 'C': code = 1; ir = 0; break;
            case 'G': code = 2; ir = 0; break;
            case 'T': code = 3; ir = 0; break;
            case 'a': code = 0; break;
            case 'c': code = 1; break;
            case 'g': code = 2; break;
            case 't': code = 3; break;
            default: break;
            }
            rd2[i >> 4] |= (uint32_t)code << (2 * (i & 15));
            if (ir) irr[i >> 5] |= 1u << (i & 31);
        }
        const int BC = seed_stride(rlen, pass);
        int64_t need = 16;
        for (int s = 0; s < 2; ++s) need = std::max<int64_t>(need, table_bytes(count_hits(ix, rd2.data(), irr.data(), 0, rlen, s, BC)));
        std::vector<uint8_t> scratch((size_t)need + 32);
        uint8_t *sp = (uint8_t *)(((uintptr_t)scratch.data() + 15) & ~(uintptr_t)15);
        SeedCand cands[kMaxCand + 1];
        const int nc = map_read_candidates(ix, rd2.data(), irr.data(), 0, rlen, pass, maxc, sp, cands);
        ncand_out[r] = nc;
        for (int i = 0; i < nc; ++i) {
            long *o = out + ((long)r * maxc + i) * 10;
            o[0] = cands[i].loc1; o[1] = cands[i].loc2; o[2] = cands[i].left1; o[3] = cands[i].left2;
            o[4] = cands[i].right1; o[5] = cands[i].right2; o[6] = cands[i].score; o[7] = cands[i].num1;
            o[8] = cands[i].num2; o[9] = cands[i].chain;
        }
    }
    return 0;
}

} // extern "C"

// ---- the CTA path (seed_cta.cuh): one emulated warp = one CTA ---------------------------------------------------
#include "../../aligngraph2_b200/csrc/seed_cta.cuh"

extern "C" {

// Same contract as emu_seed_batch; `cap` = events a strand may have before the read is reported as overflow
// (ncand_out[r] = -1 then).  Returns the number of overflowed reads.
int emu_seed_cta_batch(long ref_len, const int *cnt, const unsigned *off, const unsigned *pos, const float *vote, int cbl,
                       const char *reads, const long *offs, int n_reads, int pass, int maxc, int cap, long *out, int *ncand_out)
{
    RefIndex ix = {ref_len, cnt, off, pos, vote, cbl};
    // pack the batch the way ag2_reads_load does: every read on a 32-base boundary
    std::vector<int64_t> poff(n_reads + 1);
    std::vector<int32_t> lens(n_reads);
    int64_t p = 0;
    for (int r = 0; r < n_reads; ++r) {
        poff[r] = p;
        lens[r] = (int32_t)(offs[r + 1] - offs[r]);
        p += (lens[r] + 31) & ~31;
    }
    poff[n_reads] = p;
    std::vector<uint32_t> rd2((p >> 4) + 8, 0), irr((p >> 5) + 8, 0);
    for (int r = 0; r < n_reads; ++r)
        for (int i = 0; i < lens[r]; ++i) {
            int code = 0, ir = 1;
            switch (reads[offs[r] + i]) {
            case 'A': code = 0; ir = 0; break;
            case 'C': code = 1; ir = 0; break;
            case 'G': code = 2; ir = 0; break;
            case 'T': code = 3; ir = 0; break;
            case 'a': code = 0; break;
            case 'c': code = 1; break;
            case 'g': code = 2; break;
            case 't': code = 3; break;
            default: break;
            }
            const int64_t f = poff[r] + i;
            rd2[f >> 4] |= (uint32_t)code << (2 * (f & 15));
            if (ir) irr[f >> 5] |= 1u << (f & 31);
        }
    std::vector<SeedCand> cands((size_t)n_reads * maxc);
    std::vector<int32_t> nc(n_reads, 0), ovf(n_reads + 1, 0);
    std::vector<uint32_t> pool((size_t)(cap / (kSM + 1) + 1) * kHeavyWords + 8);
    unsigned next = 0, ovf_count = 0;
    SeedCtaArgs a = {};
    a.ix = ix;
    a.reads2 = rd2.data();
    a.irr = irr.data();
    a.read_off = poff.data();
    a.read_len = lens.data();
    a.n_work = (unsigned)n_reads;
    a.pass = pass;
    a.maxc = maxc;
    a.cap = cap;
    a.next = &next;
    a.cands = cands.data();
    a.ncand = nc.data();
    a.ovf = ovf.data();
    a.ovf_count = &ovf_count;
    a.heavy_pool = pool.data();
    std::vector<uint64_t> smem(seed_cta_smem_bytes(cap) / 8 + 2);
    blockDim.x = 32;
    blockIdx.x = 0;
    warp_emu::run_warp([&]() { seed_cta_body(a, reinterpret_cast<uint8_t *>(smem.data())); });
    for (unsigned i = 0; i < ovf_count; ++i) nc[ovf[i]] = -1;
    for (int r = 0; r < n_reads; ++r) {
        ncand_out[r] = nc[r];
        for (int i = 0; i < nc[r]; ++i) {
            long *o = out + ((long)r * maxc + i) * 10;
            const SeedCand &c = cands[(size_t)r * maxc + i];
            o[0] = c.loc1; o[1] = c.loc2; o[2] = c.left1; o[3] = c.left2; o[4] = c.right1; o[5] = c.right2;
            o[6] = c.score; o[7] = c.num1; o[8] = c.num2; o[9] = c.chain;
        }
    }
    return (int)ovf_count;
}

} // extern "C"
