// emu_map.cpp -- the product's per-read pipeline logic on the CPU: seed_device.cuh (seeding, candidates) and
// rescue_device.cuh (plan_read / finish_read), compiled unchanged with -DAG2_EMU.  The extensions themselves are
// done by the oracle here (they are tested separately); what this checks is everything AROUND them -- candidate
// order, rescue planning, linking, output choice, second pass -- by writing a `.r` thread file that must equal the
// reference's.  TEST INFRASTRUCTURE ONLY.
#define AG2_EMU 1
#include "../../aligngraph2_b200/csrc/rescue_device.cuh"
#include "../../oracle/ag2_oracle.h"

#include <string>
#include <vector>

using namespace ag2;

namespace {
struct Rec {
    Record r;
    std::string q, t;
};

void pack(const char *s, int n, std::vector<uint32_t> &w2, std::vector<uint32_t> &irr)
{
    w2.assign((n >> 4) + 4, 0);
    irr.assign((n >> 5) + 4, 0);
    for (int i = 0; i < n; ++i) {
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
        default: break;
        }
        w2[i >> 4] |= (uint32_t)code << (2 * (i & 15));
        if (ir) irr[i >> 5] |= 1u << (i & 31);
    }
}
} // namespace

extern "C" long emu_map_batch(const char *ref, long ref_len, const int *cnt, const unsigned *off, const unsigned *pos, const float *vote,
                              int cbl, const char *reads, const long *offs, const int *ids, int n_reads, int maxc, int num_output,
                              const char *r_path, long *stats /* rescue_planned pass2_reads */)
{
    RefIndex ix = {ref_len, cnt, off, pos, vote, cbl};
    orc_xdrop *x = orc_xdrop_new();
    FILE *out = fopen(r_path, "w");
    long written = 0;
    stats[0] = stats[1] = 0;
    for (int r = 0; r < n_reads; ++r) {
        const int rlen = (int)(offs[r + 1] - offs[r]);
        std::string fwd(reads + offs[r], rlen), rev(rlen, 'N');
        for (int i = 0; i < rlen; ++i) {
            char c = fwd[rlen - 1 - i];
            switch (c) {
            case 'A': c = 'T'; break;
            case 'T': c = 'A'; break;
            case 'C': c = 'G'; break;
            case 'G': c = 'C'; break;
            default: break;
            }
            rev[i] = c;
        }
        std::vector<uint32_t> rd2, irr;
        pack(fwd.data(), rlen, rd2, irr);
        auto extend = [&](int chain, int64_t loc1, int32_t loc2, int score, Rec &o) {
            long rec[4];
            orc_aln a;
            const std::string &rd = chain == 'F' ? fwd : rev;
            o.r.ok = orc_extend_candidate(x, ref, ref_len, rd.c_str(), rlen, loc1, loc2, rec, &a);
            o.r.read = r;
            o.r.strand = chain == 'F' ? 0 : 1;
            o.r.vscore = score;
            o.r.qb = (int)rec[0];
            o.r.qe = (int)rec[1];
            o.r.qs = rlen;
            o.r.sb = rec[2];
            o.r.se = rec[3];
            o.r.aln_len = a.aln_size;
            o.q.assign(a.qaln, a.aln_size);
            o.t.assign(a.taln, a.aln_size);
        };
        for (int pass = 0; pass < 2; ++pass) {
            const int BC = seed_stride(rlen, pass);
            int64_t need = 16;
            for (int s = 0; s < 2; ++s) need = std::max<int64_t>(need, table_bytes(count_hits(ix, rd2.data(), irr.data(), 0, rlen, s, BC)));
            std::vector<uint8_t> scratch((size_t)need + 32);
            uint8_t *sp = (uint8_t *)(((uintptr_t)scratch.data() + 15) & ~(uintptr_t)15);
            SeedCand cands[kMaxCand + 1];
            const int nc = map_read_candidates(ix, rd2.data(), irr.data(), 0, rlen, pass, maxc, sp, cands);
            std::vector<Rec> pool(nc + kMaxRescue);
            std::vector<Record> recs(nc);
            for (int c = 0; c < nc; ++c) {
                extend(cands[c].chain, cands[c].loc1, (int32_t)cands[c].loc2, cands[c].score, pool[c]);
                recs[c] = pool[c].r;
            }
            ReadPlan P;
            plan_read(ix, rd2.data(), irr.data(), 0, rlen, pass, recs.data(), nc, 0, sp, P);
            Record rrec[kMaxRescue];
            int64_t rref[kMaxRescue];
            for (int s = 0; s < kMaxRescue; ++s) {
                rref[s] = nc + s;
                rrec[s].ok = 0;
                if (!P.rescue[s].on) continue;
                ++stats[0];
                extend(P.rescue[s].chain, P.rescue[s].loc1, P.rescue[s].loc2, P.rescue[s].score, pool[nc + s]);
                rrec[s] = pool[nc + s].r;
            }
            int64_t refs[3 * kMaxAlns];
            const int nout = finish_read(P, rrec, rref, rlen, num_output, refs);
            for (int k = 0; k < nout; ++k) {
                const Rec &o = pool[refs[k]];
                fprintf(out, "%d\t%c\t%d\t%d\t%d\t%d\t%ld\t%ld\n%s\n%s\n", ids[r], o.r.strand ? 'R' : 'F', o.r.vscore, o.r.qb, o.r.qe,
                        o.r.qs, (long)o.r.sb, (long)o.r.se, o.q.c_str(), o.t.c_str());
                ++written;
            }
            if (P.naln_ext != 0) break;
            if (pass == 0) ++stats[1];
        }
    }
    fclose(out);
    orc_xdrop_free(x);
    return written;
}
