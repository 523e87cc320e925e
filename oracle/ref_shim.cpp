/*
 * ref_shim.cpp -- C entry points onto the UNMODIFIED reference sources (TEST INFRASTRUCTURE ONLY).
 *
 * Compiled by oracle/Makefile together with the reference's own
 *   MC/xdrop_gapalign.cpp, MC/gapalign.cpp, MC/defs.cpp, M2R/mecat2ref_aux.cpp, M2R/output.cpp
 * (read in place from /root/reference) into oracle/_ref/libref_mecat.so.  It adds no algorithm of
 * its own: every function forwards to the reference symbol named in its comment.  Used by tests/
 * and gen_golden.py to pin oracle/ag2_oracle.c, and by bench.py's `--impl reference` leg.
 */
#include "xdrop_gapalign.h"
#include "mecat2ref_aux.h"

#include <cstring>
#include <vector>

/* defined (non-static, no prototype in any header) at MC/xdrop_gapalign.cpp:10-26 */
int xdrop_align(const char* A, const int M, const char* B, const int N, int matrix[][4], int gap_open,
                int gap_extend, int x_dropoff, u8* state_array, BlastGapDP* score_array,
                u8** edit_script, int* edit_start_offset, GapPrelimEditBlock* edit_block,
                const bool forward, int& ae, int& be);

extern "C" {

/* XdropAligner::go  (MC/xdrop_gapalign.cpp:360-439).  Codes 0..3 in, ASCII strings out. */
void* ref_xdrop_new() { return new XdropAligner(0); }
void ref_xdrop_free(void* p) { delete static_cast<XdropAligner*>(p); }

int ref_xdrop_go(void* p, const char* q, int qstart, int qsize, const char* t, int tstart, int tsize,
                 int min_aln, int* out5 /* qoff qend toff tend aln_size */, char* qaln, char* taln)
{
    XdropAligner* x = static_cast<XdropAligner*>(p);
    bool ok = x->go(q, qstart, qsize, t, tstart, tsize, min_aln);
    out5[0] = x->query_start();
    out5[1] = x->query_end();
    out5[2] = x->target_start();
    out5[3] = x->target_end();
    out5[4] = x->aln_size;
    memcpy(qaln, x->query_mapped_string(), x->aln_size + 1);
    memcpy(taln, x->target_mapped_string(), x->aln_size + 1);
    return ok ? 1 : 0;
}

/* xdrop_align  (MC/xdrop_gapalign.cpp:11-213): one block DP.  ops[] receives the run-length
 * edit script expanded to one op per step, in the order the reference appends them (walk order,
 * i.e. from (ae,be) back to (0,0)). */
int ref_xdrop_block(void* p, const char* A, int M, const char* B, int N, int forward,
                    int* ae, int* be, unsigned char* ops, int* nops)
{
    XdropAligner* x = static_cast<XdropAligner*>(p);
    int a = 0, b = 0;
    int score = xdrop_align(A, M, B, N, x->score_matrix, x->param.gap_open, x->param.gap_extend,
                            x->param.x_dropoff, x->state_array, x->score_array, x->edit_script,
                            x->edit_start_offset, &x->edit_block, forward != 0, a, b);
    *ae = a;
    *be = b;
    int n = 0;
    for (int i = 0; i < x->edit_block.num_ops; ++i)
        for (int j = 0; j < x->edit_block.edit_ops[i].num; ++j)
            ops[n++] = (unsigned char)x->edit_block.edit_ops[i].op_type;
    *nops = n;
    return score;
}

/* extend_candidate  (M2R/mecat2ref_aux.cpp:210-270) incl. extract_sequences (:171-208).
 * out7 = ok qb qe qs, sb se (as two longs in outl) ... */
int ref_extend_candidate(void* p, const char* raw_ref, long ref_size, const char* fwd_read,
                         const char* rev_read, int read_len, long loc1, long loc2, int chain,
                         int score, int* outi /* qb qe qs vscore */, long* outl /* sb se */,
                         char* qmap, char* smap)
{
    XdropAligner* x = static_cast<XdropAligner*>(p);
    candidate_save can;
    memset(&can, 0, sizeof(can));
    can.loc1 = loc1;
    can.loc2 = loc2;
    can.chain = (char)chain;
    can.score = score;
    std::vector<char> qstr, tstr;
    TempResult r;
    r.qmap = qmap;
    r.smap = smap;
    int ntr = 0;
    bool ok = extend_candidate(can, x, raw_ref, ref_size, fwd_read, rev_read, qstr, tstr, 0, read_len,
                               NULL, NULL, &r, ntr);
    if (ok) {
        outi[0] = r.qb; outi[1] = r.qe; outi[2] = r.qs; outi[3] = r.vscore;
        outl[0] = r.sb; outl[1] = r.se;
    }
    return ok ? 1 : 0;
}

} /* extern "C" */
