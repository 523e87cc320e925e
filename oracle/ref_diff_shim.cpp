/*
 * ref_diff_shim.cpp -- C entry points onto the UNMODIFIED DiffAligner of vanilla MECAT2 (TEST INFRASTRUCTURE ONLY).
 *
 * AlignGraph2.py runs thirdparty/mecat's mecat2ref (seed 15, DiffAligner) for read -> contig and
 * read -> all.fasta.  Compiled by oracle/Makefile together with the reference's own
 *   thirdparty/mecat/src/common/{defs,gapalign,diff_gapalign}.cpp
 * (read in place from /root/reference) into oracle/_ref/libref_mecat_vanilla.so.  It adds no algorithm of its own: every
 * function forwards to the reference symbol named in its comment.  Used by tests/ and tests/golden/gen_diff_golden.py
 * to pin oracle/ag2_diff.c.
 */
#include "diff_gapalign.h"

#include <algorithm>
#include <cstring>

/* defined (non-static, no prototype in any header) at thirdparty/mecat/src/common/diff_gapalign.cpp:107-110 */
int Align(const char* query, const int q_len, const char* target, const int t_len, const int band_tolerance,
          const int get_aln_str, Alignment* align, int* V, int* U, DPathData2* d_path, PathPoint* aln_path,
          const int right_extend);

extern "C" {

void* ref_diff_new(int large_block) { return new DiffAligner(large_block); }
void ref_diff_free(void* p) { delete static_cast<DiffAligner*>(p); }

/* DiffAligner::go (diff_gapalign.cpp:294-349).  Codes 0..3 in, ASCII strings out. */
int ref_diff_go(void* p, const char* q, int qstart, int qsize, const char* t, int tstart, int tsize, int min_aln,
                int* out5 /* qoff qend toff tend aln_size */, char* qaln, char* taln)
{
    DiffAligner* x = static_cast<DiffAligner*>(p);
    const bool ok = x->go(q, qstart, qsize, t, tstart, tsize, min_aln);
    out5[0] = x->query_start();
    out5[1] = x->query_end();
    out5[2] = x->target_start();
    out5[3] = x->target_end();
    out5[4] = x->result->out_store_size;
    memcpy(qaln, x->query_mapped_string(), x->result->out_store_size + 1);
    memcpy(taln, x->target_mapped_string(), x->result->out_store_size + 1);
    return ok ? 1 : 0;
}

/* Align (diff_gapalign.cpp:107-219) on one block, called the way dw_in_one_direction calls it (:227-251): work arrays
 * zeroed, band tolerance 0.3 * max(block sizes), alignment strings wanted.  Q / T point at the block's first base; for
 * right_extend = 0 the block runs towards LOWER addresses.  out6 = aln_q_s aln_q_e aln_t_s aln_t_e dist aln_str_size;
 * the strings are codes 0..4 (4 = gap). */
int ref_diff_block(void* p, const char* Q, int q_len, const char* T, int t_len, int right_extend, int* out6, char* qstr, char* tstr)
{
    DiffAligner* x = static_cast<DiffAligner*>(p);
    std::fill(x->dynq, x->dynq + x->param.row_size, 0);
    std::fill(x->dynt, x->dynt + x->param.column_size, 0);
    const int rc = Align(Q, q_len, T, t_len, 0.3 * std::max(q_len, t_len), 400, x->align, x->dynq, x->dynt, x->d_path,
                         x->aln_path, right_extend);
    out6[0] = x->align->aln_q_s;
    out6[1] = x->align->aln_q_e;
    out6[2] = x->align->aln_t_s;
    out6[3] = x->align->aln_t_e;
    out6[4] = x->align->dist;
    out6[5] = x->align->aln_str_size;
    memcpy(qstr, x->align->q_aln_str, x->align->aln_str_size);
    memcpy(tstr, x->align->t_aln_str, x->align->aln_str_size);
    return rc;
}

} /* extern "C" */
