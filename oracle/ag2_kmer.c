/*
 * ag2_kmer.c -- CPU restatement of PAGraph's kmer_counter.  TEST INFRASTRUCTURE ONLY.
 *
 * Follows /root/reference/PAGraph/src/main/kmer_counter.cpp:19-96 (kmerCounter) and
 * src/tools/kmer/KmerHelper.cpp:7-25 / KmerHelper.hpp:14-31 (kmer2Code, acgt: A0 C1 G2 T3, anything else A).
 * Pinned: tests/test_kmer_counter.py compares its output with the file the unmodified reference binary
 * (oracle/_ref/kmer_counter -t 1) writes, byte for byte, and with the golden digest under tests/golden/.
 */
#include "ag2_oracle.h"

#include <stdlib.h>
#include <string.h>

static inline unsigned acgt(unsigned char ch)
{
    switch (ch) {
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return 0;
    }
}

/* Dense abundance table of every k-mer of every read, then the cut: the smallest abundance a with
 * 1 - (#bins with abundance <= a) / 4^k <= threshold (:58-77); bins with abundance >= cut are "solid" (:81-85).
 * codes_out (may be NULL) receives them in ascending order = the file order of `-t 1`.  Returns their number. */
long orc_solid_kmers(const char *reads, const long *offs, long n_reads, int k, double threshold, uint64_t *codes_out, long cap,
                     long *min_abundance)
{
    const uint64_t nbins = 1ull << (2 * k), mask = nbins - 1;
    uint32_t *tab = (uint32_t *)calloc((size_t)nbins, sizeof(uint32_t));
    for (long r = 0; r < n_reads; ++r) {
        const char *s = reads + offs[r];
        const long len = offs[r + 1] - offs[r];
        uint64_t code = 0;
        for (long i = 0; i < len; ++i) {
            code = ((code << 2) | acgt((unsigned char)s[i])) & mask;
            if (i >= k - 1) ++tab[code];
        }
    }
    /* histogram of abundances (std::map<abundance, bins> in the reference, iterated in ascending order) */
    uint32_t maxab = 0;
    for (uint64_t i = 0; i < nbins; ++i)
        if (tab[i] > maxab) maxab = tab[i];
    uint64_t *hh = (uint64_t *)calloc((size_t)maxab + 1, sizeof(uint64_t));
    for (uint64_t i = 0; i < nbins; ++i) ++hh[tab[i]];
    uint64_t sum = 0;
    long cut = 0;
    for (uint32_t a = 0; a <= maxab; ++a) {
        if (!hh[a]) continue;
        sum += hh[a];
        if (1 - sum * 1.0 / nbins <= threshold) {
            cut = a;
            break;
        }
    }
    long n = 0;
    for (uint64_t i = 0; i < nbins; ++i)
        if ((long)tab[i] >= cut) {
            if (codes_out && n < cap) codes_out[n] = i;
            ++n;
        }
    if (min_abundance) *min_abundance = cut;
    free(hh);
    free(tab);
    return n;
}
