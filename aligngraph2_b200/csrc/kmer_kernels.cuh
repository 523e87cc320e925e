// kmer_kernels.cuh -- PAGraph's kmer_counter on the GPU (PAGraph/src/main/kmer_counter.cpp:19-96,
// src/tools/kmer/KmerHelper.cpp:7-25): dense abundance table of every k-mer of every read, the abundance cut, and the
// selection of the solid k-mers.
//
// HBM-bound scatter work.  One thread per k-mer end position of the packed reads (the 2-bit packing already maps
// acgt case-insensitively and everything else to A, exactly KmerHelper::acgt); lanes of a warp that hit the same bin
// (homopolymers, tandem repeats) are merged with __match_any_sync into one atomic.
#pragma once

#include "index_kernels.cuh"

namespace ag2 {

// reads2: 2-bit packed, every read starts on a 32-base boundary (read_off), so a warp-aligned group of 32 positions
// belongs to one read: lane 0 finds it, the warp shares it.
__global__ void kmer_abundance_kernel(const uint32_t *__restrict__ reads2, const int64_t *__restrict__ read_off,
                                      const int32_t *__restrict__ read_len, int64_t n_reads, int64_t n_groups, int k,
                                      uint32_t *__restrict__ table)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const uint64_t mask = (1ull << (2 * k)) - 1;
    for (int64_t g = warp; g < n_groups; g += nwarps) {
        const int64_t pb = g << 5;
        int64_t lo = 0;
        if (lane == 0) {
            int64_t hi = n_reads - 1;
            while (lo < hi) {
                const int64_t mid = (lo + hi + 1) >> 1;
                if (read_off[mid] <= pb) lo = mid;
                else hi = mid - 1;
            }
        }
        lo = __shfl_sync(0xffffffffu, lo, 0);
        const int64_t local = pb - read_off[lo] + lane; // position inside the read
        const bool ok = local >= k - 1 && local < read_len[lo];
        uint64_t code = 0;
        if (ok) {
            const int64_t s = pb + lane - (k - 1); // first base of the k-mer (k <= 16: at most two words)
            const int64_t w = s >> 4;
            const uint64_t bits = ((uint64_t)reads2[w + 1] << 32 | reads2[w]) >> (2 * (s & 15));
#pragma unroll 4
            for (int q = 0; q < k; ++q) code = (code << 2) | ((bits >> (2 * q)) & 3u);
            code &= mask;
        }
        warp_aggregated_add(reinterpret_cast<int32_t *>(table), (int)code, 1, ok);
    }
}

// histogram of abundances in [lo, lo + kAbWindow); abundance 0 dominates, so every block first counts in shared memory
constexpr int kAbWindow = 4096;
__global__ void abundance_hist_kernel(const uint32_t *__restrict__ table, int64_t nbins, uint32_t lo, unsigned long long *hist,
                                      unsigned int *max_abundance)
{
    __shared__ unsigned int sh[kAbWindow];
    for (int i = threadIdx.x; i < kAbWindow; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    unsigned int mx = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nbins; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t a = table[i];
        mx = max(mx, a);
        if (a >= lo && a - lo < (uint32_t)kAbWindow) atomicAdd(&sh[a - lo], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kAbWindow; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], (unsigned long long)sh[i]);
    mx = __reduce_max_sync(0xffffffffu, mx);
    if ((threadIdx.x & 31) == 0 && mx) atomicMax(max_abundance, mx);
}

// solid codes in ascending order: flags -> scan (index_kernels' tiled scan over int32 flags) -> scatter
__global__ void solid_flags_kernel(const uint32_t *__restrict__ table, int64_t nbins, uint32_t cut, int32_t *flags)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nbins; i += (int64_t)gridDim.x * blockDim.x)
        flags[i] = table[i] >= cut ? 1 : 0;
}
__global__ void solid_scatter_kernel(const int32_t *__restrict__ flags, const uint32_t *__restrict__ offs, int64_t first, int64_t n,
                                     uint64_t out_base, uint64_t *out)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if (flags[i]) out[out_base + offs[i]] = (uint64_t)(first + i);
}

// dst[i] += src[i]: the sum of the per-GPU abundance tables; src may be peer memory
__global__ void kmer_table_add_kernel(uint32_t *__restrict__ dst, const uint32_t *__restrict__ src, int64_t n)
{
    const int64_t n4 = n >> 2;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        uint4 a = reinterpret_cast<uint4 *>(dst)[i];
        const uint4 b = reinterpret_cast<const uint4 *>(src)[i];
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        reinterpret_cast<uint4 *>(dst)[i] = a;
    }
    for (int64_t i = (n4 << 2) + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[i] += src[i];
}

} // namespace ag2
