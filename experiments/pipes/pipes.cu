// Microbenchmark: per-SM issue rate of the packed 16-bit instructions the pair-lane X-drop kernel is built from.
// Each kernel runs ITER x 8 independent dependency chains of one op (or a mix) per thread.
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdint>
#define ITER 4096
#define DEF(name, BODY)                                                                      \
    __global__ void name(unsigned *o, unsigned s)                                            \
    {                                                                                        \
        unsigned r[8], c = s ^ 0x3c003c00u, d = s | 0x00010001u;                             \
        for (int i = 0; i < 8; ++i) r[i] = s + threadIdx.x * 8 + i;                          \
        for (int it = 0; it < ITER; ++it) {                                                  \
            _Pragma("unroll") for (int i = 0; i < 8; ++i) { unsigned &x = r[i]; BODY; }      \
        }                                                                                    \
        unsigned a = 0;                                                                      \
        for (int i = 0; i < 8; ++i) a ^= r[i];                                               \
        if (a == 0x12345678u) o[0] = a;                                                      \
    }
__device__ __forceinline__ unsigned h2u(__half2 h) { return *reinterpret_cast<unsigned *>(&h); }
__device__ __forceinline__ __half2 u2h(unsigned u) { return *reinterpret_cast<__half2 *>(&u); }
DEF(k_lop3, x = (x & c) | (d & ~x) ^ it)
DEF(k_hadd2, x = h2u(__hadd2(u2h(x), u2h(c))))
DEF(k_hfma2, x = h2u(__hfma2(u2h(x), u2h(c), u2h(d))))
DEF(k_hmnmx2, x = h2u(__hmax2(u2h(x), u2h(c))); c += it)
DEF(k_hset2, x = __hge2_mask(u2h(x), u2h(c)) ^ d)
DEF(k_hset2only, x = __hge2_mask(u2h(x), u2h(c)); c += it)
DEF(k_vimnmx2, x = __vmaxs2(x, c); c += it)
DEF(k_viadd2, x = __vadd2(x, c))
DEF(k_imad, x = x * c + d)
DEF(k_prmt, x = __byte_perm(x, c, 0x5140 + (it & 1)))
DEF(k_mix_hadd_lop, x = h2u(__hadd2(u2h(x), u2h(c))); x = (x & d) ^ c)
DEF(k_mix_hset_lop, x = __hge2_mask(u2h(x), u2h(c)); x = (x & d) ^ c)
DEF(k_mix_hmnmx_lop, x = h2u(__hmax2(u2h(x), u2h(c))); x = (x & d) ^ c)
DEF(k_mix_hmnmx_hadd, x = h2u(__hmax2(u2h(x), u2h(c))); x = h2u(__hadd2(u2h(x), u2h(d))))
DEF(k_mix_hset_hadd, x = __hge2_mask(u2h(x), u2h(c)); x = h2u(__hadd2(u2h(x), u2h(d))))
DEF(k_mix_vimnmx_hadd, x = __vmaxs2(x, c); x = h2u(__hadd2(u2h(x), u2h(d))))
DEF(k_mix_imad_lop, x = x * c + d; x = (x & d) ^ c)
struct K { const char *n; void (*f)(unsigned *, unsigned); int ops; };
int main()
{
    K ks[] = {{"lop3", k_lop3, 2}, {"hadd2", k_hadd2, 1}, {"hfma2", k_hfma2, 1}, {"hmnmx2(+iadd)", k_hmnmx2, 1}, {"hset2+lop3", k_hset2, 2},
              {"hset2", k_hset2only, 1}, {"vimnmx.s16x2", k_vimnmx2, 1}, {"viadd.16x2", k_viadd2, 1}, {"imad", k_imad, 1}, {"prmt", k_prmt, 1},
              {"hadd2+lop3", k_mix_hadd_lop, 2}, {"hset2+lop3", k_mix_hset_lop, 2}, {"hmnmx2+lop3", k_mix_hmnmx_lop, 2},
              {"hmnmx2+hadd2", k_mix_hmnmx_hadd, 2}, {"hset2+hadd2", k_mix_hset_hadd, 2}, {"vimnmx2+hadd2", k_mix_vimnmx_hadd, 2},
              {"imad+lop3", k_mix_imad_lop, 2}};
    unsigned *o;
    cudaMalloc(&o, 4);
    int sms;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (auto &k : ks) {
        k.f<<<sms * 4, 512>>>(o, 3);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        k.f<<<sms * 4, 512>>>(o, 3);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double warp_instr = (double)sms * 4 * 16 * ITER * 8 * k.ops;
        printf("%-16s %8.3f ms  %6.2f warp-instr/clk/SM (nominal ops only, clock %d kHz)\n", k.n, ms, warp_instr / (ms * 1e-3 * clk * 1e3) / sms, clk);
    }
    return 0;
}
