// pipes3: pipes2 + the DPX 3-input s16x2 forms (VIMNMX3, VIADDMNMX: ptxas fuses the two PTX ops of each block into one SASS instruction) and VIMNMX with predicate outputs.
// Which issue pipe do the packed 16-bit ops use on sm_100a?  Exact instruction sequences via inline PTX:
// per thread 8 independent chains of op A (and, in the mixes, 8 of op B), 64 warps per SM.
#include <cstdio>
#include <cstdint>
#define ITER 2048
#define A_HSET2 "set.ge.u32.f16x2 %0, %0, %1;"
#define A_HMNMX2 "max.f16x2 %0, %0, %1;"
#define A_HADD2 "add.f16x2 %0, %0, %1;"
#define A_HFMA2 "fma.rn.f16x2 %0, %0, %1, %1;"
#define A_LOP3 "lop3.b32 %0, %0, %1, %1, 0x6a;"
#define A_VIMNMX2 "max.s16x2 %0, %0, %1;"
#define A_IMAD "mad.lo.s32 %0, %0, %1, %1;"
#define A_PRMT "prmt.b32 %0, %0, %1, 0x5140;"
#define A_IADD "add.s32 %0, %0, %1;"
#define A_SHF "shf.l.wrap.b32 %0, %0, %1, %1;"
#define A_VIADD2 "add.s16x2 %0, %0, %1;"
#define A_VIMNMX3 "{.reg .b32 t; max.s16x2 t, %0, %1; max.s16x2 %0, t, %1;}"
#define A_VIADDMNMX "{.reg .b32 t; add.s16x2 t, %0, %1; max.s16x2 %0, t, %1;}"
#define A_VIMNMX3R "{.reg .b32 t; max.s16x2.relu t, %0, %1; max.s16x2.relu %0, t, %1;}"
#define K1(name, OPA)                                                                         \
    __global__ void name(unsigned *o, unsigned s)                                             \
    {                                                                                         \
        unsigned r[8], c = s ^ 0x3c003c01u;                                                   \
        for (int i = 0; i < 8; ++i) r[i] = s + threadIdx.x * 8 + i;                           \
        for (int it = 0; it < ITER; ++it) {                                                   \
            _Pragma("unroll") for (int i = 0; i < 8; ++i) asm volatile(OPA : "+r"(r[i]) : "r"(c)); \
        }                                                                                     \
        unsigned a = 0;                                                                       \
        for (int i = 0; i < 8; ++i) a ^= r[i];                                                \
        if (a == 0x12345678u) o[0] = a;                                                       \
    }
#define K2(name, OPA, OPB)                                                                    \
    __global__ void name(unsigned *o, unsigned s)                                             \
    {                                                                                         \
        unsigned r[8], q[8], c = s ^ 0x3c003c01u;                                             \
        for (int i = 0; i < 8; ++i) { r[i] = s + threadIdx.x * 8 + i; q[i] = r[i] * 3; }      \
        for (int it = 0; it < ITER; ++it) {                                                   \
            _Pragma("unroll") for (int i = 0; i < 8; ++i) {                                   \
                asm volatile(OPA : "+r"(r[i]) : "r"(c));                                      \
                asm volatile(OPB : "+r"(q[i]) : "r"(c));                                      \
            }                                                                                 \
        }                                                                                     \
        unsigned a = 0;                                                                       \
        for (int i = 0; i < 8; ++i) a ^= r[i] ^ q[i];                                         \
        if (a == 0x12345678u) o[0] = a;                                                       \
    }
K1(s_hset2, A_HSET2) K1(s_hmnmx2, A_HMNMX2) K1(s_hadd2, A_HADD2) K1(s_hfma2, A_HFMA2) K1(s_lop3, A_LOP3)
K1(s_vimnmx2, A_VIMNMX2) K1(s_imad, A_IMAD) K1(s_prmt, A_PRMT) K1(s_iadd, A_IADD) K1(s_shf, A_SHF) K1(s_viadd2, A_VIADD2)
K1(s_vimnmx3, A_VIMNMX3) K1(s_viaddmnmx, A_VIADDMNMX) K1(s_vimnmx3r, A_VIMNMX3R)
K2(m_vimnmx3_hfma2, A_VIMNMX3, A_HFMA2) K2(m_vimnmx3_lop3, A_VIMNMX3, A_LOP3) K2(m_viaddmnmx_hfma2, A_VIADDMNMX, A_HFMA2)
K2(m_viaddmnmx_lop3, A_VIADDMNMX, A_LOP3) K2(m_vimnmx3_hmnmx2, A_VIMNMX3, A_HMNMX2) K2(m_vimnmx3_vimnmx2, A_VIMNMX3, A_VIMNMX2)
K2(m_hset2_lop3, A_HSET2, A_LOP3) K2(m_hset2_hfma2, A_HSET2, A_HFMA2) K2(m_hmnmx2_lop3, A_HMNMX2, A_LOP3)
K2(m_hmnmx2_hfma2, A_HMNMX2, A_HFMA2) K2(m_hadd2_lop3, A_HADD2, A_LOP3) K2(m_hadd2_hfma2, A_HADD2, A_HFMA2)
K2(m_vimnmx2_lop3, A_VIMNMX2, A_LOP3) K2(m_vimnmx2_hfma2, A_VIMNMX2, A_HFMA2) K2(m_imad_hfma2, A_IMAD, A_HFMA2)
K2(m_imad_lop3, A_IMAD, A_LOP3) K2(m_prmt_lop3, A_PRMT, A_LOP3) K2(m_hset2_hmnmx2, A_HSET2, A_HMNMX2)
K2(m_viadd2_lop3, A_VIADD2, A_LOP3) K2(m_viadd2_hfma2, A_VIADD2, A_HFMA2) K2(m_hset2_hadd2, A_HSET2, A_HADD2)
struct K { const char *n; void (*f)(unsigned *, unsigned); int ops; };
#define E1(x) {#x, x, 1}
#define E2(x) {#x, x, 2}
int main()
{
    K ks[] = {E1(s_hset2), E1(s_hmnmx2), E1(s_hadd2), E1(s_hfma2), E1(s_lop3), E1(s_vimnmx2), E1(s_imad), E1(s_prmt), E1(s_iadd), E1(s_shf), E1(s_viadd2),
              E1(s_vimnmx3), E1(s_viaddmnmx), E1(s_vimnmx3r), E2(m_vimnmx3_hfma2), E2(m_vimnmx3_lop3), E2(m_viaddmnmx_hfma2), E2(m_viaddmnmx_lop3), E2(m_vimnmx3_hmnmx2), E2(m_vimnmx3_vimnmx2),
              E2(m_hset2_lop3), E2(m_hset2_hfma2), E2(m_hmnmx2_lop3), E2(m_hmnmx2_hfma2), E2(m_hadd2_lop3), E2(m_hadd2_hfma2),
              E2(m_vimnmx2_lop3), E2(m_vimnmx2_hfma2), E2(m_imad_hfma2), E2(m_imad_lop3), E2(m_prmt_lop3), E2(m_hset2_hmnmx2),
              E2(m_viadd2_lop3), E2(m_viadd2_hfma2), E2(m_hset2_hadd2)};
    unsigned *o;
    cudaMalloc(&o, 4);
    int sms, clk;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
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
        const double wi = (double)sms * 4 * 16 * ITER * 8 * k.ops;
        printf("%-18s %7.3f ms  %5.2f warp-instr/clk/SM\n", k.n, ms, wi / (ms * 1e-3 * clk * 1e3) / sms);
    }
    return 0;
}
