// h2ops.cuh -- two 16-bit lanes per 32-bit register: the arithmetic the pair kernel (xdrop_pair.cuh) is made of.
//
// A value of type h2 holds two IEEE binary16 numbers (low half = chain 0 of the thread, high half = chain 1).  Every
// number the kernel ever forms is an integer of magnitude <= 2048, which binary16 represents exactly, so the results
// are the integer results.  On sm_100a each helper is one instruction (measured pipe assignment, B200,
// experiments/pipes/pipes2.cu): HADD2 / VIADD.16x2 issue on the FMA pipe; HMNMX2, HSET2 (compare -> 0xffff mask per
// half) and LOP3 on the ALU pipe.  Masks are plain bit masks, selects are LOP3.
//
// With -DAG2_EMU (CPU-side warp emulation, test infrastructure) the same functions are computed in software.
#pragma once

#include <stdint.h>
#ifdef AG2_EMU
#include "warp_emu.h"
#else
#include <cuda_fp16.h>
#endif

namespace ag2 {

typedef uint32_t h2;

// binary16 bit pattern of a small integer (|x| <= 2048), usable in constant expressions
__host__ __device__ constexpr uint32_t h16_bits(int x)
{
    if (x == 0) return 0u;
    const uint32_t sign = x < 0 ? 0x8000u : 0u;
    uint32_t m = x < 0 ? (uint32_t)(-x) : (uint32_t)x;
    int e = 0;
    for (uint32_t t = m; t > 1; t >>= 1) ++e;
    const uint32_t mant = e <= 10 ? ((m << (10 - e)) & 0x3ffu) : ((m >> (e - 10)) & 0x3ffu);
    return sign | ((uint32_t)(e + 15) << 10) | mant;
}
__host__ __device__ constexpr h2 H2C(int x) { return h16_bits(x) | (h16_bits(x) << 16); }

constexpr uint32_t kH2Sign = 0x80008000u;

#ifdef AG2_EMU
namespace h2emu {
inline int dec(uint32_t b) // exact: every value in play is an integer
{
    b &= 0xffffu;
    const int sign = (b & 0x8000u) ? -1 : 1;
    const int e = (int)((b >> 10) & 31u);
    const int mant = (int)(b & 0x3ffu);
    if (e == 0) {
        if (mant != 0) { fprintf(stderr, "h2emu: subnormal\n"); abort(); }
        return 0;
    }
    if (e == 31) { fprintf(stderr, "h2emu: inf/nan\n"); abort(); }
    const int full = mant | 0x400;
    const int sh = e - 25;
    int v;
    if (sh >= 0) v = full << sh;
    else {
        if (full & ((1 << -sh) - 1)) { fprintf(stderr, "h2emu: fraction\n"); abort(); }
        v = full >> -sh;
    }
    return sign * v;
}
inline uint32_t enc(int x) // round to nearest even above 2048, like the hardware (such values only ever meet a factor 0)
{
    if (x <= 2048 && x >= -2048) return h16_bits(x);
    const uint32_t sign = x < 0 ? 0x8000u : 0u;
    uint32_t m = x < 0 ? (uint32_t)(-x) : (uint32_t)x;
    int e = 0;
    for (uint32_t t = m; t > 1; t >>= 1) ++e;
    const int sh = e - 10;
    uint32_t q = m >> sh;
    const uint32_t rem = m & ((1u << sh) - 1u), half = 1u << (sh - 1);
    if (rem > half || (rem == half && (q & 1u))) ++q;
    if (q == 2048u) { q = 1024u; ++e; }
    if (e > 15) { fprintf(stderr, "h2emu: %d overflows binary16\n", x); abort(); }
    return sign | ((uint32_t)(e + 15) << 10) | (q & 0x3ffu);
}
inline bool neg0(uint32_t b) { return (b & 0xffffu) == 0x8000u; }
template <typename F>
inline uint32_t map2(uint32_t a, uint32_t b, F f)
{
    return (f(a & 0xffffu, b & 0xffffu) & 0xffffu) | (f(a >> 16, b >> 16) << 16);
}
} // namespace h2emu

inline h2 h2_add(h2 a, h2 b)
{
    return h2emu::map2(a, b, [](uint32_t x, uint32_t y) { return h2emu::enc(h2emu::dec(x) + h2emu::dec(y)); });
}
inline h2 h2_sub(h2 a, h2 b)
{
    return h2emu::map2(a, b, [](uint32_t x, uint32_t y) { return h2emu::enc(h2emu::dec(x) - h2emu::dec(y)); });
}
inline h2 h2_max(h2 a, h2 b)
{
    return h2emu::map2(a, b, [](uint32_t x, uint32_t y) { return h2emu::dec(x) >= h2emu::dec(y) ? x : y; });
}
inline h2 h2_min(h2 a, h2 b)
{
    return h2emu::map2(a, b, [](uint32_t x, uint32_t y) { return h2emu::dec(x) <= h2emu::dec(y) ? x : y; });
}
inline h2 h2_abs(h2 a) { return a & ~kH2Sign; }
#define AG2_H2_CMP(name, op)                                                                                         \
    inline uint32_t name(h2 a, h2 b)                                                                                 \
    {                                                                                                                \
        return h2emu::map2(a, b, [](uint32_t x, uint32_t y) { return h2emu::dec(x) op h2emu::dec(y) ? 0xffffu : 0u; }); \
    }
AG2_H2_CMP(h2_ge, >=)
AG2_H2_CMP(h2_gt, >)
AG2_H2_CMP(h2_lt, <)
AG2_H2_CMP(h2_le, <=)
AG2_H2_CMP(h2_eq, ==)
AG2_H2_CMP(h2_ne, !=)
#undef AG2_H2_CMP
#define AG2_H2_CMPF(name, op)                                                                                        \
    inline h2 name(h2 a, h2 b)                                                                                       \
    {                                                                                                                \
        return h2emu::map2(a, b, [](uint32_t x, uint32_t y) { return h2emu::dec(x) op h2emu::dec(y) ? 0x3c00u : 0u; }); \
    }
AG2_H2_CMPF(h2_gtf, >)
AG2_H2_CMPF(h2_ltf, <)
AG2_H2_CMPF(h2_eqf, ==)
#undef AG2_H2_CMPF
inline h2 h2_fma(h2 a, h2 b, h2 c) // exact only: every product and sum in play is a small integer
{
    const int lo = h2emu::dec(a) * h2emu::dec(b) + h2emu::dec(c), hi = h2emu::dec(a >> 16) * h2emu::dec(b >> 16) + h2emu::dec(c >> 16);
    return h2emu::enc(lo) | (h2emu::enc(hi) << 16);
}
inline h2 h2_mul(h2 a, h2 b) { return h2_fma(a, b, 0u); }
inline uint32_t h2_opaque(uint32_t x) { return x; }
inline uint32_t and_or(uint32_t a, uint32_t b, uint32_t c) { return (a & b) | c; }
inline uint32_t vadd2(uint32_t a, uint32_t b) { return ((a + b) & 0xffffu) | (((a >> 16) + (b >> 16)) << 16); }
inline h2 h2_from_ints(int lo, int hi) { return h2emu::enc(lo) | (h2emu::enc(hi) << 16); }
inline int h2_lo_int(h2 a) { return h2emu::dec(a); }
inline int h2_hi_int(h2 a) { return h2emu::dec(a >> 16); }
#else
__device__ __forceinline__ __half2 h2_as(h2 a) { return *reinterpret_cast<__half2 *>(&a); }
__device__ __forceinline__ h2 h2_bits(__half2 a) { return *reinterpret_cast<h2 *>(&a); }
__device__ __forceinline__ h2 h2_add(h2 a, h2 b) { return h2_bits(__hadd2(h2_as(a), h2_as(b))); }
__device__ __forceinline__ h2 h2_sub(h2 a, h2 b) { return h2_bits(__hsub2(h2_as(a), h2_as(b))); }
__device__ __forceinline__ h2 h2_max(h2 a, h2 b) { return h2_bits(__hmax2(h2_as(a), h2_as(b))); }
__device__ __forceinline__ h2 h2_min(h2 a, h2 b) { return h2_bits(__hmin2(h2_as(a), h2_as(b))); }
__device__ __forceinline__ h2 h2_abs(h2 a) { return h2_bits(__habs2(h2_as(a))); }
__device__ __forceinline__ uint32_t h2_ge(h2 a, h2 b) { return __hge2_mask(h2_as(a), h2_as(b)); }
__device__ __forceinline__ uint32_t h2_gt(h2 a, h2 b) { return __hgt2_mask(h2_as(a), h2_as(b)); }
__device__ __forceinline__ uint32_t h2_lt(h2 a, h2 b) { return __hlt2_mask(h2_as(a), h2_as(b)); }
__device__ __forceinline__ uint32_t h2_le(h2 a, h2 b) { return __hle2_mask(h2_as(a), h2_as(b)); }
__device__ __forceinline__ uint32_t h2_eq(h2 a, h2 b) { return __heq2_mask(h2_as(a), h2_as(b)); }
__device__ __forceinline__ uint32_t h2_ne(h2 a, h2 b) { return __hne2_mask(h2_as(a), h2_as(b)); }
// compares with a 1.0 / 0.0 result per half: the mask form for selects done as a*g + b on the FMA pipe
__device__ __forceinline__ h2 h2_gtf(h2 a, h2 b) { return h2_bits(__hgt2(h2_as(a), h2_as(b))); }
__device__ __forceinline__ h2 h2_ltf(h2 a, h2 b) { return h2_bits(__hlt2(h2_as(a), h2_as(b))); }
__device__ __forceinline__ h2 h2_eqf(h2 a, h2 b) { return h2_bits(__heq2(h2_as(a), h2_as(b))); }
__device__ __forceinline__ h2 h2_fma(h2 a, h2 b, h2 c) { return h2_bits(__hfma2(h2_as(a), h2_as(b), h2_as(c))); }
__device__ __forceinline__ h2 h2_mul(h2 a, h2 b) { return h2_bits(__hmul2(h2_as(a), h2_as(b))); }
// a constant the compiler must keep in a register (LOP3 takes one immediate; a second constant has to be a register)
__device__ __forceinline__ uint32_t h2_opaque(uint32_t x)
{
    asm volatile("" : "+r"(x));
    return x;
}
// (a & b) | c as ONE LOP3, whatever the compiler knows about b and c (two constants would otherwise be split into two
// instructions with an immediate each)
__device__ __forceinline__ uint32_t and_or(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t vadd2(uint32_t a, uint32_t b) { return __vadd2(a, b); }
__device__ __forceinline__ h2 h2_from_ints(int lo, int hi) { return h2_bits(__halves2half2(__int2half_rn(lo), __int2half_rn(hi))); }
__device__ __forceinline__ int h2_lo_int(h2 a) { return __half2int_rn(__low2half(h2_as(a))); }
__device__ __forceinline__ int h2_hi_int(h2 a) { return __half2int_rn(__high2half(h2_as(a))); }
#endif

// (m ? a : b) per bit
__device__ __forceinline__ uint32_t h2_sel(uint32_t m, uint32_t a, uint32_t b) { return (a & m) | (b & ~m); }

} // namespace ag2
