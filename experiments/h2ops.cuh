// h2ops.cuh -- packed 2 x fp16 arithmetic on raw 32-bit registers.
//
// The fast X-drop path keeps two adjacent DP columns in one register as fp16 values.  Every score
// that path ever forms is an integer of magnitude < 2048 (a block is at most 718 x 718 cells and
// pruned values are represented by -inf, see xdrop_fast.cuh), so fp16 add / max / compare are
// EXACT integer operations here; one HADD2 / HMNMX2 / HSET2 handles two cells.
// Under AG2_EMU (tests/emu) the same functions are implemented with _Float16 on the host.
#pragma once
#include <stdint.h>

#ifdef AG2_EMU
#include <cstring>
namespace ag2 {
namespace h2detail {
inline _Float16 lo(uint32_t v) { uint16_t b = (uint16_t)(v & 0xffffu); _Float16 f; memcpy(&f, &b, 2); return f; }
inline _Float16 hi(uint32_t v) { uint16_t b = (uint16_t)(v >> 16); _Float16 f; memcpy(&f, &b, 2); return f; }
inline uint32_t pack(_Float16 l, _Float16 h) { uint16_t a, b; memcpy(&a, &l, 2); memcpy(&b, &h, 2); return (uint32_t)a | ((uint32_t)b << 16); }
} // namespace h2detail
inline uint32_t hadd2(uint32_t a, uint32_t b) { using namespace h2detail; return pack(lo(a) + lo(b), hi(a) + hi(b)); }
inline uint32_t hsub2(uint32_t a, uint32_t b) { using namespace h2detail; return pack(lo(a) - lo(b), hi(a) - hi(b)); }
inline uint32_t hmax2(uint32_t a, uint32_t b) { using namespace h2detail; return pack(lo(a) > lo(b) ? lo(a) : lo(b), hi(a) > hi(b) ? hi(a) : hi(b)); }
#define AG2_H2_CMP(name, op)                                                                        \
    inline uint32_t name(uint32_t a, uint32_t b) { using namespace h2detail;                        \
        return (lo(a) op lo(b) ? 0xffffu : 0u) | (hi(a) op hi(b) ? 0xffff0000u : 0u); }
AG2_H2_CMP(hlt2_mask, <)
AG2_H2_CMP(hgt2_mask, >)
AG2_H2_CMP(hge2_mask, >=)
AG2_H2_CMP(heq2_mask, ==)
#undef AG2_H2_CMP
inline uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
    const uint64_t v = (uint64_t)a | ((uint64_t)b << 32);
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) r |= (uint32_t)((v >> (8 * ((sel >> (4 * i)) & 7))) & 0xff) << (8 * i);
    return r;
}
inline uint32_t h2_from_ints(int l, int h) { using namespace h2detail; return pack((_Float16)l, (_Float16)h); }
inline int h2_to_int(_Float16 f) { const float x = (float)f; return x < -65000.f ? (int)0x80000000 : (int)x; }
inline int h2_lo_int(uint32_t v) { return h2_to_int(h2detail::lo(v)); }
inline int h2_hi_int(uint32_t v) { return h2_to_int(h2detail::hi(v)); }
} // namespace ag2
#else
#include <cuda_fp16.h>
namespace ag2 {
__device__ __forceinline__ __half2 as_h2(uint32_t v) { return *reinterpret_cast<__half2 *>(&v); }
__device__ __forceinline__ uint32_t as_u32(__half2 v) { return *reinterpret_cast<uint32_t *>(&v); }
__device__ __forceinline__ uint32_t hadd2(uint32_t a, uint32_t b) { return as_u32(__hadd2(as_h2(a), as_h2(b))); }
__device__ __forceinline__ uint32_t hsub2(uint32_t a, uint32_t b) { return as_u32(__hsub2(as_h2(a), as_h2(b))); }
__device__ __forceinline__ uint32_t hmax2(uint32_t a, uint32_t b) { return as_u32(__hmax2(as_h2(a), as_h2(b))); }
__device__ __forceinline__ uint32_t hlt2_mask(uint32_t a, uint32_t b) { return __hlt2_mask(as_h2(a), as_h2(b)); }
__device__ __forceinline__ uint32_t hgt2_mask(uint32_t a, uint32_t b) { return __hgt2_mask(as_h2(a), as_h2(b)); }
__device__ __forceinline__ uint32_t hge2_mask(uint32_t a, uint32_t b) { return __hge2_mask(as_h2(a), as_h2(b)); }
__device__ __forceinline__ uint32_t heq2_mask(uint32_t a, uint32_t b) { return __heq2_mask(as_h2(a), as_h2(b)); }
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) { return __byte_perm(a, b, sel); }
__device__ __forceinline__ uint32_t h2_from_ints(int l, int h)
{
    return as_u32(__halves2half2(__int2half_rn(l), __int2half_rn(h)));
}
__device__ __forceinline__ int h2_lo_int(uint32_t v) { return __half2int_rn(__low2half(as_h2(v))); }
__device__ __forceinline__ int h2_hi_int(uint32_t v) { return __half2int_rn(__high2half(as_h2(v))); }
} // namespace ag2
#endif
