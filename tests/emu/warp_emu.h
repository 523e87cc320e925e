// warp_emu.h -- lock-step emulation of ONE CUDA warp on the CPU.  TEST INFRASTRUCTURE ONLY.
//
// Lets the CPU-side tests (no GPU in the build container) execute the product's device code
// (aligngraph2_b200/csrc/*.cuh, compiled unchanged with -DAG2_EMU) and compare it with the oracle
// before spending GPU time.  32 lanes run as ucontext fibres inside one OS thread; every warp
// collective (__shfl_sync, __ballot_sync, __reduce_*_sync, __syncwarp) is a barrier at which the
// lanes exchange values.  It is not linked into libag2_b200.so and is not a fallback.
#pragma once

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <ucontext.h>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __shared__ static
#define __launch_bounds__(...)
#define __restrict__

namespace warp_emu {

struct Dim3 { unsigned x = 1, y = 1, z = 1; };

struct State {
    ucontext_t main_ctx;
    ucontext_t lane_ctx[32];
    char *stacks[32];
    bool done[32];
    int cur = 0;
    // collective exchange
    uint64_t slot[2][32];
    int arrived = 0;
    unsigned gen = 0;
    std::function<void()> body;
};

inline State &st()
{
    static State s;
    return s;
}

inline void yield_to_main()
{
    State &s = st();
    swapcontext(&s.lane_ctx[s.cur], &s.main_ctx);
}

// Barrier + all-to-all exchange of one 64-bit value per lane.  Returns the buffer to read from.
inline const uint64_t *exchange(uint64_t v)
{
    State &s = st();
    const unsigned g = s.gen;
    uint64_t *buf = s.slot[g & 1];
    buf[s.cur] = v;
    if (++s.arrived == 32) {
        s.arrived = 0;
        ++s.gen;
    } else {
        while (s.gen == g) yield_to_main();
    }
    return buf;
}

inline void trampoline()
{
    State &s = st();
    s.body();
    s.done[s.cur] = true;
    yield_to_main();
}

// Run `body` once per lane, in lock step, as one warp.
inline void run_warp(const std::function<void()> &body)
{
    State &s = st();
    s.body = body;
    s.arrived = 0;
    const size_t kStack = 1 << 20;
    for (int l = 0; l < 32; ++l) {
        if (!s.stacks[l]) s.stacks[l] = (char *)malloc(kStack);
        getcontext(&s.lane_ctx[l]);
        s.lane_ctx[l].uc_stack.ss_sp = s.stacks[l];
        s.lane_ctx[l].uc_stack.ss_size = kStack;
        s.lane_ctx[l].uc_link = &s.main_ctx;
        makecontext(&s.lane_ctx[l], (void (*)())trampoline, 0);
        s.done[l] = false;
    }
    for (;;) {
        bool any = false;
        for (int l = 0; l < 32; ++l) {
            if (s.done[l]) continue;
            any = true;
            s.cur = l;
            swapcontext(&s.main_ctx, &s.lane_ctx[l]);
        }
        if (!any) break;
    }
}

} // namespace warp_emu

// ---- the CUDA names the device code uses ----------------------------------------------------
struct EmuIdx {
    unsigned y = 0, z = 0;
    struct X { operator unsigned() const { return (unsigned)warp_emu::st().cur; } } x;
};
static EmuIdx threadIdx;
static warp_emu::Dim3 blockIdx_storage, blockDim_storage, gridDim_storage;
#define blockIdx blockIdx_storage
#define blockDim blockDim_storage
#define gridDim gridDim_storage

template <typename T>
inline T __shfl_sync(unsigned, T v, int src)
{
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    const uint64_t *buf = warp_emu::exchange(raw);
    T out;
    memcpy(&out, &buf[src & 31], sizeof(T));
    return out;
}
template <typename T>
inline T __shfl_up_sync(unsigned m, T v, unsigned d)
{
    const int lane = warp_emu::st().cur;
    const T o = __shfl_sync(m, v, lane - (int)d);
    return lane >= (int)d ? o : v;
}
template <typename T>
inline T __shfl_down_sync(unsigned m, T v, unsigned d)
{
    const int lane = warp_emu::st().cur;
    const T o = __shfl_sync(m, v, (lane + (int)d) & 31);
    return lane + (int)d < 32 ? o : v;
}
struct uint4 { unsigned x, y, z, w; };
template <typename T>
inline T __ldcg(const T *p) { return *p; }
inline unsigned __ballot_sync(unsigned, bool p)
{
    const uint64_t *buf = warp_emu::exchange(p ? 1 : 0);
    unsigned r = 0;
    for (int l = 0; l < 32; ++l) r |= (unsigned)(buf[l] & 1) << l;
    return r;
}
inline bool __any_sync(unsigned m, bool p) { return __ballot_sync(m, p) != 0; }
inline bool __all_sync(unsigned m, bool p) { return __ballot_sync(m, p) == 0xffffffffu; }
inline int __reduce_min_sync(unsigned, int v)
{
    const uint64_t *buf = warp_emu::exchange((uint64_t)(int64_t)v);
    int r = (int)(int64_t)buf[0];
    for (int l = 1; l < 32; ++l) r = std::min(r, (int)(int64_t)buf[l]);
    return r;
}
inline int __reduce_max_sync(unsigned, int v)
{
    const uint64_t *buf = warp_emu::exchange((uint64_t)(int64_t)v);
    int r = (int)(int64_t)buf[0];
    for (int l = 1; l < 32; ++l) r = std::max(r, (int)(int64_t)buf[l]);
    return r;
}
inline int __reduce_add_sync(unsigned, int v)
{
    const uint64_t *buf = warp_emu::exchange((uint64_t)(int64_t)v);
    int r = 0;
    for (int l = 0; l < 32; ++l) r += (int)(int64_t)buf[l];
    return r;
}
inline unsigned __reduce_and_sync(unsigned, unsigned v)
{
    const uint64_t *buf = warp_emu::exchange(v);
    unsigned r = 0xffffffffu;
    for (int l = 0; l < 32; ++l) r &= (unsigned)buf[l];
    return r;
}
inline void __syncwarp(unsigned = 0xffffffffu) { warp_emu::exchange(0); }
inline void __syncthreads() { warp_emu::exchange(0); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
inline int __double2int_rz(double v) { return (int)v; }
inline long long __double2ll_rz(double v) { return (long long)v; }
inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v)
{
    const unsigned long long o = *p;
    *p = o + v;
    return o;
}
inline unsigned atomicAdd(unsigned *p, unsigned v)
{
    const unsigned o = *p;
    *p = o + v;
    return o;
}
inline unsigned atomicSub(unsigned *p, unsigned v)
{
    const unsigned o = *p;
    *p = o - v;
    return o;
}
inline unsigned atomicCAS(unsigned *p, unsigned cmp, unsigned val)
{
    const unsigned o = *p;
    if (o == cmp) *p = val;
    return o;
}
using std::max;
inline void __threadfence() {}                 // one emulated warp: program order is memory order
using std::min;
