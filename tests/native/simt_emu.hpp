// simt_emu.hpp -- TEST INFRASTRUCTURE: just enough of the CUDA warp model to run warp-cooperative device routines
// (slam_cin0051_b200/csrc/fivept_warp.cuh) on the CPU.  One warp = 32 OS threads; every __shfl / __ballot / __any /
// __syncwarp is a barrier plus an exchange through a 32-slot array, which is exactly the semantics of the *_sync
// intrinsics for code whose lanes all reach the same sequence of calls (the routines under test are written that way).
#pragma once
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>

#define __device__
#define __forceinline__ inline
#define __noinline__
#define SLAM_SIMT_EMULATION 1

struct double2 { double x, y; };
namespace simt {
struct Warp {
    std::barrier<> bar{32};
    uint64_t slot[32];
};
inline Warp* g_warp = nullptr;
struct Tid { int x = 0; };
template <class T> inline T exchange(T v, int lane, int src) {
    uint64_t b = 0;
    static_assert(sizeof(T) <= 8, "shuffle payload");
    std::memcpy(&b, &v, sizeof(T));
    g_warp->slot[lane] = b;
    g_warp->bar.arrive_and_wait();
    const uint64_t r = g_warp->slot[src & 31];
    g_warp->bar.arrive_and_wait();
    T out;
    std::memcpy(&out, &r, sizeof(T));
    return out;
}
}  // namespace simt
inline thread_local simt::Tid threadIdx;

template <class T> inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
    const int lane = threadIdx.x & 31;
    return simt::exchange(v, lane, (lane & ~(width - 1)) + (src & (width - 1)));
}
template <class T> inline T __shfl_xor_sync(unsigned, T v, int off, int width = 32) {
    const int lane = threadIdx.x & 31;
    (void)width;
    return simt::exchange(v, lane, lane ^ off);
}
template <class T> inline T __shfl_down_sync(unsigned, T v, int d, int width = 32) {
    const int lane = threadIdx.x & 31;
    return simt::exchange(v, lane, (lane & (width - 1)) + d < width ? lane + d : lane);
}
template <class T> inline T __shfl_up_sync(unsigned, T v, int d, int width = 32) {
    const int lane = threadIdx.x & 31;
    return simt::exchange(v, lane, (lane & (width - 1)) - d >= 0 ? lane - d : lane);
}
inline unsigned __ballot_sync(unsigned, bool p) {
    const int lane = threadIdx.x & 31;
    simt::g_warp->slot[lane] = p ? 1u : 0u;
    simt::g_warp->bar.arrive_and_wait();
    unsigned r = 0;
    for (int i = 0; i < 32; i++) r |= (unsigned)(simt::g_warp->slot[i] & 1u) << i;
    simt::g_warp->bar.arrive_and_wait();
    return r;
}
inline bool __any_sync(unsigned m, bool p) { return __ballot_sync(m, p) != 0u; }
inline void __syncwarp(unsigned = 0xffffffffu) { simt::g_warp->bar.arrive_and_wait(); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
