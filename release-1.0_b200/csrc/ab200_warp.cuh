// ab200_warp.cuh -- the lanes of one warp working on ONE frontier cell.
//
// K_evolve gives a cell to a whole warp. All lanes run the same control flow on the same handles (every scalar of the
// algorithm -- model handles, intervals, error bounds, the status word -- is warp-uniform and lives in registers);
// the loops over the terms of a model are strided over the lanes, and the upward-rounded error sums, which the
// reference accumulates term by term (their association order is part of the result), are carried as chains through
// warp shuffles. This header holds the few collectives the primitives are written in.
//
// Three builds of the same source:
//   device        AB_W = 32, the sm_100a warp intrinsics;
//   host, plain   AB_W = 1 (host-side callers of the primitives: jump sets in ab200_analyser.cu), collectives are trivial;
//   host, AB_WARP_EMU   development aid (tests/tools/warp_emu.h): every lane is a fiber, the collectives switch between
//                 fibers, so the cooperative code can be checked against the goldens in a GPU-less container.
#pragma once
#include "ab200_fp.cuh"

namespace ab200 {
#if !defined(__CUDA_ARCH__) && defined(AB_WARP_EMU)
namespace wemu {   // provided by tests/tools/warp_emu.h
extern int W;
int lane();
uint64_t shfl64(uint64_t v, int src);
unsigned ballot(int pred);
void sync();
}
#endif

inline namespace AB_VARIANT {

// AB_GROUP lanes work on one cell: 32 = one warp per cell; 16 or 8 = two or four cells side by side in one warp, each on its
// own lane group (own shuffle segment, own ballot bits, own barrier mask). Groups that follow the same control flow share
// the instruction stream; groups that diverge are serialised by the hardware and meet again at the reconvergence point.
#ifndef AB_GROUP
#define AB_GROUP 32
#endif
#if defined(__CUDA_ARCH__)
#define AB_W AB_GROUP
AB_HD int wlane() { return (int)(threadIdx.x & (AB_GROUP - 1)); }
#if AB_GROUP == 32
AB_HD unsigned wgmask() { return 0xFFFFFFFFu; }
AB_HD unsigned wballot(bool p) { return __ballot_sync(0xFFFFFFFFu, p); }
#else
AB_HD unsigned wgshift() { return threadIdx.x & 31u & ~(unsigned)(AB_GROUP - 1); }
AB_HD unsigned wgmask() { return ((1u << AB_GROUP) - 1u) << wgshift(); }
AB_HD unsigned wballot(bool p) { return (__ballot_sync(wgmask(), p) >> wgshift()) & ((1u << AB_GROUP) - 1u); }
#endif
AB_HD void wsync() { __syncwarp(wgmask()); }
AB_HD uint64_t wshfl(uint64_t v, int src) { return (uint64_t)__shfl_sync(wgmask(), (unsigned long long)v, src, AB_GROUP); }
AB_HD double wshfl(double v, int src) { return __shfl_sync(wgmask(), v, src, AB_GROUP); }
AB_HD int wshfl(int v, int src) { return __shfl_sync(wgmask(), v, src, AB_GROUP); }
AB_HD int popc32(unsigned m) { return __popc(m); }
AB_HD int ffs32(unsigned m) { return __ffs((int)m); }
AB_HD uint32_t cas32(uint32_t* p, uint32_t expect, uint32_t v) { return atomicCAS(p, expect, v); }
AB_HD void or32(uint32_t* p, uint32_t v) { atomicOr(p, v); }
#elif defined(AB_WARP_EMU)
#define AB_W (::ab200::wemu::W)
inline int wlane() { return wemu::lane(); }
inline void wsync() { wemu::sync(); }
inline unsigned wballot(bool p) { return wemu::ballot(p ? 1 : 0); }
inline uint64_t wshfl(uint64_t v, int src) { return wemu::shfl64(v, src); }
inline double wshfl(double v, int src) { return u2d(wemu::shfl64(d2u(v), src)); }
inline int wshfl(int v, int src) { return (int)(int64_t)wemu::shfl64((uint64_t)(int64_t)v, src); }
inline int popc32(unsigned m) { return __builtin_popcount(m); }
inline int ffs32(unsigned m) { return __builtin_ffs((int)m); }
inline uint32_t cas32(uint32_t* p, uint32_t expect, uint32_t v) { uint32_t o = *p; if (o == expect) *p = v; return o; }
inline void or32(uint32_t* p, uint32_t v) { *p |= v; }
#else
#define AB_W 1
inline int wlane() { return 0; }
inline void wsync() { }
inline unsigned wballot(bool p) { return p ? 1u : 0u; }
inline uint64_t wshfl(uint64_t v, int) { return v; }
inline double wshfl(double v, int) { return v; }
inline int wshfl(int v, int) { return v; }
inline int popc32(unsigned m) { return __builtin_popcount(m); }
inline int ffs32(unsigned m) { return __builtin_ffs((int)m); }
inline uint32_t cas32(uint32_t* p, uint32_t expect, uint32_t v) { uint32_t o = *p; if (o == expect) *p = v; return o; }
inline void or32(uint32_t* p, uint32_t v) { *p |= v; }
#endif

AB_HD bool wany(bool p) { return wballot(p) != 0u; }
AB_HD unsigned lanemask_lt(int lane) { return (lane >= 32) ? 0xFFFFFFFFu : ((1u << lane) - 1u); }

// s <- up(...up(up(s + v[0]) + v[1])... + v[cnt-1]) with v[k] the value lane k holds: the reference's term-by-term upward
// sum, one dependent DADD per term. Every lane computes the same chain, so the result is warp-uniform. Adding +0 leaves a
// non-negative sum unchanged, which lets callers pass 0 for terms that do not contribute. (One copy of each chain in the
// kernel -- not inlined -- to keep the instruction footprint of the primitives small.)
#ifdef AB_COMPOSE_STATS
extern long long g_cs[8];
#define AB_CHAIN_STAT(i, n) do { if (wlane() == 0) { g_cs[i] += (n); } } while (0)
#else
#define AB_CHAIN_STAT(i, n)
#endif
AB_HD double chain_ru(double s, double v, int cnt) {
    AB_CHAIN_STAT(5, cnt);
    int k = 0;
    #pragma unroll 1
    for (; k + 4 <= cnt; k += 4) {
        const double a = wshfl(v, k), b = wshfl(v, k + 1), cc = wshfl(v, k + 2), d = wshfl(v, k + 3);
        s = add_ru(add_ru(add_ru(add_ru(s, a), b), cc), d);
    }
    #pragma unroll 1
    for (; k < cnt; ++k) s = add_ru(s, wshfl(v, k));
    return s;
}
// the same chain in a run-time rounding mode (the reference's ambient MXCSR state)
AB_HDN double chain_m(int m, double s, double v, int cnt) {
    #pragma unroll 1
    for (int k = 0; k < cnt; ++k) s = add_m(m, s, wshfl(v, k));
    return s;
}
// Chain over the lanes flagged in `mask` only, in lane order.
AB_HD double chain_ru_masked(double s, double v, unsigned mask) {
    AB_CHAIN_STAT(6, popc32(mask)); AB_CHAIN_STAT(7, 1);
    #pragma unroll 1
    while (mask) { const int k = ffs32(mask) - 1; mask &= mask - 1; s = add_ru(s, wshfl(v, k)); }
    return s;
}

} } // namespace ab200
