// [lane build] One frontier cell per LANE: the throughput build of the evolution kernel, used only for frontiers far larger
// than the device can hold one-warp-per-cell (ab200_capi.cu picks it above AB_LANE_THRESHOLD cells). Same algorithm and
// arithmetic as the warp build one directory up, written as sequential per-lane code over lane-interleaved scratch arenas.
// ab200_fp.cuh -- directed-rounding FP64 primitives and interval arithmetic for the per-cell kernels.
//
// Device code maps every operation to the sm_100a rounding intrinsics (__dadd_rd/_ru/_rn, __dmul_*, __ddiv_*):
// DADD/DMUL carry the rounding mode in the instruction, so there is no mode register to switch and no FMA
// contraction can change a result. Semantics follow the reference's scalar helpers and Interval type:
//   include/rounding.h:32-68 (MXCSR modes), include/numeric.h:187-230 (add_up/add_down/...),
//   include/numeric.h:262-289,395-529 (Interval, add/sub/mul_ivl/med_ivl/rad_ivl), src/numeric.cpp:732-925.
// The reference also executes plain C++ arithmetic under whatever MXCSR mode the previous routine left behind
// (e.g. TaylorModel::sweep leaves "upward", src/taylor_model.cpp:972-989); Ctx::rnd mirrors that ambient mode and
// the amb_* helpers evaluate in it.
//
// A host build of the same source exists only so the kernel code can be debugged in a GPU-less container
// (tests/tools); it switches the x87/SSE rounding mode per operation and is never part of the product path.
#pragma once
#include <stdint.h>
#include <math.h>
#include <float.h>

#if defined(__CUDACC__)
#define AB_HD __host__ __device__ __forceinline__
#define AB_HDN __host__ __device__ inline __noinline__
#else
#define AB_HD inline
#define AB_HDN inline
#endif

#if !defined(__CUDA_ARCH__)
#include <fenv.h>
#include <string.h>
#endif

// The kernel sources are compiled once per MultiIndex width (AB_KW key words); each build lives in its own inline
// namespace so both can be linked into one library.
#ifndef AB_VARIANT
#define AB_VARIANT kw1
#endif

namespace ab200 { inline namespace AB_VARIANT {

enum { RN = 0, RU = 1, RD = 2 };

// Status codes surfaced through the C ABI (include/ariadne_b200.h).
enum {
    ST_OK = 0,
    ST_SQRT = 1,          // SqrtNumericException        (src/numeric.cpp:251-253)
    ST_DIVZERO = 2,       // DivideByZeroException        (src/numeric.cpp:744,829,858)
    ST_DOMAIN = 3,        // DomainException              (src/taylor_model.cpp:1847)
    ST_BOUNDER = 4,       // ARIADNE_ASSERT in flow_bounds (src/taylor_calculus.cpp:87,108,123-130)
    ST_ARENA = 5,         // per-cell scratch exhausted (no reference counterpart)
    ST_OUTPUT = 6,        // per-cell output list exhausted (no reference counterpart)
    ST_UNSUPPORTED = 7,   // operator / event path not built yet
    ST_IMPLICIT = 8,      // ImplicitFunctionException
    ST_WORKING_SETS = 9   // WorkingSetTooLargeException  (src/hybrid_evolver-image.cpp:280)
};

#if defined(__CUDA_ARCH__)
AB_HD double add_rn(double a, double b) { return __dadd_rn(a, b); }
AB_HD double add_ru(double a, double b) { return __dadd_ru(a, b); }
AB_HD double add_rd(double a, double b) { return __dadd_rd(a, b); }
AB_HD double mul_rn(double a, double b) { return __dmul_rn(a, b); }
AB_HD double mul_ru(double a, double b) { return __dmul_ru(a, b); }
AB_HD double mul_rd(double a, double b) { return __dmul_rd(a, b); }
AB_HD double div_rn(double a, double b) { return __ddiv_rn(a, b); }
AB_HD double div_ru(double a, double b) { return __ddiv_ru(a, b); }
AB_HD double div_rd(double a, double b) { return __ddiv_rd(a, b); }
AB_HD uint64_t d2u(double x) { return (uint64_t)__double_as_longlong(x); }
AB_HD double u2d(uint64_t x) { return __longlong_as_double((long long)x); }
#else
// Debug-only host path.
struct HostRound {
    static inline double op(int mode, char o, double a, double b) {
        volatile double x = a, y = b, r;
        int old = fegetround();
        fesetround(mode == RU ? FE_UPWARD : mode == RD ? FE_DOWNWARD : FE_TONEAREST);
        switch (o) { case '+': r = x + y; break; case '*': r = x * y; break; default: r = x / y; }
        fesetround(old);
        return r;
    }
};
inline double add_rn(double a, double b) { return HostRound::op(RN, '+', a, b); }
inline double add_ru(double a, double b) { return HostRound::op(RU, '+', a, b); }
inline double add_rd(double a, double b) { return HostRound::op(RD, '+', a, b); }
inline double mul_rn(double a, double b) { return HostRound::op(RN, '*', a, b); }
inline double mul_ru(double a, double b) { return HostRound::op(RU, '*', a, b); }
inline double mul_rd(double a, double b) { return HostRound::op(RD, '*', a, b); }
inline double div_rn(double a, double b) { return HostRound::op(RN, '/', a, b); }
inline double div_ru(double a, double b) { return HostRound::op(RU, '/', a, b); }
inline double div_rd(double a, double b) { return HostRound::op(RD, '/', a, b); }
inline uint64_t d2u(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
inline double u2d(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }
#endif

AB_HD double sub_rn(double a, double b) { return add_rn(a, -b); }
AB_HD double sub_ru(double a, double b) { return add_ru(a, -b); }
AB_HD double sub_rd(double a, double b) { return add_rd(a, -b); }

// Operation in a run-time rounding mode (the reference's "ambient" MXCSR state).
AB_HD double add_m(int m, double a, double b) { return m == RU ? add_ru(a, b) : (m == RD ? add_rd(a, b) : add_rn(a, b)); }
AB_HD double sub_m(int m, double a, double b) { return add_m(m, a, -b); }
AB_HD double mul_m(int m, double a, double b) { return m == RU ? mul_ru(a, b) : (m == RD ? mul_rd(a, b) : mul_rn(a, b)); }
AB_HD double div_m(int m, double a, double b) { return m == RU ? div_ru(a, b) : (m == RD ? div_rd(a, b) : div_rn(a, b)); }

AB_HD double dmin(double x, double y) { return (y < x) ? y : x; }   // std::min (include/numeric.h:113)
AB_HD double dmax(double x, double y) { return (x < y) ? y : x; }   // std::max (include/numeric.h:114)
AB_HD double dabs(double x) { return fabs(x); }
AB_HD double dinf() { return u2d(0x7FF0000000000000ull); }

// ---------------------------------------------------------------------------------------------------------------
// Interval (include/numeric.h:262-289). Plain struct: lower, upper.
struct Ivl {
    double l, u;
};
AB_HD Ivl ivl(double l, double u) { Ivl r; r.l = l; r.u = u; return r; }
AB_HD Ivl ivl(double x) { Ivl r; r.l = x; r.u = x; return r; }

AB_HD double ivl_mid(Ivl i) { return add_rn(i.l, i.u) / 2; }            // numeric.h:274 (add_approx then /2)
AB_HD double ivl_rad(Ivl i) { return sub_ru(i.u, i.l) / 2; }            // numeric.h:275
AB_HD double ivl_width(Ivl i) { return sub_ru(i.u, i.l); }              // numeric.h:276
AB_HD double ivl_mag(Ivl i) { return dmax(dabs(i.l), dabs(i.u)); }      // numeric.h:373
AB_HD bool ivl_empty(Ivl i) { return i.l > i.u; }

AB_HD Ivl ivl_neg(Ivl i) { return ivl(-i.u, -i.l); }
AB_HD Ivl ivl_add(Ivl a, Ivl b) { return ivl(add_rd(a.l, b.l), add_ru(a.u, b.u)); }     // numeric.h:424
AB_HD Ivl ivl_add(Ivl a, double x) { return ivl(add_rd(a.l, x), add_ru(a.u, x)); }
AB_HD Ivl ivl_sub(Ivl a, Ivl b) { return ivl(sub_rd(a.l, b.u), sub_ru(a.u, b.l)); }     // numeric.h:446
AB_HD Ivl ivl_sub(Ivl a, double x) { return ivl(sub_rd(a.l, x), sub_ru(a.u, x)); }
AB_HD Ivl ivl_sub(double x, Ivl b) { return ivl(sub_rd(x, b.u), sub_ru(x, b.l)); }
AB_HD Ivl add_ivl(double x, double y) { return ivl(add_rd(x, y), add_ru(x, y)); }
AB_HD Ivl sub_ivl(double x, double y) { return ivl(sub_rd(x, y), sub_ru(x, y)); }
AB_HD Ivl mul_ivl(double x, double y) { return ivl(mul_rd(x, y), mul_ru(x, y)); }
AB_HD Ivl div_ivl(double x, double y) { return ivl(div_rd(x, y), div_ru(x, y)); }
AB_HD Ivl med_ivl(Ivl i) { return add_ivl(i.l / 2, i.u / 2); }                          // numeric.h:523
AB_HD Ivl rad_ivl(Ivl i) { return sub_ivl(i.u / 2, i.l / 2); }                          // numeric.h:526

// src/numeric.cpp:749-784
AB_HD Ivl ivl_mul(Ivl a, Ivl b) {
    double rl, ru;
    if (a.l >= 0) {
        if (b.l >= 0) { rl = mul_rd(a.l, b.l); ru = mul_ru(a.u, b.u); }
        else if (b.u <= 0) { rl = mul_rd(a.u, b.l); ru = mul_ru(a.l, b.u); }
        else { rl = mul_rd(a.u, b.l); ru = mul_ru(a.u, b.u); }
    } else if (a.u <= 0) {
        if (b.l >= 0) { rl = mul_rd(a.l, b.u); ru = mul_ru(a.u, b.l); }
        else if (b.u <= 0) { rl = mul_rd(a.u, b.u); ru = mul_ru(a.l, b.l); }
        else { rl = mul_rd(a.l, b.u); ru = mul_ru(a.l, b.l); }
    } else {
        if (b.l >= 0) { rl = mul_rd(a.l, b.u); ru = mul_ru(a.u, b.u); }
        else if (b.u <= 0) { rl = mul_rd(a.u, b.l); ru = mul_ru(a.l, b.l); }
        else {
            rl = dmin(mul_rd(a.u, b.l), mul_rd(a.l, b.u));
            ru = dmax(mul_ru(a.l, b.l), mul_ru(a.u, b.u));
        }
    }
    return ivl(rl, ru);
}
// src/numeric.cpp:787-798
AB_HD Ivl ivl_mul(Ivl a, double x) {
    if (x >= 0) return ivl(mul_rd(a.l, x), mul_ru(a.u, x));
    return ivl(mul_rd(a.u, x), mul_ru(a.l, x));
}
// src/numeric.cpp:801-832; *st <- ST_DIVZERO when the divisor straddles zero
AB_HD Ivl ivl_div(Ivl a, Ivl b, int* st) {
    double rl, ru;
    if (b.l >= 0) {
        if (a.l >= 0) { rl = div_rd(a.l, b.u); ru = div_ru(a.u, b.l); }
        else if (a.u <= 0) { rl = div_rd(a.l, b.l); ru = div_ru(a.u, b.u); }
        else { rl = div_rd(a.l, b.l); ru = div_ru(a.u, b.l); }
    } else if (b.u <= 0) {
        if (a.l >= 0) { rl = div_rd(a.u, b.u); ru = div_ru(a.l, b.l); }
        else if (a.u <= 0) { rl = div_rd(a.u, b.l); ru = div_ru(a.l, b.u); }
        else { rl = div_rd(a.u, b.u); ru = div_ru(a.l, b.u); }
    } else {
        if (*st == ST_OK) *st = ST_DIVZERO;
        rl = -dinf(); ru = dinf();
    }
    return ivl(rl, ru);
}
// src/numeric.cpp:834-846
AB_HD Ivl ivl_div(Ivl a, double x) {
    if (x > 0) return ivl(div_rd(a.l, x), div_ru(a.u, x));
    if (x < 0) return ivl(div_rd(a.u, x), div_ru(a.l, x));
    return ivl(-dinf(), dinf());
}
// src/numeric.cpp:848-863
AB_HD Ivl ivl_div(double x, Ivl b, int* st) {
    if (b.l <= 0 && b.u >= 0) { if (*st == ST_OK) *st = ST_DIVZERO; return ivl(-dinf(), dinf()); }
    if (x >= 0) return ivl(div_rd(x, b.u), div_ru(x, b.l));
    return ivl(div_rd(x, b.l), div_ru(x, b.u));
}
// src/numeric.cpp:732-746
AB_HD Ivl ivl_rec(Ivl i, int* st) {
    if (i.l > 0 || i.u < 0) return ivl(div_rd(1.0, i.u), div_ru(1.0, i.l));
    if (*st == ST_OK) *st = ST_DIVZERO;
    return ivl(-dinf(), dinf());
}
// src/numeric.cpp:867-882 (the mixed-sign branch really does use mul_down(iu,iu))
AB_HD Ivl ivl_sqr(Ivl i) {
    if (i.l >= 0) return ivl(mul_rd(i.l, i.l), mul_ru(i.u, i.u));
    if (i.u <= 0) return ivl(mul_rd(i.u, i.u), mul_ru(i.l, i.l));
    return ivl(0.0, dmax(mul_ru(i.l, i.l), mul_rd(i.u, i.u)));
}
AB_HD Ivl ivl_abs(Ivl i) {                                                               // numeric.h:405
    if (i.l >= 0) return i;
    if (i.u <= 0) return ivl(-i.u, -i.l);
    return ivl(0.0, dmax(-i.l, i.u));
}
AB_HD Ivl ivl_max(Ivl a, Ivl b) { return ivl(dmax(a.l, b.l), dmax(a.u, b.u)); }          // numeric.h:397
AB_HD Ivl ivl_min(Ivl a, Ivl b) { return ivl(dmin(a.l, b.l), dmin(a.u, b.u)); }

// sqrt_rnd (src/numeric.cpp:248-268): Newton iteration, every operation in rounding mode m.
AB_HD double sqrt_m(int m, double x, int* st) {
    if (x < 0) { if (*st == ST_OK) *st = ST_SQRT; return 0.0; }
    if (x == 0.0) return 0.0;
    int n;
    double y = frexp(x, &n);
    if (n % 2) { y *= 2; n -= 1; }
    double a, b;
    a = 0.0; b = y;
    int guard = 0;
    while (a != b && guard < 200) {
        a = b;
        b = div_m(m, add_m(m, a, div_m(m, y, a)), 2.0);
        ++guard;
    }
    return ldexp(b, n / 2);
}
AB_HD Ivl ivl_sqrt(Ivl i, int* st) { return ivl(sqrt_m(RD, i.l, st), sqrt_m(RU, i.u, st)); }   // numeric.cpp:905-910

// pow_rnd(double,uint) in rounding mode m (src/numeric.cpp:228-238) and pow(Interval,uint) (:892-902)
AB_HD double pow_m(int md, double x, unsigned m) {
    if (m == 0) return 1.0;
    if (x == 0) return 0.0;
    if (x > 0.0 || (m % 2 == 0)) { double r = 1.0, p = dabs(x); while (m) { if (m % 2) r = mul_m(md, r, p); p = mul_m(md, p, p); m /= 2; } return r; }
    double r = -1.0, p = x; while (m) { if (m % 2) r = mul_m(md, -r, p); p = mul_m(md, -p, p); m /= 2; } return r;
}
AB_HD Ivl ivl_pow(Ivl i, unsigned m) {
    if (m % 2 == 0) i = ivl_abs(i);
    return ivl(pow_m(RD, i.l, m), pow_m(RU, i.u, m));
}

// closed-interval predicates (include/numeric.h:376-383)
AB_HD bool ivl_subset(Ivl a, Ivl b) { return a.l >= b.l && a.u <= b.u; }
AB_HD bool ivl_disjoint(Ivl a, Ivl b) { return a.l > b.u || a.u < b.l; }

} } // namespace ab200
