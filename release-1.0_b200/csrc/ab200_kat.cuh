// ab200_kat.cuh -- the dispatch behind the primitive-level entry points (ab200_taylor_model_op, ab200_interval_op): one
// TaylorModel operator / function or one Interval operation, by op code (include/ariadne_b200.h lists the codes). Shared by
// K_tm_op / K_interval_op and by the host replay of the known-answer vectors (tests/tools/hostkat.cpp), so that the CPU
// suite pins the same code the device runs.
#pragma once
#include "ab200_events.cuh"
#include "ab200_series.cuh"

namespace ab200 { inline namespace AB_VARIANT {

AB_HDN TM kat_tm_op(Ctx& c, int op, const TM& x, const TM& y, int k, int flag, double s, Ivl iv, Ivl* range_out) {
    TM r = tm_begin(c, x.as, acc_default());
    int st = ST_OK;
    switch (op) {
    case 0: r = tm_add(c, x, y); break;                                            // operator+(TaylorModel,TaylorModel)
    case 1: r = tm_sub(c, x, y); break;
    case 2: r = tm_mul(c, x, y); break;
    case 3: r = tm_copy(c, x); tm_acc(c, r, s); break;                              // x+Float: _acc
    case 4: r = tm_copy(c, x); tm_acc(c, r, -s); break;                             // x-Float
    case 5: r = tm_copy(c, x); tm_scal(c, r, s); break;                             // x*Float: _scal
    case 6: r = tm_copy(c, x); tm_scal(c, r, ivl_div(ivl(1.0), s)); break;          // x/Float: _scal(Interval(1)/c)
    case 7: r = tm_copy(c, x); tm_acc(c, r, iv); break;
    case 8: r = tm_copy(c, x); tm_acc(c, r, ivl_neg(iv)); break;
    case 9: r = tm_copy(c, x); tm_scal(c, r, iv); break;
    case 10: r = tm_copy(c, x); tm_scal(c, r, ivl_div(1.0, iv, &st)); break;        // x/Interval: _scal(1/c)
    case 11: r = tm_neg(c, x); break;
    case 12: if (!tm_intersection(c, x, y, r)) fail(c, ST_DOMAIN); break;
    case 13: r = tm_split(c, x, k, flag != 0); break;
    case 14: r = tm_copy(c, x); tm_antidiff(c, r, k); break;
    case 15: r = tm_rec(c, x); break;
    case 16: r = tm_sqrt(c, x); break;
    case 17: r = tm_abs(c, x); break;
    case 18: r = tm_max(c, x, y); break;
    case 19: *range_out = tm_range(c, x); r = tm_copy(c, x); break;
    case 20: r = tm_exp(c, x); break;
    case 21: r = tm_log(c, x); break;
    case 22: r = tm_sincos(c, x, true); break;
    case 23: r = tm_sincos(c, x, false); break;
    default: fail(c, ST_UNSUPPORTED);
    }
    if (st != ST_OK) fail(c, st);
    return r;
}

AB_HD Ivl kat_interval_op(int op, int m, Ivl a, Ivl b, int* st) {
    Ivl r = ivl(0.0, 0.0);
    switch (op) {
    case 0: r = ivl_add(a, b); break;
    case 1: r = ivl_sub(a, b); break;
    case 2: r = ivl_mul(a, b); break;
    case 3: r = ivl_div(a, b, st); break;
    case 4: r = ivl_add(a, b.l); break;
    case 5: r = ivl_sub(a, b.l); break;
    case 6: r = ivl_mul(a, b.l); break;
    case 7: r = ivl_div(a, b.l); break;
    case 8: r = ivl_sub(a.l, b); break;
    case 9: r = ivl_div(a.l, b, st); break;
    case 10: r = ivl_sqr(a); break;
    case 11: r = ivl_sqrt(a, st); break;
    case 12: r = ivl_pow(a, (unsigned)m); break;
    case 13: r = ivl_rec(a, st); break;
    case 14: r = ivl_abs(a); break;
    case 15: r = ivl_neg(a); break;
    case 16: r = ivl_max(a, b); break;
    case 17: r = ivl_min(a, b); break;
    case 18: r = ivl_exp(a); break;
    case 19: r = ivl_log(a, st); break;
    case 20: r = ivl_sin(a, RN); break;
    case 21: r = ivl_cos(a, RN); break;
    default: *st = ST_UNSUPPORTED;
    }
    return r;
}

} } // namespace ab200
