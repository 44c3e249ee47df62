// ab200_series.cuh -- series-based TaylorModel functions.
//   rec(x)   src/taylor_model.cpp:1882-1920 (geometric series about the range centre)
//   sqrt(x)  src/taylor_model.cpp:1840-1880   -- not built yet
//   abs/max  src/taylor_model.cpp:1486-1527   -- not built yet
#pragma once
#include "ab200_calc.cuh"

namespace ab200 { inline namespace AB_VARIANT {

// rec(x). For a constant model (every "/ constant" of the configs) eps is exactly 0 and the series has one term, so
// no libm value enters the result. For eps>0 the series degree and the truncation bound use log()/pow(), which the
// reference takes from glibc; the device values can differ in the last bit, so that branch is flagged unpinned.
AB_HDN TM tm_rec(Ctx& c, const TM& x) {
    const int m0 = c.A.top;
    Ivl r = tm_range(c, x);
    if (r.u >= 0 && r.l <= 0) { fail(c, ST_DIVZERO); return tm_copy(c, x); }
    double a = div_m(c.rnd, add_m(c.rnd, r.l, r.u), 2.0);
    double eps = dabs(div_ru(sub_ru(r.u, r.l), add_ru(r.u, r.l)));
    c.rnd = RN;
    unsigned d;
    if (eps == 0.0) d = 1;
    else d = (unsigned)(log(mul_rn(sub_rn(1.0, eps), x.acc.sweep)) / log(eps)) + 1;
    // y=1-(x/a)
    TM xa = tm_copy(c, x);
    tm_scal(c, xa, ivl_div(ivl(1.0), a));
    TM y = tm_neg(c, xa);
    tm_acc(c, y, 1.0);
    // z=(-1)^d; z=1+z*y, d times
    TM z = tm_begin(c, x.as, x.acc);
    tm_acc(c, z, (d % 2) ? -1.0 : 1.0);
    for (unsigned i = 0; i < d; ++i) {
        TM p = tm_mul(c, z, y, y.acc);
        tm_acc(c, p, 1.0);
        tm_slide(c, p, z.off);
        z = p;
    }
    double te = (eps == 0.0) ? 0.0 : div_rn(pow(eps, (double)d), sub_rn(1.0, eps));
    z.e = add_ru(te, z.e);
    c.rnd = RN;
    if (a == 0) { fail(c, ST_DIVZERO); }
    tm_scal(c, z, ivl_div(ivl(1.0), a));
    tm_slide(c, z, m0);
    return z;
}

AB_HDN TM tm_sqrt(Ctx& c, const TM& x) { fail(c, ST_UNSUPPORTED); return tm_copy(c, x); }
AB_HDN TM tm_max(Ctx& c, const TM& x, const TM& y) { fail(c, ST_UNSUPPORTED); (void)y; return tm_copy(c, x); }

} } // namespace ab200
