// ab200_series.cuh -- series-based TaylorModel functions.
//   rec(x)   src/taylor_model.cpp:1882-1920 (geometric series about the range centre)
//   sqrt(x)  src/taylor_model.cpp:1840-1880 (binomial series about the range centre, include/series.h:115-125)
//   abs/max  src/taylor_model.cpp:1486-1527 (range tests, else the reference's fixed degree-12 even polynomial +-0.035)
#pragma once
#include "ab200_calc.cuh"

namespace ab200 { inline namespace AB_VARIANT {

// log() and pow() are glibc's in the reference and CUDA's here; both are accurate to a few ulps but not bit-identical. What
// enters the result is (1) the series degree d = unsigned(log(a)/log(eps) [+1]) and (2) the truncation bound pow(eps,d)/..., which
// is added to an error term many orders larger. Each call CERTIFIES that neither can change with the last bits: the degree is
// recomputed with the quotient moved by 16 ulps either way, the error sum with the bound moved by 8 ulps either way (CUDA pow
// <= 2 ulp, glibc < 1 ulp in any rounding mode, plus the division and product after it). If either differs the call is counted
// in Ctx::libm_sensitive and reported in ab200_status.libm_sensitive: a lap that reports 0 is bit-exact by construction, not
// by luck.
AB_HD unsigned series_degree(Ctx& c, double num_log_arg, double eps, double plus_inside, unsigned plus_outside) {
    const double q = log(num_log_arg) / log(eps);
    const unsigned d = (unsigned)(q + plus_inside) + plus_outside;
    const double w = dabs(q) * 3.6e-15;                                  // 16 ulps of the quotient
    if ((unsigned)(q - w + plus_inside) + plus_outside != d || (unsigned)(q + w + plus_inside) + plus_outside != d) c.libm_sensitive++;
    return d;
}
AB_HD void series_certify_sum(Ctx& c, int mode, double err, double te, double sum) {
    const double lo = te * (1.0 - 1.8e-15), hi = te * (1.0 + 1.8e-15);    // 8 ulps of the bound
    if (add_m(mode, err, lo) != sum || add_m(mode, err, hi) != sum) c.libm_sensitive++;
}

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
    else d = series_degree(c, mul_rn(sub_rn(1.0, eps), x.acc.sweep), eps, 0.0, 1u);
    // y=1-(x/a)
    TM xa = tm_copy(c, x);
    tm_scal(c, xa, ivl_div(ivl(1.0), a));
    TM y = tm_neg(c, xa);
    tm_acc(c, y, 1.0);
    // z=(-1)^d; z=1+z*y, d times
    TM z = tm_begin(c, x.as, x.acc);
    tm_acc(c, z, (d % 2) ? -1.0 : 1.0);
    #pragma unroll 1
    for (unsigned i = 0; i < d; ++i) {
        TM p = tm_mul(c, z, y, y.acc);
        tm_acc(c, p, 1.0);
        tm_slide(c, p, z.off);
        z = p;
    }
    double te = (eps == 0.0) ? 0.0 : div_rn(pow(eps, (double)d), sub_rn(1.0, eps));
    { const double ze = z.e; z.e = add_ru(te, ze); if (eps != 0.0) series_certify_sum(c, RU, ze, te, z.e); }
    c.rnd = RN;
    if (a == 0) { fail(c, ST_DIVZERO); }
    tm_scal(c, z, ivl_div(ivl(1.0), a));
    tm_slide(c, z, m0);
    return z;
}

// sqrt(x). The series degree d and the truncation bound use log()/pow() (glibc in the reference); on the device these
// may differ in the last bit. d only changes if the quotient of logs sits within an ulp of an integer, and the
// truncation bound (~sweep threshold) is added to an error many orders larger, so results coincide in practice; the
// parity tests for sqrt systems therefore compare cell sets exactly and boxes to 1e-12 relative width.
#define AB_MAXSER 96
AB_HDN TM tm_sqrt(Ctx& c, const TM& x) {
    const int m0 = c.A.top;
    Ivl r = tm_range(c, x);
    if (r.l <= 0) { fail(c, ST_DOMAIN); return tm_copy(c, x); }
    double a = div_m(c.rnd, add_m(c.rnd, r.l, r.u), 2.0);
    double eps = div_ru(sub_ru(r.u, r.l), add_ru(r.u, r.l));
    c.rnd = RN;
    unsigned d;
    if (eps == 0.0) d = 1;
    else d = series_degree(c, mul_rn(sub_rn(1.0, eps), x.acc.sweep), eps, 1.0, 0u);
    if (d + 1 > AB_MAXSER) { fail(c, ST_UNSUPPORTED); return tm_copy(c, x); }
    // y=(x/a)-1
    TM y = tm_copy(c, x);
    tm_scal(c, y, ivl_div(ivl(1.0), a));
    tm_acc(c, y, -1.0);
    // Series<Interval>::sqrt(d,Interval(1))
    Ivl ser[AB_MAXSER];
    int st = ST_OK;
    ser[0] = ivl_sqrt(ivl(1.0), &st);
    Ivl mhr = ivl_div(-0.5, ivl(1.0), &st);
    #pragma unroll 1
    for (unsigned i = 1; i <= d; ++i) ser[i] = ivl_mul(ivl_div(ivl_mul(mhr, (double)(2 * (int)i - 3)), (double)i), ser[i - 1]);
    TM z = tm_begin(c, x.as, x.acc);
    tm_acc(c, z, ser[d - 1]);
    #pragma unroll 1
    for (unsigned i = 0; i < d; ++i) {
        TM p = tm_mul(c, z, y, y.acc);
        tm_acc(c, p, ser[d - i - 1]);
        tm_clean(c, p, p.acc);
        tm_slide(c, p, z.off);
        z = p;
    }
    double te = (eps == 0.0) ? 0.0 : mul_m(c.rnd, div_m(c.rnd, pow(eps, (double)d), sub_m(c.rnd, 1.0, eps)), ivl_mag(ser[d]));
    { const double ze = z.e; z.e = add_m(c.rnd, ze, te); if (eps != 0.0) series_certify_sum(c, c.rnd, ze, te, z.e); }
    tm_scal(c, z, ivl_sqrt(ivl(a), &st));
    if (st != ST_OK) fail(c, st);
    tm_slide(c, z, m0);
    return z;
}
// ---- exp / log / sin / cos of a model (src/taylor_model.cpp:1922-2060) ------------------------------------------------
// All working models carry x's accuracy (they are built with x.accuracy_ptr(), copies keep it, x*y takes the right operand's).
AB_HD double tm_value(Ctx& c, const TM& x) { return (x.n > 0 && key_deg(tm_key(c, x, 0)) == 0) ? tm_val(c, x, 0) : 0.0; }
// rec_fac_up (src/taylor_model.cpp:1837): 1/n! rounded upward; leaves the ambient mode upward
AB_HD double rec_fac_up(Ctx& c, unsigned n) { double r = 1.0; for (unsigned i = 1; i <= n; ++i) r = div_ru(r, (double)i); c.rnd = RU; return r; }
// _scal_exact (src/taylor_model.cpp:229-238): coefficients and error times a power of two, in the ambient mode
AB_HDN void tm_scal_exact(Ctx& c, TM& r, double s) {
    const int W = AB_W, lane = wlane(), n = r.n, m = c.rnd;
    wsync();
    #pragma unroll 1
    for (int i = lane; i < n; i += W) tm_set_val(c, r, i, mul_m(m, tm_val(c, r, i), s));
    wsync();
    r.e = mul_m(m, r.e, dabs(s));
}
// exp(x) :1949-1988 -- exp(ax+b)=exp(x)^a*exp(b): centre, scale into the unit interval, Horner, square back
AB_HDN TM tm_exp(Ctx& c, const TM& x) {
    const int m0 = c.A.top;
    TM y = tm_copy(c, x);
    const double xval = tm_value(c, x);
    tm_acc(c, y, -xval);
    const double xrad = ivl_mag(tm_range(c, y));
    unsigned sfp = 0;
    while (sfp < 30 && (double)(1 << sfp) < xrad) ++sfp;
    if (sfp >= 30) { fail(c, ST_UNSUPPORTED); return y; }
    const double sf = div_m(c.rnd, 1.0, (double)(1 << sfp));
    tm_scal_exact(c, y, sf);
    const double yrad = mul_m(c.rnd, xrad, sf);
    const int degree = x.acc.maxdeg;
    const int P = c.A.top;
    TM res = tm_begin(c, x.as, x.acc);
    res.e = pow_m(RU, yrad, (unsigned)(degree + 1));
    #pragma unroll 1
    for (int i = 0; i < degree && c.status == ST_OK; ++i) {
        tm_scal(c, res, ivl_div(ivl(1.0), (double)(degree - i)));          // res/=(degree-i)
        TM t = tm_mul(c, y, res, res.acc);                                 // res=y*res+1.0
        tm_acc(c, t, 1.0);
        tm_slide(c, t, P);
        res = t;
    }
    #pragma unroll 1
    for (unsigned i = 0; i < sfp && c.status == ST_OK; ++i) {              // _mul(square,res,res); res.swap(square)
        TM t = tm_mul(c, res, res, x.acc);
        tm_slide(c, t, P);
        const Acc a = res.acc; res = t; res.acc = a;
    }
    tm_scal(c, res, ivl_exp(ivl(xval)));                                   // res*=exp(Interval(xval))
    tm_slide(c, res, m0);
    return res;
}
// sin(x) / cos(x) :1990-2060 -- reduce by a multiple of 2 pi, then x*f(x^2) resp. f(x^2) by Horner with the factorial ratios
AB_HDN TM tm_sincos(Ctx& c, const TM& x, bool want_sin) {
    const int m0 = c.A.top;
    const double two_pi = 2 * AB_PI_APPROX;
    const int n = (int)floor(add_m(c.rnd, div_m(c.rnd, tm_value(c, x), two_pi), 0.5));
    TM y = tm_copy(c, x);
    // sin: y=x-(n*2*pi<Interval>()); cos: y=x-2*n*pi<Interval>() -- both (2n)*[pi_down,pi_up], then _acc(y,-that)
    tm_acc(c, y, ivl_neg(ivl_mul(ivl(AB_PI_DOWN, AB_PI_UP), (double)(2 * n))));
    bool wide = y.e > two_pi / 2;
    if (!wide) { const double mg = ivl_mag(tm_range(c, y)); const double rf = rec_fac_up(c, (unsigned)x.acc.maxdeg); wide = mul_m(c.rnd, mg, rf) > 1; }
    if (wide) {
        c.A.top = m0;
        TM r = tm_begin(c, x.as, x.acc); r.e = 1.0;
        return r;
    }
    TM s = tm_mul(c, y, y, x.acc);                                         // _mul(s,y,y)
    const int d = (x.acc.maxdeg + 3) / 2;
    const double srad = ivl_mag(tm_range(c, s));
    const double pw = pow_m(RU, srad, (unsigned)(d + 1));
    const double rf = rec_fac_up(c, (unsigned)((d + 1) * 2));
    const double truncation_error = mul_m(c.rnd, pw, rf);
    const int P = c.A.top;
    TM r = tm_const(c, x.as, 1.0, x.acc);                                  // r=1.0
    #pragma unroll 1
    for (int i = 0; i < d && c.status == ST_OK; ++i) {
        const int k = d - i;
        const double dv = want_sin ? (double)(-2 * k * (2 * k + 1)) : (double)(-2 * k * (2 * k - 1));
        tm_scal(c, r, ivl_div(ivl(1.0), dv));                              // r/=...
        TM t = tm_mul(c, s, r, x.acc);                                     // _mul(t,s,r); r.swap(t)
        tm_acc(c, t, 1.0);                                                 // r+=1.0
        tm_slide(c, t, P);
        r = t;
    }
    if (want_sin) { TM t = tm_mul(c, y, r, x.acc); tm_slide(c, t, P); r = t; }   // _mul(t,y,r); r.swap(t)
    r.e = add_m(c.rnd, r.e, truncation_error);
    tm_slide(c, r, m0);
    return r;
}
// log(x) :1922-1946 -- series of log(1+y), y=x/a-1 about the range centre a
AB_HDN TM tm_log(Ctx& c, const TM& x) {
    const int m0 = c.A.top;
    const Ivl rg = tm_range(c, x);
    if (rg.l <= 0) { fail(c, ST_DOMAIN); return tm_copy(c, x); }
    const double a = div_m(c.rnd, add_m(c.rnd, rg.l, rg.u), 2.0);
    const double eps = div_ru(sub_ru(rg.u, rg.l), add_ru(rg.u, rg.l));
    c.rnd = RN;
    if (!(eps > 0.0)) { fail(c, ST_UNSUPPORTED); return tm_copy(c, x); }   // a constant model: log(0) in the degree formula
    const unsigned d = series_degree(c, mul_rn(sub_rn(1.0, eps), x.acc.sweep), eps, 1.0, 0u);
    if (d < 1 || d > 4096) { fail(c, ST_UNSUPPORTED); return tm_copy(c, x); }
    TM y = tm_copy(c, x);                                                  // y=x/a-1
    tm_scal(c, y, ivl_div(ivl(1.0), a));
    tm_acc(c, y, -1.0);
    const int P = c.A.top;
    TM z = tm_begin(c, x.as, x.acc);
    tm_acc(c, z, div_m(c.rnd, (d % 2) ? -1.0 : 1.0, (double)d));           // z+=Float(d%2?-1:+1)/d
    #pragma unroll 1
    for (unsigned i = 1; i != d && c.status == ST_OK; ++i) {
        TM t = tm_mul(c, z, y, y.acc);                                     // z*y (evaluated before the Float, right to left)
        tm_acc(c, t, div_m(c.rnd, ((d - i) % 2) ? 1.0 : -1.0, (double)(d - i)));
        tm_clean(c, t, t.acc);
        tm_slide(c, t, P);
        z = t;
    }
    { TM t = tm_mul(c, z, y, y.acc); tm_slide(c, t, P); z = t; }           // z=z*y
    tm_sweep(c, z, z.acc.sweep);
    const double te = div_m(c.rnd, div_m(c.rnd, pow(eps, (double)d), sub_m(c.rnd, 1.0, eps)), (double)d);
    tm_acc(c, z, ivl(log_m(RD, a), log_m(RU, a)));                         // z+log(Interval(a))
    { const double ze = z.e; tm_acc(c, z, ivl_mul(ivl(-1.0, 1.0), te)); series_certify_sum(c, RU, ze, te, z.e); }   // +trunc_err*Interval(-1,1)
    tm_slide(c, z, m0);
    return z;
}

// abs(x) src/taylor_model.cpp:1503-1527. Result at the arena top recorded on entry.
AB_HDN TM tm_abs(Ctx& c, const TM& x) {
    const int m0 = c.A.top;
    Ivl xr = tm_range(c, x);
    if (xr.l >= 0.0) return tm_copy(c, x);
    if (xr.u <= 0.0) return tm_neg(c, x);
    const double p[7] = {0.0112167620474, 5.6963263292747541, -31.744583789655049, 100.43002481377681, -162.01366698662306, 127.45243493284417, -38.829743345344667};
    const double err = 0.035;
    double xmag = ivl_mag(xr);
    TM s0 = tm_copy(c, x);
    tm_scal(c, s0, ivl_div(ivl(1.0), xmag));                    // x/xmag
    TM s = tm_mul(c, s0, s0, s0.acc);                           // sqr(s)
    TM r = tm_begin(c, x.as, acc_default());
    tm_append(c, r, key_zero(), p[6]);
    #pragma unroll 1
    for (int i = 0; i < 6; ++i) {
        int j = 5 - i;
        TM t = tm_mul(c, s, r, r.acc);
        tm_acc(c, t, p[j]);
        tm_slide(c, t, r.off);
        r = t;
    }
    tm_acc(c, r, ivl(-err, err));
    tm_scal(c, r, xmag);
    tm_slide(c, r, m0);
    return r;
}
// max(x,y) src/taylor_model.cpp:1486-1496. Result at the arena top recorded on entry.
AB_HDN TM tm_max(Ctx& c, const TM& x, const TM& y) {
    const int m0 = c.A.top;
    Ivl xr = tm_range(c, x), yr = tm_range(c, y);
    if (xr.l >= yr.u) return tm_copy(c, x);
    if (yr.l >= xr.u) return tm_copy(c, y);
    TM sum = tm_add(c, x, y);
    TM dif = tm_sub(c, x, y);
    TM ad = tm_abs(c, dif);
    TM r = tm_add(c, sum, ad);
    tm_scal(c, r, ivl_div(ivl(1.0), 2.0));                      // .../2
    tm_slide(c, r, m0);
    return r;
}

} } // namespace ab200
