// ab200_elem.cuh -- exp / log / sin / cos on doubles in a given rounding mode and on intervals.
//
// The reference computes these itself, without libm: rounded Horner evaluations of short series after an argument reduction,
// written in plain C arithmetic that runs in whatever MXCSR mode the caller set (src/numeric.cpp:191-225 horner / next,
// :287-334 exp_rnd, :339-381 log_rnd, :442-506 cos_rnd, :530-586 the sin / cos series; exp_up = exp_rnd under upward mode
// etc., :690-723). Here the mode is a parameter and every C operation becomes the matching rounding intrinsic in the same
// order, so the results are bit-identical. "opp" forms compute with the operands negated, exactly as the reference does.
#pragma once
#include "ab200_fp.cuh"

namespace ab200 { inline namespace AB_VARIANT {

#define AB_PI_UP 3.1415926535897936
#define AB_PI_DOWN 3.1415926535897931
#define AB_PI_APPROX 3.1415926535897931
#define AB_QUARTER_PI_APPROX 0.78539816339744828
#define AB_SQRT2_APPROX 0.70710678118654757
#define AB_LOG2_APPROX 0.6931471805599453094

AB_HD double el_next_rnd(int m, double x) { double y = x; y = add_m(m, y, 1e-300); y = sub_m(m, y, 1e-300); return y; }
AB_HD double el_next_opp(int m, double x) { double y = -x; y = add_m(m, y, 1e-300); y = sub_m(m, y, 1e-300); return -y; }
AB_HD double el_pi_rnd(int m) { return m == RD ? AB_PI_DOWN : (m == RU ? AB_PI_UP : AB_PI_APPROX); }
AB_HD double el_pi_opp(int m) { return m == RD ? AB_PI_UP : (m == RU ? AB_PI_DOWN : AB_PI_APPROX); }
// y=1./c[n]; y=1.0/c[i]+x*y  (the coefficients are the integer reciprocals' denominators)
AB_HD double el_horner_rnd(int m, int n, double x, const double* c) {
    double y = div_m(m, 1.0, c[n]);
    #pragma unroll 1
    for (int i = n - 1; i >= 0; --i) y = add_m(m, div_m(m, 1.0, c[i]), mul_m(m, x, y));
    return y;
}
AB_HD double el_horner_opp(int m, int n, double x, const double* c) {
    double y = div_m(m, -1.0, c[n]);
    #pragma unroll 1
    for (int i = n - 1; i >= 0; --i) y = add_m(m, div_m(m, -1.0, c[i]), mul_m(m, x, y));
    return -y;
}
// sub_opp / mul_opp / div_opp (include/numeric.h:150-152)
AB_HD double el_sub_opp(int m, double x, double y) { return -add_m(m, -x, y); }
AB_HD double el_mul_opp(int m, double x, double y) { return -mul_m(m, -x, y); }
AB_HD double el_div_opp(int m, double x, double y) { return -div_m(m, -x, y); }

// exp_rnd (src/numeric.cpp:287-334)
AB_HDN double exp_m(int m, double x) {
    const double c[7] = { 1.0, 6.0, -360.0, 15120.0, -604800.0, 23950080.0, -946218790.0 };
    if (x == 0.0) return 1.0;
    const double n = floor(add_m(m, div_m(m, x, AB_LOG2_APPROX), 0.5));
    const double log2 = (n > 0.0) ? el_next_opp(m, AB_LOG2_APPROX) : el_next_rnd(m, AB_LOG2_APPROX);
    const double r = add_m(m, x, mul_m(m, -n, log2));
    double s, w;
    if (r < 0) { s = mul_m(m, r, r); w = add_m(m, 1.0, el_horner_rnd(m, 6, s, c)); }
    else { s = mul_m(m, -r, r); s = -s; w = el_horner_opp(m, 6, s, c); w = sub_m(m, -1.0, w); w = -w; }
    const double t = sub_m(m, r, w);
    const double y = div_m(m, add_m(m, w, r), -t);
    int e = 0;
    double z = frexp(y, &e);
    return ldexp(z, (int)(long)n + e);
}
// log_rnd (src/numeric.cpp:339-381)
AB_HDN double log_m(int m, double x) {
    const double c[12] = { 1.0, 3.0, 5.0, 7.0, 9.0, 11.0, 13.0, 15.0, 17.0, 19.0, 21.0, 23.0 };
    if (x == 1.0) return 0.0;
    int n = 0;
    double y = frexp(x, &n);
    if (y < AB_SQRT2_APPROX) { y = mul_m(m, y, 2.0); --n; }
    double t, z, s, w, ly;
    if (y >= 1.0) {
        t = sub_m(m, -1.0, y);
        z = div_m(m, sub_m(m, y, 1.0), -t);
        s = mul_m(m, z, z);
        w = el_horner_rnd(m, 10, s, c);
        ly = mul_m(m, mul_m(m, 2.0, z), w);
    } else {
        t = sub_m(m, 1.0, y);
        z = div_m(m, -t, add_m(m, 1.0, y));
        s = mul_m(m, -z, z); s = -s;
        w = el_horner_opp(m, 10, s, c);
        ly = mul_m(m, mul_m(m, 2.0, z), w);
    }
    const double log2rnd = el_next_rnd(m, AB_LOG2_APPROX);
    return add_m(m, mul_m(m, log2rnd, (double)n), ly);
}
// the series on [0, pi/4] (src/numeric.cpp:530-586)
AB_HD double el_pos_sin(int m, double x) {
    const double c[9] = { 1.0, -6.0, 120.0, -5040.0, 362880.0, -39916800.0, 6227020800.0, -1307674368000.0, 355687428096000.0 };
    const double s = mul_m(m, x, x), w = el_horner_rnd(m, 8, s, c);
    return mul_m(m, x, w);
}
AB_HD double el_neg_sin(int m, double x) {
    const double c[9] = { 1.0, -6.0, 120.0, -5040.0, 362880.0, -39916800.0, 6227020800.0, -1307674368000.0, 355687428096000.0 };
    double s = mul_m(m, -x, x); s = -s;
    const double w = el_horner_opp(m, 8, s, c);
    return mul_m(m, x, -w);
}
AB_HD double el_pos_cos(int m, double x) {
    const double c[9] = { 1.0, -2.0, 24.0, -720.0, 40320.0, -3628800.0, 479001600.0, -87178291200.0, 20922789888000.0 };
    double s = mul_m(m, -x, x); s = -s;
    return el_horner_rnd(m, 8, s, c);
}
AB_HD double el_neg_cos(int m, double x) {
    const double c[9] = { 1.0, -2.0, 24.0, -720.0, 40320.0, -3628800.0, 479001600.0, -87178291200.0, 20922789888000.0 };
    const double s = mul_m(m, x, x);
    return -el_horner_opp(m, 8, s, c);
}
// cos_rnd (src/numeric.cpp:442-506)
AB_HDN double cos_m(int m, double x) {
    const double pi_rnd = el_pi_rnd(m), pi_opp = el_pi_opp(m);
    if (x == 0.0) return 1.0;
    const long n_rnd = (long)floor(div_m(m, x, pi_opp));
    const long n_opp = (long)floor(el_div_opp(m, x, pi_rnd));
    if (n_rnd != n_opp) {
        if (n_rnd > n_opp) {
            if (n_rnd % 2 == 0) return 1.0;
            const double y1 = sub_m(m, mul_m(m, pi_rnd, (double)n_rnd), x);
            const double y2 = sub_m(m, x, mul_m(m, pi_opp, (double)n_rnd));
            return el_neg_cos(m, dmax(y1, y2));
        } else {
            if (n_rnd % 2 == 0) return -1.0;
            const double y1 = el_sub_opp(m, mul_m(m, pi_opp, (double)n_opp), x);
            const double y2 = el_sub_opp(m, x, mul_m(m, pi_rnd, (double)n_opp));
            return el_pos_cos(m, dmax(y1, y2));
        }
    }
    double y;
    if (n_rnd % 2 == 0) y = el_sub_opp(m, x, mul_m(m, (double)n_rnd, pi_rnd));
    else y = el_sub_opp(m, el_mul_opp(m, (double)(n_rnd + 1), pi_opp), x);
    const int q = (int)((long)floor(div_m(m, y, AB_QUARTER_PI_APPROX)) % 8);
    if (q == 0) return el_pos_cos(m, y);
    if (q == 1 || q == 2) {
        const double w = sub_m(m, pi_rnd / 2, y);
        return (w >= 0.0) ? el_pos_sin(m, w) : el_neg_sin(m, -w);
    }
    const double w = el_sub_opp(m, pi_opp, y);          // q == 3 || q == 4
    return el_neg_cos(m, w);
}

// Interval versions (src/numeric.cpp:912-966). `amb` is the ambient rounding mode of the caller (the plain double
// arithmetic of cos runs in it).
AB_HD Ivl ivl_exp(Ivl i) { return ivl(exp_m(RD, i.l), exp_m(RU, i.u)); }
AB_HD Ivl ivl_log(Ivl i, int* st) { if (!(i.l > 0.0)) { if (st && *st == ST_OK) *st = ST_DOMAIN; return i; } return ivl(log_m(RD, i.l), log_m(RU, i.u)); }
AB_HDN Ivl ivl_cos(Ivl i, int amb) {
    const Ivl pi = ivl(AB_PI_DOWN, AB_PI_UP);
    if (ivl_rad(i) > 2 * AB_PI_DOWN) return ivl(-1.0, 1.0);
    const double n = floor(add_m(amb, div_m(amb, i.l, 2 * AB_PI_APPROX), 0.5));
    i = ivl_sub(i, ivl_mul(pi, mul_m(amb, 2.0, n)));
    double rl, ru;
    if (i.l <= 0.0) {
        if (i.u <= 0.0) { rl = cos_m(RD, i.l); ru = cos_m(RU, i.u); }
        else if (i.u <= AB_PI_DOWN) { rl = cos_m(RD, dmax(-i.l, i.u)); ru = 1.0; }
        else return ivl(-1.0, 1.0);
    } else {
        if (i.u <= AB_PI_DOWN) { rl = cos_m(RD, i.u); ru = cos_m(RU, i.l); }
        else if (i.u <= 2 * AB_PI_DOWN) { rl = -1.0; ru = cos_m(RU, dmin(i.l, sub_rd(2 * AB_PI_DOWN, i.u))); }
        else return ivl(-1.0, 1.0);
    }
    return ivl(rl, ru);
}
AB_HD Ivl ivl_sin(Ivl i, int amb) { return ivl_cos(ivl_sub(i, ivl_div(ivl(AB_PI_DOWN, AB_PI_UP), 2.0)), amb); }

} } // namespace ab200
