// [lane build] One frontier cell per LANE: the throughput build of the evolution kernel, used only for frontiers far larger
// than the device can hold one-warp-per-cell (ab200_capi.cu picks it above AB_LANE_THRESHOLD cells). Same algorithm and
// arithmetic as the warp build one directory up, written as sequential per-lane code over lane-interleaved scratch arenas.
// ab200_calc.cuh -- op-tape evaluation, the flow bounder and the Taylor-model calculus of one time step.
//
// Reference functions restated here (file:line under /root/reference):
//   expression evaluation            src/expression.cpp:656-692,766-779; src/function.cpp:131-134
//   flow_bounds (the "Bounder")      src/taylor_calculus.cpp:55-133,584-600
//   flow_model / unchecked_flow      src/taylor_calculus.cpp:630-666; src/taylor_function.cpp:1517-1553;
//                                    src/taylor_model.cpp:3719-3783
//   _compose2                        src/taylor_model.cpp:3187-3240
//   _split1 / partial_evaluate       src/taylor_model.cpp:2497-2544, 2387-2453
//   reachability_step / integration_step   src/taylor_calculus.cpp:234-267
//   uniform_error_recondition / kuhn_recondition / recondition
//                                    src/taylor_set.cpp:499-577; src/taylor_model.cpp:3882-3956
#pragma once
#include "ab200_tm.cuh"

namespace ab200 { inline namespace AB_VARIANT {

#ifdef AB_DEBUG_DUMP
void ab_debug_dump(Ctx& c, const char* name, int idx, const TM& t);   // provided by tests/tools/hoststep.cpp
#define AB_DUMP(c, name, idx, t) ab_debug_dump(c, name, idx, t)
#define AB_DBG(...) printf(__VA_ARGS__)
#else
#define AB_DUMP(c, name, idx, t)
#define AB_DBG(...)
#endif

// ---- system description (flat op tapes) -------------------------------------------------------------------------
enum {
    OP_CNST = 0, OP_VAR = 1, OP_ADD = 2, OP_SUB = 3, OP_MUL = 4, OP_DIV = 5, OP_NEG = 6, OP_POS = 7, OP_REC = 8,
    OP_SQR = 9, OP_SQRT = 10, OP_EXP = 11, OP_LOG = 12, OP_SIN = 13, OP_COS = 14, OP_TAN = 15, OP_ABS = 16,
    OP_MAX = 17, OP_MIN = 18
};
struct TapeOp {
    int32_t op;
    int32_t arg;      // OP_VAR: variable index
    double lo, hi;    // OP_CNST: interval literal (CNST and PAR leaves)
};
struct Func { int32_t begin, end; };     // postfix slice of the op array

#define AB_MAXN 5          // continuous dimension (springs has 5)
#define AB_MAXEV 8         // events per location
#define AB_MAXLOC 16
#define AB_MAXOPS 1536
#define AB_STACK 24

enum { EV_INVARIANT = 0, EV_URGENT = 1, EV_PERMISSIVE = 2 };

struct EventDesc {
    int32_t kind;
    int32_t target;             // location index
    int32_t id;                 // event id (ordering of std::map<DiscreteEvent,...>)
    int32_t pad;
    Func guard;                 // guard or invariant function
    Func lie;                   // lie_derivative(guard,dynamic) (transitions only; grid-side jump test)
    Func reset[AB_MAXN];
};
struct LocDesc {
    int32_t n;
    int32_t n_events;
    int32_t finishing_rank;     // number of events whose name sorts before the finishing event "e-2" (std::map order)
    int32_t pad0;
    Func dyn[AB_MAXN];
    Func dyn_back[AB_MAXN];     // get_directed_dynamic(dynamic,DIRECTION_BACKWARD) = dynamic*constant(-1) (include/evolver_base.h:85-90)
    EventDesc ev[AB_MAXEV];
    double max_step;            // getHybridMaximumStepSize (src/hybrid_evolver-image.cpp:1522-1555)
    double ref_width[AB_MAXN];  // getMinimumGridCellWidths (:1506-1519)
    double grid_origin[AB_MAXN];
    double grid_length[AB_MAXN];
};
struct SystemDesc {
    int32_t n_locs;
    int32_t n_ops;
    LocDesc loc[AB_MAXLOC];
    TapeOp ops[AB_MAXOPS];
};

// ---- interval tape --------------------------------------------------------------------------------------------
AB_HDN Ivl eval_ivl(Ctx& c, const TapeOp* ops, Func f, const Ivl* x) {
    Ivl st[AB_STACK];
    int sp = 0;
    for (int p = f.begin; p < f.end; ++p) {
        const TapeOp o = ops[p];
        if (o.op != OP_CNST && o.op != OP_VAR && o.op != OP_POS && o.op != OP_NEG) AB_FLOPS(o.op == OP_MUL ? 4 : 2);
        switch (o.op) {
            case OP_CNST: st[sp++] = ivl(o.lo, o.hi); break;     // x[0]*0 then r=c (src/expression.cpp:581,773)
            case OP_VAR: st[sp++] = x[o.arg]; break;
            case OP_ADD: --sp; st[sp - 1] = ivl_add(st[sp - 1], st[sp]); break;
            case OP_SUB: --sp; st[sp - 1] = ivl_sub(st[sp - 1], st[sp]); break;
            case OP_MUL: --sp; st[sp - 1] = ivl_mul(st[sp - 1], st[sp]); break;
            case OP_DIV: --sp; st[sp - 1] = ivl_div(st[sp - 1], st[sp], &c.status); break;
            case OP_MAX: --sp; st[sp - 1] = ivl_max(st[sp - 1], st[sp]); break;
            case OP_MIN: --sp; st[sp - 1] = ivl_min(st[sp - 1], st[sp]); break;
            case OP_NEG: st[sp - 1] = ivl_neg(st[sp - 1]); break;
            case OP_POS: break;
            case OP_REC: st[sp - 1] = ivl_div(1.0, st[sp - 1], &c.status); break;
            case OP_SQR: st[sp - 1] = ivl_sqr(st[sp - 1]); break;
            case OP_SQRT: st[sp - 1] = ivl_sqrt(st[sp - 1], &c.status); break;
            case OP_ABS: st[sp - 1] = ivl_abs(st[sp - 1]); break;
            default: fail(c, ST_UNSUPPORTED); break;
        }
    }
    return st[0];
}

// ---- vectors of models ------------------------------------------------------------------------------------------
// Slide models v[0..m) (stored in ascending arena order) down to dst, packed; sets the arena top.
AB_HD void vec_slide(Ctx& c, TM* v, int m, int dst) {
    for (int i = 0; i < m; ++i) { tm_slide(c, v[i], dst); dst = c.A.top; }
    c.A.top = dst;
}

// ---- Taylor-model tape ------------------------------------------------------------------------------------------
struct TmSlot { TM v; int mark; };

AB_HDN TM tm_rec(Ctx& c, const TM& x);
AB_HDN TM tm_sqrt(Ctx& c, const TM& x);
AB_HDN TM tm_max(Ctx& c, const TM& x, const TM& y);
AB_HDN TM tm_abs(Ctx& c, const TM& x);

// Evaluate f over models x[0..nx); the result is left at the arena top position recorded on entry.
AB_HDN TM eval_tm(Ctx& c, const TapeOp* ops, Func f, const TM* x) {
    TmSlot st[AB_STACK];
    int sp = 0;
    int m0 = c.A.top;
    for (int p = f.begin; p < f.end; ++p) {
        const TapeOp o = ops[p];
        if (o.op == OP_CNST) {
            // X r=x[0]*0; r.clear(); r+=c  (src/expression.cpp:773-775,581): accuracy of the first variable
            st[sp].mark = c.A.top;
            TM r = tm_const(c, x[0].as, 0.0, x[0].acc);
            tm_acc(c, r, ivl(o.lo, o.hi));
            st[sp++].v = r;
        } else if (o.op == OP_VAR) {
            st[sp].mark = c.A.top; st[sp++].v = x[o.arg];
        } else if (o.op == OP_ADD || o.op == OP_SUB || o.op == OP_MUL || o.op == OP_DIV || o.op == OP_MAX || o.op == OP_MIN) {
            TM& a = st[sp - 2].v; TM& b = st[sp - 1].v;
            TM r;
            if (o.op == OP_ADD) r = tm_add(c, a, b);
            else if (o.op == OP_SUB) r = tm_sub(c, a, b);
            else if (o.op == OP_MUL) r = tm_mul(c, a, b, b.acc);
            else if (o.op == OP_DIV) {             // x/y: r(fresh); _mul(r,x,rec(y))  src/taylor_model.cpp:1271-1274
                int m = c.A.top;
                TM ry = tm_rec(c, b);
                r = tm_mul(c, a, ry, acc_default());
                tm_slide(c, r, m);
            } else if (o.op == OP_MAX) r = tm_max(c, a, b);
            else {                                  // min(x,y) = -max(-x,-y)  src/taylor_model.cpp:1499-1501
                int m = c.A.top;
                TM na = tm_neg(c, a); TM nb = tm_neg(c, b);
                TM mx = tm_max(c, na, nb);
                r = tm_neg(c, mx);
                tm_slide(c, r, m);
            }
            tm_slide(c, r, st[sp - 2].mark);
            --sp; st[sp - 1].v = r;
        } else {
            TM& a = st[sp - 1].v;
            TM r;
            switch (o.op) {
                case OP_NEG: r = tm_neg(c, a); break;
                case OP_POS: r = a; break;
                case OP_SQR: r = tm_mul(c, a, a, a.acc); break;     // src/taylor_model.cpp:1543-1546
                case OP_REC: {                                      // 1.0/x: r(rec(x)); _scal(r,Interval(1)/c)  :1312
                    r = tm_rec(c, a);
                    tm_scal(c, r, ivl_div(ivl(1.0), 1.0));
                } break;
                case OP_SQRT: r = tm_sqrt(c, a); break;
                case OP_ABS: r = tm_abs(c, a); break;
                default: fail(c, ST_UNSUPPORTED); r = a; break;
            }
            if (o.op != OP_POS) tm_slide(c, r, st[sp - 1].mark);
            st[sp - 1].v = r;
        }
        if (c.status != ST_OK) break;
    }
    TM res = st[0].v;
    if (res.off < m0) res = tm_copy(c, res);     // a bare variable: copy so the caller owns the result
    (void)sp;
    return res;
}

// ---- series-based functions ---------------------------------------------------------------------------------------
// rec(x) src/taylor_model.cpp:1882-1920 and sqrt(x) :1840-1880 are defined in ab200_series.cuh.

// ---- flow bounder -------------------------------------------------------------------------------------------------
// Ariadne::flow_bounds (src/taylor_calculus.cpp:55-133) preceded by the singleton widening of
// TaylorCalculus::flow_bounds (:584-600). Returns h; writes the bounding box to b[].
AB_HDN double flow_bounds(Ctx& c, const TapeOp* ops, const Func* vf, int n, const Ivl* r, double hmax, Ivl* bout) {
    const double INITIAL_MULTIPLIER = 2, MULTIPLIER = 1.125, BOX_RADIUS_MULTIPLIER = 1.03125;
    const int EXPANSION_STEPS = 8, REDUCTION_STEPS = 16, REFINEMENT_STEPS = 8;
    const double EPS = 1e-16;
    Ivl d[AB_MAXN], delta[AB_MAXN], b[AB_MAXN], nb[AB_MAXN], df[AB_MAXN], dd[AB_MAXN];
    for (int i = 0; i < n; ++i) { d[i] = r[i]; if (d[i].l == d[i].u) d[i] = ivl_add(d[i], ivl(-EPS, EPS)); }
    for (int i = 0; i < n; ++i) delta[i] = ivl_mul(ivl_sub(d[i], ivl_mid(d[i])), BOX_RADIUS_MULTIPLIER - 1);
    for (int i = 0; i < n; ++i) dd[i] = ivl_add(d[i], delta[i]);
    double hmin = hmax / (1 << REDUCTION_STEPS);
    double h = hmax;
    bool success = false;
    Ivl ih = ivl(0, h);
    while (!success) {
        if (!(h > hmin)) { fail(c, ST_BOUNDER); return h; }
        Ivl m2 = ivl_mul(ih, INITIAL_MULTIPLIER);
        for (int i = 0; i < n; ++i) df[i] = eval_ivl(c, ops, vf[i], d);
        for (int i = 0; i < n; ++i) b[i] = ivl_add(ivl_add(d[i], ivl_mul(m2, df[i])), delta[i]);
        for (int s = 0; s < EXPANSION_STEPS; ++s) {
            for (int i = 0; i < n; ++i) df[i] = eval_ivl(c, ops, vf[i], b);
            bool bounded = true, sub = true;
            for (int i = 0; i < n; ++i) {
                nb[i] = ivl_add(dd[i], ivl_mul(ih, df[i]));
                if (!(dabs(nb[i].l) <= DBL_MAX && dabs(nb[i].u) <= DBL_MAX)) bounded = false;
                if (!ivl_subset(nb[i], b[i])) sub = false;
            }
            if (!bounded) { success = false; break; }
            else if (sub) { success = true; break; }
            else {
                Ivl mi = ivl_mul(ih, MULTIPLIER);
                for (int i = 0; i < n; ++i) b[i] = ivl_add(dd[i], ivl_mul(mi, df[i]));
            }
        }
        if (!success) { h /= 2; ih = ivl(0, h); }
        if (c.status != ST_OK) return h;
    }
    for (int s = 0; s < REFINEMENT_STEPS; ++s) {
        for (int i = 0; i < n; ++i) b[i] = nb[i];
        for (int i = 0; i < n; ++i) df[i] = eval_ivl(c, ops, vf[i], b);
        for (int i = 0; i < n; ++i) {
            nb[i] = ivl_add(dd[i], ivl_mul(ih, df[i]));
            if (!ivl_subset(nb[i], b[i])) fail(c, ST_BOUNDER);
        }
    }
    for (int i = 0; i < n; ++i) { bout[i] = nb[i]; if (!ivl_subset(d[i], nb[i])) fail(c, ST_BOUNDER); }
    return h;
}

// ---- compose ------------------------------------------------------------------------------------------------------
// Power-table handles are kept in the arena (3 words each) to keep per-thread local memory small.
#define AB_MAXPOW 40
AB_HD void ph_store(Ctx& c, int slot, const TM& t) {
    aw(c, slot) = ((uint64_t)(uint32_t)t.off << 32) | (uint32_t)t.n;
    aw(c, slot + 1) = d2u(t.e);
}
AB_HD TM ph_load(Ctx& c, int slot, int as, Acc acc) {
    TM t; uint64_t w = aw(c, slot); t.off = (int)(w >> 32); t.n = (int)(w & 0xFFFFFFFFu); t.e = u2d(aw(c, slot + 1)); t.as = as; t.acc = acc; return t;
}

// _compose2 (src/taylor_model.cpp:3187-3240): r[i] = x[i] o ys, nx models in xas=nys variables.
// Results are written to r[0..nx) packed from the arena top recorded on entry.
AB_HDN void tm_compose(Ctx& c, const TM* x, int nx, const TM* ys, int nys, TM* r) {
    AB_SECTION(1);
    const int m0 = c.A.top;
    const int as = ys[0].as;
    const Acc acc0 = ys[0].acc;
    int maxp[AB_MAXARGS + 1];
    for (int j = 0; j < nys; ++j) maxp[j] = 1;
    for (int i = 0; i < nx; ++i)
        for (int t = 0; t < x[i].n; ++t) { Key a = tm_key(c, x[i], t); for (int j = 0; j < nys; ++j) { int e = key_get(a, j); if (e > maxp[j]) maxp[j] = e; } }
    // handle table
    int tab[AB_MAXARGS + 1];
    int slot = m0;
    for (int j = 0; j < nys; ++j) { tab[j] = slot; slot += 2 * (maxp[j] + 1); if (maxp[j] >= AB_MAXPOW) fail(c, ST_ARENA); }
    if (slot > c.A.limit) { fail(c, ST_ARENA); for (int i = 0; i < nx; ++i) r[i] = tm_begin(c, as, acc0); return; }
    c.A.top = slot;
    if (c.A.top > c.A.peak) c.A.peak = c.A.top;
    for (int j = 0; j < nys; ++j) {
        TM p0 = tm_copy(c, ys[j]); tm_scal(c, p0, 0.0); c.A.top = p0.off;       // powers[j][0]=ys[j]*0
        ph_store(c, tab[j], p0);
        ph_store(c, tab[j] + 2, ys[j]);                                          // powers[j][1]=ys[j]
        for (int k = 2; k <= maxp[j]; ++k) {
            TM a = ph_load(c, tab[j] + 2 * (k / 2), as, ys[j].acc);
            TM b = ph_load(c, tab[j] + 2 * ((k + 1) / 2), as, ys[j].acc);
            TM pk = tm_mul(c, a, b, ys[j].acc);
            ph_store(c, tab[j] + 2 * k, pk);
        }
    }
    const int mr = c.A.top;     // results start here; slid down to m0 at the end
    for (int i = 0; i < nx; ++i) {
        const int ri = c.A.top;
        TM ra = tm_begin(c, as, acc0);
        ra.e = x[i].e;
        for (int t0 = 0; t0 < x[i].n; t0 += AB_MAXSTREAMS) {
            const int B = (x[i].n - t0 < AB_MAXSTREAMS) ? (x[i].n - t0) : AB_MAXSTREAMS;
            TM st[AB_MAXSTREAMS];
            c.A.top = ra.off + ra.n * AB_TW;
            for (int b = 0; b < B; ++b) {
                const int t = t0 + b;
                Key a = tm_key(c, x[i], t);
                double cv = tm_val(c, x[i], t);
                TM tt = tm_const(c, as, cv, acc0);           // t=c
                for (int j = 0; j < nys; ++j) {
                    int e = key_get(a, j);
                    if (e > 0) {
                        // t=t*powers[j][e]; the factor just consumed stays below as dead space until the slide
                        // after tm_add_many reclaims the whole batch (cheaper than moving every product down)
                        TM pw = ph_load(c, tab[j] + 2 * e, as, ys[j].acc);
                        tt = tm_mul(c, tt, pw, ys[j].acc);
                    }
                }
                c.A.top = tt.off + tt.n * AB_TW;
                st[b] = tt;
            }
            TM s = tm_add_many(c, ra, st, B, acc0);          // r[i]+=t for the B terms, in order
            tm_slide(c, s, ri);
            ra = s;
        }
        c.A.top = ra.off + ra.n * AB_TW;
        r[i] = ra;
    }
    (void)mr;
    vec_slide(c, r, nx, m0);
}

// _split1(tm,k,true/false) src/taylor_model.cpp:2497-2544. Result at the arena top recorded on entry.
AB_HDN TM tm_split(Ctx& c, const TM& tm, int k, bool upper) {
    AB_SECTION(2);
    const int m0 = c.A.top;
    const int deg = tm_degree(c, tm);
    TM r = tm_copy(c, tm);
    for (int i = 0; i < r.n; ++i) {
        int ak = key_get(tm_key(c, r, i), k);
        tm_set_val(c, r, i, div_m(c.rnd, tm_val(c, r, i), (double)(1 << ak)));
    }
    // buckets by exponent of variable k, with that exponent cleared
    int boff[AB_MAXPOW], bn[AB_MAXPOW];
    if (deg + 1 > AB_MAXPOW) { fail(c, ST_ARENA); return r; }
    for (int b = 0; b <= deg; ++b) {
        boff[b] = c.A.top; bn[b] = 0;
        TM bt = tm_begin(c, tm.as, acc_default());
        for (int i = 0; i < r.n; ++i) {
            Key a = tm_key(c, r, i);
            int ak = key_get(a, k);
            if (ak == b) { key_inc(a, k, -ak); tm_append(c, bt, a, tm_val(c, r, i)); }
        }
        bn[b] = bt.n;
    }
    const int mr = c.A.top;
    TM res = tm_begin(c, tm.as, tm.acc);
    res.e = r.e;
    TM st[AB_MAXSTREAMS];
    int B = 0;
    for (int i = 0; i <= deg; ++i) {
        for (int j = i; j <= deg; ++j) {
            // sf=bin(j,i), sign flipped for the lower half when j-i is odd
            long long bn_ = 1; for (int q = 1; q <= i; ++q) bn_ = bn_ * (long long)(j - i + q) / (long long)q;
            double sf = (double)bn_;
            if (!upper && ((j - i) % 2 == 1)) sf = -sf;
            if (B == 0) c.A.top = res.off + res.n * AB_TW;
            // ary[j]*sf with the k-exponent of ary[j] already incremented i times
            TM t = tm_begin(c, tm.as, acc_default());
            for (int q = 0; q < bn[j]; ++q) {
                Key a; for (int w = 0; w < AB_KW; ++w) a.w[w] = aw(c, boff[j] + q * AB_TW + w);
                key_inc(a, k, i);
                tm_append(c, t, a, u2d(aw(c, boff[j] + q * AB_TW + AB_KW)));
            }
            tm_scal(c, t, sf);
            st[B++] = t;
            if (B == AB_MAXSTREAMS || (i == deg && j == deg)) {
                TM s = tm_add_many(c, res, st, B, tm.acc);   // r+=... for the B (i,j) pairs, in loop order
                tm_slide(c, s, mr);
                res = s;
                B = 0;
            }
        }
    }
    tm_slide(c, res, m0);
    return res;
}

// partial_evaluate(x,k,c) for c==1 (src/taylor_model.cpp:2403-2423). Result at the arena top recorded on entry.
AB_HDN TM tm_partial_evaluate_one(Ctx& c, const TM& x, int k) {
    const int m0 = c.A.top;
    const int deg = tm_degree(c, x);
    int boff[AB_MAXPOW], bn[AB_MAXPOW];
    if (deg + 1 > AB_MAXPOW) { fail(c, ST_ARENA); return tm_begin(c, x.as - 1, x.acc); }
    for (int b = 0; b <= deg; ++b) {
        boff[b] = c.A.top;
        TM bt = tm_begin(c, x.as - 1, x.acc);
        for (int i = 0; i < x.n; ++i) {
            Key a = tm_key(c, x, i);
            if (key_get(a, k) == b) tm_append(c, bt, key_remove(a, k, x.as), tm_val(c, x, i));
        }
        bn[b] = bt.n;
    }
    const int mr = c.A.top;
    TM r; r.off = boff[0]; r.n = bn[0]; r.as = x.as - 1; r.e = x.e; r.acc = x.acc;
    r = tm_copy(c, r);
    r.e = x.e;
    for (int b = 1; b <= deg; ++b) {
        TM pb; pb.off = boff[b]; pb.n = bn[b]; pb.as = x.as - 1; pb.e = 0.0; pb.acc = x.acc;
        TM s = tm_add2(c, r, pb, x.acc);
        tm_slide(c, s, mr);
        r = s;
    }
    tm_slide(c, r, m0);
    return r;
}

// ---- one flow step ------------------------------------------------------------------------------------------------
// flow_model (src/taylor_calculus.cpp:630-666) + unchecked_flow (src/taylor_function.cpp:1517-1553 and
// src/taylor_model.cpp:3719-3783). ibx: starting box, bb: bounder box, h: step. Writes n flow models in n+1
// variables (domain ibx x [0,h]) packed at the arena top recorded on entry; dom[] receives the flow domain.
AB_HDN void flow_model(Ctx& c, const TapeOp* ops, const Func* vf, int n, const Ivl* ibx, double h, const Ivl* bb,
                       int order, TM* flow, Ivl* dom) {
    AB_SECTION(3);
    const int m0 = c.A.top;
    const double EPS = 1e-16;
    Ivl bx[AB_MAXN];
    for (int i = 0; i < n; ++i) { bx[i] = ibx[i]; if (bx[i].l == bx[i].u) bx[i] = ivl_add(bx[i], ivl(-EPS, EPS)); }
    for (int i = 0; i < n; ++i) dom[i] = bx[i];
    dom[n] = ivl(0.0, h);
    TM xs[AB_MAXN], usvf[AB_MAXN], yz[AB_MAXN], y[AB_MAXN], ny[AB_MAXN], outs[AB_MAXN];
    // vector-field model on bb: tape over scalings(bb), effectively at the default accuracy
    for (int j = 0; j < n; ++j) xs[j] = tm_scaling(c, n, j, bb[j]);
    for (int i = 0; i < n; ++i) {
        usvf[i] = eval_tm(c, ops, vf[i], xs);
        // unit_scaled_vf[i]=vf[i]*(hmag/rad_ivl(b[i]))
        tm_scal(c, usvf[i], ivl_div(h, rad_ivl(bb[i]), &c.status));
    }
    vec_slide(c, usvf, n, m0);
    for (int i = 0; i < n; ++i) AB_DUMP(c, "usvf", i, usvf[i]);
    if (c.status != ST_OK) { for (int i = 0; i < n; ++i) flow[i] = tm_begin(c, n + 1, acc_default()); return; }
    // y0=scalings(d); unit_scaled_y0=embed(unscale(y0,b),1)
    for (int i = 0; i < n; ++i) {
        int m = c.A.top;
        TM y0 = tm_scaling(c, n, i, bx[i]);
        TM u = tm_unscale(c, y0, bb[i]);
        tm_slide(c, u, m);
        u.as = n + 1;
        yz[i] = u;
    }
    for (int i = 0; i < n; ++i) AB_DUMP(c, "yz", i, yz[i]);
    // model-level flow: y[i]=constant(n+1, yz[i].range()+vf[i].range()*[-1,1])
    const int ybase = c.A.top;
    for (int i = 0; i < n; ++i) {
        Ivl rg = ivl_add(tm_range(c, yz[i]), ivl_mul(tm_range(c, usvf[i]), ivl(-1, 1)));
        y[i] = tm_const(c, n + 1, 1.0, acc_default());
        tm_scal(c, y[i], rg);
    }
    for (int i = 0; i < n; ++i) AB_DUMP(c, "y0", i, y[i]);
    for (int it = 0; it < order; ++it) {
        tm_compose(c, usvf, n, y, n, ny);                 // new_y=compose(vf,y)
        int pout = c.A.top;
        for (int i = 0; i < n; ++i) {
            c.A.top = pout;
            TM w = tm_copy(c, ny[i]);
            tm_antidiff(c, w, n);
            TM s = tm_add2(c, w, yz[i], w.acc);           // new_y[i]+=yz[i]
            TM out;
            if (!tm_intersection(c, y[i], s, out)) out = s;   // IntersectionException is caught and ignored (:3766-3771)
            tm_slide(c, out, pout);
            pout = c.A.top;
            outs[i] = out;
        }
        // new_y[i].swap(y[i]) swaps expansion and error only: y[i] keeps its own accuracy
        for (int i = 0; i < n; ++i) { Acc a = y[i].acc; y[i] = outs[i]; y[i].acc = a; }
        vec_slide(c, y, n, ybase);
        for (int i = 0; i < n; ++i) AB_DUMP(c, "y", (it + 1) * 10 + i, y[i]);
        if (c.status != ST_OK) break;
    }
    for (int i = 0; i < n; ++i) tm_sweep(c, y[i], 0.0);   // leaves the ambient mode upward
    // model_flow=scale(unit_scaled_flow,b); then split(model_flow,n,true) because the time interval is [0,h]
    int pout = c.A.top;
    for (int i = 0; i < n; ++i) {
        c.A.top = pout;
        TM r = tm_copy(c, y[i]);
        tm_rescale(c, r, ivl(-1, 1), bb[i]);
        AB_DUMP(c, "scaled", i, r);
        TM sp = tm_split(c, r, n, true);
        tm_slide(c, sp, pout);
        pout = c.A.top;
        flow[i] = sp;
    }
    vec_slide(c, flow, n, m0);
}

// apply a flow model to a set: unchecked_apply(flow, combine(set.models(), time)) with
// time = ScalarTaylorFunction::variable([-h,h],0)  (src/hybrid_evolver-image.cpp:764-765; src/taylor_set.cpp:366-369).
// set: n models in p variables. Output: n models in p+1 variables at the arena top recorded on entry.
AB_HDN void flow_apply(Ctx& c, const TM* flow, int n, const Ivl* dom, const TM* set, int p, double h, TM* out) {
    AB_SECTION(4);
    const int m0 = c.A.top;
    TM ys[AB_MAXN + 1];
    for (int i = 0; i < n; ++i) {
        TM e = set[i]; e.as = p + 1;                       // embed(x1,1)
        ys[i] = tm_unscale(c, e, dom[i]);
    }
    {
        int m = c.A.top;
        TM t = tm_scaling(c, p + 1, p, ivl(-h, h));        // embed(p, scaling(1,0,[-h,h]))
        TM u = tm_unscale(c, t, dom[n]);
        tm_slide(c, u, m);
        ys[n] = u;
    }
    TM fl[AB_MAXN];
    for (int i = 0; i < n; ++i) { fl[i] = flow[i]; fl[i].as = n + 1; }
    tm_compose(c, fl, n, ys, n + 1, out);
    vec_slide(c, out, n, m0);
}

// unchecked_compose(models, ys) where the models have `nys` variables; convenience wrapper that leaves the
// result packed at the arena top recorded on entry.
AB_HD void compose_at(Ctx& c, const TM* x, int nx, const TM* ys, int nys, TM* out) {
    const int m0 = c.A.top;
    tm_compose(c, x, nx, ys, nys, out);
    vec_slide(c, out, nx, m0);
}

// integration_step(flow_set_model, scaled_time_model) src/taylor_calculus.cpp:234-245
// flow_set: n models in p+1 variables; tau: model in p variables. Output: n models in p variables, accuracy spacial.
AB_HDN void integration_step(Ctx& c, const TM* flow_set, int n, int p, const TM& tau, Acc spacial, TM* out) {
    AB_SECTION(5);
    const int m0 = c.A.top;
    TM ys[AB_MAXARGS + 1];
    for (int j = 0; j < p; ++j) ys[j] = tm_variable(c, p, j);      // BaseModelType::variables(p): default accuracy
    ys[p] = tau;
    tm_compose(c, flow_set, n, ys, p + 1, out);
    vec_slide(c, out, n, m0);
    for (int i = 0; i < n; ++i) out[i].acc = spacial;
}

// reachability_step(flow_set_model, t0, t1) src/taylor_calculus.cpp:248-267. Output: n models in p+1 variables.
AB_HDN void reachability_step(Ctx& c, const TM* flow_set, int n, int p, const TM& t0, const TM& t1, Acc spacial, TM* out) {
    AB_SECTION(6);
    const int m0 = c.A.top;
    TM ys[AB_MAXARGS + 1];
    for (int j = 0; j <= p; ++j) { ys[j] = tm_variable(c, p + 1, j); ys[j].acc = spacial; }
    TM e0 = t0; e0.as = p + 1;                                     // embed(t0,1)
    TM e1 = t1; e1.as = p + 1;
    // time_model=embedded_parameter_model[p]/2  -> operator/(TaylorModel,Float): _scal(r,Interval(1)/c)
    TM tm = tm_copy(c, ys[p]);
    tm_scal(c, tm, ivl_div(ivl(1.0), 2.0));
    // embedded_initial*(0.5-time_model)+embedded_final*(0.5+time_model)
    TM a = tm_neg(c, tm); tm_acc(c, a, 0.5);                       // 0.5-tm: r=neg(x); _acc(r,c)
    TM pa = tm_mul(c, e0, a, a.acc);
    TM b = tm_copy(c, tm); tm_acc(c, b, 0.5);                      // 0.5+tm: r(x); _acc(r,c)
    TM pb = tm_mul(c, e1, b, b.acc);
    TM s = tm_add(c, pa, pb);
    ys[p] = s;
    tm_compose(c, flow_set, n, ys, p + 1, out);
    vec_slide(c, out, n, m0);
    for (int i = 0; i < n; ++i) out[i].acc = spacial;
}

// uniform_error_recondition (src/taylor_set.cpp:499-529). set: n models in p variables, rebuilt at the arena top
// recorded on entry. Returns the new number of parameters.
AB_HDN int uniform_error_recondition(Ctx& c, TM* set, int n, int p, const double* thresholds) {
    int idx[AB_MAXN], k = 0;
    for (int i = 0; i < n; ++i) if (set[i].e > thresholds[i]) idx[k++] = i;
    if (k == 0) return p;
    const int m0 = c.A.top;
    const int np = p + k;
    TM out[AB_MAXN];
    for (int i = 0; i < n; ++i) {
        int which = -1;
        for (int q = 0; q < k; ++q) if (idx[q] == i) which = q;
        if (which < 0) { out[i] = tm_copy(c, set[i]); out[i].as = np; continue; }
        int m = c.A.top;
        TM x = set[i]; x.as = np;
        double error = x.e; x.e = 0.0;
        TM ec = tm_begin(c, np, acc_default());
        tm_append(c, ec, key_unit(p + which), error);
        TM s = tm_add(c, x, ec);                                  // operator+ : fresh default accuracy
        tm_slide(c, s, m);
        out[i] = s;
    }
    for (int i = 0; i < n; ++i) set[i] = out[i];
    (void)m0;
    return np;
}

// recondition(tm, discarded, n_err, index_of_error) src/taylor_model.cpp:3882-3956. Result at the arena top.
AB_HDN TM tm_recondition(Ctx& c, const TM& tm, const int* discarded, int nd, int n_err, int index_of_error) {
    const int nv = tm.as, nk = nv - nd;
    int kept[AB_MAXARGS + 1], kk = 0;
    for (int j = 0; j < nv; ++j) { bool dsc = false; for (int q = 0; q < nd; ++q) if (discarded[q] == j) dsc = true; if (!dsc) kept[kk++] = j; }
    double error = tm.e;
    // pass 1 (reference iterates the terms in REVERSE order): accumulate discarded magnitudes in the ambient mode
    for (int i = tm.n - 1; i >= 0; --i) {
        Key a = tm_key(c, tm, i);
        for (int q = 0; q < nd; ++q) if (key_get(a, discarded[q]) != 0) { error = add_m(c.rnd, error, dabs(tm_val(c, tm, i))); break; }
    }
    TM r = tm_begin(c, nk + n_err, tm.acc);
    bool added = false;
    if (n_err == index_of_error) { r.e = error; added = true; }
    Key ra_error = key_zero();
    if (n_err != index_of_error) key_inc(ra_error, nk + index_of_error, 1);
    int last_kept = -1;
    for (int i = 0; i < tm.n; ++i) {
        Key a = tm_key(c, tm, i);
        bool keep = true;
        for (int q = 0; q < nd; ++q) if (key_get(a, discarded[q]) != 0) { keep = false; break; }
        if (keep) last_kept = i;
    }
    for (int i = 0; i < tm.n; ++i) {
        Key a = tm_key(c, tm, i);
        bool keep = true;
        for (int q = 0; q < nd; ++q) if (key_get(a, discarded[q]) != 0) { keep = false; break; }
        if (!keep) continue;
        Key ra = key_zero();
        for (int q = 0; q < nk; ++q) { int e = key_get(a, kept[q]); if (e) key_inc(ra, q, e); }
        if (!added && key_deg(ra) > 1 && n_err != index_of_error) { tm_append(c, r, ra_error, error); added = true; }
        tm_append(c, r, ra, tm_val(c, tm, i));
        if (i == last_kept && !added) { tm_append(c, r, ra_error, error); added = true; }
    }
    c.rnd = RN;
    return r;
}

// kuhn_recondition (src/taylor_set.cpp:532-577). set: n models in p variables. Returns the number of discarded
// parameters (0 if nothing was done) and writes their indices to discarded[]; set[] is rebuilt at the arena top.
AB_HDN int kuhn_recondition(Ctx& c, TM* set, int n, int p, int* discarded) {
    const int n_kept = n, n_err = n;
    if (p <= n_kept) return 0;
    const int nd = p - n_kept;
    double dep[AB_MAXN][AB_MAXARGS + 1];
    for (int i = 0; i < n; ++i) for (int j = 0; j < p; ++j) dep[i][j] = 0.0;
    for (int i = 0; i < n; ++i)
        for (int t = 0; t < set[i].n; ++t) {
            Key a = tm_key(c, set[i], t); double av = dabs(tm_val(c, set[i], t));
            for (int j = 0; j < p; ++j) if (key_get(a, j) != 0) dep[i][j] = add_m(c.rnd, dep[i][j], av);
        }
    double cm[AB_MAXARGS + 1]; int ci[AB_MAXARGS + 1];
    for (int j = 0; j < p; ++j) { cm[j] = 0.0; ci[j] = j; for (int i = 0; i < n; ++i) cm[j] = dmax(cm[j], dep[i][j]); }
    // std::sort(..., std::greater<Pair<double,uint>>): descending by (value, index)
    for (int a = 1; a < p; ++a) {
        double v = cm[a]; int id = ci[a]; int b = a - 1;
        while (b >= 0 && (cm[b] < v || (cm[b] == v && ci[b] < id))) { cm[b + 1] = cm[b]; ci[b + 1] = ci[b]; --b; }
        cm[b + 1] = v; ci[b + 1] = id;
    }
    for (int j = 0; j < nd; ++j) discarded[j] = ci[n_kept + j];
    for (int a = 1; a < nd; ++a) { int v = discarded[a], b = a - 1; while (b >= 0 && discarded[b] > v) { discarded[b + 1] = discarded[b]; --b; } discarded[b + 1] = v; }
    TM out[AB_MAXN];
    for (int i = 0; i < n; ++i) out[i] = tm_recondition(c, set[i], discarded, nd, n_err, i);
    for (int i = 0; i < n; ++i) set[i] = out[i];
    return nd;
}

} } // namespace ab200
