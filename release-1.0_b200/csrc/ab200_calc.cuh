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
#include "ab200_elem.cuh"

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
    #pragma unroll 1
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
            case OP_EXP: st[sp - 1] = ivl_exp(st[sp - 1]); break;
            case OP_LOG: st[sp - 1] = ivl_log(st[sp - 1], &c.status); break;
            case OP_SIN: st[sp - 1] = ivl_sin(st[sp - 1], c.rnd); break;
            case OP_COS: st[sp - 1] = ivl_cos(st[sp - 1], c.rnd); break;
            default: fail(c, ST_UNSUPPORTED); break;          // tan: not implemented for intervals in the reference either
        }
    }
    return st[0];
}

// ---- vectors of models ------------------------------------------------------------------------------------------
// Slide models v[0..m) (stored in ascending arena order) down to dst, packed; sets the arena top.
AB_HD void vec_slide(Ctx& c, TM* v, int m, int dst) {
    #pragma unroll 1
    for (int i = 0; i < m; ++i) { tm_slide(c, v[i], dst); dst = c.A.top; }
    c.A.top = dst;
}

// ---- Taylor-model tape ------------------------------------------------------------------------------------------
struct TmSlot { TM v; int mark; };

AB_HDN TM tm_rec(Ctx& c, const TM& x);
AB_HDN TM tm_sqrt(Ctx& c, const TM& x);
AB_HDN TM tm_exp(Ctx& c, const TM& x);
AB_HDN TM tm_log(Ctx& c, const TM& x);
AB_HDN TM tm_sincos(Ctx& c, const TM& x, bool want_sin);
AB_HDN TM tm_max(Ctx& c, const TM& x, const TM& y);
AB_HDN TM tm_abs(Ctx& c, const TM& x);

// Evaluate f over models x[0..nx); the result is left at the arena top position recorded on entry.
AB_HDN TM eval_tm(Ctx& c, const TapeOp* ops, Func f, const TM* x) {
    TmSlot st[AB_STACK];
    int sp = 0;
    int m0 = c.A.top;
    #pragma unroll 1
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
                case OP_EXP: r = tm_exp(c, a); break;
                case OP_LOG: r = tm_log(c, a); break;
                case OP_SIN: r = tm_sincos(c, a, true); break;
                case OP_COS: r = tm_sincos(c, a, false); break;
                default: fail(c, ST_UNSUPPORTED); r = a; break;          // tan
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
// `integrator` selects IntegratorBase::flow_bounds (src/integrator.cpp:44-113): the same search with REDUCTION_STEPS 8,
// REFINEMENT_STEPS 4 and no widening of singleton coordinates.
AB_HDN double flow_bounds(Ctx& c, const TapeOp* ops, const Func* vf, int n, const Ivl* r, double hmax, Ivl* bout, bool integrator = false) {
    const double INITIAL_MULTIPLIER = 2, MULTIPLIER = 1.125, BOX_RADIUS_MULTIPLIER = 1.03125;
    const int EXPANSION_STEPS = 8, REDUCTION_STEPS = integrator ? 8 : 16, REFINEMENT_STEPS = integrator ? 4 : 8;
    const double EPS = 1e-16;
    Ivl d[AB_MAXN], delta[AB_MAXN], b[AB_MAXN], nb[AB_MAXN], df[AB_MAXN], dd[AB_MAXN];
    #pragma unroll 1
    for (int i = 0; i < n; ++i) { d[i] = r[i]; if (!integrator && d[i].l == d[i].u) d[i] = ivl_add(d[i], ivl(-EPS, EPS)); }
    #pragma unroll 1
    for (int i = 0; i < n; ++i) delta[i] = ivl_mul(ivl_sub(d[i], ivl_mid(d[i])), BOX_RADIUS_MULTIPLIER - 1);
    #pragma unroll 1
    for (int i = 0; i < n; ++i) dd[i] = ivl_add(d[i], delta[i]);
    double hmin = hmax / (1 << REDUCTION_STEPS);
    double h = hmax;
    bool success = false;
    Ivl ih = ivl(0, h);
    #pragma unroll 1
    while (!success) {
        if (!(h > hmin)) { fail(c, ST_BOUNDER); return h; }
        Ivl m2 = ivl_mul(ih, INITIAL_MULTIPLIER);
        #pragma unroll 1
        for (int i = 0; i < n; ++i) df[i] = eval_ivl(c, ops, vf[i], d);
        #pragma unroll 1
        for (int i = 0; i < n; ++i) b[i] = ivl_add(ivl_add(d[i], ivl_mul(m2, df[i])), delta[i]);
        #pragma unroll 1
        for (int s = 0; s < EXPANSION_STEPS; ++s) {
            #pragma unroll 1
            for (int i = 0; i < n; ++i) df[i] = eval_ivl(c, ops, vf[i], b);
            bool bounded = true, sub = true;
            #pragma unroll 1
            for (int i = 0; i < n; ++i) {
                nb[i] = ivl_add(dd[i], ivl_mul(ih, df[i]));
                if (!(dabs(nb[i].l) <= DBL_MAX && dabs(nb[i].u) <= DBL_MAX)) bounded = false;
                if (!ivl_subset(nb[i], b[i])) sub = false;
            }
            if (!bounded) { success = false; break; }
            else if (sub) { success = true; break; }
            else {
                Ivl mi = ivl_mul(ih, MULTIPLIER);
                #pragma unroll 1
                for (int i = 0; i < n; ++i) b[i] = ivl_add(dd[i], ivl_mul(mi, df[i]));
            }
        }
        if (!success) { h /= 2; ih = ivl(0, h); }
        if (c.status != ST_OK) return h;
    }
    #pragma unroll 1
    for (int s = 0; s < REFINEMENT_STEPS; ++s) {
        #pragma unroll 1
        for (int i = 0; i < n; ++i) b[i] = nb[i];
        #pragma unroll 1
        for (int i = 0; i < n; ++i) df[i] = eval_ivl(c, ops, vf[i], b);
        #pragma unroll 1
        for (int i = 0; i < n; ++i) {
            nb[i] = ivl_add(dd[i], ivl_mul(ih, df[i]));
            if (!ivl_subset(nb[i], b[i])) fail(c, ST_BOUNDER);
        }
    }
    #pragma unroll 1
    for (int i = 0; i < n; ++i) { bout[i] = nb[i]; if (!ivl_subset(d[i], nb[i])) fail(c, ST_BOUNDER); }
    return h;
}

// ---- compose ------------------------------------------------------------------------------------------------------
// Power-table handles are kept in the arena (3 words each) to keep per-thread local memory small.
#define AB_MAXPOW 40
#define AB_PHW 3            // words per power-table handle: offset|count, error, norm (filled on first use)
#define AB_NORM_UNSET 0x7FF8AB200AB20001ull
AB_HD void ph_store(Ctx& c, int slot, const TM& t) {
    aw(c, slot) = ((uint64_t)(uint32_t)t.off << 32) | (uint32_t)t.n;
    aw(c, slot + 1) = d2u(t.e);
    aw(c, slot + 2) = AB_NORM_UNSET;
}
// sum of the coefficient magnitudes of a power (upward, in term order, from zero: the `ys` of _mul4), computed once
AB_HD double ph_norm(Ctx& c, int slot, const TM& t) {
    const uint64_t w = aw(c, slot + 2);
    if (w != AB_NORM_UNSET) return u2d(w);
    const double nm = tm_abs_sum(c, t, 0.0, true, 0);
    wsync();                       // every lane has read the slot before it is filled
    aw(c, slot + 2) = d2u(nm);
    return nm;
}
AB_HD TM ph_load(Ctx& c, int slot, int as, Acc acc) {
    TM t; uint64_t w = aw(c, slot); t.off = (int)(w >> 32); t.n = (int)(w & 0xFFFFFFFFu); t.e = u2d(aw(c, slot + 1)); t.as = as; t.acc = acc; return t;
}

// {cv} * p with accuracy acc, where {cv} is the constant model of a nonzero coefficient (error 0) and p has at least one
// term: _mul4 (src/taylor_model.cpp:763-824) with a one-term left factor, i.e. one stream -- every term of p scaled
// (nearest), products below the sweep threshold or above the degree limit dropped into the truncation error
// tte (upward, in term order), error = up(tte+tte/2) + |cv|*p.e (the other cross terms vanish with the factor's error).
// Result at the arena top.
AB_HDN TM term_times_power(Ctx& c, double cv, const TM& p, Acc acc) {
    const int W = AB_W, lane = wlane(), n = p.n, m = c.A.top, soff = p.off;
    AB_FLOPS(5.0 * n + 3.0 + 2.0 * (1 + n) + 6.0);
    uint64_t* const base = c.A.base;
    TM r; r.off = m; r.n = 0; r.as = p.as; r.e = 0.0; r.acc = acc;
    if (!arena_room(c, m, n)) return r;
    const double sweep = acc.sweep, acv = dabs(cv); const int maxdeg = acc.maxdeg;
    wsync();
    double tte = 0.0; int rn = 0;
    #pragma unroll 1
    for (int j0 = 0; j0 < n; j0 += W) {
        const int j = j0 + lane; const bool act = j < n;
        Key ta = key_zero(); double sv = 0.0, tv = 0.0; bool disc = false;
        if (act) { const int w = soff + j * AB_TW; sv = bval(base, w); ta = bkey(base, w); tv = mul_rn(cv, sv); disc = dabs(tv) < sweep || key_deg(ta) > maxdeg; }
        const unsigned dm = wballot(disc), km = wballot(act && !disc);
        if (act && !disc) bput(base, m + (rn + popc32(km & lanemask_lt(lane))) * AB_TW, ta, tv);
        rn += popc32(km);
        if (dm) tte = chain_ru_masked(tte, mul_ru(acv, dabs(sv)), dm);
    }
    wsync();
    r.n = rn;
    r.e = add_ru(add_ru(tte, tte / 2), mul_ru(acv, p.e));
    c.rnd = RN;
    c.A.top = m + rn * AB_TW;
    arena_mark_peak(c);
    return r;
}
// t <- t * q with accuracy acc for a one-term q and t the top model on the arena (at least one term), in place: _mul4 with
// a one-term right factor -- every x-term is its own stream {x_i*q}, a dropped product adds up(tte_i+tte_i/2) to the error
// chain in x order, and the cross terms use the norm of t (needed only when q carries an error).
AB_HDN void term_times_monomial(Ctx& c, TM& t, const TM& q, Acc acc) {
    const int W = AB_W, lane = wlane(), n = t.n, off = t.off;
    AB_FLOPS(10.0 * n + 8.0);
    uint64_t* const base = c.A.base;
    const Key qa = bkey(base, q.off); const double qv = bval(base, q.off), aqv = dabs(qv), qe = q.e, te0 = t.e;
    const double sweep = acc.sweep; const int maxdeg = acc.maxdeg;
    wsync();
    double e = 0.0, norm = 0.0; int rn = 0;
    #pragma unroll 1
    for (int i0 = 0; i0 < n; i0 += W) {
        const int i = i0 + lane; const bool act = i < n;
        Key ta = key_zero(); double xv = 0.0, tv = 0.0; bool disc = false;
        if (act) { const int w = off + i * AB_TW; xv = bval(base, w); ta = key_add(bkey(base, w), qa); tv = mul_rn(xv, qv); disc = dabs(tv) < sweep || key_deg(ta) > maxdeg; }
        const unsigned dm = wballot(disc), km = wballot(act && !disc);
        wsync();          // the chunk has been read by every lane before it is compacted over
        if (act && !disc) bput(base, off + (rn + popc32(km & lanemask_lt(lane))) * AB_TW, ta, tv);
        rn += popc32(km);
        if (dm) { const double t1 = mul_ru(dabs(xv), aqv); e = chain_ru_masked(e, add_ru(t1, t1 / 2), dm); }
        if (qe != 0.0) norm = chain_ru(norm, act ? dabs(xv) : 0.0, (n - i0 < W) ? n - i0 : W);
    }
    wsync();
    t.n = rn; t.acc = acc;
    t.e = add_ru(e, add_ru(add_ru(mul_ru(norm, qe), mul_ru(aqv, te0)), mul_ru(te0, qe)));
    c.rnd = RN;
    c.A.top = off + rn * AB_TW;
}

// ---- composition on the dense accumulator -------------------------------------------------------------------------------
// State of one product chain t = c * prod_j powers[j][e_j]. While the running product has at most one term it lives in
// registers (the coefficient of a high-order term times a power usually falls below the sweep threshold entirely, or leaves
// one term); it becomes a model in the arena only when two or more terms survive.
struct ChainState {
    int kind;          // 0 = empty (every product discarded), 1 = one term in registers, 2 = model `tt` in the arena
    Key ka; double va; // the one term
    double e;          // error of the running product (kinds 0 and 1)
    TM tt;
};
// {ka: va, error xe} * p with accuracy acc (_mul4 with a one-term left factor, src/taylor_model.cpp:763-824): the products are
// folded straight into the accumulator as one stream when `fold` is set (last factor of the chain), else the survivors become
// the new state. Returns te, the gap sum of the stream (fold only); *te_out error of the product.
AB_HD double chain_times_power(Ctx& c, DAcc& DA, ChainState& st, const TM& p, int pslot, Acc acc, bool fold, double* e_out) {
    const int W = AB_W, lane = wlane(), n = p.n, soff = p.off;
    uint64_t* const base = c.A.base;
    const Key fa = st.ka; const double fv = st.va, afv = dabs(fv), xe = st.e;
    const double sweep = acc.sweep; const int maxdeg = acc.maxdeg;
    const int m = c.A.top;
    double tte = 0.0, te = 0.0; int rn = 0;
    Key k1 = key_zero(); double v1 = 0.0;      // a lone survivor
    const bool to_arena = !fold && n > W;      // survivors of a long power go to the arena as they come
    if (to_arena && !arena_room(c, m, n)) { *e_out = 0.0; return 0.0; }
    if (to_arena) wsync();
    #pragma unroll 1
    for (int j0 = 0; j0 < n; j0 += W) {
        const int j = j0 + lane; const bool act = j < n;
        Key ta = key_zero(); double sv = 0.0, tv = 0.0; bool disc = false;
        if (act) { const int w = soff + j * AB_TW; sv = bval(base, w); ta = key_add(fa, bkey(base, w)); tv = mul_rn(fv, sv); disc = dabs(tv) < sweep || key_deg(ta) > maxdeg; }
        const unsigned dm = wballot(disc);
        if (fold) {
            double g = 0.0; bool co = false;
            if (act && !disc) g = dacc_fold(DA, dacc_rank(DA, ta), ta, tv, &co);
            const unsigned gm = wballot(g != 0.0);
            if (gm) te = chain_ru_masked(te, g, gm);
#if defined(AB_COUNT_FLOPS) || defined(AB_MUL_HIST)
            { const int cc_ = popc32(wballot(co)); AB_FLOPS(5 * cc_); }
#endif
        } else {
            const unsigned km = wballot(act && !disc);
            if (to_arena) { if (act && !disc) bput(base, m + (rn + popc32(km & lanemask_lt(lane))) * AB_TW, ta, tv); }
            else if (km) {
                const int cnt = popc32(km);
                if (cnt == 1) { const int src = ffs32(km) - 1; for (int w = 0; w < AB_KW; ++w) k1.w[w] = wshfl(ta.w[w], src); v1 = wshfl(tv, src); }
                else {      // two or more survivors of a short power: they become a model
                    if (!arena_room(c, m, cnt)) { *e_out = 0.0; return 0.0; }
                    wsync();
                    if (act && !disc) bput(base, m + popc32(km & lanemask_lt(lane)) * AB_TW, ta, tv);
                }
            }
            rn += popc32(km);
        }
        if (dm) tte = chain_ru_masked(tte, mul_ru(afv, dabs(sv)), dm);
    }
    const double nm = (xe != 0.0) ? ph_norm(c, pslot, p) : 0.0;
    const double e = add_ru(add_ru(tte, tte / 2), add_ru(add_ru(mul_ru(afv, p.e), mul_ru(nm, xe)), mul_ru(xe, p.e)));
    c.rnd = RN;
    *e_out = e;
    AB_FLOPS(5.0 * n + 3.0 + 2.0 * (1 + n) + 6.0);
    if (fold) { wsync(); return te; }
    if (to_arena || rn >= 2) {
        wsync();
        st.kind = 2; st.tt.off = m; st.tt.n = rn; st.tt.as = p.as; st.tt.e = e; st.tt.acc = acc;
        c.A.top = m + rn * AB_TW; arena_mark_peak(c);
    } else if (rn == 1) { st.kind = 1; st.ka = k1; st.va = v1; st.e = e; }
    else { st.kind = 0; st.e = e; }
    return 0.0;
}
// r = x o ys for one model x on the dense accumulator; the sum stays in shared memory while the terms stream through it.
// A chain that needs a general product (both factors with two or more terms) parks the sum in the arena at ri first, since
// the product uses the accumulator itself. Result written at ri.
AB_HDN TM compose_stream_dense(Ctx& c, const TM& xi, const TM* ys, int nys, const int* tab, int as, Acc acc0, int Dc, int ri) {
    const int W = AB_W, lane = wlane();
    TM ra; ra.off = ri; ra.n = 0; ra.as = as; ra.e = xi.e; ra.acc = acc0;
    const int sbase = ri + AB_SM_DENSE * AB_TW;            // product chains are built above the room kept for the parked sum
    if (!arena_room(c, sbase, 8)) return ra;
    DAcc DA; dacc_open(c, DA, as, Dc);
    bool open = true; double e_run = xi.e;
    Key pk = key_zero(); double pv = 0.0;                  // this lane's term of the current chunk of x
    #pragma unroll 1
    for (int t = 0; t < xi.n; ++t) {
        if (t % W == 0) { const int q = t + lane; if (q < xi.n) { pk = tm_key(c, xi, q); pv = tm_val(c, xi, q); } }
        Key a; for (int w = 0; w < AB_KW; ++w) a.w[w] = wshfl(pk.w[w], t % W);
        const double cv = wshfl(pv, t % W);
        c.A.top = sbase;
        int lastj = -1;
        #pragma unroll 1
        for (int j = 0; j < nys; ++j) if (key_get(a, j) != 0) lastj = j;
        ChainState st; st.kind = (cv == 0.0) ? 0 : 1; st.ka = key_zero(); st.va = cv; st.e = 0.0;
        double te = 0.0, t_e = 0.0; bool folded = false;
        #pragma unroll 1
        for (int j = 0; j <= lastj; ++j) {
            const int e = key_get(a, j);
            if (e == 0) continue;
            const int pslot = tab[j] + AB_PHW * e;
            const TM pw = ph_load(c, pslot, as, c.cmp_acc[j]);
            const Acc accj = c.cmp_acc[j];
            if (st.kind == 0) {
                // every product so far was discarded: _mul4 with an empty left factor leaves only the error cross terms
                const double nm = (st.e != 0.0) ? ph_norm(c, pslot, pw) : 0.0;
                st.e = add_ru(0.0, add_ru(add_ru(mul_ru(0.0, pw.e), mul_ru(nm, st.e)), mul_ru(st.e, pw.e)));
                c.rnd = RN; AB_FLOPS(2.0 * pw.n + 6.0);
            } else if (st.kind == 1) {
                if (pw.n == 0) {      // one term times an empty power: only the cross term |x| * error(power) remains
                    st.e = add_ru(0.0, add_ru(add_ru(mul_ru(dabs(st.va), pw.e), mul_ru(0.0, st.e)), mul_ru(st.e, pw.e)));
                    st.kind = 0; c.rnd = RN;
                } else if (j == lastj && open) { te = chain_times_power(c, DA, st, pw, pslot, accj, true, &t_e); folded = true; }
                else { double eo; chain_times_power(c, DA, st, pw, pslot, accj, false, &eo); }
            } else {
                if (pw.n == 1 && st.tt.n >= 1 && tm_end(st.tt) == c.A.top) term_times_monomial(c, st.tt, pw, accj);
                else {
                    if (st.tt.n > 1 && pw.n > 1 && open) { ra.n = dacc_emit(c, DA, ri); ra.e = e_run; open = false; }
                    st.tt = tm_mul(c, st.tt, pw, accj);
                }
            }
            if (c.status != ST_OK) { if (open) dacc_emit(c, DA, -1); return ra; }
        }
        if (!open) { dacc_open(c, DA, as, Dc); dacc_add_model(c, DA, ra, false); open = true; }
        if (!folded) {
            if (st.kind == 1) {
                double g = 0.0;
                if (lane == 0) g = dacc_fold(DA, dacc_rank(DA, st.ka), st.ka, st.va);
                te = add_ru(0.0, wshfl(g, 0)); t_e = st.e; AB_FLOPS(5);
                wsync();
            } else if (st.kind == 0) { te = 0.0; t_e = st.e; }
            else { te = dacc_add_model(c, DA, st.tt, true); t_e = st.tt.e; }
        }
        e_run = add_ru(add_ru(e_run, t_e), te / 2); AB_FLOPS(3);
    }
    ra.n = dacc_emit(c, DA, ri);
    ra.e = e_run;
    c.rnd = RN;
    c.A.top = ri + ra.n * AB_TW;
    arena_mark_peak(c);
    return ra;
}

// _compose2 (src/taylor_model.cpp:3187-3240): r[i] = x[i] o ys, nx models in xas=nys variables.
// Results are written to r[0..nx) packed from the arena top recorded on entry.
AB_HDN void tm_compose(Ctx& c, const TM* x, int nx, const TM* ys, int nys, TM* r) {
    AB_SECTION(1);
    const int m0 = c.A.top;
    const int as = ys[0].as;
    const Acc acc0 = ys[0].acc;
    int maxp[AB_MAXARGS + 1];
    {   // largest exponent of every variable: per-byte maximum of the key words, lanes over terms, then across the lanes
        const int W = AB_W, lane = wlane();
        uint64_t mx[AB_KW]; for (int w = 0; w < AB_KW; ++w) mx[w] = 0;
        #pragma unroll 1
        for (int i = 0; i < nx; ++i)
            #pragma unroll 1
            for (int t = lane; t < x[i].n; t += W) { Key a = tm_key(c, x[i], t); for (int w = 0; w < AB_KW; ++w) mx[w] = bytes_max(mx[w], a.w[w]); }
        #pragma unroll 1
        for (int d = 1; d < W; d <<= 1) for (int w = 0; w < AB_KW; ++w) mx[w] = bytes_max(mx[w], wshfl(mx[w], lane ^ d));
        Key km; for (int w = 0; w < AB_KW; ++w) km.w[w] = mx[w];
        #pragma unroll 1
        for (int j = 0; j < nys; ++j) { const int e = key_get(km, j); maxp[j] = e > 1 ? e : 1; }
    }
    // handle table
    int* const tab = c.cmp_tab;
    int slot = m0;
    #pragma unroll 1
    for (int j = 0; j < nys; ++j) { tab[j] = slot; slot += AB_PHW * (maxp[j] + 1); if (maxp[j] >= AB_MAXPOW) fail(c, ST_ARENA); }
    if (slot > c.A.limit) { fail(c, ST_ARENA); for (int i = 0; i < nx; ++i) r[i] = tm_begin(c, as, acc0); return; }
    c.A.top = slot;
    if (c.A.top > c.A.peak) c.A.peak = c.A.top;
    #pragma unroll 1
    for (int j = 0; j < nys; ++j) c.cmp_acc[j] = ys[j].acc;
    #pragma unroll 1
    for (int j = 0; j < nys; ++j) {
        TM p0 = tm_copy(c, ys[j]); tm_scal(c, p0, 0.0); c.A.top = p0.off;       // powers[j][0]=ys[j]*0
        ph_store(c, tab[j], p0);
        ph_store(c, tab[j] + AB_PHW, ys[j]);                                          // powers[j][1]=ys[j]
        #pragma unroll 1
        for (int k = 2; k <= maxp[j]; ++k) {
            TM a = ph_load(c, tab[j] + AB_PHW * (k / 2), as, ys[j].acc);
            TM b = ph_load(c, tab[j] + AB_PHW * ((k + 1) / 2), as, ys[j].acc);
            TM pk = tm_mul(c, a, b, ys[j].acc);
            ph_store(c, tab[j] + AB_PHW * k, pk);
        }
    }
    const int mr = c.A.top;     // results start here; slid down to m0 at the end
#ifdef AB_COMPOSE_STATS
    if (wlane() == 0) { extern long long g_cs[8]; g_cs[0]++; for (int i = 0; i < nx; ++i) g_cs[1] += x[i].n; }
#endif
    #pragma unroll 1
    for (int i = 0; i < nx; ++i) {
        const int ri = c.A.top;
        TM ra = tm_begin(c, as, acc0);
        ra.e = x[i].e;
        // r[i] = sum over the terms of x[i], in term order, of t = c * prod_j powers[j][e_j]. The running sum stays in the
        // shared-memory accumulator while the terms stream through it (no intermediate merges); a term whose product chain
        // needs a general product -- which uses the accumulator itself -- first parks the sum in the arena at ri.
        bool streamed = false;
        int Dc0 = 0; for (int j = 0; j < nys; ++j) if (ys[j].acc.maxdeg > Dc0) Dc0 = ys[j].acc.maxdeg;
        if (dacc_fits(c, as, Dc0) && (long long)ri + (long long)(AB_SM_DENSE + 8) * AB_TW < (long long)c.A.limit) {
            ra = compose_stream_dense(c, x[i], ys, nys, tab, as, acc0, Dc0, ri);
            if (c.status != ST_OK) { for (int q = i; q < nx; ++q) r[q] = tm_begin(c, as, acc0); return; }
            streamed = true;
        } else
        if (c.sm && (long long)ri + (long long)(AB_SM_CAP + 8) * AB_TW < (long long)c.A.limit) {
            const int sbase = ri + AB_SM_CAP * AB_TW;           // product chains are built above the room kept for the parked sum
            AccTab T; bool open = false, failed = false; int dmax = 0; double e_run = x[i].e;
            T.overflow = false;
            Key pk = key_zero(); double pv = 0.0;               // this lane's term of the current chunk of x[i]
            const int W = AB_W, lane = wlane();
            #pragma unroll 1
            for (int t = 0; t < x[i].n && !failed; ++t) {
                if (t % W == 0) { const int q = t + lane; if (q < x[i].n) { pk = tm_key(c, x[i], q); pv = tm_val(c, x[i], q); } }
                Key a; for (int w = 0; w < AB_KW; ++w) a.w[w] = wshfl(pk.w[w], t % W);
                const double cv = wshfl(pv, t % W);
                c.A.top = sbase;
                TM tt; bool started = false;
                if (cv == 0.0) { tt = tm_const(c, as, cv, acc0); started = true; }
                #pragma unroll 1
                for (int j = 0; j < nys; ++j) {
                    const int e = key_get(a, j);
                    if (e == 0) continue;
                    const TM pw = ph_load(c, tab[j] + AB_PHW * e, as, c.cmp_acc[j]);
                    if (!started) {
                        if (pw.n == 0) { tt = tm_const(c, as, cv, acc0); tt = tm_mul(c, tt, pw, c.cmp_acc[j]); }
                        else tt = term_times_power(c, cv, pw, c.cmp_acc[j]);
                        started = true;
                    } else if (pw.n == 1 && tt.n >= 1 && tm_end(tt) == c.A.top) term_times_monomial(c, tt, pw, c.cmp_acc[j]);
                    else if (tt.n == 0) {
                        // every product so far was discarded: _mul4 with an empty left factor leaves only the error cross terms
                        const double nm = (tt.e != 0.0) ? ph_norm(c, tab[j] + AB_PHW * e, pw) : 0.0;
                        tt.e = add_ru(0.0, add_ru(add_ru(mul_ru(0.0, pw.e), mul_ru(nm, tt.e)), mul_ru(tt.e, pw.e)));
                        tt.acc = c.cmp_acc[j]; c.rnd = RN;
                    } else {
                        if (tt.n > 1 && pw.n > 1 && open) {      // general product ahead: it needs the accumulator
                            ra.n = acc_emit(c, T, as, dmax, ri); ra.e = e_run; open = false;
                        }
                        tt = tm_mul(c, tt, pw, c.cmp_acc[j]);
                    }
                }
                if (!started) tt = tm_const(c, as, cv, acc0);
                if (c.status != ST_OK) { failed = true; break; }
                if (!open) {
                    const int keep = c.A.top;
                    if (!acc_open(c, T, 0, keep, AB_SM_CAP)) { failed = true; break; }
                    acc_add_model(c, T, ra);
                    open = true;
                }
                const double te = T.overflow ? 0.0 : acc_add_model(c, T, tt);
                if (T.overflow) { failed = true; break; }
                { const int d = tm_degree(c, tt); if (d > dmax) dmax = d; }
                e_run = add_ru(add_ru(e_run, tt.e), te / 2);
            }
            if (!failed) {
                if (open) { ra.n = acc_emit(c, T, as, dmax, ri); }
                ra.e = e_run; ra.off = ri;
                c.rnd = RN;
                streamed = true;
            } else if (c.status != ST_OK) { for (int q = i; q < nx; ++q) r[q] = tm_begin(c, as, acc0); return; }
        }
        if (!streamed) {
        // batched form (no shared area, or a sum too large for it): products of up to AB_MAXSTREAMS terms, then one addition
        c.A.top = ri;
        ra = tm_begin(c, as, acc0);
        ra.e = x[i].e;
        #pragma unroll 1
        for (int t0 = 0; t0 < x[i].n; t0 += AB_MAXSTREAMS) {
            const int B = (x[i].n - t0 < AB_MAXSTREAMS) ? (x[i].n - t0) : AB_MAXSTREAMS;
            TM st[AB_MAXSTREAMS];
            c.A.top = ra.off + ra.n * AB_TW;
            #pragma unroll 1
            for (int b = 0; b < B; ++b) {
                const int t = t0 + b;
                const Key a = tm_key(c, x[i], t);
                const double cv = tm_val(c, x[i], t);
                // t=c; t=t*powers[j][e] for every variable with a positive exponent. The first product scales a power by the
                // coefficient, and a product with a one-term power only shifts and scales the terms: both are done without
                // building the constant model or leaving a copy behind (same arithmetic as _mul4 on those operands, see
                // term_times_power / term_times_monomial).
                TM tt; bool started = false;
                if (cv == 0.0) { tt = tm_const(c, as, cv, acc0); started = true; }
                #pragma unroll 1
                for (int j = 0; j < nys; ++j) {
                    const int e = key_get(a, j);
                    if (e == 0) continue;
                    const TM pw = ph_load(c, tab[j] + AB_PHW * e, as, ys[j].acc);
#ifdef AB_COMPOSE_STATS
                    if (wlane() == 0) { extern long long g_cs[8]; g_cs[2]++; g_cs[3] += (long long)(started ? tt.n : 1) * pw.n; if (!started || tt.n <= 1 || pw.n <= 1) g_cs[4]++; }
#endif
                    if (!started) {
                        if (pw.n == 0) { tt = tm_const(c, as, cv, acc0); tt = tm_mul(c, tt, pw, ys[j].acc); }
                        else tt = term_times_power(c, cv, pw, ys[j].acc);
                        started = true;
                    } else if (pw.n == 1 && tt.n >= 1 && tm_end(tt) == c.A.top) term_times_monomial(c, tt, pw, ys[j].acc);
                    else tt = tm_mul(c, tt, pw, ys[j].acc);      // the factor just consumed stays below as dead space until the slide
                }
                if (!started) tt = tm_const(c, as, cv, acc0);
                c.A.top = tt.off + tt.n * AB_TW;
                st[b] = tt;
            }
            TM s = tm_add_many(c, ra, st, B, acc0);          // r[i]+=t for the B terms, in order
            tm_slide(c, s, ri);
            ra = s;
        }
        }
        c.A.top = ra.off + ra.n * AB_TW;
        r[i] = ra;
    }
    (void)mr;
    vec_slide(c, r, nx, m0);
}

// _split1(tm,k,true/false) src/taylor_model.cpp:2497-2544. Result at the arena top recorded on entry.
// The reference computes r = sum_i sum_{j>=i} ary[j] * (+-bin(j,i)) x_k^i as (deg+1)(deg+2)/2 scaled copies added one after
// another. Stream (i,j) only meets keys whose exponent of x_k is i, and such a key receives its contributions in j order, so
// the streams are fed to one accumulator bucket by bucket (j), all i of a bucket side by side; each stream's two
// upward-rounded sums (the scaling gaps of _scal and the coincidence gaps of _add2, both in term order) are recorded as
// exponent codes and summed by one lane per stream, and the error chain e <- up(up(e + t.e) + te/2) is applied at the end
// in the reference's (i major, j minor) order.
AB_HDN TM tm_split(Ctx& c, const TM& tm, int k, bool upper) {
    AB_SECTION(2);
    const int m0 = c.A.top;
    const int deg = tm_degree(c, tm);
    const int W = AB_W, lane = wlane();
    uint64_t* const base = c.A.base;
    TM res = tm_begin(c, tm.as, tm.acc);
    res.e = tm.e;
    if (deg + 1 > AB_MAXPOW) { fail(c, ST_ARENA); return res; }
    // buckets by exponent of variable k, with that exponent cleared and the coefficient divided by 2^exponent
    int boff[AB_MAXPOW], bn[AB_MAXPOW];
    const int md = c.rnd;
    long long want = 0;
    #pragma unroll 1
    for (int b = 0; b <= deg; ++b) {
        boff[b] = c.A.top;
        TM bt = tm_begin(c, tm.as, acc_default());
        AB_FILTER_COPY(c, bt, tm, key_get(a, k) == b, key_with(a, k, 0), div_m(md, v, (double)(1 << b)));
        bn[b] = bt.n;
        want += (long long)bt.n * (b + 1);
    }
    { const long long R = monomial_count(c.binom, deg, tm.as); if (want > R) want = R; }
    const int mr = c.A.top;
    // per-stream sums ts[i][j], te[i][j], kept above the buckets
    const int npair = (deg + 1) * (deg + 1);
    if (!arena_room(c, mr, npair + 2)) return res;
    double* const TS = (double*)(base + mr); double* const TE = TS + npair;
    const int sums_end = mr + 2 * npair;
    if (sums_end > c.A.peak) c.A.peak = sums_end;
    AccTab T; DAcc DA;
    const bool dense = dacc_fits(c, tm.as, deg);
    T.overflow = false;
    #pragma unroll 1
    for (int attempt = 0; attempt < 2; ++attempt) {
        if (dense) dacc_open(c, DA, tm.as, deg);
        else if (!acc_open(c, T, attempt, sums_end + (int)want * AB_TW, want)) return res;
        // gap codes of one bucket: one 32-bit cell per (i, term): low half scaling gap, high half coincidence gap
        uint32_t* G; int gcap;
        if (c.sm) { G = (uint32_t*)(c.sm + AB_SM_OFF_G16); gcap = AB_SM_G16 / 2; }
        else {
            const long long at = dense ? sums_end : (T.val + T.cap) - base; int mx = 1; for (int b = 0; b <= deg; ++b) if (bn[b] > mx) mx = bn[b];
            if (at + (mx + 1) / 2 + 1 > (long long)c.A.limit) { fail(c, ST_ARENA); return res; }
            G = (uint32_t*)(base + at); gcap = mx;
        }
        bool badgap = false;
        #pragma unroll 1
        for (int j = 0; j <= deg && !T.overflow; ++j) {
            const int nb = bn[j];
            if (nb == 0) continue;
            int ib = gcap / nb; if (ib < 1) { fail(c, ST_ARENA); return res; }
            #pragma unroll 1
            for (int ia = 0; ia <= j && !T.overflow; ia += ib) {
                const int ni = (j + 1 - ia < ib) ? (j + 1 - ia) : ib;
                const int nitems = ni * nb;
                #pragma unroll 1
                for (int w0 = 0; w0 < nitems; w0 += W) {
                    const int w = w0 + lane; const bool act = w < nitems;
                    Key key = key_zero(); double tv = 0.0, gs = 0.0;
                    if (act) {
                        const int ii = w / nb, q = w - ii * nb, i = ia + ii;
                        const Key a = bkey(base, boff[j] + q * AB_TW); const double v = bval(base, boff[j] + q * AB_TW);
                        key = key_with(a, k, i);
                        // sf=bin(j,i), sign flipped for the lower half when j-i is odd; t=ary[j]*sf is _scal(Float)
                        double sf;
                        if (j <= 33 && c.binom) sf = (double)binom_at(c.binom, j, (i < j - i) ? i : j - i);
                        else { long long bn_ = 1; for (int z = 1; z <= i; ++z) bn_ = bn_ * (long long)(j - i + z) / (long long)z; sf = (double)bn_; }
                        if (!upper && ((j - i) % 2 == 1)) sf = -sf;
                        if (sf == 1) tv = v;
                        else if (sf == 2 || sf == 0.5) tv = mul_m(md, v, sf);
                        else { gs = add_ru(mul_ru(v, sf), mul_ru(v, -sf)); tv = mul_rn(v, sf); }
                    }
                    const int slot = dense ? 0 : acc_slot(T, act, key, lane);
                    if (T.overflow) break;
                    if (act) {
                        const double ga = dense ? dacc_fold(DA, dacc_rank(DA, key), key, tv) : acc_fold(T, slot, tv);
                        G[w] = gap_code(gs, &badgap) | (gap_code(ga, &badgap) << 16);
                    }
                }
                wsync();
                if (T.overflow) break;
                #pragma unroll 1
                for (int ii = lane; ii < ni; ii += W) {
                    double ts = 0.0, te = 0.0;
                    const uint32_t* row = G + ii * nb;
                    #pragma unroll 1
                    for (int q = 0; q < nb; ++q) { const uint32_t cd = row[q]; if (cd & 0xFFFFu) ts = add_ru(ts, gap_value(cd & 0xFFFFu)); if (cd >> 16) te = add_ru(te, gap_value(cd >> 16)); }
                    TS[(ia + ii) * (deg + 1) + j] = ts; TE[(ia + ii) * (deg + 1) + j] = te;
                }
                wsync();
            }
        }
        if (wany(badgap)) fail(c, ST_UNSUPPORTED);
        if (!T.overflow) break;
        if (attempt == 1) { fail(c, ST_ARENA); return res; }
    }
#if defined(AB_COUNT_FLOPS) || defined(AB_MUL_HIST)
    for (int j = 0; j <= deg; ++j) if (bn[j]) AB_FLOPS((double)(j + 1) * (4.0 * bn[j] + 3.0 + 3.0 + 5.0 * bn[j] * 0.5));   // _scal + _add2 per stream (about half the terms coincide)
#endif
    // error chain in the reference's stream order; an empty bucket contributes nothing
    double e = tm.e;
    #pragma unroll 1
    for (int i = 0; i <= deg; ++i)
        #pragma unroll 1
        for (int j = i; j <= deg; ++j) {
            if (bn[j] == 0) continue;
            const double ts = TS[i * (deg + 1) + j], te = TE[i * (deg + 1) + j];
            e = add_ru(add_ru(e, ts / 2), te / 2);       // t.error() = ts/2 (zero for the exact scalings), then _add2's chain
        }
    wsync();
    res.off = m0;
    res.n = dense ? dacc_emit(c, DA, m0) : acc_emit(c, T, tm.as, deg, m0);      // the buckets are no longer needed: the result lands at m0
    res.e = e;
    c.A.top = m0 + res.n * AB_TW;
    arena_mark_peak(c);
    c.rnd = RN;
    return res;
}

// partial_evaluate(x,k,c) for c==1 (src/taylor_model.cpp:2403-2423). Result at the arena top recorded on entry.
AB_HDN TM tm_partial_evaluate_one(Ctx& c, const TM& x, int k) {
    const int m0 = c.A.top;
    const int deg = tm_degree(c, x);
    int boff[AB_MAXPOW], bn[AB_MAXPOW];
    if (deg + 1 > AB_MAXPOW) { fail(c, ST_ARENA); return tm_begin(c, x.as - 1, x.acc); }
    #pragma unroll 1
    for (int b = 0; b <= deg; ++b) {
        boff[b] = c.A.top;
        TM bt = tm_begin(c, x.as - 1, x.acc);
        AB_FILTER_COPY(c, bt, x, key_get(a, k) == b, key_remove(a, k, x.as), v);
        bn[b] = bt.n;
    }
    const int mr = c.A.top;
    TM r; r.off = boff[0]; r.n = bn[0]; r.as = x.as - 1; r.e = x.e; r.acc = x.acc;
    r = tm_copy(c, r);
    r.e = x.e;
    #pragma unroll 1
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
    #pragma unroll 1
    for (int i = 0; i < n; ++i) { bx[i] = ibx[i]; if (bx[i].l == bx[i].u) bx[i] = ivl_add(bx[i], ivl(-EPS, EPS)); }
    #pragma unroll 1
    for (int i = 0; i < n; ++i) dom[i] = bx[i];
    dom[n] = ivl(0.0, h);
    TM xs[AB_MAXN], usvf[AB_MAXN], yz[AB_MAXN], y[AB_MAXN], ny[AB_MAXN], outs[AB_MAXN];
    // vector-field model on bb: tape over scalings(bb), effectively at the default accuracy
    cta_align(c, 3);
    #pragma unroll 1
    for (int j = 0; j < n; ++j) xs[j] = tm_scaling(c, n, j, bb[j]);
    #pragma unroll 1
    for (int i = 0; i < n; ++i) {
        usvf[i] = eval_tm(c, ops, vf[i], xs);
        // unit_scaled_vf[i]=vf[i]*(hmag/rad_ivl(b[i]))
        tm_scal(c, usvf[i], ivl_div(h, rad_ivl(bb[i]), &c.status));
    }
    vec_slide(c, usvf, n, m0);
    #pragma unroll 1
    for (int i = 0; i < n; ++i) AB_DUMP(c, "usvf", i, usvf[i]);
    if (c.status != ST_OK) { for (int i = 0; i < n; ++i) flow[i] = tm_begin(c, n + 1, acc_default()); return; }
    // y0=scalings(d); unit_scaled_y0=embed(unscale(y0,b),1)
    #pragma unroll 1
    for (int i = 0; i < n; ++i) {
        int m = c.A.top;
        TM y0 = tm_scaling(c, n, i, bx[i]);
        TM u = tm_unscale(c, y0, bb[i]);
        tm_slide(c, u, m);
        u.as = n + 1;
        yz[i] = u;
    }
    #pragma unroll 1
    for (int i = 0; i < n; ++i) AB_DUMP(c, "yz", i, yz[i]);
    // model-level flow: y[i]=constant(n+1, yz[i].range()+vf[i].range()*[-1,1])
    const int ybase = c.A.top;
    #pragma unroll 1
    for (int i = 0; i < n; ++i) {
        Ivl rg = ivl_add(tm_range(c, yz[i]), ivl_mul(tm_range(c, usvf[i]), ivl(-1, 1)));
        y[i] = tm_const(c, n + 1, 1.0, acc_default());
        tm_scal(c, y[i], rg);
    }
    #pragma unroll 1
    for (int i = 0; i < n; ++i) AB_DUMP(c, "y0", i, y[i]);
    #pragma unroll 1
    for (int it = 0; it < order; ++it) {
        cta_align(c, it < 8 ? 4 + 3 * it : 28);
        tm_compose(c, usvf, n, y, n, ny);                 // new_y=compose(vf,y)
        cta_align(c, it < 8 ? 5 + 3 * it : 28);
        int pout = c.A.top;
        #pragma unroll 1
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
        #pragma unroll 1
        for (int i = 0; i < n; ++i) { Acc a = y[i].acc; y[i] = outs[i]; y[i].acc = a; }
        vec_slide(c, y, n, ybase);
        #pragma unroll 1
        for (int i = 0; i < n; ++i) AB_DUMP(c, "y", (it + 1) * 10 + i, y[i]);
        if (c.status != ST_OK) break;
    }
    cta_align(c, 29);
    #pragma unroll 1
    for (int i = 0; i < n; ++i) tm_sweep(c, y[i], 0.0);   // leaves the ambient mode upward
    // model_flow=scale(unit_scaled_flow,b); then split(model_flow,n,true) because the time interval is [0,h]
    int pout = c.A.top;
    #pragma unroll 1
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
    #pragma unroll 1
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
    #pragma unroll 1
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
    #pragma unroll 1
    for (int j = 0; j < p; ++j) ys[j] = tm_variable(c, p, j);      // BaseModelType::variables(p): default accuracy
    ys[p] = tau;
    tm_compose(c, flow_set, n, ys, p + 1, out);
    vec_slide(c, out, n, m0);
    #pragma unroll 1
    for (int i = 0; i < n; ++i) out[i].acc = spacial;
}

// reachability_step(flow_set_model, t0, t1) src/taylor_calculus.cpp:248-267. Output: n models in p+1 variables.
AB_HDN void reachability_step(Ctx& c, const TM* flow_set, int n, int p, const TM& t0, const TM& t1, Acc spacial, TM* out) {
    AB_SECTION(6);
    const int m0 = c.A.top;
    TM ys[AB_MAXARGS + 1];
    #pragma unroll 1
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
    #pragma unroll 1
    for (int i = 0; i < n; ++i) out[i].acc = spacial;
}

// uniform_error_recondition (src/taylor_set.cpp:499-529). set: n models in p variables, rebuilt at the arena top
// recorded on entry. Returns the new number of parameters.
AB_HDN int uniform_error_recondition(Ctx& c, TM* set, int n, int p, const double* thresholds) {
    int idx[AB_MAXN], k = 0;
    #pragma unroll 1
    for (int i = 0; i < n; ++i) if (set[i].e > thresholds[i]) idx[k++] = i;
    if (k == 0) return p;
    const int m0 = c.A.top;
    const int np = p + k;
    TM out[AB_MAXN];
    #pragma unroll 1
    for (int i = 0; i < n; ++i) {
        int which = -1;
        #pragma unroll 1
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
    #pragma unroll 1
    for (int i = 0; i < n; ++i) set[i] = out[i];
    (void)m0;
    return np;
}

// recondition(tm, discarded, n_err, index_of_error) src/taylor_model.cpp:3882-3956. Result at the arena top.
AB_HDN TM tm_recondition(Ctx& c, const TM& tm, const int* discarded, int nd, int n_err, int index_of_error) {
    const int nv = tm.as, nk = nv - nd;
    int kept[AB_MAXARGS + 1], kk = 0;
    #pragma unroll 1
    for (int j = 0; j < nv; ++j) { bool dsc = false; for (int q = 0; q < nd; ++q) if (discarded[q] == j) dsc = true; if (!dsc) kept[kk++] = j; }
    double error = tm.e;
    // pass 1 (reference iterates the terms in REVERSE order): accumulate discarded magnitudes in the ambient mode
    #pragma unroll 1
    for (int i = tm.n - 1; i >= 0; --i) {
        Key a = tm_key(c, tm, i);
        #pragma unroll 1
        for (int q = 0; q < nd; ++q) if (key_get(a, discarded[q]) != 0) { error = add_m(c.rnd, error, dabs(tm_val(c, tm, i))); break; }
    }
    TM r = tm_begin(c, nk + n_err, tm.acc);
    bool added = false;
    if (n_err == index_of_error) { r.e = error; added = true; }
    Key ra_error = key_zero();
    if (n_err != index_of_error) key_inc(ra_error, nk + index_of_error, 1);
    int last_kept = -1;
    #pragma unroll 1
    for (int i = 0; i < tm.n; ++i) {
        Key a = tm_key(c, tm, i);
        bool keep = true;
        #pragma unroll 1
        for (int q = 0; q < nd; ++q) if (key_get(a, discarded[q]) != 0) { keep = false; break; }
        if (keep) last_kept = i;
    }
    #pragma unroll 1
    for (int i = 0; i < tm.n; ++i) {
        Key a = tm_key(c, tm, i);
        bool keep = true;
        #pragma unroll 1
        for (int q = 0; q < nd; ++q) if (key_get(a, discarded[q]) != 0) { keep = false; break; }
        if (!keep) continue;
        Key ra = key_zero();
        #pragma unroll 1
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
    #pragma unroll 1
    for (int i = 0; i < n; ++i) for (int j = 0; j < p; ++j) dep[i][j] = 0.0;
    #pragma unroll 1
    for (int i = 0; i < n; ++i)
        #pragma unroll 1
        for (int t = 0; t < set[i].n; ++t) {
            Key a = tm_key(c, set[i], t); double av = dabs(tm_val(c, set[i], t));
            #pragma unroll 1
            for (int j = 0; j < p; ++j) if (key_get(a, j) != 0) dep[i][j] = add_m(c.rnd, dep[i][j], av);
        }
    double cm[AB_MAXARGS + 1]; int ci[AB_MAXARGS + 1];
    #pragma unroll 1
    for (int j = 0; j < p; ++j) { cm[j] = 0.0; ci[j] = j; for (int i = 0; i < n; ++i) cm[j] = dmax(cm[j], dep[i][j]); }
    // std::sort(..., std::greater<Pair<double,uint>>): descending by (value, index)
    #pragma unroll 1
    for (int a = 1; a < p; ++a) {
        double v = cm[a]; int id = ci[a]; int b = a - 1;
        #pragma unroll 1
        while (b >= 0 && (cm[b] < v || (cm[b] == v && ci[b] < id))) { cm[b + 1] = cm[b]; ci[b + 1] = ci[b]; --b; }
        cm[b + 1] = v; ci[b + 1] = id;
    }
    #pragma unroll 1
    for (int j = 0; j < nd; ++j) discarded[j] = ci[n_kept + j];
    #pragma unroll 1
    for (int a = 1; a < nd; ++a) { int v = discarded[a], b = a - 1; while (b >= 0 && discarded[b] > v) { discarded[b + 1] = discarded[b]; --b; } discarded[b + 1] = v; }
    TM out[AB_MAXN];
    #pragma unroll 1
    for (int i = 0; i < n; ++i) out[i] = tm_recondition(c, set[i], discarded, nd, n_err, i);
    #pragma unroll 1
    for (int i = 0; i < n; ++i) set[i] = out[i];
    return nd;
}

// TaylorIntegrator::flow (src/integrator.cpp:116-152): flow_bounds, then `order` Picard iterations on function level over the
// domain dx x [-h,h]: phi <- antiderivative(compose(f,phi), time) + phi0 with phi0 = the coordinate functions and phi
// starting from the constant bounding box. No intersection with the previous iterate, no sweep beyond the models' own
// (default) accuracy. Writes the n models (n+1 variables) packed at the arena top recorded on entry; returns h.
AB_HDN double integrator_flow(Ctx& c, const TapeOp* ops, const Func* vf, int n, const Ivl* dx, double hmax, int order, TM* phi, Ivl* dom) {
    const int m0 = c.A.top;
    Ivl bx[AB_MAXN];
    const double h = flow_bounds(c, ops, vf, n, dx, hmax, bx, true);
    for (int i = 0; i < n; ++i) dom[i] = dx[i];
    dom[n] = ivl(-h, h);
    if (c.status != ST_OK) { for (int i = 0; i < n; ++i) phi[i] = tm_begin(c, n + 1, acc_default()); return h; }
    TM phi0[AB_MAXN], fphi[AB_MAXN], nphi[AB_MAXN];
    for (int i = 0; i < n; ++i) phi0[i] = tm_scaling(c, n + 1, i, dom[i]);                     // ScalarTaylorFunction::variable(dom,i)
    const int pbase = c.A.top;
    for (int i = 0; i < n; ++i) { phi[i] = tm_const(c, n + 1, 1.0, acc_default()); tm_scal(c, phi[i], bx[i]); }   // constant(dom,bx[i])
    const Ivl sf = rad_ivl(dom[n]);
    #pragma unroll 1
    for (int k = 0; k < order; ++k) {
        for (int i = 0; i < n; ++i) fphi[i] = eval_tm(c, ops, vf[i], phi);                   // compose(f,phi) = f.evaluate(phi.models())
        for (int i = 0; i < n; ++i) {
            TM w = tm_copy(c, fphi[i]);
            tm_antidiff(c, w, n);                                                             // antiderivative(model,n) ...
            tm_scal(c, w, sf);                                                                // ... * rad_ivl(dom[n])
            nphi[i] = tm_add(c, w, phi0[i]);                                                  // + phi0[i]
        }
        for (int i = 0; i < n; ++i) phi[i] = nphi[i];
        vec_slide(c, phi, n, pbase);
        if (c.status != ST_OK) break;
    }
    vec_slide(c, phi, n, m0);
    return h;
}

} } // namespace ab200
