// [lane build] One frontier cell per LANE: the throughput build of the evolution kernel, used only for frontiers far larger
// than the device can hold one-warp-per-cell (ab200_capi.cu picks it above AB_LANE_THRESHOLD cells). Same algorithm and
// arithmetic as the warp build one directory up, written as sequential per-lane code over lane-interleaved scratch arenas.
// ab200_events.cuh -- guard / invariant tests and blocking-time selection for one step.
// Reference: src/hybrid_evolver-image.cpp:581-681 (initially active events, discard test), :770-883 (blocking times),
// :1085-1127 (_compute_blocking_info); src/taylor_calculus.cpp:603-611 and include/calculus_base.h:88-99,295-311
// (active).
#pragma once
#include "ab200_calc.cuh"

namespace ab200 { inline namespace AB_VARIANT {

struct WorkingSet;

struct BlockingInfo {
    TM time;           // blocking_time_model (scaled to [0,1])
    bool finishing;    // blocking_events == {finishing_event}
    int n_blocking;    // number of blocking events other than finishing
    int blocking[AB_MAXEV];
    int n_nontransverse;
    int nontransverse[AB_MAXEV];
};

enum { TRI_FALSE = 0, TRI_TRUE = 1, TRI_INDET = 2 };
AB_HD int tribool_of(Ivl r) { return r.l > 0 ? TRI_TRUE : (r.u < 0 ? TRI_FALSE : TRI_INDET); }   // calculus_base.h:88-99
AB_HD bool tri_possibly(int t) { return t != TRI_FALSE; }
AB_HD bool tri_definitely(int t) { return t == TRI_TRUE; }

AB_HD bool is_urgent(const EventDesc& e) { return e.kind == EV_INVARIANT || e.kind == EV_URGENT; }

// active(guard,set_model): range of the guard tape over the set's models (include/calculus_base.h:306-311)
AB_HDN int guard_active(Ctx& c, const SystemDesc& S, Func g, const TM* models) {
    const int m = c.A.top;
    TM gm = eval_tm(c, S.ops, g, models);
    Ivl r = tm_range(c, gm);
    c.A.top = m;
    return tribool_of(r);
}

// ---- implicit crossing-time models ------------------------------------------------------------------------------------
// Scalar case (one guard): f has `has`+1 variables, the last one is the scaled time; h(x) solves f(x,h(x))=0.

// jacobian2_range (src/taylor_model.cpp:3106-3126), rs=1
AB_HDN Ivl tm_jacobian2_range(Ctx& c, const TM& f) {
    const int k = f.as - 1;
    Ivl J = ivl(0.0, 0.0);
    for (int t = 0; t < f.n; ++t) {
        Key a = tm_key(c, f, t);
        int e = key_get(a, k);
        if (e > 0) {
            double x = tm_val(c, f, t);
            if (key_deg(a) == 1) J = ivl_add(J, x);
            else J = ivl_add(J, ivl_mul(ivl_mul(ivl(-1, 1), x), (double)e));
        }
    }
    return J;
}
// jacobian2(f,x) (src/taylor_model.cpp:2990-3006): derivative w.r.t. the last variable at the interval point x, by
// evaluating the expansion over degree-1 Differential<Interval> numbers (include/expansion.h:507-530,
// include/differential.h:367-410,550-568). (v,g) = value and derivative; absent terms behave as [0,0].
AB_HDN Ivl tm_jacobian2_at(Ctx& c, const TM& f, const Ivl* x) {
    const int as = f.as, k = as - 1;
    Ivl rv = ivl_mul(x[0], 0.0), rg = ivl(0.0, 0.0);          // zero = y[0]; zero*=0
    const Ivl onev = ivl_add(rv, 1.0);                          // one = zero; one+=1
    for (int t = 0; t < f.n; ++t) {
        Key a = tm_key(c, f, t);
        Ivl tv = onev, tg = ivl(0.0, 0.0);
        for (int j = 0; j < as; ++j) {
            int e = key_get(a, j);
            for (int l = 0; l < e; ++l) {
                // t=t*y[j] with y[j]=(x[j], j==k ? 1 : absent)
                Ivl nv = ivl_mul(tv, x[j]);
                Ivl ng = (j == k) ? ivl_add(ivl_mul(tv, ivl(1.0)), ivl_mul(tg, x[j])) : ivl_mul(tg, x[j]);
                tv = nv; tg = ng;
            }
        }
        double cf = tm_val(c, f, t);
        tv = ivl_mul(tv, cf); tg = ivl_mul(tg, cf);             // t*=c
        rv = ivl_add(rv, tv); rg = ivl_add(rg, tg);             // r+=t
    }
    return rg;
}
// refines(tv1,tv2) src/taylor_model.cpp:2602-2611
AB_HDN bool tm_refines(Ctx& c, const TM& t1, const TM& t2) {
    const int m = c.A.top;
    TM d = t2; d.e = 0.0;
    TM n1 = tm_neg(c, t1);
    TM s = tm_add2(c, d, n1, t2.acc);
    double nm = tm_norm(c, s);
    c.A.top = m;
    return nm <= t2.e;
}
// disjoint(tv1,tv2) src/taylor_model.cpp:2612-2627
AB_HDN bool tm_disjoint(Ctx& c, const TM& t1, const TM& t2) {
    const int m = c.A.top;
    TM a = t1; a.e = 0.0; TM b = t2; b.e = 0.0;
    TM d = tm_sub(c, b, a);
    double nm = tm_norm(c, d);
    c.A.top = m;
    return nm > add_ru(t1.e, t2.e);
}
// implicit(f) src/taylor_model.cpp:3562-3665 with _implicit5 (:3496-3558), scalar case. Returns false where the reference
// throws ImplicitFunctionException (singular Jacobian, disjoint iterate, contraction not verified). The solution is left at
// the arena top recorded on entry.
AB_HDN bool tm_implicit(Ctx& c, const TM& f, TM& out) {
    const int m0 = c.A.top;
    const int fas = f.as, has = fas - 1;
    // D2f=jacobian2_range(f); inverse throws if it contains zero
    { Ivl D = tm_jacobian2_range(c, f); AB_DBG("implicit: J2range %a %a\n", D.l, D.u); if (D.l <= 0 && D.u >= 0) return false; }
    // ---- _implicit5(f,6)
    TM ys[AB_MAXARGS + 1];
    for (int j = 0; j < has; ++j) ys[j] = tm_variable(c, has, j);
    // D2f value = coefficient of the time variable
    double d2f = 0.0;
    { Key u = key_unit(has); for (int t = 0; t < f.n; ++t) if (key_eq(tm_key(c, f, t), u)) d2f = tm_val(c, f, t); }
    // g=-f with that coefficient set to 0 (the entry is inserted if absent), then g=J*g, clean
    TM g = tm_begin(c, fas, acc_default());
    {
        Key u = key_unit(has); bool done = false;
        for (int t = 0; t < f.n; ++t) {
            Key a = tm_key(c, f, t);
            if (!done && !key_less(a, u)) { if (key_eq(a, u)) { tm_append(c, g, a, 0.0); done = true; continue; } tm_append(c, g, u, 0.0); done = true; }
            tm_append(c, g, a, -tm_val(c, f, t));
        }
        if (!done) tm_append(c, g, u, 0.0);
        g.e = f.e;
    }
    const double J = div_m(c.rnd, 1.0, d2f);
    tm_scal(c, g, J);                                            // J(0,0)*g[0]; t+=... keeps the value, error unchanged
    g.e = add_ru(add_ru(0.0, g.e), 0.0);
    c.rnd = RN;
    tm_clean(c, g, acc_default());
    // h is the last persistent model on the arena: it grows in place while everything above it is scratch
    TM h = tm_const(c, has, 1.0, acc_default());
    tm_scal(c, h, ivl(-1, 1));                                   // constants(has,[-1,1])
    bool contracting = false;
    for (int k = 0; k < 6; ++k) {
        const int mk = c.A.top;
        ys[has] = h;
        TM hn;
        tm_compose(c, &g, 1, ys, has + 1, &hn);
        if (!contracting) {
            if (tm_refines(c, hn, h)) contracting = true;
            AB_DUMP(c, "hn", k, hn); AB_DBG("implicit: k=%d contracting=%d\n", k, (int)contracting);
            if (tm_disjoint(c, hn, h)) { AB_DBG("implicit: disjoint\n"); c.A.top = m0; return false; }
        }
        TM hi;
        if (!tm_intersection(c, h, hn, hi)) { c.rnd = RN; fail(c, ST_IMPLICIT); c.A.top = m0; return false; }
        // keep hi in place of h (h sits just below mk)
        tm_slide(c, hi, h.off);
        h = hi;
        (void)mk;
    }
    if (!contracting) { AB_DBG("implicit: not contracting\n"); c.A.top = m0; return false; }
    // ---- three Newton refinements
    for (int it = 0; it < 3; ++it) {
        Ivl x[AB_MAXARGS + 1];
        for (int j = 0; j < has; ++j) x[j] = ivl(-1, 1);
        x[has] = tm_range(c, h);
        Ivl D = tm_jacobian2_at(c, f, x);
        if (D.l <= 0 && D.u >= 0) { c.A.top = m0; return false; }
        Ivl Dinv = ivl_mul(ivl(1.0), ivl_div(1.0, D, &c.status));
        h.e = 0.0;                                               // clobber(h)
        ys[has] = h;
        TM fh;
        tm_compose(c, &f, 1, ys, has + 1, &fh);
        // dh=prod(D2finv,fidh): t=TaylorModel(0); t+=Dinv*fidh  -> fresh model, same expansion, error rounded through _add2
        TM dh = tm_copy(c, fh);
        tm_scal(c, dh, Dinv);
        dh.e = add_ru(add_ru(0.0, dh.e), 0.0);
        c.rnd = RN;
        // h-=dh: _acc(h,neg(dh)) keeps h's accuracy
        TM nd = tm_neg(c, dh);
        TM hs = tm_add2(c, h, nd, h.acc);
        tm_slide(c, hs, h.off);
        h = hs;
    }
    tm_slide(c, h, m0);
    out = h;
    return true;
}

// TaylorModel::evaluate(Vector<Interval>) src/taylor_model.cpp:1650-1666
AB_HDN Ivl tm_evaluate(Ctx& c, const TM& f, const Ivl* v) {
    Ivl r = ivl_mul(ivl(-1, 1), f.e);
    for (int t = 0; t < f.n; ++t) {
        Key a = tm_key(c, f, t);
        Ivl tv = ivl(tm_val(c, f, t));
        for (int j = 0; j < f.as; ++j) tv = ivl_mul(tv, ivl_pow(v[j], (unsigned)key_get(a, j)));
        r = ivl_add(r, tv);
    }
    return r;
}
// scaled_touching_time_interval (src/taylor_calculus.cpp:524-578): five bisections each for the first time the guard
// may vanish and the first time it is definitely positive.
AB_HDN Ivl touching_time_interval(Ctx& c, const TM& gf) {
    const int as = gf.as;
    Ivl x[AB_MAXARGS + 1];
    for (int j = 0; j < as - 1; ++j) x[j] = ivl(-1, 1);
    const double tmin = 0.0, tmax = 1.0;
    x[as - 1] = ivl(tmin);
    Ivl init = tm_evaluate(c, gf, x);
    double lower_time = tmin, upper_time = tmax;
    if (!(init.u < 0 || init.l > 0)) lower_time = tmin;                  // possibly(initial_guard_range==0)
    else {
        double mn = tmin, mx = tmax;
        for (int i = 0; i < 5; ++i) {
            double nt = add_rn(mn, mx) / 2;
            x[as - 1] = ivl(tmin, nt);
            Ivl gr = tm_evaluate(c, gf, x);
            if (!(gr.u < 0 || gr.l > 0)) mx = nt; else { mn = nt; lower_time = nt; }
        }
    }
    if (init.l > 0 || init.u < 0) upper_time = tmax;
    else {
        double mn = tmin, mx = tmax;
        for (int i = 0; i < 5; ++i) {
            double nt = add_rn(mn, mx) / 2;
            x[as - 1] = ivl(nt);
            if (tm_evaluate(c, gf, x).l > 0.0) { mx = nt; upper_time = nt; } else mn = nt;
        }
    }
    return ivl(lower_time, upper_time);
}

// time<blocking_time for TaylorModels: (sd-*this)>0  (include/taylor_model.h:290-295, include/numeric.h:579-583)
AB_HDN int tm_less(Ctx& c, const TM& time, const TM& blocking) {
    const int m = c.A.top;
    TM d = tm_sub(c, blocking, time);
    Ivl r = tm_range(c, d);
    c.A.top = m;
    return r.l > 0 ? TRI_TRUE : (r.u <= 0 ? TRI_FALSE : TRI_INDET);
}

// _compute_blocking_info (src/hybrid_evolver-image.cpp:1085-1127) with
// compute_eventBlockingTimes_and_nonTransverseEvents (:770-843) and compute_blockingTime_and_relatedEvents (:846-883).
AB_HDN void compute_blocking_info(Ctx& c, const SystemDesc& S, const LocDesc& L, const TM* fset, int n, int p,
                                  const TM& tstep, Acc spacial, BlockingInfo& B) {
    (void)spacial;
    B.time = tstep; B.finishing = true; B.n_blocking = 0; B.n_nontransverse = 0;
    bool any = false;
    for (int e = 0; e < L.n_events; ++e) if (is_urgent(L.ev[e])) any = true;
    if (!any) return;
    // event_blocking_times in std::map<DiscreteEvent,...> order: events sorted by name, the finishing event "e-2" among them
    TM times[AB_MAXEV + 1]; int who[AB_MAXEV + 1]; int nt = 0;      // who: event index, -1 = finishing
    const int m = c.A.top;
    TM pos[AB_MAXN];
    for (int i = 0; i < n; ++i) pos[i] = tm_split(c, fset[i], n, true);   // split(models,dimension,1) (sic)
    bool fin_done = false, have_fin_override = false; TM fin_override = tstep; int fin_slot = -1;
    for (int e = 0; e <= L.n_events; ++e) {
        if (!fin_done && e == L.finishing_rank) { times[nt] = tstep; who[nt] = -1; fin_slot = nt; ++nt; fin_done = true; }
        if (e == L.n_events) break;
        if (!is_urgent(L.ev[e])) continue;
        int act = guard_active(c, S, L.ev[e].guard, pos);
        if (c.status != ST_OK) return;
        if (!tri_possibly(act)) continue;
        TM gf = eval_tm(c, S.ops, L.ev[e].guard, fset);                // apply(guard,flow_set_model)[0]
        gf.as = p + 1;
        TM cross;
        AB_DUMP(c, "gf", e, gf);
        if (tm_implicit(c, gf, cross)) {
            AB_DUMP(c, "cross", e, cross);
            Ivl nd = tm_jacobian2_range(c, gf);
            if (!(nd.l > 0 || nd.u < 0)) { fail(c, ST_BOUNDER); return; }      // assert in the reference
            if (nd.l > 0) { times[nt] = cross; who[nt] = e; ++nt; }
        } else {
            if (c.status != ST_OK) return;
            // DegenerateCrossingException: non-transverse crossing (:803-838)
            Ivl tti = touching_time_interval(c, gf);
            if (tti.u >= 0 && tti.l <= 1.0) {
                TM fin[AB_MAXN];
                for (int i = 0; i < n; ++i) fin[i] = tm_partial_evaluate_one(c, fset[i], n);   // partial_evaluate(models,dimension,1.0) (sic)
                int fa = guard_active(c, S, L.ev[e].guard, fin);
                if (c.status != ST_OK) return;
                if (fa == TRI_TRUE) {
                    TM tt = tm_const(c, p, 1.0, spacial); tm_scal(c, tt, tti);             // time_model(Interval,space_domain)
                    times[nt] = tt; who[nt] = e; ++nt;
                } else if (fa == TRI_INDET && tti.l > 1. / 16) {
                    TM lt = tm_begin(c, p, spacial); tm_append(c, lt, key_zero(), tti.l);  // time_model(lower,space_domain)
                    // event_blocking_times[finishing_event]=lower_touching_time_model
                    fin_override = lt; have_fin_override = true;
                } else B.nontransverse[B.n_nontransverse++] = e;
            }
        }
    }
    if (!fin_done) { times[nt] = tstep; who[nt] = -1; fin_slot = nt; ++nt; }
    if (have_fin_override) times[fin_slot] = fin_override;
    // compute_blockingTime_and_relatedEvents
    unsigned bset = 1u << 0; TM bt = times[0];
    for (int k = 1; k < nt; ++k) {
        int first = tm_less(c, times[k], bt);
        if (first == TRI_TRUE) { bset = 1u << k; bt = times[k]; }
        else if (first == TRI_INDET) {
            unsigned nset = 1u << k;
            for (int q = 0; q < k; ++q) if ((bset >> q) & 1u) { if (tri_possibly(tm_less(c, times[q], times[k]))) nset |= 1u << q; }
            // blocking_time.swap(new_blocking_time): expansion and error of `time`, accuracy kept
            Acc a = bt.acc; bt = times[k]; bt.acc = a;
            bset = nset;
        }
    }
    int count = 0; for (int k = 0; k < nt; ++k) if ((bset >> k) & 1u) ++count;
    const double SMALL = 1. / 16;
    Ivl btr = tm_range(c, bt);
    bt = tm_copy(c, bt);
    if (count != 1 && btr.u > SMALL) {
        bset = 0; B.finishing = true; B.n_blocking = 0;
        if (btr.l > 0) { tm_acc(c, bt, -bt.e); bt.e = 0.0; }
        tm_scal(c, bt, 1 - SMALL);
    } else {
        B.finishing = false;
        for (int k = 0; k < nt; ++k) if ((bset >> k) & 1u) { if (who[k] < 0) B.finishing = true; else B.blocking[B.n_blocking++] = who[k]; }
    }
    tm_slide(c, bt, m);
    B.time = bt;
}

// is_enclosure_to_be_discarded (:652-681): a blocking event definitely active on the starting set; with LOWER semantics
// also a possibly active one whose guard may be crossed in the positive direction (has_nonnegative_crossing :612-632,
// positively_crossing :637-650: the Lie derivative of the guard over the set's bounding box).
AB_HDN bool enclosure_to_be_discarded(Ctx& c, const SystemDesc& S, const LocDesc& L, const TM* set, int n, int p, int lower, int backward) {
    (void)p;
    int blocking_active = TRI_FALSE;
    for (int e = 0; e < L.n_events; ++e) {
        if (!is_urgent(L.ev[e])) continue;
        Ivl inf[AB_MAXN];
        for (int i = 0; i < n; ++i) inf[i] = ivl(-DBL_MAX, DBL_MAX);
        int st = c.status;
        Ivl gr = eval_ivl(c, S.ops, L.ev[e].guard, inf);
        c.status = st;                                   // the reference evaluates without range checks here
        if (tri_definitely(tribool_of(gr))) continue;    // "everywhere active": ignored (:596-598)
        int act = guard_active(c, S, L.ev[e].guard, set);
        if (tri_possibly(act)) {
            // blocking_event_initially_active = blocking_event_initially_active || initially_active  (Kleene or)
            if (act == TRI_TRUE || blocking_active == TRI_TRUE) blocking_active = TRI_TRUE;
            else if (act == TRI_INDET || blocking_active == TRI_INDET) blocking_active = TRI_INDET;
        }
    }
    if (blocking_active == TRI_TRUE) return true;
    if (blocking_active == TRI_INDET && lower) {
        Ivl bb[AB_MAXN];
        for (int i = 0; i < n; ++i) bb[i] = tm_range(c, set[i]);
        for (int e = 0; e < L.n_events; ++e) {
            if (!is_urgent(L.ev[e])) continue;
            int act = tribool_of(eval_ivl(c, S.ops, L.ev[e].guard, bb));
            if (L.ev[e].lie.end <= L.ev[e].lie.begin) { fail(c, ST_UNSUPPORTED); return true; }   // lie_derivative throws in the reference
            Ivl d = eval_ivl(c, S.ops, L.ev[e].lie, bb);
            if (backward) d = ivl_neg(d);                 // Lie derivative along dynamic*(-1)
            int pc = d.l > 0 ? TRI_TRUE : (d.u < 0 ? TRI_FALSE : TRI_INDET);
            if (tri_possibly(act) && tri_possibly(pc)) return true;
        }
    }
    return false;
}
AB_HD bool blocking_has_transition(const LocDesc& L, const BlockingInfo& B) {
    for (int k = 0; k < B.n_blocking; ++k) if (L.ev[B.blocking[k]].kind != EV_INVARIANT) return true;
    return false;
}

// ---- activation times of non-blocking events -----------------------------------------------------------------------
// _compute_activation_info + compute_activationTimes (src/hybrid_evolver-image.cpp:884-975,1130-1157), UPPER semantics:
// permissive events, plus urgent transitions whose crossing was non-transverse, get a [lower,upper] window of scaled
// times during which the jump may be taken. Events are visited in name order (std::map<DiscreteEvent,..>).
struct Activation { int ev; TM lower, upper; };

AB_HDN int compute_activation_times(Ctx& c, const SystemDesc& S, const LocDesc& L, const BlockingInfo& B, const TM* fset,
                                    int n, int p, int lower, Activation* act) {
    bool wanted[AB_MAXEV]; bool any = false;
    for (int e = 0; e < L.n_events; ++e) { wanted[e] = (L.ev[e].kind == EV_PERMISSIVE); any = any || wanted[e]; }
    for (int k = 0; k < B.n_nontransverse; ++k) if (L.ev[B.nontransverse[k]].kind != EV_INVARIANT) { wanted[B.nontransverse[k]] = true; any = true; }
    const int m0 = c.A.top;
    // The reference builds the initial/final set models even when the activation map is empty; their only lasting
    // effect then is the rounding mode left by partial_evaluate's additions (round-to-nearest once a model has degree
    // >= 1), so without candidates only that is reproduced.
    if (!any) {
        for (int i = 0; i < n; ++i) if (tm_degree(c, fset[i]) >= 1) c.rnd = RN;
        return 0;
    }
    TM init[AB_MAXN], fin[AB_MAXN];
    for (int i = 0; i < n; ++i) {                                   // partial_evaluate(models,p,0.0): terms free of the time variable
        TM r = tm_begin(c, p, fset[i].acc);
        for (int t = 0; t < fset[i].n; ++t) { Key a = tm_key(c, fset[i], t); if (key_get(a, p) == 0) tm_append(c, r, key_remove(a, p, p + 1), tm_val(c, fset[i], t)); }
        r.e = fset[i].e;
        init[i] = r;
    }
    for (int i = 0; i < n; ++i) fin[i] = tm_partial_evaluate_one(c, fset[i], p);
    TM zero = B.time; zero.n = 0; zero.e = 0.0;                     // blocking_time_model*0.0
    int na = 0;
    for (int e = 0; e < L.n_events && c.status == ST_OK; ++e) {
        if (!wanted[e]) continue;
        int active = guard_active(c, S, L.ev[e].guard, fset);
        if (active == TRI_TRUE) { act[na].ev = e; act[na].lower = zero; act[na].upper = B.time; ++na; continue; }
        if (active == TRI_FALSE) continue;
        int ia = guard_active(c, S, L.ev[e].guard, init);
        int fa = guard_active(c, S, L.ev[e].guard, fin);
        // crossing_time (:548-566): implicit crossing, or the touching-time interval for a degenerate crossing
        TM gf = eval_tm(c, S.ops, L.ev[e].guard, fset);
        gf.as = p + 1;
        TM cross;
        if (!tm_implicit(c, gf, cross)) {
            if (c.status != ST_OK) break;
            Ivl tti = touching_time_interval(c, gf);
            cross = tm_const(c, p, 1.0, fset[0].acc); tm_scal(c, cross, tti);
        }
        TM lo = tm_copy(c, cross); tm_acc(c, lo, -cross.e); lo.e = 0.0;
        TM up = tm_copy(c, cross); tm_acc(c, up, cross.e); up.e = 0.0;
        act[na].ev = e;
        if (!lower) {
            act[na].lower = (ia == TRI_FALSE) ? lo : zero;
            act[na].upper = (fa == TRI_FALSE) ? up : B.time;
            ++na;
        } else {
            // LOWER semantics (:927-969): the window shrinks to what is certainly active, and is kept only if non-empty
            act[na].lower = (ia == TRI_TRUE) ? zero : up;
            act[na].upper = (fa == TRI_FALSE) ? lo : ((fa == TRI_TRUE) ? B.time : up);
            if (tm_less(c, act[na].lower, act[na].upper) == TRI_TRUE) ++na;
        }
    }
    if (c.status != ST_OK) return 0;
    // keep only the windows: copy them to the top in order, then slide the block down to m0
    TM keep[2 * AB_MAXEV];
    for (int k = 0; k < na; ++k) { keep[2 * k] = tm_copy(c, act[k].lower); keep[2 * k + 1] = tm_copy(c, act[k].upper); }
    vec_slide(c, keep, 2 * na, m0);
    for (int k = 0; k < na; ++k) { act[k].lower = keep[2 * k]; act[k].upper = keep[2 * k + 1]; }
    return na;
}

// reachability_time (src/taylor_calculus.cpp:278-296): t0 + s*(t1-t0) with s the new last variable scaled to [0,1].
// Result (p+1 variables) at the arena top recorded on entry.
AB_HDN TM reachability_time(Ctx& c, const TM& t0, const TM& t1, int p, Acc spacial) {
    const int m0 = c.A.top;
    TM a = t0; a.as = p + 1;
    TM b = t1; b.as = p + 1;
    TM s = tm_scaling(c, p + 1, p, ivl(0.0, 1.0));
    TM d = tm_sub(c, b, a);
    TM sd = tm_mul(c, s, d);
    TM r = tm_add(c, a, sd);
    tm_slide(c, r, m0);
    r.acc = spacial;
    return r;
}


} } // namespace ab200
