// ab200_events.cuh -- guard / invariant tests and blocking-time selection for one step.
// Reference: src/hybrid_evolver-image.cpp:581-681 (initially active events, discard test), :770-883 (blocking times),
// :1085-1127 (_compute_blocking_info); src/taylor_calculus.cpp:603-611 and include/calculus_base.h:88-99,295-311
// (active).
#pragma once
#include "ab200_calc.cuh"

namespace ab200 { inline namespace AB_VARIANT {

struct WorkingSet;

// Pending working sets of one cell (std::list<EvolutionData> in the reference, popped from the front). Jumps inside
// the evolver are not built yet, so the queue never holds more than the current set.
struct WsQueue { int count; };
AB_HD void wsq_init(WsQueue& q) { q.count = 0; }
AB_HD bool wsq_pop(Ctx& c, WsQueue& q, WorkingSet& ws) { (void)c; (void)q; (void)ws; return false; }

struct BlockingInfo {
    TM time;           // blocking_time_model (scaled to [0,1])
    bool finishing;    // blocking_events == {finishing_event}
    int n_blocking;    // number of blocking events other than finishing
    int blocking[AB_MAXEV];
    int n_nontransverse;
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

// _compute_blocking_info (src/hybrid_evolver-image.cpp:1085-1127) with
// compute_eventBlockingTimes_and_nonTransverseEvents (:770-843). Crossing-time models (implicit function of the guard
// along the flow) are not built yet: a step whose flow tube possibly meets an urgent guard reports UNSUPPORTED.
AB_HDN void compute_blocking_info(Ctx& c, const SystemDesc& S, const LocDesc& L, const TM* fset, int n, int p,
                                  const TM& tstep, Acc spacial, BlockingInfo& B) {
    (void)p; (void)spacial;
    B.time = tstep; B.finishing = true; B.n_blocking = 0; B.n_nontransverse = 0;
    bool any = false;
    for (int e = 0; e < L.n_events; ++e) if (is_urgent(L.ev[e])) any = true;
    if (!any) return;
    // positive_flow_set_model=split(flow_set_model.models(),dimension,1)  (sic: variable index = dimension)
    const int m = c.A.top;
    TM pos[AB_MAXN];
    for (int i = 0; i < n; ++i) pos[i] = tm_split(c, fset[i], n, true);
    for (int e = 0; e < L.n_events; ++e) {
        if (!is_urgent(L.ev[e])) continue;
        int act = guard_active(c, S, L.ev[e].guard, pos);
        if (c.status != ST_OK) break;
        if (tri_possibly(act)) { fail(c, ST_UNSUPPORTED); break; }
    }
    c.A.top = m;
}

// is_enclosure_to_be_discarded (:652-681) for UPPER semantics: a blocking event definitely active on the starting set.
AB_HDN bool enclosure_to_be_discarded(Ctx& c, const SystemDesc& S, const LocDesc& L, const TM* set, int n, int p) {
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
    return blocking_active == TRI_TRUE;
}
AB_HD bool has_permissive_or_nontransverse(const LocDesc& L, const BlockingInfo& B) {
    if (B.n_nontransverse > 0) return true;
    for (int e = 0; e < L.n_events; ++e) if (L.ev[e].kind == EV_PERMISSIVE) return true;
    return false;
}
AB_HD bool blocking_has_transition(const LocDesc& L, const BlockingInfo& B) {
    for (int k = 0; k < B.n_blocking; ++k) if (L.ev[B.blocking[k]].kind != EV_INVARIANT) return true;
    return false;
}

} } // namespace ab200
