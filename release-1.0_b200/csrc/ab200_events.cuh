// ab200_events.cuh -- guard / invariant tests and blocking-time selection for one step.
// Reference: src/hybrid_evolver-image.cpp:581-681,770-883,1085-1127; src/taylor_calculus.cpp:603-611.
#pragma once
#include "ab200_calc.cuh"

namespace ab200 {

struct WorkingSet;

// Pending working sets of one cell (std::list<EvolutionData> in the reference, popped from the front).
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

// _compute_blocking_info (src/hybrid_evolver-image.cpp:1085-1127)
AB_HDN void compute_blocking_info(Ctx& c, const SystemDesc& S, const LocDesc& L, const TM* fset, int n, int p,
                                  const TM& tstep, Acc spacial, BlockingInfo& B) {
    (void)S; (void)fset; (void)n; (void)p; (void)spacial;
    B.time = tstep; B.finishing = true; B.n_blocking = 0; B.n_nontransverse = 0;
    for (int e = 0; e < L.n_events; ++e) {
        if (L.ev[e].kind == EV_INVARIANT || L.ev[e].kind == EV_URGENT) { fail(c, ST_UNSUPPORTED); return; }
    }
}
AB_HD bool enclosure_to_be_discarded(Ctx& c, const SystemDesc& S, const LocDesc& L, const TM* set, int n, int p) {
    (void)c; (void)S; (void)L; (void)set; (void)n; (void)p; return false;
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

} // namespace ab200
