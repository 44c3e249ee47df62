// ab200_jump.cuh -- ForwardDiscreteJumpSetChecker::check on one cell box (reference: src/set_checker.cpp:66-203), written
// over the kernels' own TaylorModel / interval routines so that the same code serves K_jump_check (one warp per box) and the
// host-side checker that pins it.
#pragma once
#include "ab200_events.cuh"
#include "ab200_series.cuh"

namespace ab200 { inline namespace AB_VARIANT {

AB_HD int jtri_not(int t) { return t == TRI_INDET ? TRI_INDET : (t == TRI_TRUE ? TRI_FALSE : TRI_TRUE); }
AB_HD int jtri_and(int a, int b) { if (a == TRI_FALSE || b == TRI_FALSE) return TRI_FALSE; if (a == TRI_TRUE && b == TRI_TRUE) return TRI_TRUE; return TRI_INDET; }

// _is_outside_any_invariant (src/set_checker.cpp:111-135)
AB_HDN int jump_outside_any_invariant(Ctx& c, const SystemDesc& S, int l, const Ivl* bx) {
    const LocDesc& L = S.loc[l];
    int result = TRI_FALSE;
    #pragma unroll 1
    for (int e = 0; e < L.n_events; ++e) {
        if (L.ev[e].kind != EV_INVARIANT && L.ev[e].kind != EV_URGENT) continue;
        int act = tribool_of(eval_ivl(c, S.ops, L.ev[e].guard, bx));
        if (act == TRI_TRUE) return TRI_TRUE;
        if (act == TRI_INDET) result = TRI_INDET;
    }
    return result;
}
// _is_transition_taken (src/set_checker.cpp:66-108); *lie_missing is set when the guard has no Lie derivative (the reference
// throws there, src/function.cpp:1329)
AB_HDN int jump_transition_taken(Ctx& c, const SystemDesc& S, int l, int e, const TM* src, int n, int* lie_missing) {
    const EventDesc& E = S.loc[l].ev[e];
    const bool urgent = (E.kind == EV_URGENT);
    int act = guard_active(c, S, E.guard, src);
    if (act == TRI_TRUE && urgent) return TRI_FALSE;
    if (act != TRI_FALSE && !urgent) return act;
    if (act != TRI_FALSE && urgent) {
        // _is_positively_crossing: Lie derivative of the guard over the enclosure's bounding box
        if (E.lie.end <= E.lie.begin) { *lie_missing = 1; fail(c, ST_UNSUPPORTED); return TRI_INDET; }
        Ivl bb[AB_MAXN];
        #pragma unroll 1
        for (int i = 0; i < n; ++i) bb[i] = tm_range(c, src[i]);
        Ivl d = eval_ivl(c, S.ops, E.lie, bb);
        if (d.u < 0) return TRI_FALSE;
        return TRI_INDET;
    }
    return TRI_FALSE;
}
// ForwardDiscreteJumpSetChecker::check (src/set_checker.cpp:171-203) on the box lo/hi per coordinate
AB_HDN int jump_check(Ctx& c, const SystemDesc& S, int l, int e, const double* box, int* lie_missing) {
    const int n = S.loc[l].n;
    const int m = c.A.top;
    Ivl bx[AB_MAXN];
    #pragma unroll 1
    for (int i = 0; i < n; ++i) bx[i] = ivl(box[2 * i], box[2 * i + 1]);
    int out_src = jump_outside_any_invariant(c, S, l, bx);
    if (out_src == TRI_TRUE) return TRI_FALSE;
    int result = TRI_FALSE;
    TM src[AB_MAXN];
    #pragma unroll 1
    for (int i = 0; i < n; ++i) src[i] = tm_scaling(c, n, i, bx[i]);           // TaylorSet(Box) (src/taylor_set.cpp:69-72)
    int taken = jump_transition_taken(c, S, l, e, src, n, lie_missing);
    if (taken != TRI_FALSE) {
        const EventDesc& E = S.loc[l].ev[e];
        const int nt = S.loc[E.target].n;
        Ivl tb[AB_MAXN];
        #pragma unroll 1
        for (int i = 0; i < nt; ++i) { TM t = eval_tm(c, S.ops, E.reset[i], src); tb[i] = tm_range(c, t); }
        int out_trg = jump_outside_any_invariant(c, S, E.target, tb);
        result = jtri_and(jtri_and(jtri_not(out_src), taken), jtri_not(out_trg));
    }
    c.A.top = m;
    return result;
}

} } // namespace ab200
