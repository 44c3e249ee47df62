// ab200_evolve.cuh -- evolution of one frontier cell for one lock-to-grid lap.
//
// Restates ImageSetHybridEvolver::_evolution / _evolution_step (reference: src/hybrid_evolver-image.cpp:258-547,
// 752-766, 1000-1127, 1159-1178, 1253-1284) for UPPER_SEMANTICS with the shipped settings
// (src/hybrid_evolver-image.cpp:1347-1359: reconditioning on, subdivisions off, premature termination on,
// error-rate enforcement off, boxing off) and TaylorCalculus defaults (src/taylor_calculus.cpp:136-146).
// Output per cell: the bounding boxes of every reach set and of every final set, which is all the discretiser
// consumes (include/hybrid_set.h:976-996).
#pragma once
#include "ab200_calc.cuh"
#include "ab200_series.cuh"
#include "ab200_events.cuh"

namespace ab200 { inline namespace AB_VARIANT {

struct EvolveParams {
    double T;                  // continuous time of the lap (HybridTime::continuous_time)
    int32_t max_events;        // HybridTime::discrete_time
    int32_t ignore_activations;
    int32_t max_boxes;         // capacity of each per-cell output list
    int32_t temporal_order;    // 6
    double ratio;              // maximum_enclosure_widths_ratio = 5.0
    double spacial_sweep;      // 1e-12
    int32_t spacial_order;     // 6
    int32_t max_working_sets;  // per-cell queue capacity (device limit, not a reference setting)
    int32_t lower;             // Semantics: 0 = UPPER_SEMANTICS, 1 = LOWER_SEMANTICS (include/evolver_interface.h:46-49)
    int32_t backward;          // ContinuousEvolutionDirection: 0 = DIRECTION_FORWARD, 1 = DIRECTION_BACKWARD
};

struct CellOutput {
    // boxes are [lo0,hi0,lo1,hi1,...] records of 2*n doubles; loc index stored separately
    double* reach;  int32_t* reach_loc;  int32_t n_reach;
    double* final_; int32_t* final_loc;  int32_t n_final;
    int32_t stride;            // doubles per box record (2*AB_MAXN)
    int32_t steps;             // number of _evolution_step calls (work counter)
    // final sets as enclosures (set models, not only their boxes) for callers that hand them to the next lap
    // (LowerReachEpsilonWorker, include/workers.h:299-302); null = not wanted
    uint64_t* enc; int32_t enc_cap; int32_t enc_used; int32_t n_enc;
};

// The reference hands enclosures to the discretiser through HybridSetConstIterator::dereference
// (include/hybrid_set.h:883-887), which replaces every TaylorSet by TaylorSet(its box()), i.e. scalings(bbox); the
// discretiser then takes bounding_box() of THAT (include/hybrid_set.h:991-992). The emitted box is therefore
// range(scaling(range(set))) per coordinate, a few ulps wider than range(set).
AB_HD void out_box(Ctx& c, double* dst, int32_t* dloc, int32_t& count, int cap, int stride, int loc, const Ivl* b, int n) {
    if (count >= cap) { fail(c, ST_OUTPUT); return; }
    double* p = dst + (size_t)count * stride;
    #pragma unroll 1
    for (int i = 0; i < n; ++i) {
        const int m = c.A.top;
        TM s = tm_scaling(c, n, i, b[i]);
        Ivl r = tm_range(c, s);
        c.A.top = m;
        p[2 * i] = r.l; p[2 * i + 1] = r.u;
    }
    dloc[count] = loc;
    ++count;
}

// Enclosure record (ab200_reach_evolve_enclosures): the TaylorSet of a final set as the reference keeps it in
// final_sets.adjoin(loc,_toolbox->enclosure(set_model)) (src/hybrid_evolver-image.cpp:304; enclosure() is the identity,
// src/taylor_calculus.cpp:758-763). Words: [loc<<32 | p] [n<<32 | words of the whole record] then per model
// [terms<<32 | argument size] [error bits] and the terms as they lie in the arena (AB_TW words each). The record is opaque
// to the host and specific to the key width of the build that wrote it.
struct WorkingSet;
AB_HD int enc_words(const TM* set, int n) { int w = 2; for (int i = 0; i < n; ++i) w += 2 + set[i].n * AB_TW; return w; }

// A working set (EvolutionData, include/hybrid_evolver-image.h): location, event count, set model (n models in p
// variables) and time model, stored packed in the arena starting at `base`.
struct WorkingSet {
    int loc;
    int n_events;
    int p;
    TM set[AB_MAXN];
    TM time;
};

AB_HD void out_enclosure(Ctx& c, CellOutput& O, const WorkingSet& ws, int n) {
    if (!O.enc) return;
    const int need = enc_words(ws.set, n);
    if (O.enc_used + need > O.enc_cap) { fail(c, ST_OUTPUT); return; }
    uint64_t* q = O.enc + O.enc_used;
    const int W = AB_W, lane = wlane();
    wsync();
    if (lane == 0) { q[0] = ((uint64_t)(uint32_t)ws.loc << 32) | (uint32_t)ws.p; q[1] = ((uint64_t)(uint32_t)n << 32) | (uint32_t)need; }
    int o = 2;
    #pragma unroll 1
    for (int i = 0; i < n; ++i) {
        const TM& m = ws.set[i];
        if (lane == 0) { q[o] = ((uint64_t)(uint32_t)m.n << 32) | (uint32_t)m.as; q[o + 1] = d2u(m.e); }
        const int nw = m.n * AB_TW;
        const uint64_t* src = c.A.base + m.off;
        for (int k = lane; k < nw; k += W) q[o + 2 + k] = src[k];
        o += 2 + nw;
    }
    wsync();
    O.enc_used += need; O.n_enc++;
}

// _is_enclosure_too_large  src/hybrid_evolver-image.cpp:1253-1284
AB_HD bool enclosure_too_large(Ctx& c, const LocDesc& L, const Ivl* range, const double* init_w, double ratio) {
    const int n = L.n;
    bool use_init = false;
    #pragma unroll 1
    for (int i = 0; i < n; ++i) if (init_w[i] >= L.ref_width[i]) { use_init = true; break; }
    #pragma unroll 1
    for (int i = 0; i < n; ++i) {
        double w = ivl_width(range[i]);
        if (use_init) { if (div_m(c.rnd, w, init_w[i]) >= ratio) return true; }
        else { if (w >= mul_m(c.rnd, L.ref_width[i], ratio)) return true; }
    }
    return false;
}

// Pending working sets of one cell (std::list<EvolutionData> in the reference, unbounded by default). Their handles live in
// a table at the bottom of the cell's arena (AB_QW words each), so the capacity is a setting (EvolveParams::max_working_sets,
// at most AB_MAXQ) rather than a register-file limit; WorkingSetTooLarge beyond it.
#define AB_MAXQ 256
#define AB_QW 26
AB_HD void tmh_store(uint64_t* w, const TM& t) {
    w[0] = ((uint64_t)(uint32_t)t.off << 32) | (uint32_t)t.n; w[1] = ((uint64_t)(uint32_t)t.as << 32) | (uint32_t)t.acc.maxdeg; w[2] = d2u(t.e); w[3] = d2u(t.acc.sweep);
}
AB_HD TM tmh_load(const uint64_t* w) {
    TM t; t.off = (int)(w[0] >> 32); t.n = (int)(w[0] & 0xFFFFFFFFu); t.as = (int)(w[1] >> 32); t.acc.maxdeg = (int)(w[1] & 0xFFFFFFFFu); t.e = u2d(w[2]); t.acc.sweep = u2d(w[3]); return t;
}
// warp-uniform: every lane writes / reads the same words
AB_HD void ws_store(Ctx& c, int qbase, int k, const WorkingSet& w) {
    uint64_t* p = c.A.base + qbase + k * AB_QW;
    p[0] = ((uint64_t)(uint32_t)w.loc << 32) | (uint32_t)w.n_events; p[1] = (uint64_t)(uint32_t)w.p;
    #pragma unroll 1
    for (int i = 0; i < AB_MAXN; ++i) tmh_store(p + 2 + 4 * i, w.set[i]);
    tmh_store(p + 2 + 4 * AB_MAXN, w.time);
}
AB_HD void ws_load(Ctx& c, int qbase, int k, WorkingSet& w) {
    const uint64_t* p = c.A.base + qbase + k * AB_QW;
    w.loc = (int)(p[0] >> 32); w.n_events = (int)(p[0] & 0xFFFFFFFFu); w.p = (int)p[1];
    #pragma unroll 1
    for (int i = 0; i < AB_MAXN; ++i) w.set[i] = tmh_load(p + 2 + 4 * i);
    w.time = tmh_load(p + 2 + 4 * AB_MAXN);
}

// Slide the models of working set w (set[0..n), time; stored in ascending arena order) down to dst, packed.
AB_HD int ws_pack(Ctx& c, WorkingSet& w, int n, int dst) {
    #pragma unroll 1
    for (int i = 0; i < n; ++i) { tm_slide(c, w.set[i], dst); dst = c.A.top; }
    tm_slide(c, w.time, dst);
    return c.A.top;
}

// Evolution of one enclosure (a box in location loc0) for time P.T mirrors _evolution: a FIFO of working sets
// (std::list<EvolutionData>, popped from the front, successors pushed at the back), kept packed at the bottom of the
// cell's arena in queue order; everything above the last queued set is scratch. The loop is split into cell_begin /
// cell_advance (one queue iteration = at most one _evolution_step).
struct CellState {
    int nq;
    int qcap;
    int qbase;           // handle table
    int base0;           // models of the queued sets start here
    int guard_iter;
    double init_w[AB_MAXN];
};

// enc: null, or the enclosure record the evolution starts from instead of the box (_evolution_add_initialSet with a TaylorSet:
// set_model(enclosure) is a copy with the spacial accuracy, src/taylor_calculus.cpp:746-753; time_model(0.0) over its p parameters)
AB_HDN void cell_begin(Ctx& c, const SystemDesc& S, int loc0, const double* box, const EvolveParams& P, CellState& st, const uint64_t* enc = 0) {
    Acc spacial; spacial.sweep = P.spacial_sweep; spacial.maxdeg = P.spacial_order;
    const int n0 = S.loc[loc0].n;
    ctx_stage_tables(c);
    st.qcap = (P.max_working_sets > 0 && P.max_working_sets < AB_MAXQ) ? P.max_working_sets : AB_MAXQ;
    st.qbase = c.A.top;
    if ((long long)st.qbase + (long long)(st.qcap + AB_MAXEV + 1) * AB_QW + 64 > (long long)c.A.limit) { fail(c, ST_ARENA); st.nq = 0; return; }
    c.A.top = st.qbase + (st.qcap + AB_MAXEV + 1) * AB_QW;      // room for the successors of one step past the capacity check
    st.base0 = c.A.top;
    st.nq = 1;
    // ---- _evolution_add_initialSet: set_model(box)=scalings(box) with spacial accuracy; time_model(0.0)
    WorkingSet w0;
    w0.loc = loc0; w0.n_events = 0; w0.p = n0;
    #pragma unroll 1
    for (int i = 0; i < AB_MAXN; ++i) { w0.set[i].off = 0; w0.set[i].n = 0; w0.set[i].as = 0; w0.set[i].e = 0.0; w0.set[i].acc = spacial; }
    if (enc) {
        const int pe = (int)(enc[0] & 0xFFFFFFFFu), ne = (int)(enc[1] >> 32);
        if ((int)(enc[0] >> 32) != loc0 || ne != n0 || pe < 0 || pe > AB_MAXARGS) { fail(c, ST_UNSUPPORTED); st.nq = 0; return; }
        w0.p = pe;
        const int W = AB_W, lane = wlane();
        int o = 2;
        #pragma unroll 1
        for (int i = 0; i < n0; ++i) {
            TM m = tm_begin(c, pe, spacial);
            m.n = (int)(enc[o] >> 32); m.as = (int)(enc[o] & 0xFFFFFFFFu); m.e = u2d(enc[o + 1]);
            const int nw = m.n * AB_TW;
            if (m.as != pe || !arena_room(c, m.off, m.n)) { if (c.status == ST_OK) fail(c, ST_UNSUPPORTED); st.nq = 0; return; }
            uint64_t* dst = c.A.base + m.off;
            for (int k = lane; k < nw; k += W) dst[k] = enc[o + 2 + k];
            c.A.top = m.off + nw; arena_mark_peak(c);
            wsync();
            w0.set[i] = m;
            o += 2 + nw;
        }
    } else {
    #pragma unroll 1
    for (int i = 0; i < n0; ++i) { w0.set[i] = tm_scaling(c, n0, i, ivl(box[2 * i], box[2 * i + 1])); w0.set[i].acc = spacial; }
    }
    w0.time = tm_begin(c, w0.p, spacial); tm_append(c, w0.time, key_zero(), 0.0);          // constant(as,0.0): explicit zero term
    #pragma unroll 1
    for (int i = 0; i < n0; ++i) st.init_w[i] = ivl_width(tm_range(c, w0.set[i]));   // _indexed_set_models_widths
    ws_store(c, st.qbase, 0, w0);
    st.guard_iter = 0;
}

// Pops the front working set: emits it as a final set or takes one _evolution_step. Returns false once the cell is
// finished (queue empty) or has failed (c.status).
AB_HDN bool cell_advance(Ctx& c, const SystemDesc& S, const EvolveParams& P, CellState& st, CellOutput& O) {
    if (st.nq <= 0 || c.status != ST_OK) return false;
    if (++st.guard_iter > 100000) { fail(c, ST_OUTPUT); return false; }
    Acc spacial; spacial.sweep = P.spacial_sweep; spacial.maxdeg = P.spacial_order;
    int& nq = st.nq;
    const double* init_w = st.init_w;
    const int base0 = st.base0, qbase = st.qbase;
    do {
        WorkingSet ws; ws_load(c, qbase, 0, ws);
        const int qend = c.A.top;                      // end of the last queued set; scratch starts here
        const LocDesc& L = S.loc[ws.loc];
        const int n = L.n, p = ws.p;
        int ns = 0;                                    // successors are parked in the handle table past the queued sets
        WorkingSet w;
        #pragma unroll 1
        for (int i = 0; i < AB_MAXN; ++i) { w.set[i].off = 0; w.set[i].n = 0; w.set[i].as = 0; w.set[i].e = 0.0; w.set[i].acc = spacial; }
        Ivl range[AB_MAXN];
        #pragma unroll 1
        for (int i = 0; i < n; ++i) range[i] = tm_range(c, ws.set[i]);
        Ivl trange = tm_range(c, ws.time);
        bool too_large = enclosure_too_large(c, L, range, init_w, P.ratio);
        if (trange.l >= P.T || ws.n_events >= P.max_events) {
            out_box(c, O.final_, O.final_loc, O.n_final, P.max_boxes, O.stride, ws.loc, range, n);
            out_enclosure(c, O, ws, n);
        } else if (too_large) {
            // premature termination on enclosure size; UPPER semantics keeps the set as final, LOWER drops it (:308-313)
            if (!P.lower) { out_box(c, O.final_, O.final_loc, O.n_final, P.max_boxes, O.stride, ws.loc, range, n); out_enclosure(c, O, ws, n); }
        } else {
            // ---- _evolution_step ------------------------------------------------------------------------------
            O.steps++;
#ifdef AB_DEBUG_DUMP
            printf("STEP %d\n", O.steps - 1);
            #pragma unroll 1
            for (int i = 0; i < n; ++i) AB_DUMP(c, "set", i, ws.set[i]);
            AB_DUMP(c, "time", -1, ws.time);
#endif
            double remaining = sub_m(c.rnd, P.T, trange.l);
            double hprop = dmin(remaining, L.max_step);
            Ivl bb[AB_MAXN], dom[AB_MAXN + 1];
            const Func* vf = P.backward ? L.dyn_back : L.dyn;      // get_directed_dynamic (:470)
            cta_align(c, 1);
            double h = flow_bounds(c, S.ops, vf, n, range, hprop, bb);
            if (c.status != ST_OK) break;
            TM flow[AB_MAXN], fset[AB_MAXN];
            flow_model(c, S.ops, vf, n, range, h, bb, P.temporal_order, flow, dom);
            cta_align(c, 32);
            flow_apply(c, flow, n, dom, ws.set, p, h, fset);
#ifdef AB_DEBUG_DUMP
            printf("h %a bb", h); for (int i = 0; i < n; ++i) printf(" %a %a", bb[i].l, bb[i].u); printf("\n");
            #pragma unroll 1
            for (int i = 0; i < n; ++i) AB_DUMP(c, "flow", i, flow[i]);
#endif
            vec_slide(c, fset, n, qend);              // drop the flow model, keep the flow set model
            #pragma unroll 1
            for (int i = 0; i < n; ++i) AB_DUMP(c, "fset", i, fset[i]);
            if (c.status != ST_OK) break;

            // blocking info (src/hybrid_evolver-image.cpp:1085-1127)
            TM tstep = tm_begin(c, p, spacial); tm_append(c, tstep, key_zero(), 1.0);     // time_model(1.0,Box(p))
            TM zero_t = tm_begin(c, p, spacial); tm_append(c, zero_t, key_zero(), 0.0);   // time_model(0.0,Box(p))
            BlockingInfo B;
            cta_align(c, 34);
            compute_blocking_info(c, S, L, fset, n, p, tstep, spacial, B);
            if (c.status != ST_OK) break;
            if (!enclosure_to_be_discarded(c, S, L, ws.set, n, p, P.lower, P.backward)) {
                if (c.status != ST_OK) break;
                Activation acts[AB_MAXEV]; int na = 0;
                if (!P.ignore_activations) {
                    if (p + 2 > AB_MAXARGS) { fail(c, ST_UNSUPPORTED); break; }     // parameter count beyond the packed key width
                    na = compute_activation_times(c, S, L, B, fset, n, p, P.lower, acts);
                    if (c.status != ST_OK) break;
                }
                // reachable set (:1159-1178)
                {
                    const int m = c.A.top;
                    TM reach[AB_MAXN];
                    cta_align(c, 36);
                    reachability_step(c, fset, n, p, zero_t, B.time, spacial, reach);
                    Ivl rb[AB_MAXN];
                    #pragma unroll 1
                    for (int i = 0; i < n; ++i) AB_DUMP(c, "reach", i, reach[i]);
                    #pragma unroll 1
                    for (int i = 0; i < n; ++i) rb[i] = tm_range(c, reach[i]);
                    out_box(c, O.reach, O.reach_loc, O.n_reach, P.max_boxes, O.stride, ws.loc, rb, n);
                    c.A.top = m;
                }
                // _computeEvolutionForEvents (:1000-1082); with LOWER semantics only when the blocking event is unique (:544)
                if (!P.lower || ((B.finishing ? 1 : 0) + B.n_blocking) == 1) {
                TM ev[AB_MAXN];
                cta_align(c, 38);
                integration_step(c, fset, n, p, B.time, spacial, ev);
                #pragma unroll 1
                for (int i = 0; i < n; ++i) AB_DUMP(c, "evolved", i, ev[i]);
                // final_time_model=time_model+blocking_time_model*time_step
                TM ft;
                {
                    const int m = c.A.top;
                    TM sc = tm_copy(c, B.time); tm_scal(c, sc, h);
                    ft = tm_add(c, ws.time, sc);
                    tm_slide(c, ft, m);
                }
                int np = p;
                cta_align(c, 40);
                {   // reconditioning (enable_reconditioning)
                    double thr[AB_MAXN];
                    #pragma unroll 1
                    for (int i = 0; i < n; ++i) thr[i] = div_m(c.rnd, L.ref_width[i], 4.0);
                    np = uniform_error_recondition(c, ev, n, p, thr);
                    if (np > p) ft.as = np;                                   // embed(final_time_model,diff)
                    int disc[AB_MAXARGS + 1];
                    int nd = kuhn_recondition(c, ev, n, np, disc);
                    if (nd > 0) {
                        ft = tm_recondition(c, ft, disc, nd, n, n);
                        np = 2 * n;
                    }
                }
                // successors, in the order of the blocking_events set (std::set<DiscreteEvent>: by name, with the
                // finishing event "e-2" at its rank)
                bool fin_done = false;
                #pragma unroll 1
                for (int k = 0; k <= B.n_blocking && c.status == ST_OK; ++k) {
                    const int e = (k < B.n_blocking) ? B.blocking[k] : -1;
                    if (B.finishing && !fin_done && (e < 0 || e >= L.finishing_rank)) {
                        w.loc = ws.loc; w.n_events = ws.n_events; w.p = np;
                        #pragma unroll 1
                        for (int i = 0; i < n; ++i) w.set[i] = ev[i];
                        w.time = ft;
                        ws_store(c, qbase, nq + ns, w); ++ns;
                        fin_done = true;
                    }
                    if (e < 0) break;
                    const EventDesc& E = L.ev[e];
                    if (E.kind != EV_INVARIANT && !P.ignore_activations) {
                        // jump: apply(reset,evolved_set_model) into the target location (:1049-1058)
                        const int nt = S.loc[E.target].n;
                        w.loc = E.target; w.n_events = ws.n_events + 1; w.p = np;
                        #pragma unroll 1
                        for (int i = 0; i < nt; ++i) w.set[i] = eval_tm(c, S.ops, E.reset[i], ev);
                        w.time = ft;
                        ws_store(c, qbase, nq + ns, w); ++ns;
                    }
                }
                // non-blocking events (:1060-1081): jump from the part of the flow tube inside the activation window;
                // the new working set carries one more parameter (the time within the window) and is not reconditioned
                #pragma unroll 1
                for (int a = 0; a < na && c.status == ST_OK; ++a) {
                    const EventDesc& E = L.ev[acts[a].ev];
                    const int nt = S.loc[E.target].n;
                    TM aset[AB_MAXN];
                    reachability_step(c, fset, n, p, acts[a].lower, acts[a].upper, spacial, aset);
                    w.loc = E.target; w.n_events = ws.n_events + 1; w.p = p + 1;
                    #pragma unroll 1
                    for (int i = 0; i < nt; ++i) w.set[i] = eval_tm(c, S.ops, E.reset[i], aset);
                    const int m = c.A.top;
                    TM up = tm_copy(c, acts[a].upper); tm_scal(c, up, h); TM t1 = tm_add(c, ws.time, up);
                    TM lo = tm_copy(c, acts[a].lower); tm_scal(c, lo, h); TM t0 = tm_add(c, ws.time, lo);
                    TM at = reachability_time(c, t0, t1, p, spacial);
                    tm_slide(c, at, m);
                    w.time = at;
                    ws_store(c, qbase, nq + ns, w); ++ns;
                }
                }
            }
        }
        if (c.status != ST_OK) break;
        // ---- pop the front, push the successors at the back, re-pack the queue at the bottom of the arena
        if (nq - 1 + ns > st.qcap) { fail(c, ST_WORKING_SETS); break; }
        cta_align(c, 44);
        wsync();
        // 1. copy the successors to the top in canonical order (their models may be anywhere above qend)
        #pragma unroll 1
        for (int k = 0; k < ns; ++k) {
            ws_load(c, qbase, nq + k, w);
            const int nn = S.loc[w.loc].n;
            #pragma unroll 1
            for (int i = 0; i < nn; ++i) w.set[i] = tm_copy(c, w.set[i]);
            w.time = tm_copy(c, w.time);
            ws_store(c, qbase, nq + k, w);
        }
        // 2. slide everything after the front set down over it
        int dst = base0;
        #pragma unroll 1
        for (int k = 1; k < nq + ns; ++k) { ws_load(c, qbase, k, w); wsync(); dst = ws_pack(c, w, S.loc[w.loc].n, dst); ws_store(c, qbase, k - 1, w); }
        nq = nq - 1 + ns;
        c.A.top = dst;
        wsync();
    } while (0);
    return nq > 0 && c.status == ST_OK;
}

AB_HDN void evolve_cell(Ctx& c, const SystemDesc& S, int loc0, const double* box, const EvolveParams& P, CellOutput& O, const uint64_t* enc = 0) {
    CellState st;
    cell_begin(c, S, loc0, box, P, st, enc);
    #pragma unroll 1
    while (cell_advance(c, S, P, st, O)) { }
}

} } // namespace ab200
