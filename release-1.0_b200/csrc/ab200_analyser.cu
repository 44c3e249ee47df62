// ab200_analyser.cu -- host mirror of HybridReachabilityAnalyser over the C ABI lap.
//
// Reference: HybridReachabilityAnalyser(system,domain,accuracy) src/reachability_analyser.cpp:83-100;
// getHybridMidpointAbsoluteDerivatives src/hybrid_set.cpp:485-515; getHybridGrid src/reachability_analyser.cpp:1028-1089;
// getLockToGridTime :1141-1170; ImageSetHybridEvolver::tune_settings src/hybrid_evolver-image.cpp:99-113,1506-1555;
// upper_reach_evolve :309-377; outer_chain_reach :380-493; ReachabilityRestriction src/reachability_restriction.cpp.
// These are one-off host computations per analyser; the per-cell work is all on the device.
#include <vector>
#include <string>
#include <algorithm>
#include <cstring>
#include <cmath>
#include <limits>
#include "../../include/ariadne_b200.h"
#include "ab200_events.cuh"
#include "ab200_series.cuh"
#include "ab200_jump.cuh"
#include "ab200_system.h"
#include "ab200_paving.h"

using namespace ab200;

const HostSystem* ab200_ctx_system(const ab200_ctx* ctx);
const KeyLayout* ab200_ctx_layout(const ab200_ctx* ctx);
void ab200_ctx_set_error(ab200_ctx* ctx, const char* msg);
// ForwardDiscreteJumpSetChecker::check for a batch of cell boxes (stride AB200_BOX_STRIDE) of transition `event` out of
// location `loc`, on the device (K_jump_check, one warp per box): tri[i] = 0 false / 1 true / 2 indeterminate. Returns an
// AB200 code; *lie_missing is set if the guard has no Lie derivative.
int ab200_ctx_jump_check(ab200_ctx* ctx, int loc, int event, size_t n, const double* boxes, int* tri, int* lie_missing);
std::vector<uint64_t>& ab200_cellset_keys(ab200_cellset* cs);
ab200_cellset* ab200_cellset_new();

struct ab200_analyser {
    ab200_ctx* ctx;
    int accuracy;
    int n_locs, dim;
    std::vector<GridDesc> grid;
    std::vector<std::vector<double> > domain;     // [loc][2*dim]
    std::vector<double> max_step;
    double lock_time;
    KeyLayout K;
    std::vector<CellSet> restriction;              // domain discretisation per location (lazy)
    std::vector<char> discretised;
    int64_t cells_evolved;
};

typedef std::vector<CellSet> HybridCells;          // one CellSet per location

static HybridCells empty_cells(const ab200_analyser* an) {
    HybridCells h;
    for (int l = 0; l < an->n_locs; ++l) h.push_back(CellSet(an->grid[l], an->K, l));
    return h;
}
static bool all_empty(const HybridCells& h) { for (size_t l = 0; l < h.size(); ++l) if (!h[l].empty()) return false; return true; }
static void split_keys(const ab200_analyser* an, const std::vector<uint64_t>& keys, HybridCells& out) {
    // keys are sorted and the location is the top field, so each location is one contiguous run
    int shift = 0; for (int i = 0; i < an->K.dim; ++i) shift += an->K.bits[i];
    size_t i = 0;
    while (i < keys.size()) {
        int l = (int)(keys[i] >> shift);
        size_t j = i; while (j < keys.size() && (int)(keys[j] >> shift) == l) ++j;
        out[l].keys.assign(keys.begin() + i, keys.begin() + j);
        i = j;
    }
}
static void join_keys(const HybridCells& h, std::vector<uint64_t>& keys) {
    keys.clear();
    for (size_t l = 0; l < h.size(); ++l) keys.insert(keys.end(), h[l].keys.begin(), h[l].keys.end());
}

// ReachabilityRestriction::apply_to (src/reachability_restriction.cpp:162-187): intersect with the outer
// approximation of the domain box.
static bool ensure_restriction(ab200_analyser* an, int l) {
    if (an->discretised[l]) return true;
    CellSet s(an->grid[l], an->K, l);
    bool ok = s.adjoin_outer_approximation(an->domain[l].data());
    an->restriction[l] = s; an->discretised[l] = 1;
    return ok;
}
static bool apply_restriction(ab200_analyser* an, HybridCells& h) {
    bool ok = true;
    const double divider = (double)(1 << an->accuracy);
    for (int l = 0; l < an->n_locs; ++l) {
        if (h[l].empty()) continue;
        if (!an->discretised[l]) {
            // "If the domain covers the result, no restriction is required" -- and the domain stays undiscretised,
            // which matters because the evolver tuning depends on it (see tune_evolver).
            double sb[2 * AB_GRID_MAXN];
            bool covers = h[l].bounding_box(sb);
            for (int i = 0; covers && i < an->dim; ++i) {
                double inaccuracy = an->grid[l].lengths[i] / divider;
                double lo = an->domain[l][2 * i] - inaccuracy, hi = an->domain[l][2 * i + 1] + inaccuracy;
                if (!(sb[2 * i] > lo && sb[2 * i + 1] < hi)) covers = false;        // Box::covers is strict
            }
            if (covers) continue;
        }
        ok = ensure_restriction(an, l) && ok;
        h[l].restrict(an->restriction[l]);
    }
    return ok;
}

extern "C" int ab200_analyser_create(ab200_analyser** out, ab200_ctx* ctx, const double* domain_boxes, int32_t accuracy) {
    if (!out || !ctx || !domain_boxes) return AB200_E_ARGUMENT;
    *out = 0;
    const HostSystem* hs = ab200_ctx_system(ctx);
    if (!hs) { ab200_ctx_set_error(ctx, "no system uploaded"); return AB200_E_ARGUMENT; }
    const SystemDesc& S = hs->desc;
    const int L = S.n_locs, dim = S.loc[0].n;
    ab200_analyser* an = new ab200_analyser();
    an->ctx = ctx; an->accuracy = accuracy; an->n_locs = L; an->dim = dim; an->cells_evolved = 0;
    an->domain.resize(L); an->grid.resize(L); an->max_step.resize(L);
    std::vector<std::vector<double> > hmad(L, std::vector<double>(dim, 0.0));
    Ctx c; ctx_init_host(c, 0, 0);
    double lock = 0.0;
    for (int l = 0; l < L; ++l) {
        an->domain[l].assign(domain_boxes + (size_t)l * AB200_BOX_STRIDE, domain_boxes + (size_t)l * AB200_BOX_STRIDE + 2 * dim);
        Ivl d[AB_MAXN];
        bool empty = false;
        for (int i = 0; i < dim; ++i) { d[i] = ivl(an->domain[l][2 * i], an->domain[l][2 * i + 1]); if (d[i].l > d[i].u) empty = true; }
        if (empty) continue;
        for (int i = 0; i < dim; ++i) {
            Ivl der = eval_ivl(c, S.ops, S.loc[l].dyn[i], d);
            Ivl a = ivl_abs(der);
            hmad[l][i] = ivl_mid(a);                                   // abs(der_bbx[i]).midpoint()
            double maxAbsDer = a.u;                                    // getLockToGridTime
            if (maxAbsDer > 0) lock = std::max(lock, ivl_width(d[i]) / maxAbsDer);
        }
    }
    if (c.status != ST_OK) { delete an; ab200_ctx_set_error(ctx, "dynamics cannot be evaluated on the domain"); return c.status; }
    an->lock_time = lock;
    // getHybridGrid(hmad,domain)
    std::vector<double> minDom(dim, std::numeric_limits<double>::infinity());
    for (int l = 0; l < L; ++l) for (int i = 0; i < dim; ++i) minDom[i] = std::min(minDom[i], ivl_width(ivl(an->domain[l][2 * i], an->domain[l][2 * i + 1])));
    std::vector<double> minNonZero = minDom;
    double maxratio = 0.0;
    for (int l = 0; l < L; ++l) for (int i = 0; i < dim; ++i) maxratio = std::max(maxratio, hmad[l][i] / minDom[i]);
    for (int l = 0; l < L; ++l) for (int i = 0; i < dim; ++i) if (hmad[l][i] > 0) minNonZero[i] = std::min(minNonZero[i], hmad[l][i] / maxratio);
    const double divider = (double)(1 << accuracy);
    for (int l = 0; l < L; ++l) {
        GridDesc g; g.dim = dim;
        for (int i = 0; i < dim; ++i) {
            g.lengths[i] = (hmad[l][i] > 0) ? hmad[l][i] / maxratio : minNonZero[i];
            g.origin[i] = ivl_mid(ivl(an->domain[l][2 * i], an->domain[l][2 * i + 1]));     // domain centre
        }
        an->grid[l] = g;
        // getHybridMaximumStepSize
        double mss = 0.0;
        for (int i = 0; i < dim; ++i) if (hmad[l][i] > 0) mss = std::max(mss, g.lengths[i] / divider / hmad[l][i]);
        if (mss == 0) mss = std::numeric_limits<double>::infinity();
        an->max_step[l] = 2.0 * mss;
        double refw[AB_MAXN];
        for (int i = 0; i < dim; ++i) refw[i] = g.lengths[i] / divider;                      // getMinimumGridCellWidths
        int rc = ab200_set_location_tuning(ctx, l, an->max_step[l], refw, g.origin, g.lengths);
        if (rc != AB200_OK) { delete an; return rc; }
    }
    int rc = ab200_set_accuracy_for_domain(ctx, accuracy, domain_boxes);
    if (rc != AB200_OK) { delete an; return rc; }
    an->K = *ab200_ctx_layout(ctx);
    an->restriction.resize(L); an->discretised.assign(L, 0);
    *out = an;
    return AB200_OK;
}
extern "C" void ab200_analyser_destroy(ab200_analyser* an) { delete an; }
extern "C" int ab200_analyser_info(const ab200_analyser* an, double* lock, double* origin, double* lengths, double* max_step) {
    if (!an) return AB200_E_ARGUMENT;
    if (lock) *lock = an->lock_time;
    for (int l = 0; l < an->n_locs; ++l) {
        for (int i = 0; i < an->dim; ++i) { if (origin) origin[l * AB200_MAX_DIM + i] = an->grid[l].origin[i]; if (lengths) lengths[l * AB200_MAX_DIM + i] = an->grid[l].lengths[i]; }
        if (max_step) max_step[l] = an->max_step[l];
    }
    return AB200_OK;
}
extern "C" int ab200_analyser_set_lock_time(ab200_analyser* an, double t) { if (!an || !(t > 0)) return AB200_E_ARGUMENT; an->lock_time = t; return AB200_OK; }
extern "C" int64_t ab200_analyser_cells_evolved(const ab200_analyser* an) { return an ? an->cells_evolved : 0; }

// _get_tuned_evolver (src/reachability_analyser.cpp:125-142): every lap builds a fresh evolver tuned on
// ReachabilityRestriction::bounding_box(), which is the domain widened by one cell per side until the domain has been
// discretised and the bounding box of the discretised cells afterwards (src/reachability_restriction.cpp:597-621).
static int tune_evolver(ab200_analyser* an) {
    const HostSystem* hs = ab200_ctx_system(an->ctx);
    const SystemDesc& S = hs->desc;
    const int dim = an->dim;
    const double divider = (double)(1 << an->accuracy);
    Ctx c; ctx_init_host(c, 0, 0);
    for (int l = 0; l < an->n_locs; ++l) {
        double bb[2 * AB_GRID_MAXN];
        bool have = false;
        if (an->discretised[l]) have = an->restriction[l].bounding_box(bb);
        if (!have) {
            for (int i = 0; i < dim; ++i) {
                double inaccuracy = an->grid[l].lengths[i] / divider;
                bb[2 * i] = an->domain[l][2 * i] - inaccuracy; bb[2 * i + 1] = an->domain[l][2 * i + 1] + inaccuracy;
            }
        }
        Ivl d[AB_MAXN]; bool empty = false;
        for (int i = 0; i < dim; ++i) { d[i] = ivl(bb[2 * i], bb[2 * i + 1]); if (d[i].l > d[i].u) empty = true; }
        double mss = 0.0;
        if (!empty) {
            for (int i = 0; i < dim; ++i) {
                Ivl der = eval_ivl(c, S.ops, S.loc[l].dyn[i], d);
                double hm = ivl_mid(ivl_abs(der));
                if (hm > 0) mss = std::max(mss, an->grid[l].lengths[i] / divider / hm);
            }
        }
        if (c.status != ST_OK) { ab200_ctx_set_error(an->ctx, "dynamics cannot be evaluated on the restriction box"); return c.status; }
        if (mss == 0) mss = std::numeric_limits<double>::infinity();
        an->max_step[l] = 2.0 * mss;
        double refw[AB_MAXN];
        for (int i = 0; i < dim; ++i) refw[i] = an->grid[l].lengths[i] / divider;
        int rc = ab200_set_location_tuning(an->ctx, l, an->max_step[l], refw, an->grid[l].origin, an->grid[l].lengths);
        if (rc != AB200_OK) return rc;
    }
    return AB200_OK;
}

// to_enclosures (src/reachability_analyser.cpp:1244-1253) + UpperReachEvolveWorker::get_result.
// The reference has no scratch limit (its TaylorModels live on the heap). The analyser therefore answers AB200_E_ARENA by raising
// the context's arena_words eightfold (up to AB_ARENA_GROW_MAX words per cell) and running the lap again; the raised value
// stays for the laps that follow. The evolver-level entry points keep the strict error (ab200_settings.arena_words is theirs).
#define AB_ARENA_GROW_MAX (4 * 1024 * 1024)
extern "C" int ab200_ctx_get_settings(const ab200_ctx* ctx, ab200_settings* s);
template<class F> static int with_arena_growth(ab200_analyser* an, F call) {
    int rc = call();
    while (rc == AB200_E_ARENA) {
        ab200_settings s; ab200_ctx_get_settings(an->ctx, &s);
        if (s.arena_words >= AB_ARENA_GROW_MAX) break;
        s.arena_words = s.arena_words * 8 < AB_ARENA_GROW_MAX ? s.arena_words * 8 : AB_ARENA_GROW_MAX;
        if (ab200_set_settings(an->ctx, &s) != AB200_OK) break;
        rc = call();
    }
    return rc;
}

static int lap(ab200_analyser* an, const HybridCells& frontier, int mince, double T, int events, int ignore_activations, int direction,
               HybridCells& reach, HybridCells& evolve, ab200_status* st_acc) {
    std::vector<int32_t> locs; std::vector<double> boxes;
    for (int l = 0; l < an->n_locs; ++l) {
        std::vector<CellBox> cells;
        frontier[l].enumerate(cells, mince ? (1 << 30) : -1);   // mince(accuracy): down to the leaves (bdd_set.cpp:1881-1883)
        for (size_t k = 0; k < cells.size(); ++k) {
            locs.push_back(l);
            size_t o = boxes.size(); boxes.resize(o + AB200_BOX_STRIDE, 0.0);
            for (int i = 0; i < 2 * an->dim; ++i) boxes[o + i] = cells[k].b[i];
        }
    }
    an->cells_evolved += (int64_t)locs.size();
    { int trc = tune_evolver(an); if (trc != AB200_OK) return trc; }
    ab200_cellset *r = 0, *e = 0; ab200_status st; memset(&st, 0, sizeof(st));
    int rc = with_arena_growth(an, [&]() { return ab200_upper_reach_evolve(an->ctx, locs.size(), locs.data(), boxes.data(), T, events, ignore_activations, direction, &r, &e, &st); });
    if (st_acc) {
        st_acc->steps += st.steps; st_acc->launches += st.launches; st_acc->keys_emitted += st.keys_emitted; st_acc->evolve_ms += st.evolve_ms; st_acc->raster_ms += st.raster_ms;
        st_acc->arena_peak = std::max(st_acc->arena_peak, st.arena_peak);
        if (rc != AB200_OK) { st_acc->code = st.code ? st.code : rc; st_acc->cell = st.cell; }
    }
    if (rc != AB200_OK) return rc;
    reach = empty_cells(an); evolve = empty_cells(an);
    split_keys(an, ab200_cellset_keys(r), reach);
    split_keys(an, ab200_cellset_keys(e), evolve);
    ab200_cellset_free(r); ab200_cellset_free(e);
    return AB200_OK;
}

// ---- grid-side jump sets -------------------------------------------------------------------------------------------
// ReachabilityRestriction::forward_jump_set (src/reachability_restriction.cpp:340-375) with
// ForwardDiscreteJumpSetChecker (src/set_checker.cpp:50-226). One-off host work per lap over the cell sets (the
// per-cell flows stay on the device); it uses the same Taylor-model routines as the kernel, compiled for the host.
namespace {
struct JumpCalc {
    const SystemDesc& S;
    std::vector<uint64_t> arena;
    Ctx c;
    bool lie_missing;
    explicit JumpCalc(const SystemDesc& s) : S(s), arena(1u << 22), lie_missing(false) {
        ctx_init_host(c, arena.data(), (int)arena.size());
    }
    static int tri_not(int t) { return t == TRI_INDET ? TRI_INDET : (t == TRI_TRUE ? TRI_FALSE : TRI_TRUE); }
    static int tri_and(int a, int b) { if (a == TRI_FALSE || b == TRI_FALSE) return TRI_FALSE; if (a == TRI_TRUE && b == TRI_TRUE) return TRI_TRUE; return TRI_INDET; }

    // the checks themselves live in ab200_jump.cuh (shared with K_jump_check)
    int outside_any_invariant(int l, const Ivl* bx) { return jump_outside_any_invariant(c, S, l, bx); }
    // TaylorSet(Box) (src/taylor_set.cpp:69-72)
    void box_set(const Ivl* bx, int n, TM* set) { for (int i = 0; i < n; ++i) set[i] = tm_scaling(c, n, i, bx[i]); }
    int transition_taken(int l, int e, const TM* src, int n) { int lm = 0; int r = jump_transition_taken(c, S, l, e, src, n, &lm); if (lm) lie_missing = true; return r; }
    // ForwardDiscreteJumpSetChecker::check (src/set_checker.cpp:171-203)
    int check(int l, int e, const double* box) { int lm = 0; int r = jump_check(c, S, l, e, box, &lm); if (lm) lie_missing = true; return r; }

    // BackwardDiscreteJumpSetChecker (src/set_checker.cpp:229-352): a source cell is kept if some transition may be taken
    // from it into the set reached so far (`start`, one tree per location, with the bounding boxes of its cells).
    bool has_feasible_transitions(int l, const TM* src, int n, const std::vector<CellTree>& start, const std::vector<std::vector<double> >& start_bb) {
        const LocDesc& L = S.loc[l];
        for (int e = 0; e < L.n_events; ++e) {
            const EventDesc& E = L.ev[e];
            if (E.kind != EV_URGENT && E.kind != EV_PERMISSIVE) continue;
            if (start[E.target].empty()) continue;
            const int nt = S.loc[E.target].n;
            const int m = c.A.top;
            bool overlap = true;
            for (int i = 0; i < nt; ++i) {
                TM t = eval_tm(c, S.ops, E.reset[i], src); Ivl r = tm_range(c, t);
                const double bl = start_bb[E.target][2 * i], bu = start_bb[E.target][2 * i + 1];
                if (!(bl < r.u && bu > r.l)) overlap = false;                     // Box::overlaps (open)
            }
            c.A.top = m;
            (void)n;
            if (overlap) return true;
        }
        return false;
    }
    int check_backward(int l, const double* box, const std::vector<CellTree>& start, const std::vector<std::vector<double> >& start_bb) {
        const LocDesc& L = S.loc[l];
        const int n = L.n;
        const int m = c.A.top;
        Ivl bx[AB_MAXN];
        for (int i = 0; i < n; ++i) bx[i] = ivl(box[2 * i], box[2 * i + 1]);
        int out_src = outside_any_invariant(l, bx);
        if (out_src == TRI_TRUE) return TRI_FALSE;
        TM src[AB_MAXN]; box_set(bx, n, src);
        if (!has_feasible_transitions(l, src, n, start, start_bb)) { c.A.top = m; return TRI_FALSE; }
        int result = TRI_FALSE;
        for (int e = 0; e < L.n_events && c.status == ST_OK; ++e) {
            const EventDesc& E = L.ev[e];
            if (E.kind != EV_URGENT && E.kind != EV_PERMISSIVE) continue;
            int taken = transition_taken(l, e, src, n);
            if (taken == TRI_FALSE) continue;
            const int nt = S.loc[E.target].n;
            Ivl tb[AB_MAXN]; double tbox[2 * AB_MAXN];
            for (int i = 0; i < nt; ++i) { TM t = eval_tm(c, S.ops, E.reset[i], src); tb[i] = tm_range(c, t); tbox[2 * i] = tb[i].l; tbox[2 * i + 1] = tb[i].u; }
            int out_trg = outside_any_invariant(E.target, tb);
            if (tri_not(out_trg) == TRI_FALSE) continue;                         // possibly(!is_outside_any_trg_invariant)
            int covered = start[E.target].superset(tbox) ? TRI_TRUE : (start[E.target].disjoint(tbox) ? TRI_FALSE : TRI_INDET);
            int ok = tri_and(tri_and(tri_not(out_src), taken), tri_and(covered, tri_not(out_trg)));
            if (ok == TRI_TRUE) { c.A.top = m; return TRI_TRUE; }
            if (ok == TRI_INDET) result = TRI_INDET;
        }
        c.A.top = m;
        return result;
    }

    // TaylorSet::split() (src/taylor_set.cpp:408-452) into two sets begun at the arena top.
    void split(const TM* set, int n, TM* a, TM* b) {
        const int ng = set[0].as;
        double rmax = 0.0; int armax = 0;
        for (int j = 0; j < ng; ++j) {
            double r = 0.0; Key u = key_unit(j);
            for (int i = 0; i < n; ++i) {
                double g = 0.0;
                for (int t = 0; t < set[i].n; ++t) { Key k = tm_key(c, set[i], t); if (key_eq(k, u)) { g = tm_val(c, set[i], t); break; } if (key_deg(k) > 1) break; }
                r = add_m(c.rnd, r, dabs(g));
            }
            if (r > rmax) { rmax = r; armax = j; }
        }
        double emax = 0.0; int aemax = 0;
        for (int i = 0; i < n; ++i) if (set[i].e > emax) { emax = set[i].e; aemax = i; }
        if (emax > rmax) {
            for (int i = 0; i < n; ++i) if (i != aemax) { a[i] = set[i]; b[i] = set[i]; }
            a[aemax] = tm_copy(c, set[aemax]); a[aemax].e = emax / 2; tm_acc(c, a[aemax], -(emax / 2));
            b[aemax] = tm_copy(c, set[aemax]); b[aemax].e = emax / 2; tm_acc(c, b[aemax], emax / 2);
        } else {
            for (int i = 0; i < n; ++i) { a[i] = tm_split(c, set[i], armax, false); b[i] = tm_split(c, set[i], armax, true); }
        }
    }
    // TaylorSet::disjoint(Box) (src/taylor_set.cpp:170-207). Note the box handed down the recursion is the one already
    // widened by the errors (as the reference does).
    int disjoint(const TM* set, int n, const Ivl* bx, int level = 0) {
        Ivl ebx[AB_MAXN], bbx[AB_MAXN];
        for (int i = 0; i < n; ++i) ebx[i] = ivl_add(bx[i], ivl(-set[i].e, set[i].e));
        for (int i = 0; i < n; ++i) { TM ef = set[i]; ef.e = 0.0; bbx[i] = tm_range(c, ef); }
        bool dis = false, sub = true, cen = true;
        double rb = 0.0, re = 0.0;
        for (int i = 0; i < n; ++i) {
            if (bbx[i].l > ebx[i].u || bbx[i].u < ebx[i].l) dis = true;
            if (!(bbx[i].l >= ebx[i].l && bbx[i].u <= ebx[i].u)) sub = false;
            double v = (set[i].n > 0 && key_deg(tm_key(c, set[i], 0)) == 0) ? tm_val(c, set[i], 0) : 0.0;
            if (!(ebx[i].l <= v && v <= ebx[i].u)) cen = false;
            rb = dmax(rb, sub_ru(bbx[i].u, bbx[i].l)); re = dmax(re, sub_ru(ebx[i].u, ebx[i].l));
        }
        if (dis) return TRI_TRUE;
        if (sub) return TRI_FALSE;
        if (cen) return TRI_FALSE;
        if ((rb / 2) * 2.0 < re / 2) return TRI_INDET;
        if (level > 64 || c.status != ST_OK) { fail(c, ST_UNSUPPORTED); return TRI_INDET; }
        const int m = c.A.top;
        TM a[AB_MAXN], b[AB_MAXN];
        split(set, n, a, b);
        int ra = disjoint(a, n, ebx, level + 1), rbb = disjoint(b, n, ebx, level + 1);
        c.A.top = m;
        return tri_and(ra, rbb);
    }
};
}

static int forward_jump_set(ab200_analyser* an, const HybridCells& reach, HybridCells& result) {
    const HostSystem* hs = ab200_ctx_system(an->ctx);
    const SystemDesc& S = hs->desc;
    result = empty_cells(an);
    JumpCalc J(S);
    bool ok = true;
    for (int l = 0; l < an->n_locs; ++l) {
        const LocDesc& L = S.loc[l];
        for (int e = 0; e < L.n_events; ++e) {
            if (L.ev[e].kind != EV_URGENT && L.ev[e].kind != EV_PERMISSIVE) continue;
            if (reach[l].empty()) continue;
            const int trg = L.ev[e].target;
            const int n = L.n, nt = S.loc[trg].n;
            CellSet feasible = reach[l];
            // the activation test of every cell box of a level of the paving goes to the device in one batch; AB200_JUMP_HOST=1
            // runs the same descent with the host build of the same check (the checker the tests compare against)
            const bool on_host = getenv("AB200_JUMP_HOST") != 0;
            if (on_host) feasible.outer_restrict([&](const double* box) { return J.check(l, e, box); });
            else {
                int brc = AB200_OK; int lm = 0;
                feasible.outer_restrict_batched([&](const std::vector<CellBox>& boxes, std::vector<int>& tri) {
                    std::vector<double> flat(boxes.size() * AB200_BOX_STRIDE, 0.0);
                    for (size_t k = 0; k < boxes.size(); ++k) for (int i = 0; i < 2 * n; ++i) flat[k * AB200_BOX_STRIDE + i] = boxes[k].b[i];
                    tri.assign(boxes.size(), TRI_INDET);
                    if (brc == AB200_OK) brc = ab200_ctx_jump_check(an->ctx, l, e, boxes.size(), flat.data(), tri.data(), &lm);
                });
                if (lm) { J.lie_missing = true; fail(J.c, ST_UNSUPPORTED); }
                else if (brc != AB200_OK) return brc;
            }
            // get_reset (src/set_checker.cpp:206-226)
            CellSet image(an->grid[trg], an->K, trg);
            std::vector<CellBox> cells; feasible.enumerate(cells);
            for (size_t k = 0; k < cells.size() && J.c.status == ST_OK; ++k) {
                J.c.A.top = 0;
                Ivl bx[AB_MAXN]; TM src[AB_MAXN], t[AB_MAXN];
                for (int i = 0; i < n; ++i) bx[i] = ivl(cells[k].b[2 * i], cells[k].b[2 * i + 1]);
                J.box_set(bx, n, src);
                double bb[2 * AB_MAXN];
                for (int i = 0; i < nt; ++i) { t[i] = eval_tm(J.c, S.ops, L.ev[e].reset[i], src); Ivl r = tm_range(J.c, t[i]); bb[2 * i] = r.l; bb[2 * i + 1] = r.u; }
                ok = image.adjoin_outer_approximation(bb, [&](const double* cb) {
                    Ivl cx[AB_MAXN]; for (int i = 0; i < nt; ++i) cx[i] = ivl(cb[2 * i], cb[2 * i + 1]);
                    return J.disjoint(t, nt, cx);
                }) && ok;
            }
            if (J.c.status != ST_OK) {
                ab200_ctx_set_error(an->ctx, J.lie_missing
                    ? "forward_jump_set: Failed to compute Lie derivative of a guard (the reference throws here too, src/function.cpp:1329)"
                    : "forward_jump_set: guard/reset cannot be evaluated on the host (unsupported function or arena)");
                return J.c.status == ST_OK ? AB200_E_UNSUPPORTED : J.c.status;
            }
            result[trg].adjoin(image);
        }
    }
    ok = apply_restriction(an, result) && ok;
    if (!ok) { ab200_ctx_set_error(an->ctx, "jump set lies outside the packed-key window"); return AB200_E_KEYSPACE; }
    return AB200_OK;
}

// ReachabilityRestriction::backward_jump_set (src/reachability_restriction.cpp:381-408): every location's domain
// discretisation (created now if it was still lazy, which also changes what later laps are tuned on) restricted to the
// cells from which a transition may lead into `reach`.
static int backward_jump_set(ab200_analyser* an, const HybridCells& reach, HybridCells& result) {
    const HostSystem* hs = ab200_ctx_system(an->ctx);
    const SystemDesc& S = hs->desc;
    result = empty_cells(an);
    JumpCalc J(S);
    std::vector<CellTree> start; std::vector<std::vector<double> > start_bb(an->n_locs, std::vector<double>(2 * AB_GRID_MAXN, 0.0));
    for (int l = 0; l < an->n_locs; ++l) { start.push_back(CellTree(reach[l])); if (!reach[l].empty()) reach[l].bounding_box(start_bb[l].data()); }
    bool ok = true;
    for (int l = 0; l < an->n_locs; ++l) {
        ok = ensure_restriction(an, l) && ok;
        CellSet src = an->restriction[l];
        src.outer_restrict([&](const double* box) { return J.check_backward(l, box, start, start_bb); });
        if (J.c.status != ST_OK) {
            ab200_ctx_set_error(an->ctx, J.lie_missing
                ? "backward_jump_set: Failed to compute Lie derivative of a guard (the reference throws here too, src/function.cpp:1329)"
                : "backward_jump_set: guard/reset cannot be evaluated on the host (unsupported function or arena)");
            return J.c.status;
        }
        result[l] = src;
    }
    if (!ok) { ab200_ctx_set_error(an->ctx, "domain lies outside the packed-key window"); return AB200_E_KEYSPACE; }
    return AB200_OK;
}

// ReachabilityRestriction::outer_intersection_with(HybridBoundedConstraintSet) for box-only sets
// (src/reachability_restriction.cpp:522-535).
static int discretise_initial(ab200_analyser* an, uint64_t n_init, const int32_t* loc, const double* boxes, HybridCells& initial) {
    initial = empty_cells(an);
    bool ok = true;
    for (uint64_t k = 0; k < n_init; ++k) {
        if (loc[k] < 0 || loc[k] >= an->n_locs) return AB200_E_ARGUMENT;
        CellSet s(an->grid[loc[k]], an->K, loc[k]);
        ok = s.adjoin_outer_approximation(boxes + k * AB200_BOX_STRIDE) && ok;
        initial[loc[k]].adjoin(s);
    }
    ok = apply_restriction(an, initial) && ok;
    if (!ok) { ab200_ctx_set_error(an->ctx, "initial set or domain lies outside the packed-key window"); return AB200_E_KEYSPACE; }
    return AB200_OK;
}

static ab200_cellset* to_cellset(const HybridCells& h);
extern "C" int ab200_analyser_discretise(ab200_analyser* an, uint64_t n_init, const int32_t* loc, const double* boxes, ab200_cellset** out) {
    if (!an || !out || (n_init && (!loc || !boxes))) return AB200_E_ARGUMENT;
    HybridCells initial;
    int rc = discretise_initial(an, n_init, loc, boxes, initial);
    if (rc != AB200_OK) return rc;
    *out = to_cellset(initial);
    return AB200_OK;
}
static ab200_cellset* to_cellset(const HybridCells& h) { ab200_cellset* cs = ab200_cellset_new(); join_keys(h, ab200_cellset_keys(cs)); return cs; }

// ---- constraint-set initial sets ---------------------------------------------------------------------------------------
// ReachabilityRestriction::outer_intersection_with(HybridBoundedConstraintSet) (src/reachability_restriction.cpp:522-533):
// per location the cells of the set's domain box, outer-restricted to the constraint
// (BDDTreeSet::outer_restrict(OpenSetInterface), src/bdd_set.cpp:589-637,1670-1686: a cell is dropped when the set
// definitely does not overlap it, kept whole when the set definitely covers it, kept at the leaves otherwise), then the
// analyser's own restriction. ConstraintSet::overlaps / covers are src/function_set.cpp:218-235: overlaps through the
// subdividing test of include/geometry.h:93-109 to a tolerance of radius(codomain)/8, covers by one interval evaluation.
namespace {
enum { T_FALSE = 0, T_TRUE = 1, T_INDET = 2 };
struct ConstraintEval {
    const HostConstraints* H; const ConstraintSetDesc* S; int n;
    void image(const Ivl* d, Ivl* out, int* st) const {
        Ctx c; ctx_init_host(c, 0, 0);
        for (int i = 0; i < S->m; ++i) out[i] = eval_ivl(c, H->ops.data(), S->f[i], d);
        if (c.status != ST_OK) *st = c.status;
    }
    // disjoint(d, f, b, eps) of include/geometry.h:93-109
    int disjoint(const Ivl* d, double eps, int* st, int depth) const {
        Ivl fd[AB_MAXN], fc[AB_MAXN], mid[AB_MAXN];
        image(d, fd, st);
        for (int i = 0; i < n; ++i) mid[i] = ivl(ivl_mid(d[i]));
        image(mid, fc, st);
        for (int i = 0; i < S->m; ++i) { const Ivl b = ivl(S->codomain[2 * i], S->codomain[2 * i + 1]); if (fd[i].l > b.u || fd[i].u < b.l) return T_TRUE; }
        bool in = true;
        for (int i = 0; i < S->m; ++i) { const Ivl b = ivl(S->codomain[2 * i], S->codomain[2 * i + 1]); if (!(fc[i].l > b.l && fc[i].u < b.u)) in = false; }
        if (in) return T_FALSE;
        double dmax = 0.0; int imax = 0;                                   // Box::radius (include/box.h:93-99), irmax (geometry.h:39-50)
        for (int i = 0; i < n; ++i) { const double w = ivl_width(d[i]); if (w > dmax) { dmax = w; imax = i; } }
        const double half = dmax / 2, rad = half > 0 ? half * (1 + 2e-16) : half * (1 - 2e-16);   // up(dmax/2), include/numeric.h:111
        if (eps == 0 || rad < eps || depth > 60) return T_INDET;
        Ivl l[AB_MAXN], r[AB_MAXN];
        for (int i = 0; i < n; ++i) { l[i] = d[i]; r[i] = d[i]; }
        const double cm = ivl_mid(d[imax]);                                 // med_approx
        l[imax].u = cm; r[imax].l = cm;
        const int a = disjoint(l, eps, st, depth + 1), b = disjoint(r, eps, st, depth + 1);
        if (a == T_FALSE || b == T_FALSE) return T_FALSE;                  // tribool &&
        if (a == T_TRUE && b == T_TRUE) return T_TRUE;
        return T_INDET;
    }
    int check(const double* box, int* st) const {
        Ivl d[AB_MAXN];
        for (int i = 0; i < n; ++i) d[i] = ivl(box[2 * i], box[2 * i + 1]);
        double rc = 0.0;                                                    // radius(codomain): max of the interval radii (src/vector.cpp:312-319)
        for (int i = 0; i < S->m; ++i) rc = std::max(rc, ivl_rad(ivl(S->codomain[2 * i], S->codomain[2 * i + 1])));
        if (disjoint(d, rc / 8, st, 0) == T_TRUE) return 0;                // !possibly(overlaps)
        Ivl fd[AB_MAXN];
        image(d, fd, st);
        bool covers = true;                                                 // definitely(approximate_inside(f(bx),codomain))
        for (int i = 0; i < S->m; ++i) if (!(fd[i].l > S->codomain[2 * i] && fd[i].u < S->codomain[2 * i + 1])) covers = false;
        return covers ? 1 : 2;
    }
};
}
// the discretised initial set of a constraint-set text (shared by ab200_analyser_discretise_constraint and the lower analyses)
static int discretise_constraint_sets(ab200_analyser* an, const char* text, HostConstraints& H, HybridCells& initial) {
    const HostSystem* hs = ab200_ctx_system(an->ctx);
    std::string err = parse_constraints(text, *hs, H);
    if (!err.empty()) { ab200_ctx_set_error(an->ctx, ("constraint set: " + err).c_str()); return AB200_E_ARGUMENT; }
    initial = empty_cells(an);
    bool ok = true; int st = ST_OK;
    for (size_t k = 0; k < H.sets.size(); ++k) {
        const ConstraintSetDesc& S = H.sets[k];
        const int l = S.loc, n = hs->desc.loc[l].n;
        bool empty = false;
        for (int i = 0; i < n; ++i) if (S.domain[2 * i] > S.domain[2 * i + 1]) empty = true;
        if (empty) continue;                                                // has_empty(loc): an empty set in this location
        double dom[AB200_BOX_STRIDE] = {0};
        for (int i = 0; i < 2 * n; ++i) dom[i] = S.domain[i];
        CellSet s(an->grid[l], an->K, l);
        ok = s.adjoin_outer_approximation(dom) && ok;                       // _insert_domain_discretisation
        if (S.m > 0) {                                                      // !unconstrained()
            ConstraintEval E; E.H = &H; E.S = &S; E.n = n;
            s.outer_restrict([&](const double* box) { return E.check(box, &st); }, true);     // (down to the BDD's own depth)
        }
        initial[l].adjoin(s);
    }
    if (st != ST_OK) { ab200_ctx_set_error(an->ctx, "constraint function cannot be evaluated on a cell"); return st; }
    ok = apply_restriction(an, initial) && ok;
    if (!ok) { ab200_ctx_set_error(an->ctx, "initial set or domain lies outside the packed-key window"); return AB200_E_KEYSPACE; }
    return AB200_OK;
}
extern "C" int ab200_analyser_discretise_constraint(ab200_analyser* an, const char* text, ab200_cellset** out) {
    if (!an || !text || !out) return AB200_E_ARGUMENT;
    *out = 0;
    HostConstraints H; HybridCells initial;
    int rc = discretise_constraint_sets(an, text, H, initial);
    if (rc != AB200_OK) return rc;
    *out = to_cellset(initial);
    return AB200_OK;
}
// _enclosures_from_discretised_initial_set_midpoints (src/reachability_analyser.cpp:1218-1243) for constraint-set initial sets:
// the centre of (cell ^ domain of the set) is kept when the set is unconstrained or definitely covers it
// (BoundedConstraintSet::covers, src/function_set.cpp:327-333: the domain covers the point strictly and f(point) lies strictly
// inside the codomain)
static int constraint_midpoint_enclosures(ab200_analyser* an, const HostConstraints& H, const HybridCells& initial,
                                          std::vector<int32_t>& elocs, std::vector<double>& eboxes) {
    const HostSystem* hs = ab200_ctx_system(an->ctx);
    int st = ST_OK;
    for (int l = 0; l < an->n_locs; ++l) {
        const ConstraintSetDesc* S = 0;
        for (size_t k = 0; k < H.sets.size(); ++k) if (H.sets[k].loc == l) { S = &H.sets[k]; break; }
        if (!S) continue;
        const int n = hs->desc.loc[l].n;
        std::vector<CellBox> cells; initial[l].enumerate(cells);
        ConstraintEval E; E.H = &H; E.S = S; E.n = n;
        for (size_t k = 0; k < cells.size(); ++k) {
            double mb[AB200_BOX_STRIDE] = {0}; Ivl d[AB_MAXN];
            bool inside = true;
            for (int i = 0; i < n; ++i) {
                const double lo = std::max(cells[k].b[2 * i], S->domain[2 * i]), hi = std::min(cells[k].b[2 * i + 1], S->domain[2 * i + 1]);
                const double mid = (lo + hi) / 2;
                mb[2 * i] = mid; mb[2 * i + 1] = mid; d[i] = ivl(mid);
                if (!(S->domain[2 * i] < mid && S->domain[2 * i + 1] > mid)) inside = false;      // Ariadne::covers(domain,bx) is strict
            }
            if (S->m > 0) {
                if (!inside) continue;
                Ivl fd[AB_MAXN]; E.image(d, fd, &st);
                bool covers = true;
                for (int i = 0; i < S->m; ++i) if (!(fd[i].l > S->codomain[2 * i] && fd[i].u < S->codomain[2 * i + 1])) covers = false;
                if (!covers) continue;
            }
            elocs.push_back(l); eboxes.insert(eboxes.end(), mb, mb + AB200_BOX_STRIDE);
        }
    }
    if (st != ST_OK) { ab200_ctx_set_error(an->ctx, "constraint function cannot be evaluated on a midpoint"); return st; }
    return AB200_OK;
}

// upper_reach_evolve (src/reachability_analyser.cpp:309-377) from an already discretised initial set
static int upper_reach_evolve_cells(ab200_analyser* an, const HybridCells& initial, double T, int32_t max_events, int mince,
                                    ab200_cellset** reach_out, ab200_cellset** evolve_out, ab200_status* status) {
    *reach_out = 0; *evolve_out = 0;
    an->cells_evolved = 0;
    ab200_status st; memset(&st, 0, sizeof(st)); st.cell = -1;
    double lock = an->lock_time;
    unsigned time_steps = (unsigned)(T / lock);
    double remaining = T - time_steps * lock;
    if (time_steps == 0) { time_steps = 1; remaining = 0.0; lock = T; }
    int rc;
    HybridCells reach = empty_cells(an), evolve = empty_cells(an), found, next;
    // first lap: evolver->orbit per initial enclosure (:348-353)
    rc = lap(an, initial, mince, lock, max_events, 0, 0, found, next, &st);
    if (rc != AB200_OK) { if (status) *status = st; return rc; }
    for (int l = 0; l < an->n_locs; ++l) { reach[l].adjoin(found[l]); evolve[l].adjoin(next[l]); }
    for (unsigned i = 1; i < time_steps; ++i) {
        rc = lap(an, evolve, 0, lock, max_events, 0, 0, found, next, &st);
        if (rc != AB200_OK) { if (status) *status = st; return rc; }
        evolve = next;
        for (int l = 0; l < an->n_locs; ++l) reach[l].adjoin(found[l]);
    }
    if (!all_empty(evolve) && remaining > 0) {
        rc = lap(an, evolve, 0, remaining, max_events, 0, 0, found, next, &st);
        if (rc != AB200_OK) { if (status) *status = st; return rc; }
        evolve = next;
        for (int l = 0; l < an->n_locs; ++l) reach[l].adjoin(found[l]);
    }
    apply_restriction(an, reach);
    apply_restriction(an, evolve);
    *reach_out = to_cellset(reach); *evolve_out = to_cellset(evolve);
    if (status) *status = st;
    return AB200_OK;
}
extern "C" int ab200_analyser_upper_reach_evolve(ab200_analyser* an, uint64_t n_init, const int32_t* loc, const double* boxes,
                                                 double T, int32_t max_events, int mince,
                                                 ab200_cellset** reach_out, ab200_cellset** evolve_out, ab200_status* status) {
    if (!an || !reach_out || !evolve_out) return AB200_E_ARGUMENT;
    *reach_out = 0; *evolve_out = 0;
    HybridCells initial;
    int rc = discretise_initial(an, n_init, loc, boxes, initial);
    if (rc != AB200_OK) return rc;
    return upper_reach_evolve_cells(an, initial, T, max_events, mince, reach_out, evolve_out, status);
}
extern "C" int ab200_analyser_upper_reach_evolve_from(ab200_analyser* an, const ab200_cellset* initial_cells, double T, int32_t max_events, int mince,
                                                      ab200_cellset** reach_out, ab200_cellset** evolve_out, ab200_status* status) {
    if (!an || !initial_cells || !reach_out || !evolve_out) return AB200_E_ARGUMENT;
    HybridCells initial = empty_cells(an);
    split_keys(an, ab200_cellset_keys(const_cast<ab200_cellset*>(initial_cells)), initial);
    return upper_reach_evolve_cells(an, initial, T, max_events, mince, reach_out, evolve_out, status);
}

// _enclosures_from_discretised_initial_set_midpoints (src/reachability_analyser.cpp:1218-1243), box-only initial sets: the
// centre of every discretised initial cell intersected with the initial box of its location, as a degenerate box
static void midpoint_enclosures(ab200_analyser* an, uint64_t n_init, const int32_t* loc, const double* boxes, const HybridCells& initial,
                                std::vector<int32_t>& elocs, std::vector<double>& eboxes) {
    const int dim = an->dim;
    for (int l = 0; l < an->n_locs; ++l) {
        const double* ib = 0;
        for (uint64_t k = 0; k < n_init; ++k) if (loc[k] == l) { ib = boxes + k * AB200_BOX_STRIDE; break; }
        if (!ib) continue;
        std::vector<CellBox> cells; initial[l].enumerate(cells);
        for (size_t k = 0; k < cells.size(); ++k) {
            elocs.push_back(l);
            size_t o = eboxes.size(); eboxes.resize(o + AB200_BOX_STRIDE, 0.0);
            for (int i = 0; i < dim; ++i) {
                double lo = std::max(cells[k].b[2 * i], ib[2 * i]), hi = std::min(cells[k].b[2 * i + 1], ib[2 * i + 1]);   // intersection(cell,domain)
                double mid = (lo + hi) / 2;                                                                          // Box::centre
                eboxes[o + 2 * i] = mid; eboxes[o + 2 * i + 1] = mid;
            }
        }
    }
}

// lower_reach_evolve (src/reachability_analyser.cpp:230-275): the midpoints of the discretised initial cells
// (_enclosures_from_discretised_initial_set_midpoints :1218-1243, box-only initial sets) are evolved for the whole time
// with LOWER_SEMANTICS; every orbit is discretised on its own, checked against the outer domain widened by
// get_epsilon (:1255-1275) and restricted (the evolve set is not). The evolution runs on the device; the per-orbit
// discretisation and the domain check are host work on the returned enclosure boxes -- this is not the hot path.
extern "C" int ab200_ctx_get_settings(const ab200_ctx* ctx, ab200_settings* s);
static int lower_reach_evolve_units(ab200_analyser* an, const std::vector<int32_t>& elocs, const std::vector<double>& eboxes, double T, int32_t max_events,
                                    ab200_cellset** reach_out, ab200_cellset** evolve_out, ab200_status* status);
extern "C" int ab200_analyser_lower_reach_evolve(ab200_analyser* an, uint64_t n_init, const int32_t* loc, const double* boxes,
                                                 double T, int32_t max_events,
                                                 ab200_cellset** reach_out, ab200_cellset** evolve_out, ab200_status* status) {
    if (!an || !reach_out || !evolve_out) return AB200_E_ARGUMENT;
    *reach_out = 0; *evolve_out = 0;
    an->cells_evolved = 0;
    HybridCells initial;
    int rc = discretise_initial(an, n_init, loc, boxes, initial);
    if (rc != AB200_OK) return rc;
    std::vector<int32_t> elocs; std::vector<double> eboxes;
    midpoint_enclosures(an, n_init, loc, boxes, initial, elocs, eboxes);
    return lower_reach_evolve_units(an, elocs, eboxes, T, max_events, reach_out, evolve_out, status);
}
// The same from a constraint-set initial set (text of ab200_analyser_discretise_constraint).
extern "C" int ab200_analyser_lower_reach_evolve_constraint(ab200_analyser* an, const char* constraint_text, double T, int32_t max_events,
                                                            ab200_cellset** reach_out, ab200_cellset** evolve_out, ab200_status* status) {
    if (!an || !constraint_text || !reach_out || !evolve_out) return AB200_E_ARGUMENT;
    *reach_out = 0; *evolve_out = 0;
    an->cells_evolved = 0;
    HostConstraints H; HybridCells initial;
    int rc = discretise_constraint_sets(an, constraint_text, H, initial);
    if (rc != AB200_OK) return rc;
    std::vector<int32_t> elocs; std::vector<double> eboxes;
    rc = constraint_midpoint_enclosures(an, H, initial, elocs, eboxes);
    if (rc != AB200_OK) return rc;
    return lower_reach_evolve_units(an, elocs, eboxes, T, max_events, reach_out, evolve_out, status);
}
static int lower_reach_evolve_units(ab200_analyser* an, const std::vector<int32_t>& elocs, const std::vector<double>& eboxes, double T, int32_t max_events,
                                    ab200_cellset** reach_out, ab200_cellset** evolve_out, ab200_status* status) {
    const int dim = an->dim;
    int rc;
    an->cells_evolved = (int64_t)elocs.size();
    { int trc = tune_evolver(an); if (trc != AB200_OK) return trc; }
    ab200_settings saved, low;
    ab200_ctx_get_settings(an->ctx, &saved);
    low = saved; low.semantics = 1;
    rc = ab200_set_settings(an->ctx, &low);
    if (rc != AB200_OK) return rc;
    ab200_boxlist *rb = 0, *fb = 0; ab200_status st; memset(&st, 0, sizeof(st));
    rc = with_arena_growth(an, [&]() { return ab200_reach_evolve_boxes(an->ctx, elocs.size(), elocs.data(), eboxes.data(), T, max_events, 0, 0, &rb, &fb, &st); });
    { ab200_settings now; ab200_ctx_get_settings(an->ctx, &now); saved.arena_words = now.arena_words; }     // (keep a raised scratch limit)
    ab200_set_settings(an->ctx, &saved);
    if (status) *status = st;
    if (rc != AB200_OK) return rc;
    HybridCells reach = empty_cells(an), evolve = empty_cells(an);
    const double divider = (double)(1 << an->accuracy);
    bool ok = true; int result = AB200_OK;
    const uint64_t *roff, *foff; const double *rbx, *fbx; const int32_t *rl, *fl;
    ab200_boxlist_view(rb, &roff, &rbx, &rl); ab200_boxlist_view(fb, &foff, &fbx, &fl);
    for (size_t e = 0; e < elocs.size() && result == AB200_OK; ++e) {
        HybridCells local_reach = empty_cells(an), local_evolve = empty_cells(an);
        std::vector<std::vector<double> > eps(an->n_locs, std::vector<double>(dim, 0.0));
        for (uint64_t k = roff[e]; k < roff[e + 1]; ++k) {
            const double* b = rbx + k * AB200_BOX_STRIDE;
            ok = local_reach[rl[k]].adjoin_outer_approximation(b) && ok;
            for (int i = 0; i < dim; ++i) eps[rl[k]][i] = std::max(eps[rl[k]][i], ivl_width(ivl(b[2 * i], b[2 * i + 1])));
        }
        for (uint64_t k = foff[e]; k < foff[e + 1]; ++k) ok = local_evolve[fl[k]].adjoin_outer_approximation(fbx + k * AB200_BOX_STRIDE) && ok;
        for (int l = 0; l < an->n_locs; ++l) {
            double bb[2 * AB_GRID_MAXN];
            if (!local_reach[l].bounding_box(bb)) continue;
            for (int i = 0; i < dim; ++i) {
                const double cell = an->grid[l].lengths[i] / divider;
                const double e_i = eps[l][i] + cell;                                   // get_epsilon
                const double lo = (an->domain[l][2 * i] - cell) - e_i, hi = (an->domain[l][2 * i + 1] + cell) + e_i;   // widen(outer_domain_box)
                if (!(bb[2 * i] >= lo && bb[2 * i + 1] <= hi)) {
                    ab200_ctx_set_error(an->ctx, "The lower reach is not inside the over-approximated domain. Check the domain used for variables.");
                    result = AB200_E_DOMAIN;
                }
            }
        }
        apply_restriction(an, local_reach);
        for (int l = 0; l < an->n_locs; ++l) { reach[l].adjoin(local_reach[l]); evolve[l].adjoin(local_evolve[l]); }
    }
    ab200_boxlist_free(rb); ab200_boxlist_free(fb);
    if (result != AB200_OK) return result;
    if (!ok) { ab200_ctx_set_error(an->ctx, "a reached cell lies outside the packed-key window"); return AB200_E_KEYSPACE; }
    *reach_out = to_cellset(reach); *evolve_out = to_cellset(evolve);
    return AB200_OK;
}

// glibc's rand() for a process that never called srand(): the TYPE_3 additive feedback generator of random_r.c seeded with 1
// (r[i] = r[i-31] + r[i-3] after a Lehmer-seeded table and 310 discarded draws; the result is the sum shifted right by one).
// _filter_enclosures (src/reachability_analyser.cpp:704-727) draws from it.
struct GlibcRand {
    std::vector<uint32_t> r;
    GlibcRand() {
        r.resize(34);
        r[0] = 1;
        for (int i = 1; i < 31; ++i) { long long v = (16807LL * (long long)(int32_t)r[i - 1]) % 2147483647LL; if (v < 0) v += 2147483647LL; r[i] = (uint32_t)v; }
        for (int i = 31; i < 34; ++i) r[i] = r[i - 31];
        for (int i = 34; i < 344; ++i) r.push_back(r[i - 31] + r[i - 3]);
    }
    int next() { const size_t k = r.size(); r.push_back(r[k - 31] + r[k - 3]); return (int)(r.back() >> 1); }
};

// epsilon_lower_chain_reach / _epsilon_lower_chain_reach (src/reachability_analyser.cpp:497-627) with LowerReachEpsilonWorker
// (include/workers.h:171-329): see include/ariadne_b200.h. The evolution of a lap's enclosures runs on the device
// (ab200_reach_evolve_enclosures: one warp per enclosure, final set models handed to the next lap); discretising the returned
// boxes, the widened-domain check, the set algebra and the pruning draw are host work on a few hundred enclosures per lap.
int ab200_ctx_swap_system(ab200_ctx* ctx, HostSystem* other);     // ab200_capi.cu: exchange the context's system with *other and upload it
struct SubstitutedSystem {      // _system->substitute_all(midpoints) ... substitute_all(original_parameters)
    ab200_ctx* ctx; HostSystem other; bool active;
    SubstitutedSystem() : ctx(0), active(false) { }
    ~SubstitutedSystem() { if (active) ab200_ctx_swap_system(ctx, &other); }
};
static int epsilon_lower_chain_reach_impl(ab200_analyser* an, uint64_t n_init, const int32_t* loc, const double* boxes, const char* constraint_text,
                                         int32_t prune, const char* substituted_system_text,
                                         ab200_cellset** reach_out, double* epsilon_out, int32_t* laps_out, ab200_status* status);
extern "C" int ab200_analyser_epsilon_lower_chain_reach(ab200_analyser* an, uint64_t n_init, const int32_t* loc, const double* boxes, int32_t prune,
                                                        const char* substituted_system_text,
                                                        ab200_cellset** reach_out, double* epsilon_out, int32_t* laps_out, ab200_status* status) {
    if (!an || !reach_out || (n_init && (!loc || !boxes))) return AB200_E_ARGUMENT;
    return epsilon_lower_chain_reach_impl(an, n_init, loc, boxes, 0, prune, substituted_system_text, reach_out, epsilon_out, laps_out, status);
}
// The same from a constraint-set initial set (text of ab200_analyser_discretise_constraint).
extern "C" int ab200_analyser_epsilon_lower_chain_reach_constraint(ab200_analyser* an, const char* constraint_text, int32_t prune,
                                                                   const char* substituted_system_text,
                                                                   ab200_cellset** reach_out, double* epsilon_out, int32_t* laps_out, ab200_status* status) {
    if (!an || !reach_out || !constraint_text) return AB200_E_ARGUMENT;
    return epsilon_lower_chain_reach_impl(an, 0, 0, 0, constraint_text, prune, substituted_system_text, reach_out, epsilon_out, laps_out, status);
}
static int epsilon_lower_chain_reach_impl(ab200_analyser* an, uint64_t n_init, const int32_t* loc, const double* boxes, const char* constraint_text,
                                         int32_t prune, const char* substituted_system_text,
                                         ab200_cellset** reach_out, double* epsilon_out, int32_t* laps_out, ab200_status* status) {
    *reach_out = 0;
    an->cells_evolved = 0;
    SubstitutedSystem sub;
    if (substituted_system_text) {
        const HostSystem* hs = ab200_ctx_system(an->ctx);
        std::string e = parse_system(substituted_system_text, sub.other);
        if (!e.empty()) { ab200_ctx_set_error(an->ctx, ("substituted system: " + e).c_str()); return AB200_E_ARGUMENT; }
        if (sub.other.loc_names != hs->loc_names || sub.other.desc.loc[0].n != hs->desc.loc[0].n) { ab200_ctx_set_error(an->ctx, "the substituted system has other locations or variables"); return AB200_E_ARGUMENT; }
        sub.ctx = an->ctx;
        int src = ab200_ctx_swap_system(an->ctx, &sub.other);
        if (src != AB200_OK) return src;
        sub.active = true;
    }
    int rc;
    const int dim = an->dim, L = an->n_locs;
    const double divider = (double)(1 << an->accuracy);
    std::vector<int32_t> elocs; std::vector<double> eboxes;
    if (constraint_text) {
        HostConstraints H; HybridCells initial;
        rc = discretise_constraint_sets(an, constraint_text, H, initial);
        if (rc == AB200_OK) rc = constraint_midpoint_enclosures(an, H, initial, elocs, eboxes);
        if (rc != AB200_OK) return rc;
    } else {
        HybridCells initial;
        rc = discretise_initial(an, n_init, loc, boxes, initial);
        if (rc != AB200_OK) return rc;
        midpoint_enclosures(an, n_init, loc, boxes, initial, elocs, eboxes);
    }
    // _restriction->filter(initial_enclosures) (src/reachability_restriction.cpp:186-228): an enclosure whose box the
    // (undiscretised) outer domain covers is kept; otherwise it is kept if its outer approximation meets the domain cells
    {
        std::vector<int32_t> kl; std::vector<double> kb;
        for (size_t k = 0; k < elocs.size(); ++k) {
            const int l = elocs[k]; const double* b = &eboxes[k * AB200_BOX_STRIDE];
            bool keep = false;
            if (!an->discretised[l]) {
                keep = true;
                for (int i = 0; keep && i < dim; ++i) {
                    const double cell = an->grid[l].lengths[i] / divider;
                    if (!(b[2 * i] > an->domain[l][2 * i] - cell && b[2 * i + 1] < an->domain[l][2 * i + 1] + cell)) keep = false;    // Box::covers is strict
                }
            }
            if (!keep) {
                ensure_restriction(an, l);
                CellSet t(an->grid[l], an->K, l);
                t.adjoin_outer_approximation(b); t.restrict(an->restriction[l]);
                keep = !t.empty();
            }
            if (keep) { kl.push_back(l); kb.insert(kb.end(), b, b + AB200_BOX_STRIDE); }
        }
        elocs.swap(kl); eboxes.swap(kb);
    }
    { int trc = tune_evolver(an); if (trc != AB200_OK) return trc; }
    ab200_settings saved, low;
    ab200_ctx_get_settings(an->ctx, &saved);
    low = saved; low.semantics = 1;
    HybridCells reach = empty_cells(an);
    std::vector<std::vector<double> > epsilon(L, std::vector<double>(dim, 0.0));
    ab200_enclosures* current = 0;            // null: the first lap starts from the midpoint boxes
    size_t n_current = elocs.size();
    ab200_status total; memset(&total, 0, sizeof(total)); total.cell = -1;
    GlibcRand rnd;
    int laps = 0, result = AB200_OK; bool ok = true;
    while (n_current > 0 && result == AB200_OK) {
        ++laps;
        if (getenv("AB200_EPS_DEBUG")) fprintf(stderr, "lap %d initial enclosures %zu\n", laps, n_current);
        an->cells_evolved += (int64_t)n_current;
        ab200_boxlist *rb = 0, *fb = 0; ab200_enclosures* fe = 0; ab200_status st; memset(&st, 0, sizeof(st));
        rc = ab200_set_settings(an->ctx, &low);
        if (rc == AB200_OK) rc = with_arena_growth(an, [&]() { return ab200_reach_evolve_enclosures(an->ctx, elocs.size(), elocs.data(), eboxes.data(), current, an->lock_time, 1 /* lock_to_grid_steps */, 0, 0, 0, &rb, &fb, &fe, &st); });
        { ab200_settings now; ab200_ctx_get_settings(an->ctx, &now); saved.arena_words = now.arena_words; low.arena_words = now.arena_words; }   // (keep a raised scratch limit)
        ab200_set_settings(an->ctx, &saved);
        total.steps += st.steps; total.launches += st.launches; total.evolve_ms += st.evolve_ms; total.libm_sensitive += st.libm_sensitive; total.alg_flops += st.alg_flops;
        if (st.arena_peak > total.arena_peak) total.arena_peak = st.arena_peak;
        total.build = st.build;
        if (rc != AB200_OK) { total.code = st.code; total.cell = st.cell; result = rc; break; }
        // LowerReachEpsilonWorker::_compute, in unit order: discretise every orbit, collect widths, cell counts and final sets
        HybridCells local_reach = empty_cells(an), evolve_global = empty_cells(an);
        std::vector<std::vector<double> > local_eps(L, std::vector<double>(dim, 0.0));
        std::vector<uint64_t> superposed(L, 0);
        const uint64_t *roff, *foff; const double *rbx, *fbx; const int32_t *rl, *fl;
        ab200_boxlist_view(rb, &roff, &rbx, &rl); ab200_boxlist_view(fb, &foff, &fbx, &fl);
        for (size_t e = 0; e < n_current; ++e) {
            HybridCells cur_reach = empty_cells(an), cur_evolve = empty_cells(an);
            for (uint64_t k = roff[e]; k < roff[e + 1]; ++k) {
                const double* b = rbx + k * AB200_BOX_STRIDE;
                ok = cur_reach[rl[k]].adjoin_outer_approximation(b) && ok;
                for (int i = 0; i < dim; ++i) local_eps[rl[k]][i] = std::max(local_eps[rl[k]][i], ivl_width(ivl(b[2 * i], b[2 * i + 1])));
            }
            for (uint64_t k = foff[e]; k < foff[e + 1]; ++k) ok = cur_evolve[fl[k]].adjoin_outer_approximation(fbx + k * AB200_BOX_STRIDE) && ok;
            for (int l = 0; l < L; ++l) {
                local_reach[l].adjoin(cur_reach[l]); evolve_global[l].adjoin(cur_evolve[l]);
                if (!cur_evolve[l].empty()) { std::vector<CellBox> cells; cur_evolve[l].enumerate(cells); superposed[l] += cells.size(); }
            }
        }
        const uint64_t n_rb = roff[n_current], n_fb = foff[n_current];
        ab200_boxlist_free(rb); ab200_boxlist_free(fb);
        for (int l = 0; l < L; ++l) for (int i = 0; i < dim; ++i) {
            local_eps[l][i] += an->grid[l].lengths[i] / divider;                           // get_result: plus the minimum cell lengths
            epsilon[l][i] = std::max(epsilon[l][i], local_eps[l][i]);
        }
        // widen(outer_domain_box, local_epsilon).superset(local_reach.bounding_box())
        for (int l = 0; l < L && result == AB200_OK; ++l) {
            double bb[2 * AB_GRID_MAXN];
            if (!local_reach[l].bounding_box(bb)) continue;
            for (int i = 0; i < dim; ++i) {
                const double cell = an->grid[l].lengths[i] / divider;
                const double lo = (an->domain[l][2 * i] - cell) - local_eps[l][i], hi = (an->domain[l][2 * i + 1] + cell) + local_eps[l][i];
                if (!(bb[2 * i] >= lo && bb[2 * i + 1] <= hi)) {
                    ab200_ctx_set_error(an->ctx, "The lower reach is not inside the over-approximated domain. Check the domain used for variables.");
                    result = AB200_E_DOMAIN;
                }
            }
        }
        if (result != AB200_OK) { ab200_enclosures_free(fe); break; }
        apply_restriction(an, local_reach);
        if (getenv("AB200_EPS_DEBUG")) { size_t c = 0; for (int l = 0; l < L; ++l) { std::vector<CellBox> cells; local_reach[l].enumerate(cells); c += cells.size(); } fprintf(stderr, "  reach boxes %llu final boxes %llu reach size before removal %zu\n", (unsigned long long)n_rb, (unsigned long long)n_fb, c); }
        // (no constraint set: _epsilon_lower_chain_reach_constraint_check finds no infeasible cell)
        bool any_new = false;
        for (int l = 0; l < L; ++l) { local_reach[l].remove(reach[l]); if (!local_reach[l].empty()) any_new = true; }
        if (!any_new) { ab200_enclosures_free(fe); break; }
        for (int l = 0; l < L; ++l) reach[l].adjoin(local_reach[l]);
        // _filter_enclosures: the final enclosures become the next initial ones, all of them or a random share
        uint64_t n_final = 0, n_units = 0; const int32_t* flocs = 0; const uint64_t* uoff = 0;
        ab200_enclosures_view(fe, &n_final, &flocs, &uoff, &n_units);
        // the worker appends a unit's final sets location by location (ListSet's std::map, ordered by location name)
        std::vector<uint64_t> order;
        {
            const HostSystem* hs = ab200_ctx_system(an->ctx);
            for (uint64_t u = 0; u < n_units; ++u) {
                const size_t o = order.size();
                for (uint64_t k = uoff[u]; k < uoff[u + 1]; ++k) order.push_back(k);
                std::stable_sort(order.begin() + o, order.end(), [&](uint64_t a, uint64_t b) { return hs->loc_names[flocs[a]] < hs->loc_names[flocs[b]]; });
            }
        }
        std::vector<uint64_t> adjoined(L, 0);
        for (int l = 0; l < L; ++l) if (!evolve_global[l].empty()) { std::vector<CellBox> cells; evolve_global[l].enumerate(cells); adjoined[l] = cells.size(); }
        std::vector<uint64_t> keep;
        for (size_t kk = 0; kk < order.size(); ++kk) {
            const uint64_t k = order[kk];
            if (prune) {
                const int l = flocs[k];
                const double coverage_ratio = (double)adjoined[l] / (double)superposed[l];
                if (keep.size() <= 2 || (double)rnd.next() < 2 * coverage_ratio * 2147483647.0) keep.push_back(k);
            } else keep.push_back(k);
        }
        ab200_enclosures* next = 0;
        rc = ab200_enclosures_select(fe, keep.size(), keep.data(), &next);
        ab200_enclosures_free(fe);
        if (current) ab200_enclosures_free(current);
        current = next; n_current = keep.size();
        if (rc != AB200_OK) { result = rc; break; }
    }
    if (current) ab200_enclosures_free(current);
    if (status) *status = total;
    if (laps_out) *laps_out = laps;
    if (result != AB200_OK) return result;
    if (!ok) { ab200_ctx_set_error(an->ctx, "a reached cell lies outside the packed-key window"); return AB200_E_KEYSPACE; }
    if (epsilon_out) for (int l = 0; l < L; ++l) for (int i = 0; i < AB200_MAX_DIM; ++i) epsilon_out[(size_t)l * AB200_MAX_DIM + i] = i < dim ? epsilon[l][i] : 0.0;
    *reach_out = to_cellset(reach);
    return AB200_OK;
}

static int outer_chain_reach_cells(ab200_analyser* an, HybridCells current, int32_t direction, ab200_cellset** reach_out, ab200_status* status);
extern "C" int ab200_analyser_outer_chain_reach(ab200_analyser* an, uint64_t n_init, const int32_t* loc, const double* boxes,
                                                int32_t direction, ab200_cellset** reach_out, ab200_status* status) {
    if (!an || !reach_out) return AB200_E_ARGUMENT;
    *reach_out = 0;
    HybridCells current;
    int rc = discretise_initial(an, n_init, loc, boxes, current);
    if (rc != AB200_OK) return rc;
    return outer_chain_reach_cells(an, current, direction, reach_out, status);
}
extern "C" int ab200_analyser_outer_chain_reach_from(ab200_analyser* an, const ab200_cellset* initial_cells, int32_t direction,
                                                     ab200_cellset** reach_out, ab200_status* status) {
    if (!an || !initial_cells || !reach_out) return AB200_E_ARGUMENT;
    HybridCells current = empty_cells(an);
    split_keys(an, ab200_cellset_keys(const_cast<ab200_cellset*>(initial_cells)), current);
    return outer_chain_reach_cells(an, current, direction, reach_out, status);
}
// outer_chain_reach (src/reachability_analyser.cpp:380-493) from an already discretised initial set
static int outer_chain_reach_cells(ab200_analyser* an, HybridCells current, int32_t direction, ab200_cellset** reach_out, ab200_status* status) {
    *reach_out = 0;
    an->cells_evolved = 0;
    ab200_status st; memset(&st, 0, sizeof(st)); st.cell = -1;
    int rc;
    HybridCells reach = empty_cells(an), fin = empty_cells(an), new_reach, new_final;
    int guard = 0;
    while (!all_empty(current) && guard++ < 100000) {
        rc = lap(an, current, 0, an->lock_time, 1, 1, direction, new_reach, new_final, &st);       // lock_to_grid_steps=1, ignore_activations
        if (rc != AB200_OK) { if (status) *status = st; return rc; }
        for (int l = 0; l < an->n_locs; ++l) { new_final[l].remove(fin[l]); new_reach[l].remove(reach[l]); }
        apply_restriction(an, new_final); apply_restriction(an, new_reach);
        for (int l = 0; l < an->n_locs; ++l) { reach[l].adjoin(new_reach[l]); fin[l].adjoin(new_final[l]); }
        rc = direction ? backward_jump_set(an, new_reach, current) : forward_jump_set(an, new_reach, current);
        if (rc != AB200_OK) { if (status) { *status = st; status->code = rc; } return rc; }
        for (int l = 0; l < an->n_locs; ++l) current[l].adjoin(new_final[l]);
    }
    *reach_out = to_cellset(reach);
    if (status) *status = st;
    return AB200_OK;
}

extern "C" int ab200_analyser_enumerate(const ab200_analyser* an, const ab200_cellset* cs, int mince, uint64_t* count,
                                        int32_t* locs, double* boxes) {
    if (!an || !cs || !count) return AB200_E_ARGUMENT;
    HybridCells h = empty_cells(an);
    split_keys(an, ab200_cellset_keys(const_cast<ab200_cellset*>(cs)), h);
    uint64_t n = 0;
    for (int l = 0; l < an->n_locs; ++l) {
        std::vector<CellBox> cells; h[l].enumerate(cells, mince ? (1 << 30) : -1);
        for (size_t k = 0; k < cells.size(); ++k) {
            if (locs) locs[n] = l;
            if (boxes) { for (int i = 0; i < AB200_BOX_STRIDE; ++i) boxes[n * AB200_BOX_STRIDE + i] = (i < 2 * an->dim) ? cells[k].b[i] : 0.0; }
            ++n;
        }
    }
    *count = n;
    return AB200_OK;
}
