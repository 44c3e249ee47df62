/* ariadne_b200.h -- C ABI of libariadne_b200.so, the B200 implementation of Ariadne 1.0's grid-reachability lap.
 *
 * The reference has no FFI: its seams are C++ virtual interfaces. Each entry point below names
 * the reference interface it replaces (paths relative to the reference tree). Conventions: every call returns
 * AB200_OK or an AB200_E_* code; inputs are borrowed for the duration of the call; outputs are library-owned until
 * the matching *_free; one host thread drives a context; all GPU work is stream-ordered. There is no CPU fallback:
 * every compute entry point fails with AB200_E_CUDA when no sm_100 device is usable.
 *
 * Boxes are passed as [lo0,hi0,lo1,hi1,...] with AB200_BOX_STRIDE doubles per box regardless of the dimension.
 */
#ifndef ARIADNE_B200_H
#define ARIADNE_B200_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AB200_MAX_DIM 5
#define AB200_BOX_STRIDE (2 * AB200_MAX_DIM)

/* Error codes and the reference exception each one stands for (include/exceptions.h, include/workers.h:129-139). */
enum {
    AB200_OK = 0,
    AB200_E_SQRT = 1,          /* SqrtNumericException          src/numeric.cpp:251-253 */
    AB200_E_DIVZERO = 2,       /* DivideByZeroException         src/numeric.cpp:744,829,858 */
    AB200_E_DOMAIN = 3,        /* DomainException               src/taylor_model.cpp:1847 */
    AB200_E_BOUNDER = 4,       /* ARIADNE_ASSERT in flow_bounds src/taylor_calculus.cpp:87,108,123-130 -> std::runtime_error */
    AB200_E_ARENA = 5,         /* per-cell scratch exhausted: raise ab200_settings.arena_words */
    AB200_E_OUTPUT = 6,        /* per-cell box list exhausted: raise ab200_settings.max_boxes */
    AB200_E_UNSUPPORTED = 7,   /* operator or event path not implemented on the device yet */
    AB200_E_IMPLICIT = 8,      /* ImplicitFunctionException     src/taylor_model.cpp:3562-3665 */
    AB200_E_WORKING_SETS = 9,  /* WorkingSetTooLargeException   src/hybrid_evolver-image.cpp:280 */
    AB200_E_CUDA = 100,        /* CUDA runtime failure or no usable GPU */
    AB200_E_ARGUMENT = 101,    /* malformed argument */
    AB200_E_KEYSPACE = 102,    /* a reached cell fell outside the packed-key window */
    AB200_E_TIMEOUT = 103      /* TimeoutException              include/interruptible.h:56-63 */
};

typedef struct ab200_ctx ab200_ctx;
typedef struct ab200_cellset ab200_cellset;   /* sorted, duplicate-free packed cell keys */
typedef struct ab200_boxlist ab200_boxlist;   /* per-enclosure bounding boxes (parity / debugging) */
typedef struct ab200_analyser ab200_analyser;
typedef struct ab200_enclosures ab200_enclosures;   /* TaylorSet enclosures (set models) in the device's record format */

/* ImageSetHybridEvolverSettings + TaylorCalculus orders (src/hybrid_evolver-image.cpp:1347-1359,
 * src/taylor_calculus.cpp:136-146) plus device sizing. ab200_default_settings fills the reference defaults. */
typedef struct ab200_settings {
    int32_t temporal_order;             /* 6 */
    int32_t spacial_order;              /* 6 */
    double  sweep_threshold;            /* 1e-12 */
    double  maximum_enclosure_widths_ratio;   /* 5.0 */
    int32_t max_boxes;                  /* capacity of each per-cell reach / final list */
    int32_t arena_words;                /* 64-bit words of scratch per cell */
    int32_t hash_log2;                  /* log2 of the de-duplication table size */
    int32_t batch_cells;                /* cells per launch */
    double  ttl_seconds;                /* Interruptible::ttl; <=0 disables */
    int32_t semantics;                  /* 0 = UPPER_SEMANTICS (default), 1 = LOWER_SEMANTICS: the Semantics argument of
                                           EvolverInterface::reach_evolve / orbit (include/evolver_interface.h:46-49,114-124) */
    int32_t max_working_sets;           /* ImageSetHybridEvolverSettings::maximum_number_of_working_sets
                                           (src/hybrid_evolver-image.cpp:280,1358). The reference default 0 means unlimited; here 0
                                           means the device default of 64 pending sets per cell, and the largest accepted value
                                           is 256 (the handles live in the cell's arena). AB200_E_WORKING_SETS beyond it. */
} ab200_settings;

typedef struct ab200_status {
    int32_t code;             /* first failing cell's code (first error wins, include/workers.h:129-139) */
    int64_t cell;             /* index of that cell in the batch, -1 if none */
    int64_t steps;            /* total _evolution_step calls executed */
    int64_t keys_emitted;     /* lattice cells written by the rasteriser before de-duplication */
    int32_t arena_peak;       /* largest per-cell scratch use, words */
    double  evolve_ms;        /* device time of the evolution kernel(s) */
    double  raster_ms;        /* device time of rasterise + compact + sort */
    int64_t launches;         /* kernels launched by this call */
    int32_t build;            /* evolution build used by the last batch: 0 = one warp per cell (one-word MultiIndex keys),
                                 1 = one warp per cell (two-word keys), 2 = one lane per cell (very large frontiers) */
    int32_t libm_sensitive;   /* rec / sqrt model calls of this call whose series degree or truncation bound could change with
                                 the last bits of log() / pow() (glibc in the reference, CUDA on the device). 0 certifies that no
                                 result of the call depends on them. Counted by the warp builds (build 0 / 1); the lane build
                                 (build 2) does not count and always reports 0 */
    double  alg_flops;        /* FP64 operations of the REFERENCE algorithm executed by this call, counted per primitive on the
                                 actual term counts with the formulas of SURVEY.md 8(d) (0 when the lane build ran) */
} ab200_status;

void ab200_default_settings(ab200_settings* s);

/* Lifetime. One context per GPU (device = CUDA ordinal). Replaces the evolver object created by
 * HybridReachabilityAnalyser::_get_tuned_evolver (src/reachability_analyser.cpp:125-142). */
int  ab200_ctx_create(ab200_ctx** out, int device);
void ab200_ctx_destroy(ab200_ctx* ctx);
const char* ab200_last_error(const ab200_ctx* ctx);

/* System upload: flat op tapes of every location's dynamics, guards, invariants and resets, in the text form
 * documented in release-1.0_b200/csrc/ab200_system.h. Replaces the per-call function reconstruction of
 * HybridAutomatonInterface::dynamic_function/guard_function/invariant_function/reset_function
 * (include/hybrid_automaton_interface.h:96-141, src/hybrid_io_automaton.cpp:938-1026). */
int  ab200_set_system_text(ab200_ctx* ctx, const char* text);
int  ab200_system_info(const ab200_ctx* ctx, int32_t* n_locations, int32_t* dims /* [n_locations] or NULL */);
/* Host-only validation of a system text (no context, no GPU): parses it exactly as ab200_set_system_text does and
 * reports the number of locations, events and tape operations, or the parser's message. Lets the adapter that walks
 * HybridIOAutomaton::modes()/transitions() be checked on a machine without a device. */
int  ab200_system_check(const char* text, int32_t* n_locations, int32_t* n_events, int32_t* n_ops, char* err, int32_t errlen);
int  ab200_location_index(const ab200_ctx* ctx, const char* name);

int  ab200_set_settings(ab200_ctx* ctx, const ab200_settings* s);
int  ab200_ctx_get_settings(const ab200_ctx* ctx, ab200_settings* s);

/* ImageSetHybridEvolver::tune_settings (src/hybrid_evolver-image.cpp:99-113): per-location maximum step size and
 * reference enclosure widths, and the grid + accuracy used by outer_approximation. */
int  ab200_set_location_tuning(ab200_ctx* ctx, int32_t loc, double max_step, const double* ref_widths,
                               const double* grid_origin, const double* grid_lengths);
int  ab200_set_accuracy(ab200_ctx* ctx, int32_t accuracy);
/* The same with the packed cell keys sized to the analysed domain (n_locations boxes of AB200_BOX_STRIDE doubles, lo/hi per
 * coordinate; what HybridReachabilityAnalyser is constructed with, src/reachability_analyser.cpp:83-100). A BDDTreeSet has
 * no index range; a 64-bit key has, and the uniform split of ab200_set_accuracy does not hold a 5-variable system at
 * accuracy 10 (BASELINE configs[3]). The uniform layout is kept whenever it fits, so keys do not change for smaller systems. */
int  ab200_set_accuracy_for_domain(ab200_ctx* ctx, int32_t accuracy, const double* domain_boxes);

/* The hot call == UpperReachEvolveWorker::get_result() (include/workers.h:51-81,102-153): evolve every enclosure for
 * (t_continuous, max_events), discretise reach and final sets at the context's accuracy. */
int  ab200_upper_reach_evolve(ab200_ctx* ctx, uint64_t n_cells, const int32_t* loc, const double* boxes,
                              double t_continuous, int32_t max_events, int ignore_activations, int direction,
                              ab200_cellset** reach_out, ab200_cellset** evolve_out, ab200_status* status);
int  ab200_cellset_view(const ab200_cellset* cs, const uint64_t** keys, uint64_t* count);
void ab200_cellset_free(ab200_cellset* cs);
/* Decode a packed key into (location, lattice index per coordinate at resolution 2^-accuracy of the base grid). */
int  ab200_key_decode(const ab200_ctx* ctx, uint64_t key, int32_t* loc, int64_t* idx);

/* == EvolverInterface::reach_evolve (include/evolver_interface.h:114-124) for a batch, returning what the
 * discretiser consumes: the bounding box of every reach and final enclosure (include/hybrid_set.h:976-996). */
int  ab200_reach_evolve_boxes(ab200_ctx* ctx, uint64_t n_cells, const int32_t* loc, const double* boxes,
                              double t_continuous, int32_t max_events, int ignore_activations, int direction,
                              ab200_boxlist** reach_out, ab200_boxlist** final_out, ab200_status* status);
/* offsets has n_cells+1 entries; boxes/locs have offsets[n_cells] entries. */
int  ab200_boxlist_view(const ab200_boxlist* bl, const uint64_t** offsets, const double** boxes, const int32_t** locs);
void ab200_boxlist_free(ab200_boxlist* bl);

/* == EvolverInterface::reach_evolve(initial_enclosure, time, semantics) (include/evolver_interface.h:95-124) with the
 * enclosures themselves handed over, as LowerReachEpsilonWorker::_compute needs (include/workers.h:264-303: the final
 * TaylorSets of one lap are the initial sets of the next). Unit i starts from initial enclosure i when `initial` is given
 * (n_units, loc and boxes are then ignored) and from box i otherwise. final_enclosures receives every final set as a
 * TaylorSet -- the set models as the reference keeps them in Orbit::final(), src/hybrid_evolver-image.cpp:304 -- in unit
 * order; the box lists are those of ab200_reach_evolve_boxes. max_enclosure_words bounds the records of one unit
 * (0 = 16384 words; AB200_E_OUTPUT beyond it). The semantics is ab200_settings.semantics. Warp builds only. */
int  ab200_reach_evolve_enclosures(ab200_ctx* ctx, uint64_t n_units, const int32_t* loc, const double* boxes, const ab200_enclosures* initial,
                                   double t_continuous, int32_t max_events, int ignore_activations, int direction, int32_t max_enclosure_words,
                                   ab200_boxlist** reach_out, ab200_boxlist** final_out, ab200_enclosures** final_enclosures, ab200_status* status);
/* count enclosures with their locations; unit_offsets has n_units+1 entries (enclosures of unit u: [unit_offsets[u], unit_offsets[u+1])). */
int  ab200_enclosures_view(const ab200_enclosures* e, uint64_t* count, const int32_t** locs, const uint64_t** unit_offsets, uint64_t* n_units);
/* Coordinate model k of enclosure i as a TaylorModel: exps is [n_terms][nargs] (MultiIndex), coefs [n_terms]; pass NULL for
 * exps / coefs to query nargs, n_terms and the uniform error only. */
int  ab200_enclosures_model(const ab200_enclosures* e, uint64_t i, int32_t k, int32_t* nargs, int32_t* n_terms, double* error, uint8_t* exps, double* coefs);
/* Enclosures from the caller's TaylorModels (the reverse of ab200_enclosures_model), for an adapter that receives TaylorSets
 * from the reference (LowerReachEpsilonWorker's initial_enclosures, include/workers.h:171-200). Enclosure i lies in location
 * locs[i], has nargs[i] parameters and one model per state variable of that location; models are numbered across all
 * enclosures in order (n_terms[m], errors[m]); exps ([terms][nargs of the enclosure]) and coefs run on across the models.
 * Terms in the reference's Expansion order (graded, include/multi_index.h:346-358); AB200_E_ARGUMENT otherwise. */
int  ab200_enclosures_create(ab200_ctx* ctx, uint64_t count, const int32_t* locs, const int32_t* nargs, const int32_t* n_terms,
                             const double* errors, const uint8_t* exps, const double* coefs, ab200_enclosures** out);
/* The enclosures idx[0..n) of src as a new list (ReachabilityRestriction::filter, _filter_enclosures keep a subset). */
int  ab200_enclosures_select(const ab200_enclosures* src, uint64_t n, const uint64_t* idx, ab200_enclosures** out);
void ab200_enclosures_free(ab200_enclosures* e);

/* == IntegratorInterface::flow_bounds (include/integrator_interface.h:45-50) with the TaylorCalculus constants
 * (src/taylor_calculus.cpp:55-133,584-600). */
int  ab200_flow_bounds(ab200_ctx* ctx, uint64_t n, const int32_t* loc, const double* boxes, double hmax,
                       double* h_out, double* bounds_out, int32_t* status_out);

/* == IntegratorBase::flow_bounds (src/integrator.cpp:44-113): the integrator's own constants (8 step halvings, 4 refinements,
 * no widening of singleton coordinates); this is what IntegratorInterface::flow_bounds of a TaylorIntegrator computes. */
int  ab200_integrator_flow_bounds(ab200_ctx* ctx, uint64_t n, const int32_t* loc, const double* boxes, double hmax,
                                  double* h_out, double* bounds_out, int32_t* status_out);
/* == TaylorIntegrator(temporal_order)::flow(vector_field, state_domain, suggested_time_step)
 * (include/integrator.h:56-68, src/integrator.cpp:116-152) for the dynamics of location `loc`: the VectorTaylorFunction of
 * the flow over domain = box x [-h,h] (written to domain_out, 2*(n+1) doubles), returned as one polynomial model per state
 * variable in n+1 scaled arguments (exponent bytes per term, coefficient, uniform error). */
typedef struct ab200_taylor_models ab200_taylor_models;
int  ab200_integrator_flow(ab200_ctx* ctx, int32_t loc, const double* box, double hmax, int32_t temporal_order,
                           double* h_out, double* domain_out, ab200_taylor_models** out);
int  ab200_taylor_models_view(const ab200_taylor_models* m, int32_t i, int32_t* nargs, int32_t* n_terms, double* error,
                              const uint8_t** exponents /* [n_terms][nargs] */, const double** coefficients);
void ab200_taylor_models_free(ab200_taylor_models* m);

/* Primitive-level entry points (rows a8 / a15 of the scope table): ONE TaylorModel operator or function of
 * include/taylor_model.h / src/taylor_model.cpp, run by one warp with the kernels' own primitives, and a batch of Interval
 * operations of include/numeric.h. They exist so that the reference's unit-test vectors (test/test_taylor_model.cpp:164-270,
 * test/test_interval.cpp:152-254) can be replayed on the device bit for bit; models are given and returned as
 * (exponent bytes [n_terms][nargs] in the reference's graded order, coefficients, uniform error), default Accuracy.
 * op: 0 x+y, 1 x-y, 2 x*y, 3 x+s, 4 x-s, 5 x*s, 6 x/s, 7 x+I, 8 x-I, 9 x*I, 10 x/I, 11 -x, 12 intersection(x,y),
 *     13 split(x,k,flag), 14 antiderivative(x,k), 15 rec(x), 16 sqrt(x), 17 abs(x), 18 max(x,y), 19 x.range() -> range_out,
 *     20 exp(x), 21 log(x), 22 sin(x), 23 cos(x). */
int  ab200_taylor_model_op(ab200_ctx* ctx, int32_t op, int32_t nargs,
                           int32_t x_terms, const uint8_t* x_exps, const double* x_coefs, double x_error,
                           int32_t y_terms, const uint8_t* y_exps, const double* y_coefs, double y_error,
                           int32_t k, int32_t flag, double s, const double* interval /* [2] */, double* range_out /* [2] */,
                           ab200_taylor_models** out);
/* ops[i] = {op, m, a.lower, a.upper, b.lower, b.upper}, results[i] = {lower, upper, status}.
 * op: 0 a+b, 1 a-b, 2 a*b, 3 a/b, 4 a+Float(b.lower), 5 a-Float, 6 a*Float, 7 a/Float, 8 Float(a.lower)-b, 9 Float/b,
 *     10 sqr(a), 11 sqrt(a), 12 pow(a,m), 13 rec(a), 14 abs(a), 15 -a, 16 max(a,b), 17 min(a,b), 18 exp(a), 19 log(a),
 *     20 sin(a), 21 cos(a). */
int  ab200_interval_op(ab200_ctx* ctx, int32_t n, const double* ops, double* results);

/* Host mirror of HybridReachabilityAnalyser (include/reachability_analyser.h:122-172): grid, lock-to-grid time and
 * evolver tuning derived from the domain exactly as HybridReachabilityAnalyser(system,domain,accuracy) does
 * (src/reachability_analyser.cpp:83-100,1028-1170; src/hybrid_set.cpp:485-515), laps driven through
 * ab200_upper_reach_evolve, set algebra on sorted key arrays. domain: one box per location. */
int  ab200_analyser_create(ab200_analyser** out, ab200_ctx* ctx, const double* domain_boxes, int32_t accuracy);
void ab200_analyser_destroy(ab200_analyser* an);
int  ab200_analyser_info(const ab200_analyser* an, double* lock_to_grid_time, double* grid_origin /* [L][MAX_DIM] */,
                         double* grid_lengths, double* max_step /* [L] */);
int  ab200_analyser_set_lock_time(ab200_analyser* an, double lock_to_grid_time);
/* ReachabilityRestriction::outer_intersection_with(initial set) (src/reachability_restriction.cpp:522-535): the
 * discretised initial set the first lap starts from. */
int  ab200_analyser_discretise(ab200_analyser* an, uint64_t n_init, const int32_t* loc, const double* boxes,
                               ab200_cellset** out);
/* upper_reach_evolve (src/reachability_analyser.cpp:309-377). initial: n_init boxes with their locations.
 * mince != 0 enumerates the initial set as uniform cells (BDDTreeSet::mince) instead of merged cells. */
int  ab200_analyser_upper_reach_evolve(ab200_analyser* an, uint64_t n_init, const int32_t* loc, const double* boxes,
                                       double t_continuous, int32_t max_events, int mince,
                                       ab200_cellset** reach_out, ab200_cellset** evolve_out, ab200_status* status);
/* outer_chain_reach (src/reachability_analyser.cpp:380-493). */
/* HybridReachabilityAnalyser::lower_reach_evolve (src/reachability_analyser.cpp:230-275): midpoints of the discretised
 * initial cells evolved with LOWER_SEMANTICS for the whole time; AB200_E_DOMAIN if an orbit leaves the widened domain
 * (the reference's DomainException). lower_reach / lower_evolve are its two halves (:208-227). */
int  ab200_analyser_lower_reach_evolve(ab200_analyser* an, uint64_t n_init, const int32_t* loc, const double* boxes,
                                       double T, int32_t max_events,
                                       ab200_cellset** reach, ab200_cellset** evolve, ab200_status* status);
/* HybridReachabilityAnalyser::epsilon_lower_chain_reach (src/reachability_analyser.cpp:497-627) with
 * LowerReachEpsilonWorker (include/workers.h:171-329) on the device, for an empty constraint set and singleton parameters
 * (the verifier's parameter splitting and constraint checking are out of scope): the midpoints of the discretised initial
 * cells are evolved with LOWER_SEMANTICS one lock-to-grid lap at a time, the final TaylorSets of a lap being the initial
 * sets of the next, until a lap reaches no new cell. epsilon_out is [n_locations][AB200_MAX_DIM]: per location the largest
 * reach-enclosure widths plus one cell (HybridFloatVector). prune = Settings::enable_lower_pruning (the reference default is
 * 1): a final enclosure is kept with probability 2 * adjoined / superposed evolve cells of its location, drawn from glibc's
 * rand() sequence for seed 1 in unit order -- what the reference draws in a fresh process whose worker runs one thread;
 * with more threads the reference's own result depends on their completion order. AB200_E_DOMAIN as in lower_reach_evolve.
 * substituted_system_text (or NULL): the system with its nonsingleton parameters at their midpoints, in the format of
 * ab200_set_system_text. The reference evolves that system (_system->substitute_all(midpoints), :517-527) while the grid and
 * the lock time stay those of the analyser's own system; the context's system is put back before the call returns. */
int  ab200_analyser_epsilon_lower_chain_reach(ab200_analyser* an, uint64_t n_init, const int32_t* loc, const double* boxes, int32_t prune,
                                              const char* substituted_system_text,
                                              ab200_cellset** reach_out, double* epsilon_out, int32_t* laps_out, ab200_status* status);
/* lower_reach_evolve / epsilon_lower_chain_reach from a constraint-set initial set (HybridBoundedConstraintSet, the text of
 * ab200_analyser_discretise_constraint): the midpoint of (initial cell ^ domain of the set) starts an evolution when the set
 * is unconstrained or definitely covers it (_enclosures_from_discretised_initial_set_midpoints,
 * src/reachability_analyser.cpp:1218-1243; BoundedConstraintSet::covers, src/function_set.cpp:327-333). */
int  ab200_analyser_lower_reach_evolve_constraint(ab200_analyser* an, const char* constraint_text, double T, int32_t max_events,
                                                  ab200_cellset** reach, ab200_cellset** evolve, ab200_status* status);
int  ab200_analyser_epsilon_lower_chain_reach_constraint(ab200_analyser* an, const char* constraint_text, int32_t prune,
                                                         const char* substituted_system_text,
                                                         ab200_cellset** reach_out, double* epsilon_out, int32_t* laps_out, ab200_status* status);
/* direction: 0 = DIRECTION_FORWARD (forward_jump_set between laps), 1 = DIRECTION_BACKWARD (backward_jump_set,
 * src/reachability_restriction.cpp:381-408, src/set_checker.cpp:229-352). */
int  ab200_analyser_outer_chain_reach(ab200_analyser* an, uint64_t n_init, const int32_t* loc, const double* boxes,
                                      int32_t direction, ab200_cellset** reach_out, ab200_status* status);
/* Number of frontier enclosures consumed by the evolver in the last analyser call (the metric's unit of work). */
int64_t ab200_analyser_cells_evolved(const ab200_analyser* an);
/* Enumerate a cell set as the reference does (maximal dyadic cells, boxes by bisection, src/bdd_set.cpp:800-891).
 * Pass boxes==NULL to query the count. */
/* Constraint-set initial sets: ReachabilityRestriction::outer_intersection_with(HybridBoundedConstraintSet)
 * (src/reachability_restriction.cpp:413-434,522-533) with ConstraintSet::overlaps / covers (src/function_set.cpp:218-235,
 * include/geometry.h:93-109) deciding every cell, i.e. what upper_reach / outer_chain_reach do with their initial_set
 * argument before the first lap. `text` holds one `set` per location (domain box, m constraint functions as op tapes over the
 * location's variables, codomain box; release-1.0_b200/csrc/ab200_system.h). The *_from calls run the analyses from such a set. */
int  ab200_analyser_discretise_constraint(ab200_analyser* an, const char* text, ab200_cellset** out);
int  ab200_analyser_upper_reach_evolve_from(ab200_analyser* an, const ab200_cellset* initial, double t_continuous, int32_t max_events, int mince,
                                            ab200_cellset** reach_out, ab200_cellset** evolve_out, ab200_status* status);
int  ab200_analyser_outer_chain_reach_from(ab200_analyser* an, const ab200_cellset* initial, int32_t direction,
                                           ab200_cellset** reach_out, ab200_status* status);
int  ab200_analyser_enumerate(const ab200_analyser* an, const ab200_cellset* cs, int mince, uint64_t* count,
                              int32_t* locs, double* boxes);

/* Multi-GPU laps (replaces the thread farm of UpperReachEvolveWorker, include/workers.h:102-153, across GPUs).
 * One process per GPU, one context per process. Rank 0 obtains an id with ab200_comm_unique_id and hands the 128 bytes to
 * the other ranks by whatever the host program uses (MPI, torch.distributed, a file); every rank then calls
 * ab200_comm_init. From then on every lap of the context (ab200_upper_reach_evolve and the analyser calls built on it) is
 * sharded: all ranks pass the same frontier, blocks of 64 cells are dealt round-robin, each rank evolves and rasterises its
 * share, and the sorted cell keys are all-gathered with NCCL and merged on the device, so every rank receives the reach /
 * evolve sets of the whole lap. A failing cell on any rank fails the lap on every rank. ab200_status counts this rank's
 * share. NCCL (libnccl.so.2) is loaded on first use. */
int  ab200_comm_unique_id(uint8_t* id128);
int  ab200_comm_init(ab200_ctx* ctx, int32_t n_ranks, int32_t rank, const uint8_t* id128);
void ab200_comm_destroy(ab200_ctx* ctx);
int  ab200_comm_info(const ab200_ctx* ctx, int32_t* n_ranks, int32_t* rank);

/* Device probes used by bench.py for the roofline denominators. */
int  ab200_probe_fp64(ab200_ctx* ctx, double* gflops_rounded_ops);      /* DMUL.RD + DADD.RU chains: what this library can issue */
int  ab200_probe_fp64_fma(ab200_ctx* ctx, double* gflops_fma);          /* DFMA chains, 2 operations each: the nominal peak */

#ifdef __cplusplus
}
#endif
#endif
