#!/bin/sh
# Regenerates the golden vectors from the compiled reference (oracle/_ref/ref_harness, built by `make -C oracle all
# harness` from the unmodified sources under /root/reference). Run from the repository root in the build container.
set -e
H=oracle/_ref/ref_harness
G=tests/golden
# per-enclosure reach/final boxes + per-enclosure discretisation, van der Pol, accuracy 6, 24 cells spread over the domain
$H vanderpol cells acc=6 mince=1 ncells=24 stride=1050 disc=1 out=$G/vanderpol_cells_acc6.txt
# system descriptions (op tapes) extracted from the reference automata through the public expression API
for s in vanderpol bouncingball watertank springs boost; do $H $s system out=$G/$s.sys; done
# same at accuracy 3 (large cells: long models, early termination on enclosure size)
$H vanderpol cells acc=3 init=1.0,1.3,1.0,1.6 disc=1 out=$G/vanderpol_cells_acc3.txt
# analyser level: upper_reach_evolve over 2 laps, merged frontier between laps
$H vanderpol reach acc=3 init=1.0,1.3,1.0,1.6 laps=2 out=$G/vanderpol_reach_acc3.txt
$H vanderpol reach acc=4 init=-1.6,-1.3,0.3,0.8 laps=1 out=$G/vanderpol_reach_acc4.txt
gzip -9 -f -n $G/vanderpol_cells_acc6.txt
# bouncing ball, chain-reach mode (ignore_activations, 1 event): urgent guard, touching-time and implicit crossing steps
$H bouncingball cells acc=5 init=-3.0,3.0,0.0,2.0 ncells=60 stride=37 out=$G/bouncingball_cells_acc5.txt
$H bouncingball cells acc=4 init=0.0,0.5,1.0,2.0 out=$G/bouncingball_cells_acc4.txt
# bouncing ball, full outer_chain_reach (laps + grid-side jump sets through the urgent bounce), configs[1]
$H bouncingball reach acc=5 out=$G/bouncingball_chain_acc5.txt
$H bouncingball reach acc=6 init=-2.0,0.5,0.5,1.5 out=$G/bouncingball_chain_acc6.txt
# watertank (configs[0] system, short horizon): sqrt dynamics, invariants, permissive open/close jumps through the
# activation-time windows, 4 locations
$H watertank reach acc=3 T=0.5 initr=1.0,1.0,7.6,7.7 initf=0.0,0.0,5.7,5.8 out=$G/watertank_reach_acc3.txt
# springs (configs[3] system): 5 variables (wide keys), urgent sticking with non-transverse crossings
$H springs reach acc=2 out=$G/springs_reach_acc2.txt && gzip -9 -f -n $G/springs_reach_acc2.txt
# boost converter (configs[4] system): 3 locations, min() guard, urgent resets; upper reach over 3 laps. Its
# outer_chain_reach throws in the reference itself (Lie derivative of min(), src/function.cpp:1329)
$H boost reach acc=4 chain=0 T=0.011 out=$G/boost_reach_acc4.txt
# the bench workload itself (configs[2], accuracy 7): 64 frontier cells spread over the 101,520-cell frontier; also the
# sample the algorithmic FP64 operation count of bench.py is taken from (tests/tools/hostdbg.cpp built with -DAB_MUL_HIST)
$H vanderpol cells acc=7 mince=1 ncells=64 stride=1580 init=-2.2,2.2,-2.8,2.8 out=$G/vanderpol_cells_acc7.txt && gzip -9 -f -n $G/vanderpol_cells_acc7.txt
# LOWER_SEMANTICS (EvolverInterface::reach_evolve's Semantics argument): discard on possibly active blocking events with a
# non-negative crossing, no final set on enclosure-size termination, unique blocking event, shrunken activation windows
$H bouncingball cells acc=4 init=0.0,0.5,1.0,2.0 lower=1 out=$G/bouncingball_cells_acc4_lower.txt
$H bouncingball cells acc=4 chain=0 T=1.5 init=0.0,0.5,1.0,2.0 lower=1 out=$G/bouncingball_cells_acc4_lower_events.txt
$H watertank cells acc=3 T=0.5 initr=1.0,1.0,7.6,7.7 initf=0.0,0.0,5.7,5.8 lower=1 out=$G/watertank_cells_acc3_lower.txt
# analyser level: lower_reach_evolve (midpoints of the initial cells, whole horizon, LOWER_SEMANTICS)
$H bouncingball reach acc=4 chain=0 lower=1 T=1.5 init=0.0,0.5,1.0,2.0 out=$G/bouncingball_lower_reach_acc4.txt
$H vanderpol reach acc=4 lower=1 T=1.5 init=1.0,1.5,1.0,1.6 out=$G/vanderpol_lower_reach_acc4.txt
$H watertank reach acc=3 lower=1 T=0.5 initr=1.0,1.0,7.6,7.7 initf=0.0,0.0,5.7,5.8 out=$G/watertank_lower_reach_acc3.txt
# DIRECTION_BACKWARD (the dynamics times constant(-1), include/evolver_base.h:85-90)
$H vanderpol cells acc=4 init=1.0,1.5,1.0,1.6 backward=1 out=$G/vanderpol_cells_acc4_backward.txt
$H bouncingball cells acc=4 init=0.0,0.5,1.0,2.0 backward=1 out=$G/bouncingball_cells_acc4_backward.txt
# configs[0] as examples/watertank-hysteresis.cc sets it up (both idle locations, x in [6,7.5]), horizon 2: several
# transitions per trajectory; replayed by the host-build CPU test (too slow per cell for the GPU suite)
$H watertank reach acc=3 T=2 out=$G/watertank_reach_acc3_T2.txt
$H springs reach acc=3 out=$G/springs_reach_acc3.txt && gzip -9 -f -n $G/springs_reach_acc3.txt
# outer_chain_reach(initial, DIRECTION_BACKWARD): backward laps plus backward_jump_set between them
$H bouncingball reach acc=5 backward=1 init=2.0,4.0,0.0,0.3 out=$G/bouncingball_backward_chain_acc5.txt
# epsilon_lower_chain_reach (src/reachability_analyser.cpp:497-627): lock-to-grid laps of the midpoints with LOWER_SEMANTICS, the
# final TaylorSets of a lap starting the next. One worker thread, so that the pruning draws (rand()) come in unit order.
ORACLE_THREADS=1 $H vanderpol epslower acc=3 prune=0 init=1.0,1.3,1.0,1.6 out=$G/vanderpol_epslower_acc3.txt
ORACLE_THREADS=1 $H bouncingball epslower acc=4 out=$G/bouncingball_epslower_acc4.txt
ORACLE_THREADS=1 $H watertank epslower acc=3 prune=0 out=$G/watertank_epslower_acc3.txt
ORACLE_THREADS=1 $H watertank epslower acc=3 prune=1 out=$G/watertank_epslower_acc3_prune.txt
# five variables (two-word MultiIndex keys in the enclosure records); lock time 1 keeps the reference run short
ORACLE_THREADS=1 $H springs epslower acc=2 prune=1 lock=1.0 out=$G/springs_epslower_acc2_lock1.txt
# the reference's own failure: the lower reach of the boost converter leaves the widened domain in the first lap
ORACLE_THREADS=1 $H boost epslower acc=3 prune=1 out=$G/boost_epslower_acc3_domain_error.txt
# the system the laps evolve: interval parameters at their midpoints (getMidpointsSet, :1000-1014)
$H watertank system midparams=1 out=$G/watertank_midparams.sys
# IntegratorInterface::flow_bounds / TaylorIntegrator::flow (src/integrator.cpp:44-152): the vector field of
# test/test_integrator.cpp (x'=2, y'=3) with its box and step, and the van der Pol / bouncing-ball dynamics
$H constder system out=$G/constder.sys
$H constder integrator box=0.0,1.0,-0.5,1.5 h=1.0 order=4 out=$G/integrator_constder.txt
$H constder integrator box=0.0,1.0,-0.5,1.5 h=1.0 order=1 out=$G/integrator_constder_o1.txt
$H vanderpol integrator box=1.0,1.25,0.5,0.75 h=0.25 order=4 out=$G/integrator_vanderpol_o4.txt
$H vanderpol integrator box=-2.0,-1.9375,0.0,0.0625 h=0.125 order=6 out=$G/integrator_vanderpol_o6.txt
$H bouncingball integrator box=0.0,0.5,1.0,2.0 h=0.5 order=3 out=$G/integrator_bouncingball.txt
# ---- BASELINE depth (configs[0] grid depth 6, configs[3] grid depth 10, configs[4] at accuracy 6): per-enclosure boxes of
# sampled frontier cells over a horizon the compiled reference finishes in seconds, plus the fingerprint of the union of
# their discretisations (what UpperReachEvolveWorker adjoins). A full lock-to-grid lap at these depths costs the reference hours.
D=${TMPDIR:-/tmp}
$H watertank cells acc=6 T=3.0 ncells=24 stride=14 disc=2 out=$D/wt6.txt && python $G/fingerprint_unions.py $D/wt6.txt $G/watertank_cells_acc6.txt && gzip -9 -f -n $G/watertank_cells_acc6.txt
$H springs cells acc=10 T=6.0 ncells=2 disc=2 out=$D/sp10.txt && python $G/fingerprint_unions.py $D/sp10.txt $G/springs_cells_acc10.txt && gzip -9 -f -n $G/springs_cells_acc10.txt
$H boost cells acc=6 chain=0 T=0.011 disc=2 out=$D/bo6.txt && python $G/fingerprint_unions.py $D/bo6.txt $G/boost_cells_acc6.txt && gzip -9 -f -n $G/boost_cells_acc6.txt
# ---- constraint-set initial sets (ReachabilityRestriction::outer_intersection_with(HybridBoundedConstraintSet),
# src/reachability_restriction.cpp:413-434,522-533): a disc and a two-function band as initial sets of upper_reach_evolve and
# outer_chain_reach; the dumps hold the constraint as text (cons| lines), the initial cells and the analysis result
$H vanderpol reach acc=4 cons=disc init=0.9,1.5,1.0,1.6 laps=1 out=$G/vanderpol_cons_disc_acc4.txt
$H vanderpol reach acc=6 cons=band init=0.5,2.0,0.5,2.0 T=0.02 out=$G/vanderpol_cons_band_acc6.txt
$H bouncingball reach acc=5 cons=disc init=-0.5,0.5,1.5,2.0 cc=0.0,1.8 cr=0.2 out=$G/bouncingball_cons_disc_chain_acc5.txt
# a constraint set over the default (large) initial box: the finest levels of the box's paving are merged, so
# BDDTreeSet::outer_restrict stops above the accuracy (BDDTreeSet::depth()); lower analyses start from the midpoints the set covers
ORACLE_THREADS=1 $H vanderpol reach lower=1 cons=disc acc=4 T=1.5 out=$G/vanderpol_cons_disc_lower_reach_acc4.txt
ORACLE_THREADS=1 $H vanderpol epslower cons=disc acc=4 prune=1 out=$G/vanderpol_cons_disc_epslower_acc4.txt
# exp / log / sin / cos on the tapes: a pendulum with transcendental terms (not a BASELINE config), three cells, T=1
$H pendulum system out=$G/pendulum.sys
$H pendulum cells acc=4 T=1.0 ncells=12 stride=1 out=$G/pendulum_cells_acc4.txt
