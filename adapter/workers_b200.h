// adapter/workers_b200.h -- the binding a maintainer adds to Ariadne 1.0 to run the grid-reachability lap on
// libariadne_b200.so (include/ariadne_b200.h). Compiled against the reference's own headers; nothing of the reference is
// modified: the worker class below has the constructor and get_result() of UpperReachEvolveWorker
// (/root/reference/include/workers.h:51-81,102-153), and src/reachability_analyser.cpp:192-198 is pointed at it by
// adapter/shadow/workers.h, which oracle/Makefile (target `adapter`) puts in front of the reference's include directory.
#ifndef ARIADNE_WORKERS_B200_H
#define ARIADNE_WORKERS_B200_H

#include <list>
#include <string>
#include <utility>

#include "ariadne.h"
#include "hybrid_evolver-image.h"
// (include the reference's workers.h before this header: it has no include guard)
#include "../include/ariadne_b200.h"

namespace Ariadne {

// The automaton as the op-tape text of release-1.0_b200/csrc/ab200_system.h, through the public expression API only
// (HybridIOAutomaton::modes()/transitions(), DiscreteIOMode::dynamics()/invariants(), DiscreteIOTransition::activation()/
// reset(), Expression<Real>::op()/subexpressions(); constants via evaluate(e,Vector<Interval>(1))). Events are emitted in
// std::set<DiscreteEvent> order, which is the order ab200_set_system_text requires.
std::string b200_system_text(const HybridIOAutomaton& sys);

// A constraint-set initial set in one location, { x in domain : functions(x) in codomain } (BoundedConstraintSet,
// include/function_set.h:143-200), as the `set` block ab200_analyser_discretise_constraint reads. The functions are given
// as the expressions the set was built from (a VectorFunction cannot be walked through the public API).
std::string b200_constraint_text(const HybridIOAutomaton& sys, const DiscreteLocation& loc, const Box& domain,
                                 const List<RealExpression>& functions, const Box& codomain);

class B200UpperReachEvolveWorker
{
  public:
    typedef HybridDenotableSet HDS;
    typedef HybridEvolver::EnclosureType EnclosureType;
    typedef boost::shared_ptr<EvolverInterface<HybridAutomatonInterface,EnclosureType> > EvolverType;

    // Same arguments as UpperReachEvolveWorker (include/workers.h:51-58). The evolver must be the tuned
    // ImageSetHybridEvolver the analyser hands to its workers (its settings carry the step sizes and reference widths).
    B200UpperReachEvolveWorker(const EvolverType& evolver, std::list<EnclosureType>& initial_enclosures, const HybridTime& time,
                               const HybridGrid& grid, const int& accuracy, const bool& ignore_activations,
                               const ContinuousEvolutionDirection& continuous_direction);
    // Consumes the enclosure list and returns (reach, evolve) discretised on `grid` at `accuracy`, as
    // UpperReachEvolveWorker::get_result() does; throws the reference's exceptions for the library's error codes.
    std::pair<HDS,HDS> get_result();

  private:
    HDS _to_denotable(const ab200_cellset* cs) const;
    const EvolverType& _evolver;
    std::list<EnclosureType>& _initial_enclosures;
    const HybridTime& _time;
    const HybridGrid& _grid;
    const int& _accuracy;
    const bool& _ignore_activations;
    const ContinuousEvolutionDirection& _continuous_direction;
};

// LowerReachEpsilonWorker (/root/reference/include/workers.h:171-329) on the library: same constructor and get_result().
// One call evolves every initial TaylorSet for one lock-to-grid lap with LOWER_SEMANTICS on the device
// (ab200_reach_evolve_enclosures: the sets go over as enclosure records, ab200_enclosures_create, and the final sets come
// back as TaylorSets built from ab200_enclosures_model); discretisation, epsilon and the cell counts are taken in the
// order of the input list, i.e. what the reference computes with one worker thread.
class B200LowerReachEpsilonWorker
{
  public:
    typedef HybridDenotableSet HDS;
    typedef HybridEvolver::EnclosureType EnclosureType;
    typedef HybridEvolver::ContinuousEnclosureType CE;
    typedef std::list<EnclosureType> EL;
    typedef std::map<DiscreteLocation,uint> HUM;
    typedef boost::shared_ptr<EvolverInterface<HybridAutomatonInterface,EnclosureType> > EvolverType;

    B200LowerReachEpsilonWorker(const EvolverType& evolver, EL& initial_enclosures, const HybridTime& time,
                                const HybridGrid& grid, const int& accuracy);
    tuple<std::pair<HUM,HUM>,EL,HDS,HybridFloatVector> get_result();

  private:
    const EvolverType& _evolver;
    EL& _initial_enclosures;
    const HybridTime& _time;
    const HybridGrid& _grid;
    const int& _accuracy;
};

} // namespace Ariadne

#endif
