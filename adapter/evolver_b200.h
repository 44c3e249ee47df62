// adapter/evolver_b200.h -- the EvolverInterface seam: an ImageSetHybridEvolver whose _evolution
// (/root/reference/include/hybrid_evolver-image.h:124-127, src/hybrid_evolver-image.cpp:280-322) runs on libariadne_b200.so.
// Everything else -- settings, tune_settings, orbit / evolve / reach / reach_evolve of EvolverBase
// (include/evolver_base.h:55-79, include/evolver_interface.h:51-124) -- is inherited unchanged, so the object drops into any
// place that takes an EvolverInterface<HybridAutomatonInterface,LocalisedEnclosureType>.
//
// What comes back: the device evolves the enclosure with the reference's own arithmetic and returns, per reach / final
// enclosure, the box the reference's discretiser would see (TaylorSet(bounding_box()).bounding_box()); the lists handed back
// hold TaylorSet(that box). So every returned enclosure CONTAINS the reference's enclosure, and its outer approximation on a
// grid is the same cell set in all but the edge case of a box face within a few ulps of a grid line. The intermediate list
// stays empty (the device does not keep the per-step evolved sets apart from the final ones).
// Initial enclosures must be boxes (TaylorSet(Box), what to_enclosures() and every analyser call produce); for a general
// TaylorSet the bounding box is evolved under UPPER_SEMANTICS and an exception is thrown under LOWER_SEMANTICS.
#ifndef ARIADNE_EVOLVER_B200_H
#define ARIADNE_EVOLVER_B200_H

#include "ariadne.h"
#include "hybrid_evolver-image.h"
#include "../include/ariadne_b200.h"

namespace Ariadne {

class B200ImageSetHybridEvolver : public ImageSetHybridEvolver
{
  public:
    B200ImageSetHybridEvolver(const SystemType& system) : ImageSetHybridEvolver(system) { }
    B200ImageSetHybridEvolver* clone() const { return new B200ImageSetHybridEvolver(*this); }
  protected:
    virtual void _evolution(EnclosureListType& final, EnclosureListType& reachable, EnclosureListType& intermediate,
                            const EnclosureType& initial, const TimeType& time, bool ignore_activations,
                            ContinuousEvolutionDirection direction, Semantics semantics) const;
};

} // namespace Ariadne

#endif
