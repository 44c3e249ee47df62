// adapter/shadow/workers.h -- placed in front of the reference's include directory when src/reachability_analyser.cpp is
// compiled for the B200 (oracle/Makefile, target `adapter`): it includes the reference's own workers.h unchanged, then the
// B200 worker, and points the name the analyser instantiates (src/reachability_analyser.cpp:195) at it. The reference
// source file itself is compiled unmodified. (In a patched reference tree this is the two-line change of INTEGRATION.md.)
#ifndef ARIADNE_WORKERS_SHADOW_H
#define ARIADNE_WORKERS_SHADOW_H
#include_next "workers.h"
#include "workers_b200.h"
#define UpperReachEvolveWorker B200UpperReachEvolveWorker
#define LowerReachEpsilonWorker B200LowerReachEpsilonWorker     /* src/reachability_analyser.cpp:583 */
#endif
