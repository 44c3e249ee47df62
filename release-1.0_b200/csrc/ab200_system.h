// ab200_system.h -- host-side construction of the flat op-tape system description (SystemDesc).
//
// The reference rebuilds ScalarFunction/VectorFunction objects from expression trees on every call
// (src/hybrid_io_automaton.cpp:938-1026) and interprets them with dynamic_cast dispatch
// (src/expression.cpp:766-779). Here every DiscreteLocation's dynamics, guards, invariants and resets are
// compiled ONCE into postfix tapes that the kernels read from constant memory.
//
// Text form accepted by parse_system() (one token stream, '#' starts a comment):
//   system <name>
//   locations <L>
//   location <idx> <name> n <n> events <E>
//     dyn <i> <tape...> ;
//     event <k> <name> kind <invariant|urgent|permissive> target <loc idx> id <int>
//       guard <tape...> ;
//       lie <tape...> ;                  (transitions only: Lie derivative of the guard along the dynamics)
//       reset <i> <tape...> ;            (transitions only)
//   The events of a location must be listed in ascending order of their names (the order of std::set<DiscreteEvent>, which the
//   reference's evolver iterates; adapter/workers_b200.cpp emits them that way); parse_system rejects any other order.
//   tape tokens: V<j> | C<lo>,<hi> (hex or decimal doubles) | ADD SUB MUL DIV NEG POS REC SQR SQRT EXP LOG SIN COS TAN ABS MAX MIN
#pragma once
#include <string>
#include <vector>
#include "ab200_calc.cuh"

namespace ab200 {

struct HostSystem {
    std::string name;
    std::vector<std::string> loc_names;
    std::vector<std::vector<std::string> > event_names;
    SystemDesc desc;
};

// Returns an empty string on success, else an error message.
std::string parse_system(const std::string& text, HostSystem& out);

// A HybridBoundedConstraintSet (include/hybrid_set.h:367-410): per location a domain box, a function (m tapes over the
// location's n variables) and a codomain box; m = 0 is the plain box set. Text form:
//   constraints
//   set <location name> m <m>
//     domain <lo> <hi> ... (n pairs)
//     codomain <lo> <hi> ... (m pairs; inf / -inf allowed)
//     f <i> <tape...> ;
struct ConstraintSetDesc {
    int loc;
    int m;
    double domain[2 * AB_MAXN];
    double codomain[2 * AB_MAXN];
    Func f[AB_MAXN];
};
struct HostConstraints {
    std::vector<ConstraintSetDesc> sets;
    std::vector<TapeOp> ops;
};
std::string parse_constraints(const std::string& text, const HostSystem& sys, HostConstraints& out);

} // namespace ab200
