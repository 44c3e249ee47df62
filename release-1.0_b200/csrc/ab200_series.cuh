// ab200_series.cuh -- series-based TaylorModel functions (rec, sqrt, abs/max). See the definitions for citations.
#pragma once
#include "ab200_calc.cuh"

namespace ab200 {

AB_HDN TM tm_rec(Ctx& c, const TM& x) { fail(c, ST_UNSUPPORTED); return tm_copy(c, x); }
AB_HDN TM tm_sqrt(Ctx& c, const TM& x) { fail(c, ST_UNSUPPORTED); return tm_copy(c, x); }
AB_HDN TM tm_max(Ctx& c, const TM& x, const TM& y) { fail(c, ST_UNSUPPORTED); (void)y; return tm_copy(c, x); }

} // namespace ab200
