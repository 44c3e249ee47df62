// ab200_kernels_s2.cu -- tuning build of the evolution kernel (one-word keys) with __launch_bounds__(64,12) and 2-way
// merges (half the register-resident stream heads). Selected at run time with AB200_TUNE=s2.
#define AB_KW 1
#define AB_VARIANT s2
#define AB_MINBLOCKS 12
#define AB_MAXSTREAMS 2
#define AB_MULSTREAMS 2
#include <cuda_runtime.h>
#include <stdio.h>
#include "ab200_evolve.cuh"
#include "ab200_device.h"

namespace ab200 { inline namespace AB_VARIANT {
#include "ab200_evolve_kernel.inc"
} }
