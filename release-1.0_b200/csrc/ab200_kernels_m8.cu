// ab200_kernels_m8.cu -- tuning build of the evolution kernel (one-word keys) with __launch_bounds__(64,12) and 8-way
// merges (half the register-resident stream heads). Selected at run time with AB200_TUNE=m8.
#define AB_KW 1
#define AB_VARIANT m8
#define AB_MINBLOCKS 12
#define AB_MAXSTREAMS 8
#define AB_MULSTREAMS 8
#include <cuda_runtime.h>
#include <stdio.h>
#include "ab200_evolve.cuh"
#include "ab200_device.h"

namespace ab200 { inline namespace AB_VARIANT {
#include "ab200_evolve_kernel.inc"
} }
