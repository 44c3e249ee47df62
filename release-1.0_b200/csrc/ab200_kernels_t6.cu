// ab200_kernels_t6.cu -- tuning build of the evolution kernel (one-word keys) with __launch_bounds__(64,6): <=168
// registers, 12 resident warps per SM. Selected at run time with AB200_TUNE=t6.
#define AB_KW 1
#define AB_VARIANT t6
#define AB_MINBLOCKS 6
#include <cuda_runtime.h>
#include <stdio.h>
#include "ab200_evolve.cuh"
#include "ab200_device.h"

namespace ab200 { inline namespace AB_VARIANT {
#include "ab200_evolve_kernel.inc"
} }
