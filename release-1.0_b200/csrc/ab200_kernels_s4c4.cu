// ab200_kernels_s4c4.cu -- tuning build: __launch_bounds__(64,12), 4-way merges, arena chunk 4 (sector per lane).
// Selected at run time with AB200_TUNE=s4c4.
#define AB_KW 1
#define AB_CHUNK 4
#define AB_VARIANT s4c4
#define AB_MINBLOCKS 12
#define AB_MAXSTREAMS 4
#define AB_MULSTREAMS 4
#include <cuda_runtime.h>
#include <stdio.h>
#include "ab200_evolve.cuh"
#include "ab200_device.h"

namespace ab200 { inline namespace AB_VARIANT {
#include "ab200_evolve_kernel.inc"
} }
