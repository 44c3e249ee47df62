// ab200_kernels_c4.cu -- tuning build: default register budget, 8-way merges, arena chunk 4 (sector per lane).
// Selected at run time with AB200_TUNE=c4.
#define AB_KW 1
#define AB_CHUNK 4
#define AB_VARIANT c4
#define AB_MINBLOCKS 1
#include <cuda_runtime.h>
#include <stdio.h>
#include "ab200_evolve.cuh"
#include "ab200_device.h"

namespace ab200 { inline namespace AB_VARIANT {
#include "ab200_evolve_kernel.inc"
} }
