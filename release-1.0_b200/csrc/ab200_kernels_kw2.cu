// ab200_kernels_kw2.cu -- the evolution kernel built with two-word MultiIndex keys (up to 15 model variables), used for
// systems of dimension 3..5 (set models have up to 3n parameters: springs n=5, boost converter n=3).
#define AB_KW 2
#define AB_VARIANT kw2
#define AB_COUNT_FLOPS 1
#include <cuda_runtime.h>
#include <stdio.h>
#include "ab200_evolve.cuh"
#include "ab200_device.h"

namespace ab200 { inline namespace AB_VARIANT {
#include "ab200_evolve_kernel.inc"
} }
