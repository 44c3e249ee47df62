// ab200_kernels_r128.cu -- build of the evolution kernel (one-word keys) for frontiers that fill the device:
// __launch_bounds__(64,8), i.e. <=128 registers and 16 resident warps per SM, 8-way merges. Twice the cells in flight of
// the 255-register build at the same merge width (profiles/r01_variants.md). AB200_TUNE=r128 forces it.
#define AB_KW 1
#define AB_STREAMING_ACC 1   // accumulator traffic of products marked evict-first (+2 % at bench scale)
#define AB_VARIANT r128
#define AB_MINBLOCKS 8
#define AB_MAXSTREAMS 8
#define AB_MULSTREAMS 8
#include <cuda_runtime.h>
#include <stdio.h>
#include "ab200_evolve.cuh"
#include "ab200_device.h"

namespace ab200 { inline namespace AB_VARIANT {
#include "ab200_evolve_kernel.inc"
} }
